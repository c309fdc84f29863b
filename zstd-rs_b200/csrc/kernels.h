// kernels.h -- launch interface between the host library (api.cpp / plan.cpp) and kernels.cu
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "b200z_types.h"

namespace b200z {

struct PipelineArgs {
    const BlockDesc *descs;   // [nblocks]  device
    BlockAux *aux;            // [nblocks]  device
    const FrameDesc *frames;  // [nframes]  device
    FrameState *states;       // [nframes]  device, in/out
    const uint8_t *input;     // device, readable up to the next 4-byte boundary past the last frame
    uint8_t *lit_scratch;     // device
    uint32_t *seq_scratch;    // device, 3 x u32 per sequence
    uint8_t *output;          // device
    uint64_t output_cap;
    uint32_t nblocks;
    uint32_t nframes;
};

int init_kernels();  // per-device function attributes (dynamic shared memory); call once per context
int launch_predefined(FseSlot *predef, cudaStream_t s);
constexpr int kNumStages = 4;
extern const char *const kStageNames[kNumStages];
int launch_stage(const PipelineArgs &a, int stage, cudaStream_t s);
int launch_pipeline(const PipelineArgs &a, cudaStream_t s);
int launch_checksum(const PipelineArgs &a, cudaStream_t s);   // optional 5th stage: XXH64 of every frame's plaintext
struct PipelineStreams { cudaStream_t main, side; cudaEvent_t fork, join; };   // only `main` is used since k_exec became a programmatic dependent of k_fse
int launch_pipeline_overlapped(const PipelineArgs &a, const PipelineStreams &ps);   // k_exec beside k_fse (programmatic dependent launch)
int launch_fse_exec(const PipelineArgs &a, cudaStream_t s);                          // the overlapped pair alone
uint32_t pipeline_launch_count(const PipelineArgs &a);

}  // namespace b200z
