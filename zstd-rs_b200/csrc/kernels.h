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
    // execution scheduling (one device buffer: ticket counter, resume[nframes], then the frame list of k_exec_cta)
    uint32_t *ticket;             // 16 bytes; zeroed with resume[] before every pass (sched_bytes)
    uint32_t *resume;             // [nframes] first block k_exec still has to execute (written by k_exec_cta)
    const uint32_t *cta_frames;   // [n_cta_frames] frames executed by k_exec_cta
    uint32_t n_cta_frames;
    uint32_t sched_bytes;         // 16 + 4 * nframes
};

int init_kernels();  // per-device function attributes (dynamic shared memory); call once per context
int launch_predefined(FseSlot *predef, cudaStream_t s);
constexpr int kNumStages = 5;
extern const char *const kStageNames[kNumStages];
int launch_stage(const PipelineArgs &a, int stage, cudaStream_t s);
int launch_pipeline(const PipelineArgs &a, cudaStream_t s);
int launch_checksum(const PipelineArgs &a, cudaStream_t s);   // optional 5th stage: XXH64 of every frame's plaintext
struct PipelineStreams { cudaStream_t main, side; cudaEvent_t fork, join; };
int launch_pipeline_overlapped(const PipelineArgs &a, const PipelineStreams &ps);   // k_huf on `side` beside k_fse on `main`
uint32_t pipeline_launch_count(const PipelineArgs &a);

}  // namespace b200z
