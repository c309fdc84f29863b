#!/usr/bin/env python3
"""Diagnostic: device-resident pass time per rank before / after NCCL initialisation (run under torchrun)."""
import os, sys, time
import numpy as np
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import _pkg, datagen as G
rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
pkg = _pkg.load()
fs = G.config_c2b(total_bytes=2048 * 131072, frame_bytes=131072, cache=False)
ctx = pkg.Context(local)
stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local))
batch = pkg.Batch(ctx, fs.comp, fs.frames_io())
d_out = torch.empty(fs.D + 64, dtype=torch.uint8, device="cuda")
def timeit(tag, n=10):
    for _ in range(3): batch.run(d_out)
    stream.synchronize(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(stream)
    for _ in range(n): batch.run(d_out)
    e1.record(stream)
    t1 = time.perf_counter()
    stream.synchronize(); torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"rank {rank} {tag}: {e0.elapsed_time(e1) / n:.3f} ms/pass (host enqueue {1e3 * (t1 - t0) / n:.3f} ms/pass, wall {1e3 * (t2 - t0) / n:.3f})", flush=True)
print(rank, {k: v for k, v in os.environ.items() if any(x in k for x in ("CUDA", "NCCL", "OMP", "TORCH_NCCL"))}, flush=True)
timeit("before init_process_group")
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    timeit("after init_process_group")
    dist.barrier(); torch.cuda.synchronize()
    timeit("after barrier")
    t = torch.ones(1, device="cuda"); dist.all_reduce(t); torch.cuda.synchronize()
    timeit("after all_reduce")
    dist.destroy_process_group()
    timeit("after destroy")
