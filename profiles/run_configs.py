#!/usr/bin/env python3
"""Runs the other BASELINE.json configs (C2a, C3, C4, C5) at moderate size on one GPU: bit-exact check + device-resident GB/s.
These are parity-test configs, not bench lines; the numbers go to DESIGN.md for orientation."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import _pkg
import datagen as G

pkg = _pkg.load()
ctx = pkg.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream())
out = {}
sets = {
    "C2a_64x16MiB_chained": lambda: G.config_c2a(total_bytes=1 << 30, nframes=64, cache=False),
    "C2a_1frame_64MiB_chained": lambda: G.config_c2a(total_bytes=64 << 20, nframes=1, cache=False),
    "C3_10000x64KiB_huffman": lambda: G.config_c3(nframes=10000, cache=False),
    "C4_1024x1MiB_silesia_mix": lambda: G.config_c4(nframes=1024, cache=False),
    "C5_20000_small_dict": lambda: G.config_c5(nframes=20000, cache=False),
}
only = sys.argv[1:] or list(sets)
for name in only:
    t0 = time.time()
    fs = sets[name]()
    gen_s = time.time() - t0
    D = pkg.Dictionary.raw_content(ctx, 1, fs.raw_dict.tobytes()) if fs.raw_dict is not None else None
    b = pkg.Batch(ctx, fs.comp, fs.frames_io(), forced_dict=D)
    d_out = torch.zeros(fs.D + 64, dtype=torch.uint8, device="cuda")
    for _ in range(2):
        b.run(d_out)
    stream.synchronize()
    res = b.finish()
    ok = bool((res["status"] == 0).all()) and bool(np.array_equal(d_out[:fs.D].cpu().numpy(), fs.plain))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 5
    e0.record(stream)
    for _ in range(n):
        b.run(d_out)
    e1.record(stream)
    stream.synchronize()
    ms = e0.elapsed_time(e1) / n
    prof = b.run_profile(d_out)
    info = b.info()
    out[name] = {"bit_exact": ok, "frames": fs.nframes, "blocks": info["blocks"], "C": fs.C, "D": fs.D, "ratio": fs.D / fs.C, "ms": ms,
                 "decompressed_GBps": fs.D / ms / 1e6, "kernel_ms": prof, "gen_s": gen_s}
    print(name, json.dumps(out[name]), flush=True)
    b.close()
    del d_out
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "configs.json"), "w"), indent=1)
