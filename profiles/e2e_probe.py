import os, sys, time, numpy as np, torch
sys.path.insert(0, os.getcwd())
import _pkg, datagen as G
pkg = _pkg.load()
fs = G.config_c2b(cache=False)
ctx = pkg.Context(0)
h_in = torch.from_numpy(fs.comp.copy()).pin_memory()
h_out = torch.empty(fs.D + 64, dtype=torch.uint8).pin_memory()
io = fs.frames_io()
for flags in (0, 1):
    ctx.set_flags(flags)
    ts = []
    for i in range(8):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = pkg.decode_frames(ctx, h_in, io, h_out)
        ts.append((time.perf_counter() - t0) * 1e3)
    print("flags", flags, "e2e ms", [round(t, 1) for t in ts], flush=True)
os.environ["B200Z_TRACE"] = "1"
pkg.decode_frames(ctx, h_in, io, h_out)
for cb in (128 << 20, 512 << 20):
    os.environ["B200Z_PIPELINE_CHUNK_BYTES"] = str(cb); os.environ.pop("B200Z_TRACE", None)
    ts = []
    for i in range(5):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        pkg.decode_frames(ctx, h_in, io, h_out)
        ts.append((time.perf_counter() - t0) * 1e3)
    print("chunk", cb >> 20, "MiB e2e ms", [round(t, 1) for t in ts], flush=True)
