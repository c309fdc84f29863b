#!/usr/bin/env python3
"""Development probe: where k_exec sits in time relative to k_fse (device timestamps).  Needs a library built with -DB200Z_PROBE:
    profiles/variants.sh probe "-DB200Z_PROBE" ; B200Z_LIB=zstd-rs_b200/variants/libb200zstd_probe.so python profiles/probe_overlap.py [config]"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import _pkg
import datagen as G

pkg = _pkg.load()
ctx = pkg.Context(0)
L = pkg.lib()
L.b200z_probe_read.restype = C.c_int
L.b200z_probe_read.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]
stream = torch.cuda.ExternalStream(ctx.stream())
name = sys.argv[1] if len(sys.argv) > 1 else "c2b"
fs = {"c2b": lambda: G.config_c2b(cache=False), "c4_4k": lambda: G.config_c4(nframes=4096, cache=False),
      "c2b_half": lambda: G.config_c2b(total_bytes=512 << 20, cache=False)}[name]()
b = pkg.Batch(ctx, fs.comp, fs.frames_io())
d_out = torch.zeros(fs.D + 64, dtype=torch.uint8, device="cuda")
for _ in range(3):
    b.run(d_out)
stream.synchronize()
rows = []
for _ in range(5):
    buf = (C.c_ulonglong * 8)()
    assert L.b200z_probe_read(None, 1) == 0
    b.run(d_out)
    stream.synchronize()
    assert L.b200z_probe_read(buf, 0) == 0
    t0 = buf[0]
    rows.append({"fse_end": (buf[1] - t0) / 1e6, "exec_first_start": (buf[2] - t0) / 1e6, "exec_last_start": (buf[4] - t0) / 1e6,
                 "exec_first_end": (buf[5] - t0) / 1e6, "exec_end": (buf[3] - t0) / 1e6})
print(name, os.environ.get("B200Z_LIB", ""), json.dumps({k: round(float(np.median([r[k] for r in rows])), 3) for k in rows[0]}))
