#!/usr/bin/env python3
"""Dynamic warp-instructions per CUDA source line (needs -lineinfo + --import-source on). usage: line_mix.py rep [topN]"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
agg = collections.Counter(); src = {}; stall = collections.Counter()
hdr = None; cur_file = ""
for r in rows:
    if r and r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": hdr = {h: i for i, h in enumerate(r)}; continue
    if hdr is None or len(r) < len(hdr) - 2: continue
    try:
        ln = int(r[0]); ex = int(r[hdr["Instructions Executed"]])
    except Exception:
        continue
    key = (cur_file, ln)
    agg[key] += ex; src[key] = r[1].strip()
    try: stall[key] += int(r[hdr["# Samples"]])
    except Exception: pass
tot = sum(agg.values()); st = sum(stall.values()) or 1
print("total dynamic warp instr", tot)
for key, c in agg.most_common(topn):
    print(f"{c/tot*100:5.1f}%  samp {stall[key]/st*100:5.1f}%  {key[0]}:{key[1]:4d}  {src[key][:110]}")
