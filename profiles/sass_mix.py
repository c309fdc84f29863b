#!/usr/bin/env python3
"""Instruction mix of the hot loop of a kernel from an ncu report (source page). usage: sass_mix.py report.ncu-rep [min_frac]"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]; frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.12
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]; ci = {h: i for i, h in enumerate(hdr)}
S, E, T = ci["Source"], ci["Instructions Executed"], ci["Thread Instructions Executed"]
body = [r for r in rows[2:] if len(r) > T and r[E].isdigit()]
mx = max(int(r[E]) for r in body)
hot = [r for r in body if int(r[E]) >= frac * mx]
agg = collections.Counter(); tot = 0
for r in hot:
    toks = r[S].split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    agg[op.split(".")[0]] += int(r[E]); tot += int(r[E])
alltot = sum(int(r[E]) for r in body)
print(f"max exec/instr {mx}; hot static instrs {len(hot)} of {len(body)}; hot dynamic {tot} of {alltot} ({tot/alltot:.2%})")
for op, c in agg.most_common(30):
    print(f"  {op:10s} {c:12d} {c/tot*100:5.1f}%")
if len(sys.argv) > 3:
    for r in hot: print(r[E].rjust(10), r[ci['stall_wait']].rjust(5) if 'stall_wait' in ci else '', r[S])
