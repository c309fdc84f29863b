#!/bin/bash
# profiles/quick.sh [variant names...]: short device-resident bench per library variant ("base" = the in-tree build)
for V in "$@"; do
  if [ "$V" = base ]; then unset B200Z_LIB; else export B200Z_LIB=$PWD/zstd-rs_b200/variants/libb200zstd_$V.so; fi
  python bench.py --steps 10 --skip-cpu --e2e-steps 0 2>&1 | tail -1 | python -c "
import json,sys
l=sys.stdin.read()
try:
    j=json.loads(l); r=j['roofline']
    print('$V', 'ms/step %.3f' % j['ms_per_step'], 'GB/s %.1f' % j['value'], 'bit_exact', j['bit_exact'], {k: round(v,3) for k,v in r['kernel_ms'].items()}, 'done@', {k: round(v,2) for k,v in r['overlapped_completion_ms'].items()})
except Exception as e:
    print('$V FAILED', l[-600:])
"
done
