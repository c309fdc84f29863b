#!/bin/bash
# Run under gpurun (one GPU).  Captures (a) every launch of one bench pass with its device time, (b) a full-set
# profile of each hot kernel.  Outputs land in gpurun_out/; the summaries judged are copied to profiles/.
set -x
mkdir -p gpurun_out
TAG=${1:-r01}
FRAMES=${2:-8192}
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 3 --skip-cpu --e2e-steps 0 --frames $FRAMES > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
for K in k_fse k_huf k_exec k_setup; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o gpurun_out/prof_${K}_${TAG} \
      python bench.py --steps 1 --warmup 3 --skip-cpu --e2e-steps 0 --frames $FRAMES > gpurun_out/ncu_${K}_${TAG}.log 2>&1
done
ls -la gpurun_out
