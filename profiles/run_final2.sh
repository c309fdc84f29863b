#!/bin/bash
# Refresh of the round-end evidence after a change that touches one kernel: bench line, launch list, one full capture of $2.
set -x
mkdir -p gpurun_out
TAG=${1:-r01h}; K=${2:-k_setup}
python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 3 --skip-cpu --e2e-steps 0 > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o gpurun_out/prof_${K}_${TAG} \
    python bench.py --steps 1 --warmup 3 --skip-cpu --e2e-steps 0 > gpurun_out/ncu_${K}_${TAG}.log 2>&1
cut -c1-300 gpurun_out/bench_${TAG}.json
