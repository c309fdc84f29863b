#!/bin/bash
# Round-end evidence on one GPU: parity tests, the bench line, the ncu launch list of the same command, one full-set capture per kernel.
set -x
mkdir -p gpurun_out
TAG=${1:-r01g}
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/pytest_gpu_${TAG}.log
python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2>> gpurun_out/bench_${TAG}.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 3 --skip-cpu --e2e-steps 0 > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
for K in k_exec k_fse k_huf k_setup; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o gpurun_out/prof_${K}_${TAG} \
      python bench.py --steps 1 --warmup 3 --skip-cpu --e2e-steps 0 > gpurun_out/ncu_${K}_${TAG}.log 2>&1
done
python profiles/run_configs.py > gpurun_out/configs_${TAG}.log 2>&1
tail -3 gpurun_out/pytest_gpu_${TAG}.log; cut -c1-600 gpurun_out/bench_${TAG}.json; ls gpurun_out | head -30
