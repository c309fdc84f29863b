set -x
mkdir -p gpurun_out
TAG=${1:-r01e}
ncu --set full --clock-control none --import-source on -k regex:k_exec -s 3 -c 1 -f -o gpurun_out/prof_k_exec_${TAG} \
      python bench.py --steps 1 --warmup 3 --skip-cpu --e2e-steps 0 > gpurun_out/ncu_k_exec_${TAG}.log 2>&1
