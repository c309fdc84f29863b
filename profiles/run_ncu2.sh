set -x
mkdir -p gpurun_out
for K in k_exec k_fse; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o gpurun_out/prof_${K}_r01d \
      python bench.py --steps 1 --warmup 3 --skip-cpu --e2e-steps 0 > gpurun_out/ncu_${K}_r01d.log 2>&1
done
ls -la gpurun_out
