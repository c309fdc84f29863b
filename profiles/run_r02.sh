#!/bin/bash
# Round-2 evidence on one GPU: parity tests, the bench line (+ reference arm), the ncu launch list of the bench command,
# one full-set ncu capture per kernel, the other BASELINE configs at full size.  Everything lands in gpurun_out/.
set -x
mkdir -p gpurun_out
TAG=${1:-r02}
python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -4 > gpurun_out/pytest_gpu_${TAG}.log
python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2>> gpurun_out/bench_${TAG}.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 3 --skip-cpu --e2e-steps 0 > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
# per-kernel captures: sequential launches (quick2 -> run_profile) so that every kernel is alone on the device
for K in k_exec k_fse k_huf k_setup k_xxh64; do
  QUICK_MODES=auto B200Z_QUICK_CHECKSUM=1 ncu --set full --clock-control none --import-source on -k regex:"${K}\$" -s 8 -c 1 -f -o gpurun_out/prof_${K}_${TAG} \
      python profiles/quick2.py c2b > gpurun_out/ncu_${K}_${TAG}.log 2>&1
done
QUICK_MODES=auto ncu --set full --clock-control none --import-source on -k regex:"k_exec_cta\$" -s 2 -c 1 -f -o gpurun_out/prof_k_exec_cta_${TAG} \
    python profiles/quick2.py c2a64 > gpurun_out/ncu_k_exec_cta_${TAG}.log 2>&1
# where k_exec sits in time relative to k_fse (device timestamps; needs profiles/variants.sh probe "-DB200Z_PROBE" built beforehand)
if [ -f zstd-rs_b200/variants/libb200zstd_probe.so ]; then
  for P in 1 0; do
    B200Z_EXEC_PDL=$P B200Z_LIB=$PWD/zstd-rs_b200/variants/libb200zstd_probe.so python profiles/probe_overlap.py c2b 2>&1 | tail -1 | sed "s/^/PDL=$P /" >> gpurun_out/probe_${TAG}.txt
  done
  B200Z_LIB=$PWD/zstd-rs_b200/variants/libb200zstd_probe.so python profiles/probe_overlap.py c4_4k 2>&1 | tail -1 >> gpurun_out/probe_${TAG}.txt
fi
for C in c2a c3 c4 c5; do
  python bench.py --config $C --skip-cpu --e2e-steps 2 --steps 5 >> gpurun_out/configs_${TAG}.jsonl 2>> gpurun_out/bench_${TAG}.err
done
tail -3 gpurun_out/pytest_gpu_${TAG}.log; cut -c1-700 gpurun_out/bench_${TAG}.json; ls gpurun_out | head -40
