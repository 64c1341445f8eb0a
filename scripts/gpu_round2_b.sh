#!/bin/bash
# GPU session B of round 2: eigensolver probe at n = 50 000, the new parity tests (adaptive plane count, GSL tails), the four
# bench workloads on a Haar U with the polygenic phenotype, a sections-only ncu capture of the per-SNP kernel.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( time timeout 900 python -m pytest tests/test_gpu_scale_parity.py tests/test_gpu_parity.py -q -x -k "scale or n50000 or n10000 or i8 or tails or null_model or all_modes" ) > gpurun_out/b_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/b_pytest.log
( time timeout 900 python bench.py --u-source qr --steps 4 --warmup 3 ) > gpurun_out/b_bench_lmm_qr.json 2> gpurun_out/b_bench_lmm_qr.err
( time timeout 600 python bench.py --u-source qr --steps 3 --warmup 3 --slices 5 --no-e2e --no-cpu-baseline --no-gk ) > gpurun_out/b_bench_lmm_qr_T5.json 2> gpurun_out/b_bench_lmm_qr_T5.err
( time timeout 600 python bench.py --workload gk --steps 4 --warmup 3 ) > gpurun_out/b_bench_gk.json 2> gpurun_out/b_bench_gk.err
( time timeout 900 python bench.py --workload lmm1 --steps 4 --warmup 3 ) > gpurun_out/b_bench_lmm1.json 2> gpurun_out/b_bench_lmm1.err
( time timeout 900 python bench.py --workload mv --steps 3 --warmup 3 ) > gpurun_out/b_bench_mv.json 2> gpurun_out/b_bench_mv.err
( time timeout 900 scripts/eig_probe 50000 ) > gpurun_out/b_eig_probe.log 2>&1
timeout 900 ncu --section SpeedOfLight --section ComputeWorkloadAnalysis --section MemoryWorkloadAnalysis --section WarpStateStats --section LaunchStats --section Occupancy --section SchedulerStats \
  --clock-control none --import-source on -k regex:'lmm_assoc_v2_kernel' -s 4 -c 1 -o gpurun_out/b_prof_lmm \
  python bench.py --u-source qr --batch 8192 --steps 1 --warmup 3 --no-e2e --no-parity --no-cpu-baseline --no-gk > gpurun_out/b_ncu_lmm.log 2>&1
ls -la gpurun_out | tail -20
tail -3 gpurun_out/b_pytest.log
for f in gpurun_out/b_bench_*.json; do echo "== $f"; head -c 700 $f; echo; done
tail -30 gpurun_out/b_eig_probe.log
