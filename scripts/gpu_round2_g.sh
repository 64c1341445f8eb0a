#!/bin/bash
# GPU session G: rotating TMA producer in the per-SNP kernel, unrolled hole gather; ncu of the MAIN projection launch (every batch now
# has a second, usually empty, hole-pass launch of the same kernel: -s counts both).
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale_parity.py -q -x -k "cta_pair or assoc or i8 or properties or n50000 or n10000 or subbatch or nan_rule or exact_x or dosage or cuda_path or mouse_hs1940_gk or bxd" ) > gpurun_out/g_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/g_pytest.log
( time timeout 600 python bench.py --u-source qr --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-gk ) > gpurun_out/g_bench_lmm_qr.json 2> gpurun_out/g_bench_lmm_qr.err
( time timeout 600 python bench.py --u-source qr --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-gk --no-parity --miss 0.001 ) > gpurun_out/g_bench_lmm_qr_miss01pct.json 2> gpurun_out/g_bench_lmm_qr_miss01pct.err
( time timeout 900 python bench.py --workload lmm1 --steps 4 --warmup 3 --no-cpu-baseline ) > gpurun_out/g_bench_lmm1.json 2> gpurun_out/g_bench_lmm1.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'i8_gemm_pair_kernel' -s 6 -c 1 -o gpurun_out/g_prof_gemm \
  python bench.py --u-source qr --batch 8192 --steps 1 --warmup 3 --no-e2e --no-parity --no-cpu-baseline --no-gk > gpurun_out/g_ncu_gemm.log 2>&1
timeout 900 ncu --section SpeedOfLight --section ComputeWorkloadAnalysis --section MemoryWorkloadAnalysis --section WarpStateStats --section SchedulerStats --section LaunchStats --section Occupancy \
  --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'lmm_assoc_v2_kernel' -s 2 -c 1 -o gpurun_out/g_prof_lmm \
  python bench.py --u-source qr --batch 8192 --steps 1 --warmup 3 --no-e2e --no-parity --no-cpu-baseline --no-gk > gpurun_out/g_ncu_lmm.log 2>&1
du -sh gpurun_out; tail -4 gpurun_out/g_pytest.log
for f in gpurun_out/g_bench_*.json; do echo "== $f"; head -c 300 $f; echo; done
