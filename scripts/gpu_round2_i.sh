#!/bin/bash
# GPU session I: per-SNP kernel with one chunk of lookahead (two chunks of slack in the stage ring), suspend-hinted waits, register double buffer.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale_parity.py -q -x -k "assoc or properties or n50000 or n10000 or exact_x or cuda_path or nan_rule or subbatch or bxd" ) > gpurun_out/i_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/i_pytest.log
( time timeout 600 python bench.py --u-source qr --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-gk ) > gpurun_out/i_bench_lmm_qr.json 2> gpurun_out/i_bench_lmm_qr.err
( time timeout 900 python bench.py --workload lmm1 --steps 4 --warmup 3 --no-cpu-baseline ) > gpurun_out/i_bench_lmm1.json 2> gpurun_out/i_bench_lmm1.err
timeout 900 ncu --section SpeedOfLight --section ComputeWorkloadAnalysis --section MemoryWorkloadAnalysis --section WarpStateStats --section SchedulerStats --section LaunchStats --section Occupancy \
  --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'lmm_assoc_v2_kernel' -s 2 -c 1 -o gpurun_out/i_prof_lmm \
  python bench.py --u-source qr --batch 8192 --steps 1 --warmup 3 --no-e2e --no-parity --no-cpu-baseline --no-gk > gpurun_out/i_ncu_lmm.log 2>&1
du -sh gpurun_out; tail -4 gpurun_out/i_pytest.log
for f in gpurun_out/i_bench_*.json; do echo "== $f"; head -c 300 $f; echo; done
