#!/bin/bash
# GPU session Q (final measurements of round 2, 3-plane projection with exact linear x-sums): whole GPU suite, ncu launch list + captures of both hot kernels, one line per BASELINE config,
# the reference arm, the missing-genotype variants.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/q_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/q_pytest.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'i8_|bed_|miss_|lmm_|slice|col_|hole' -c 60 --csv --log-file gpurun_out/q_launches.csv \
  python bench.py --u-source qr --batch 8192 --steps 2 --warmup 3 --no-e2e --no-parity --no-cpu-baseline --no-gk > gpurun_out/q_ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'i8_gemm_pair_kernel' -s 6 -c 1 -o gpurun_out/q_prof_gemm \
  python bench.py --u-source qr --batch 8192 --steps 1 --warmup 3 --no-e2e --no-parity --no-cpu-baseline --no-gk > gpurun_out/q_ncu_gemm.log 2>&1
timeout 900 ncu --section SpeedOfLight --section ComputeWorkloadAnalysis --section MemoryWorkloadAnalysis --section WarpStateStats --section SchedulerStats --section LaunchStats --section Occupancy \
  --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'lmm_assoc_v2_kernel' -s 2 -c 1 -o gpurun_out/q_prof_lmm \
  python bench.py --u-source qr --batch 8192 --steps 1 --warmup 3 --no-e2e --no-parity --no-cpu-baseline --no-gk > gpurun_out/q_ncu_lmm.log 2>&1
ls -la gpurun_out/*.ncu-rep
( time timeout 1500 python bench.py --steps 16 --warmup 3 ) > gpurun_out/q_bench_lmm.json 2> gpurun_out/q_bench_lmm.err
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/q_bench_reference.json 2> gpurun_out/q_bench_reference.err
( time timeout 900 python bench.py --workload lmm1 --steps 8 --warmup 3 ) > gpurun_out/q_bench_lmm1.json 2> gpurun_out/q_bench_lmm1.err
( time timeout 900 python bench.py --workload mv --steps 6 --warmup 3 ) > gpurun_out/q_bench_mv.json 2> gpurun_out/q_bench_mv.err
( time timeout 900 python bench.py --workload gk --steps 8 --warmup 3 ) > gpurun_out/q_bench_gk.json 2> gpurun_out/q_bench_gk.err
( time timeout 600 python bench.py --u-source qr --steps 4 --warmup 3 --no-cpu-baseline --no-gk ) > gpurun_out/q_bench_lmm_qr.json 2> gpurun_out/q_bench_lmm_qr.err
( time timeout 600 python bench.py --u-source qr --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-gk --miss 0.01 ) > gpurun_out/q_bench_lmm_qr_miss1pct.json 2> gpurun_out/q_bench_lmm_qr_miss1pct.err
( time timeout 600 python bench.py --u-source qr --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-gk --miss 0.001 ) > gpurun_out/q_bench_lmm_qr_miss01pct.json 2> gpurun_out/q_bench_lmm_qr_miss01pct.err
du -sh gpurun_out; tail -4 gpurun_out/q_pytest.log
for f in gpurun_out/q_bench_*.json; do echo "== $f"; head -c 260 $f; echo; done
