#!/bin/bash
# GPU session D of round 2: whole GPU suite, ncu (launch list + full capture of the projection GEMM + sections of the per-SNP kernel) on the
# shipped configuration, then the headline bench on the eigendecomposition-derived U (n = 50 000, cusolverMgSyevd).
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/d_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/d_pytest.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'i8_|bed_|miss_|lmm_|slice|col_' -c 60 --csv --log-file gpurun_out/d_launches.csv \
  python bench.py --u-source qr --batch 8192 --steps 2 --warmup 3 --no-e2e --no-parity --no-cpu-baseline --no-gk > gpurun_out/d_ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'i8_gemm_pair_kernel' -s 3 -c 1 -o gpurun_out/d_prof_gemm \
  python bench.py --u-source qr --batch 8192 --steps 1 --warmup 3 --no-e2e --no-parity --no-cpu-baseline --no-gk > gpurun_out/d_ncu_gemm.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'lmm_assoc_v2_kernel' -s 2 -c 1 -o gpurun_out/d_prof_lmm \
  python bench.py --u-source qr --batch 8192 --steps 1 --warmup 3 --no-e2e --no-parity --no-cpu-baseline --no-gk > gpurun_out/d_ncu_lmm.log 2>&1
( time timeout 1500 python bench.py --steps 8 --warmup 3 ) > gpurun_out/d_bench_lmm.json 2> gpurun_out/d_bench_lmm.err
ls -la gpurun_out | tail -12
tail -5 gpurun_out/d_pytest.log
head -c 600 gpurun_out/d_bench_lmm.json; echo
tail -5 gpurun_out/d_bench_lmm.err
