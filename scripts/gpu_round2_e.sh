#!/bin/bash
# GPU session E of round 2: barrier-free stage ring in the per-SNP kernel, wave-synchronised projection GEMM, dosage digit rows.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( time timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale_parity.py -q -x -k "dosage or assoc or null_model or i8 or cuda_path or properties or scale or n50000 or n10000 or nan_rule or subbatch or mouse_hs1940_gk or bxd" ) > gpurun_out/e_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/e_pytest.log
( time timeout 600 python bench.py --u-source qr --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-gk ) > gpurun_out/e_bench_lmm_qr.json 2> gpurun_out/e_bench_lmm_qr.err
( time timeout 600 python bench.py --u-source qr --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-gk --no-parity --opt gemm_wave_sync=0 ) > gpurun_out/e_bench_lmm_qr_nosync.json 2> gpurun_out/e_bench_lmm_qr_nosync.err
( time timeout 600 python bench.py --u-source qr --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-gk --no-parity --opt gemm_panel=12 ) > gpurun_out/e_bench_lmm_qr_panel12.json 2> gpurun_out/e_bench_lmm_qr_panel12.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'i8_gemm_pair_kernel' -s 3 -c 1 -o gpurun_out/e_prof_gemm \
  python bench.py --u-source qr --batch 8192 --steps 1 --warmup 3 --no-e2e --no-parity --no-cpu-baseline --no-gk > gpurun_out/e_ncu_gemm.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'lmm_assoc_v2_kernel' -s 2 -c 1 -o gpurun_out/e_prof_lmm \
  python bench.py --u-source qr --batch 8192 --steps 1 --warmup 3 --no-e2e --no-parity --no-cpu-baseline --no-gk > gpurun_out/e_ncu_lmm.log 2>&1
( time timeout 900 python bench.py --workload lmm1 --steps 4 --warmup 3 --no-cpu-baseline ) > gpurun_out/e_bench_lmm1.json 2> gpurun_out/e_bench_lmm1.err
ls -la gpurun_out | tail -12
tail -5 gpurun_out/e_pytest.log
for f in gpurun_out/e_bench_*.json; do echo "== $f"; head -c 300 $f; echo; done
