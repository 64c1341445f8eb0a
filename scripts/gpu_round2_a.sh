#!/bin/bash
# GPU session A of round 2: full GPU test suite, the new scale-parity tests, the headline bench on the eigh-derived U,
# kernel-isolated power probe, ncu launch list + full capture of the two hot kernels.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/a_pytest.log
( time timeout 1500 python bench.py --steps 8 --warmup 3 ) > gpurun_out/a_bench_lmm.json 2> gpurun_out/a_bench_lmm.err
timeout 600 python scripts/power_probe.py > gpurun_out/a_power.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'i8_|bed_|miss_|lmm_|slice|col_' -c 80 --csv --log-file gpurun_out/a_launches.csv \
  python bench.py --u-source qr --batch 8192 --steps 2 --warmup 3 --no-e2e --no-parity --no-cpu-baseline --no-gk > gpurun_out/a_ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'i8_gemm_pair_kernel|lmm_assoc_v2_kernel' -s 6 -c 2 -o gpurun_out/a_prof \
  python bench.py --u-source qr --batch 8192 --steps 1 --warmup 3 --no-e2e --no-parity --no-cpu-baseline --no-gk > gpurun_out/a_ncu_full.log 2>&1
timeout 300 python scripts/check_pair2.py > gpurun_out/a_pair2_check.log 2>&1; echo "check_pair2 rc=$?" >> gpurun_out/a_pair2_check.log
if grep -q "pair2 ok" gpurun_out/a_pair2_check.log; then
  timeout 600 python scripts/power_probe.py --opt gemm_groups=2 --out gpurun_out/power_probe_groups2.json > gpurun_out/a_power2.log 2>&1
  timeout 600 python scripts/power_probe.py --opt gemm_groups=2 --opt gemm_panel=4 --out gpurun_out/power_probe_groups2_p4.json > gpurun_out/a_power2b.log 2>&1
fi
ls -la gpurun_out | tail -20
tail -3 gpurun_out/a_pytest.log
head -c 600 gpurun_out/a_bench_lmm.json
