#!/bin/bash
# GPU session H: whole GPU suite on the current library; source-level ncu capture of the per-SNP kernel; missing-genotype benches with the new gather.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/h_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/h_pytest.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'lmm_assoc_v2_kernel' -s 2 -c 1 -o gpurun_out/h_prof_lmm \
  python bench.py --u-source qr --batch 8192 --steps 1 --warmup 3 --no-e2e --no-parity --no-cpu-baseline --no-gk > gpurun_out/h_ncu_lmm.log 2>&1
( time timeout 600 python bench.py --u-source qr --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-gk --miss 0.001 ) > gpurun_out/h_bench_lmm_qr_miss01pct.json 2> gpurun_out/h_bench_lmm_qr_miss01pct.err
( time timeout 600 python bench.py --u-source qr --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-gk --miss 0.01 ) > gpurun_out/h_bench_lmm_qr_miss1pct.json 2> gpurun_out/h_bench_lmm_qr_miss1pct.err
du -sh gpurun_out; tail -4 gpurun_out/h_pytest.log
for f in gpurun_out/h_bench_*.json; do echo "== $f"; head -c 300 $f; echo; done
