#!/bin/bash
# GPU session K: 10-node x interpolants + one exact final pass in the per-SNP kernel.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale_parity.py -q -x -k "not cli" ) > gpurun_out/k_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/k_pytest.log
( time timeout 600 python bench.py --u-source qr --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-gk ) > gpurun_out/k_bench_lmm_qr.json 2> gpurun_out/k_bench_lmm_qr.err
( time timeout 900 python bench.py --workload lmm1 --steps 4 --warmup 3 --no-cpu-baseline ) > gpurun_out/k_bench_lmm1.json 2> gpurun_out/k_bench_lmm1.err
( time timeout 1500 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-gk ) > gpurun_out/k_bench_lmm.json 2> gpurun_out/k_bench_lmm.err
du -sh gpurun_out; tail -4 gpurun_out/k_pytest.log
for f in gpurun_out/k_bench_*.json; do echo "== $f"; head -c 300 $f; echo; done
