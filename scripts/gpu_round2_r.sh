#!/bin/bash
# GPU session R: hoisted passes accumulate x'x only (linear sums from the side GEMM), final pass takes its linear sums from the exact interpolants.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale_parity.py -q -x -k "not cli" ) > gpurun_out/r_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r_pytest.log
( time timeout 600 python bench.py --u-source qr --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-gk ) > gpurun_out/r_bench_lmm_qr.json 2> gpurun_out/r_bench_lmm_qr.err
( time timeout 900 python bench.py --workload lmm1 --steps 4 --warmup 3 --no-cpu-baseline ) > gpurun_out/r_bench_lmm1.json 2> gpurun_out/r_bench_lmm1.err
( time timeout 1500 python bench.py --steps 8 --warmup 3 --no-gk ) > gpurun_out/r_bench_lmm.json 2> gpurun_out/r_bench_lmm.err
tail -4 gpurun_out/r_pytest.log
for f in gpurun_out/r_bench_*.json; do echo "== $f"; head -c 300 $f; echo; tail -2 ${f%.json}.err; done
