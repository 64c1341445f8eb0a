#!/bin/bash
# GPU session C of round 2: whole GPU suite on the interpolated-refinement kernel + cusolverMg eigensolver, the headline bench on the
# eigendecomposition-derived U (n = 50 000), A/B of the interpolated refinement, the -lmm 1 line.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/c_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c_pytest.log
( time timeout 600 python bench.py --u-source qr --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-gk ) > gpurun_out/c_bench_lmm_qr_interp.json 2> gpurun_out/c_bench_lmm_qr_interp.err
( time timeout 600 python bench.py --u-source qr --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-gk --no-parity --opt lmm_interp=0 ) > gpurun_out/c_bench_lmm_qr_exact.json 2> gpurun_out/c_bench_lmm_qr_exact.err
( time timeout 900 python bench.py --workload lmm1 --steps 4 --warmup 3 ) > gpurun_out/c_bench_lmm1.json 2> gpurun_out/c_bench_lmm1.err
( time timeout 1500 python bench.py --steps 8 --warmup 3 ) > gpurun_out/c_bench_lmm.json 2> gpurun_out/c_bench_lmm.err
ls -la gpurun_out | tail -12
tail -5 gpurun_out/c_pytest.log
for f in gpurun_out/c_bench_*.json; do echo "== $f"; head -c 400 $f; echo; done
tail -5 gpurun_out/c_bench_lmm.err
