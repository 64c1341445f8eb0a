#!/bin/bash
# GPU session P: exact linear x-sums from a side GEMM (LmmConst::xsum) + 3-plane projection.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale_parity.py -q -x -k "n50000 or n10000 or exact_x or assoc_kernel_all or properties or i8_path or nan_rule or cuda_path" ) > gpurun_out/p_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/p_pytest.log
( time timeout 600 python bench.py --u-source qr --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-gk ) > gpurun_out/p_bench_lmm_qr.json 2> gpurun_out/p_bench_lmm_qr.err
( time timeout 600 python bench.py --u-source qr --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-gk --miss 0.01 ) > gpurun_out/p_bench_lmm_qr_miss1pct.json 2> gpurun_out/p_bench_lmm_qr_miss1pct.err
( time timeout 900 python bench.py --workload lmm1 --steps 4 --warmup 3 --no-cpu-baseline ) > gpurun_out/p_bench_lmm1.json 2> gpurun_out/p_bench_lmm1.err
tail -4 gpurun_out/p_pytest.log
for f in gpurun_out/p_bench_*.json; do echo "== $f"; head -c 300 $f; echo; tail -2 ${f%.json}.err; done
