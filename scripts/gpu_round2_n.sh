#!/bin/bash
# GPU session N: double-buffered FP64 DMMA GEMM (seam tests, BIMBAM kinship, G x E, -lm), exact x-sums with covariates, gk e2e with pinned K.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( time timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "dgemm or kinship or exact_x or gxe or lm_entry or eigh or null_model or mvlmm or state_errors or dosage" ) > gpurun_out/n_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/n_pytest.log
timeout 300 python scripts/dgemm_bench.py > gpurun_out/n_dgemm_bench.json 2> gpurun_out/n_dgemm_bench.err
( time timeout 900 python bench.py --workload gk --steps 6 --warmup 3 ) > gpurun_out/n_bench_gk.json 2> gpurun_out/n_bench_gk.err
tail -4 gpurun_out/n_pytest.log; cat gpurun_out/n_dgemm_bench.json; head -c 300 gpurun_out/n_bench_gk.json
