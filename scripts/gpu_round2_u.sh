#!/bin/bash
# final sanity of the committed tree: smoke() + the CLI / plugin tests that session R did not re-run
mkdir -p gpurun_out
( time timeout 120 python __graft_entry__.py smoke ) > gpurun_out/u_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/u_smoke.log
( time timeout 230 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cli or plugin" ) > gpurun_out/u_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/u_pytest.log
tail -3 gpurun_out/u_smoke.log; tail -4 gpurun_out/u_pytest.log
