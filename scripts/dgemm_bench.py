#!/usr/bin/env python
"""FP64 GEMM seam (gb200_dgemm, DMMA kernel) kernel-only rate: C = A^T B and C = A B at 6144^3, and the lower-triangle kinship form.
  python scripts/dgemm_bench.py   -> one JSON line"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gemma_b200
n = int(sys.argv[1]) if len(sys.argv) > 1 else 6144
rng = np.random.default_rng(1)
A = rng.standard_normal((n, n)); B = rng.standard_normal((n, n)); C = np.zeros((n, n))
ctx = gemma_b200.Context(0)
out = {"n": n}
ref = None
for ta, tb in (("N", "N"), ("T", "N"), ("N", "T")):
    ctx.dgemm(ta, tb, 1.0, A, B, 0.0, C)                    # warm-up
    ctx.profile_enable(True); ctx.profile_reset()
    for _ in range(3):
        ctx.dgemm(ta, tb, 1.0, A, B, 0.0, C)
    ms, k = ctx.profile_get("dgemm")
    ctx.profile_enable(False)
    out[ta + tb] = {"tflops": 2.0 * n ** 3 * k / (ms * 1e-3) / 1e12, "ms": ms / k}
    if ta == "N" and tb == "N":
        ref = A[:64] @ B
        out["max_err_vs_numpy"] = float(np.abs(C[:64] - ref).max())
print(json.dumps(out))
