"""Throughput of the two-phenotype mvLMM entry point on synthetic data (BASELINE config 5 shape: n = 10 000).
usage: python scripts/mv_bench.py [n] [snps]   -> one JSON line"""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import gemma_b200  # noqa: E402
from gemma_b200 import synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
l = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
rng = np.random.default_rng(5)
v = rng.standard_normal(n); v /= np.linalg.norm(v)
U = np.eye(n) - 2.0 * np.outer(v, v)
ev = synth.spectrum_like_kinship(n, 7)
bed, G = synth.make_bed(n, l + 64, seed=99)
g = U @ (np.sqrt(ev)[:, None] * rng.standard_normal((n, 2)))
Y = 0.8 * g @ np.array([[1.0, 0.3], [0.0, 0.8]]) + rng.standard_normal((n, 2))
Y[:, 0] += 0.05 * (G[l] - G[l].mean())
W = np.ones((n, 1))
ctx = gemma_b200.Context(0)
t0 = time.time(); ctx.mvlmm_setup(U, ev, W, Y); t_setup = time.time() - t0
t0 = time.time(); nm = ctx.mvlmm_null(); t_null = time.time() - t0
ctx.mvlmm_batch_bed(bed[:256], n)                                   # warm-up (U planes, kernels)
ctx.profile_enable(True); ctx.profile_reset()
t0 = time.time(); out = ctx.mvlmm_batch_bed(bed[:l], n); t_batch = time.time() - t0
lmm_ms, _ = ctx.profile_get("lmm"); utx_ms, _ = ctx.profile_get("utx")
print(json.dumps({"what": "mvLMM, 2 phenotypes, -lmm 1 (Wald), c=1, PLINK rows through gb200_mvlmm_batch_bed", "n": n, "snps": l,
                  "setup_s": round(t_setup, 2), "null_model_s": round(t_null, 3), "batch_s_host_call": round(t_batch, 3),
                  "per_snp_kernel_ms": round(lmm_ms, 2), "projection_ms": round(utx_ms, 2), "snps_per_s": round(l / t_batch, 1),
                  "snps_per_s_kernel_only": round(l / (lmm_ms * 1e-3), 1), "frac_p_lt_1e-3": float((out[:, 5] < 1e-3).mean())}))
