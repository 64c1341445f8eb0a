#!/usr/bin/env python
"""A/B correctness of the projection kernels on one GPU: gemm_groups = 2 (two eigenvector groups per tile, hole pass on the tensor
pipe) against gemm_groups = 1 (one group, FP64 hole fix-up) and the FP64 path, with missing genotypes and a masked cohort."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gemma_b200
from gemma_b200 import synth
from oracle import oracle as O

for n_total, n_drop, l, miss in ((2500, 3, 517, 0.01), (1100, 0, 130, 0.0), (3000, 11, 700, 0.05)):
    rng = np.random.default_rng(n_total)
    mask = np.ones(n_total, dtype=np.uint8)
    if n_drop:
        mask[rng.choice(n_total, n_drop, replace=False)] = 0
    n = int(mask.sum())
    Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    ev = synth.spectrum_like_kinship(n, 5)
    c = gemma_b200.Context(0)
    c.lmm_setup(Q, ev, np.ones((n, 1)), rng.standard_normal(n))
    bed, G = synth.make_bed(n_total, l, seed=n_total + 1, miss_rate=miss)
    X = O.lmm_impute(np.where(G < 0, np.nan, G)[:, mask == 1])
    ref = (Q.T @ X).T
    scale = np.abs(ref).max()
    c.set_option("utx_path", 2)
    out = {}
    for gg in (1, 2):
        c.set_option("gemm_groups", gg)
        got = c.lmm_project_bed(bed, n_total, mask if n_drop else None)
        out[gg] = got
        err = np.abs(got - ref).max() / scale
        print("n=%d l=%d miss=%g gemm_groups=%d: max err vs FP64 reference %.3e" % (n, l, miss, gg, err), flush=True)
        assert err < 1e-9, err
    print("   groups 1 vs 2: %.3e" % (np.abs(out[1] - out[2]).max() / scale), flush=True)
    c.close()
print("pair2 ok")
