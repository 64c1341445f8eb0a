"""Accuracy of the int8 digit-plane projection vs the FP64 projection at full size (n = 50 000), for the
DESIGN.md precision table: max / rms error of U^T x relative to max|U^T x|, and the induced change of the
-lmm 4 outputs, for T = 4..8 planes.  Run on the GPU box:  python scripts/i8_accuracy.py [n] [l]"""
import sys, os, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gemma_b200
from gemma_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
l = int(sys.argv[2]) if len(sys.argv) > 2 else 256
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(7)
U = torch.zeros((n, n), dtype=torch.float64, device=dev)
for s in range(0, n, 2048):
    e = min(n, s + 2048)
    q, _ = torch.linalg.qr(torch.randn((e - s, e - s), dtype=torch.float64, device=dev, generator=g))
    U[s:e, s:e] = q
v = torch.randn(n, dtype=torch.float64, device=dev, generator=g); v /= v.norm()
w = v @ U
for s in range(0, n, 4096):
    e = min(n, s + 4096)
    U[s:e] -= 2.0 * v[s:e, None] * w[None, :]
ev_h = synth.spectrum_like_kinship(n, 3)
ev = torch.from_numpy(ev_h).to(dev)
bed, G = synth.make_bed(n, l, seed=99, miss_rate=0.002)
y_h = synth.phenotype(n, np.where(G[:16] < 0, 0, G[:16]), 5)
y = torch.from_numpy(y_h).to(dev)
UtWt = (torch.ones((1, n), dtype=torch.float64, device=dev) @ U).contiguous()
Uty = (y @ U).contiguous()
ctx = gemma_b200.Context(0)
ctx.lmm_setup_rotated_dev(n, 1, U.data_ptr(), ev.data_ptr(), UtWt.data_ptr(), Uty.data_ptr())
nm = ctx.lmm_null(float(ev_h.mean()))
ctx.lmm_params(4, l_mle_null=nm["l_mle_null"], logl_mle_H0=nm["logl_mle_H0"])
ctx.set_option("utx_path", 1)
ref = ctx.lmm_project_bed(bed, n)
sref = ctx.lmm_batch_bed(bed, n)
scale = np.abs(ref).max()
out = {"n": n, "l": l, "max_abs_utx": float(scale), "planes": {}}
for T in (4, 5, 6, 7, 8):
    ctx.set_option("utx_path", 2); ctx.set_option("n_slices", T)
    got = ctx.lmm_project_bed(bed, n)
    s = ctx.lmm_batch_bed(bed, n)
    rel = lambda a, b: float(np.nanmax(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))
    out["planes"][T] = {"utx_max_err_over_max": float(np.abs(got - ref).max() / scale),
                        "utx_rms_err_over_rms": float(np.sqrt(((got - ref) ** 2).mean() / (ref ** 2).mean())),
                        "beta_rel": rel(s["beta"], sref["beta"]), "se_rel": rel(s["se"], sref["se"]),
                        "p_wald_rel": rel(s["p_wald"], sref["p_wald"]), "p_lrt_rel": rel(s["p_lrt"], sref["p_lrt"]),
                        "p_score_rel": rel(s["p_score"], sref["p_score"]), "min_p_wald": float(np.nanmin(sref["p_wald"]))}
print(json.dumps(out, indent=1))
