"""CPU checks of two numerical claims behind the round-2 per-SNP kernel and 3-plane projection (DESIGN.md 4.1 a, d) -- numpy
restatements of OUR formulation, no GPU:

 (a) the sums S^k(lambda) = sum_i h_i^k z_i, h_i = 1 / (lambda delta_i + 1), are analytic in log(lambda) with poles at distance pi from
     the real axis, so a Chebyshev interpolant over one grid interval (one decade + 2 x 0.15) through 20 / 15 nodes reproduces them to
     ~1e-13 / ~1e-9, and S^3 = S^2 + 1/2 dS^2/dt (dh/dt = h^2 - h);
 (b) with U^T x rounded to 3 digit planes, sums LINEAR in x taken from the genotype-space product x . (U phi) are exact while the
     quadratic sum x'Hx only sees the rounding averaged over the eigenvectors -- except for a constant eigenvector, whose rounding is
     coherent and has to be patched with its exact projection."""
import numpy as np
from numpy.polynomial import chebyshev as Ch

from gemma_b200 import synth
from oracle import i8_planes as P


def _nodes(lo, hi, M):
    tau = np.cos(np.pi * (np.arange(M) + 0.5) / M)
    return tau, 0.5 * (lo + hi) + 0.5 * (hi - lo) * tau


def _coef(f, M):
    j = np.arange(M)
    c = (2.0 / M) * (f[None, :] * np.cos(np.pi * np.arange(M)[:, None] * (j[None, :] + 0.5) / M)).sum(axis=1)
    c[0] *= 0.5
    return c


def test_chebyshev_interpolants_of_the_lambda_sums():
    rng = np.random.default_rng(1)
    n = 6000
    ev = synth.spectrum_like_kinship(n, 3)
    x = rng.standard_normal(n) * 0.65
    y = rng.standard_normal(n) * np.sqrt(0.5 * ev + 0.5)
    marg = 0.15
    worst = {20: 0.0, 15: 0.0, "s3": 0.0}
    for g in range(10):
        lo = np.log(1e-5) + g * np.log(10.0) - marg
        hi = lo + np.log(10.0) + 2 * marg
        tt = np.linspace(lo, hi, 33)                                   # incl. the margins
        ta = (2 * tt - (lo + hi)) / (hi - lo)
        for z in (x * x, x * y, y * y):
            S = lambda t, k: np.array([np.sum(z / (np.exp(q) * ev + 1.0) ** k) for q in np.atleast_1d(t)])
            scale = np.array([np.sum(np.abs(z) / (np.exp(q) * ev + 1.0)) for q in tt])
            for M in (20, 15):
                tau, t = _nodes(lo, hi, M)
                for k in (1, 2):
                    c = _coef(S(t, k), M)
                    worst[M] = max(worst[M], float(np.max(np.abs(Ch.chebval(ta, c) - S(tt, k)) / scale)))
                    if k == 2 and M == 20:
                        d = Ch.chebder(c) * (2.0 / (hi - lo))
                        s3 = Ch.chebval(ta, c) + 0.5 * Ch.chebval(ta, d)   # S^3 = S^2 + 1/2 dS^2/dt
                        worst["s3"] = max(worst["s3"], float(np.max(np.abs(s3 - S(tt, 3)) / scale)))
    assert worst[20] < 2e-13, worst
    assert worst[15] < 5e-10, worst
    assert worst["s3"] < 5e-11, worst


def test_three_planes_exact_linear_sums_and_the_coherent_constant_eigenvector():
    rng = np.random.default_rng(7)
    n, l = 1024, 64
    # an orthogonal U whose FIRST column is the constant vector (the null eigenvector of a centred kinship matrix)
    A = rng.standard_normal((n, n)); A[:, 0] = 1.0
    U, _ = np.linalg.qr(A)
    assert np.allclose(np.abs(U[:, 0]), 1.0 / np.sqrt(n))
    ev = synth.spectrum_like_kinship(n, 5)
    G = rng.binomial(2, rng.uniform(0.05, 0.5, l)[None, :], size=(n, l)).astype(np.int64)
    exact = U.T @ G.astype(np.float64)
    planes, scale = P.slice_planes(U, 3)
    noisy = P.project(planes, scale, G)
    eps = np.abs(noisy - exact)[1:].max() / np.sqrt((G * G).mean())     # per projected value, eigenvectors 1.. (incoherent rounding)
    assert eps < 6e-6                                                     # ~ colmax sqrt(n) / (sqrt(12) 127.4 256^2) at this small n
    h = 1.0 / (2.0 * ev + 1.0)
    yt = rng.standard_normal(n)
    # (1) linear sums: sum_i h_i (U^T x)_i y_i == x . (U (h * y)) -- exact in genotype space, whatever the planes do
    v = U @ (h * yt)
    lin_exact = (h * yt) @ exact
    assert np.allclose(G.T.astype(np.float64) @ v, lin_exact, rtol=1e-12, atol=1e-12)
    lin_noisy = (h * yt) @ noisy
    assert np.max(np.abs(lin_noisy - lin_exact)) > 1e3 * np.max(np.abs(G.T @ v - lin_exact))    # what the side GEMM buys
    # (2) quadratic sums: the incoherent rounding averages out ...
    q_exact = (h[:, None] * exact ** 2).sum(axis=0)
    nz = noisy.copy(); nz[0] = exact[0]                                  # constant eigenvector patched with its exact projection
    q_patched = (h[:, None] * nz ** 2).sum(axis=0)
    assert np.max(np.abs(q_patched - q_exact) / q_exact) < 8 * eps / np.sqrt(n)
    # ... but the constant eigenvector's rounding is coherent: ONE common relative error on every SNP's projection, so its share of
    # x'Hx (the squared genotype mean, the largest single term) shifts by the same relative amount on every SNP instead of averaging
    # out -- and that share does not shrink with n, unlike the 1/sqrt(n) of the incoherent part.  The product patches it exactly.
    rel0 = (noisy[0] - exact[0]) / exact[0]
    assert np.allclose(rel0, rel0[0], rtol=1e-6) and 0 < abs(rel0[0]) < 0.5 / (127.4 * 256.0 ** 2) * 1.01
    q_noisy = (h[:, None] * noisy ** 2).sum(axis=0)
    shift = h[0] * exact[0] ** 2 * ((1.0 + rel0[0]) ** 2 - 1.0)
    assert np.allclose(q_noisy - q_patched, shift, rtol=1e-6, atol=1e-9 * np.abs(shift).max())
