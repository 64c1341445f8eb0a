"""The moment formulation of the multivariate LMM that the CUDA kernels run (gemma_b200/csrc/mvlmm_core.cuh), instantiated for the
host (tests/host/mvlmm_check.cpp, one "lane"), against the numpy restatement oracle/mvlmm_oracle.py (itself pinned against the
compiled reference CLI in tests/test_oracle_vs_ref.py).  CPU only."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import scipy.linalg

from oracle import oracle as O
from oracle import mvlmm_oracle as MV

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_dp = C.POINTER(C.c_double)


def _p(a):
    return a.ctypes.data_as(_dp)


@pytest.fixture(scope="module")
def core(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("mvh") / "libmvh.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, os.path.join(ROOT, "tests", "host", "mvlmm_check.cpp")])
    return C.CDLL(so)


def _problem(n, c, seed):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((n, 3 * n))
    K = O.center_matrix(A @ A.T / (3 * n))
    ev, U = scipy.linalg.eigh(K)
    ev, _ = O.zero_small_eval(ev)
    W = np.ones((n, c))
    if c > 1:
        W[:, :c - 1] = rng.standard_normal((n, c - 1))
    G = rng.binomial(2, rng.uniform(0.1, 0.5, 12)[:, None], size=(12, n)).astype(float)
    g = U @ (np.sqrt(ev)[:, None] * rng.standard_normal((n, 2)))
    E = rng.standard_normal((n, 2)) @ np.array([[1.0, 0.3], [0.0, 0.8]])
    Y = 0.9 * g @ np.array([[1.0, 0.4], [0.0, 0.7]]) + E
    Y[:, 0] += 0.9 * (G[0] - G[0].mean()); Y[:, 1] -= 0.7 * (G[0] - G[0].mean())       # a strong SNP -> Newton-Raphson branch
    return dict(ev=ev, U=U, UtW=U.T @ W, UtY=U.T @ Y, UtX=(U.T @ G.T).T)


@pytest.mark.parametrize("n,c,seed", [(240, 1, 1), (300, 2, 2), (260, 3, 3)])
def test_moment_form_matches_restatement(core, n, c, seed):
    pb = _problem(n, c, seed)
    ev = np.ascontiguousarray(pb["ev"]); X = np.ascontiguousarray(pb["UtW"].T); Y = np.ascontiguousarray(pb["UtY"].T)
    Vg0, Ve0, _ = MV.mph_initial(ev, X, Y)
    out = np.zeros(64)
    k = core.mvh_null(n, c, _p(ev), _p(X), _p(Y), _p(np.ascontiguousarray(Vg0)), _p(np.ascontiguousarray(Ve0)), _p(out))
    assert k == 18 + 2 * c
    nm = MV.null_model(ev, pb["UtW"], pb["UtY"])
    assert np.allclose(out[0:4], nm["Vg_remle"].ravel(), rtol=1e-7, atol=1e-10) and np.allclose(out[4:8], nm["Ve_remle"].ravel(), rtol=1e-7, atol=1e-10)
    assert out[8] == pytest.approx(nm["logl_remle_H0"], rel=1e-10)
    assert np.allclose(out[9:13], nm["Vg_mle"].ravel(), rtol=1e-7, atol=1e-10) and np.allclose(out[13:17], nm["Ve_mle"].ravel(), rtol=1e-7, atol=1e-10)
    assert out[17] == pytest.approx(nm["logl_mle_H0"], rel=1e-10)
    assert np.allclose(out[18:18 + 2 * c], nm["B_mle"].ravel(), rtol=1e-6, atol=1e-9)
    Vg = np.ascontiguousarray(out[9:13]); Ve = np.ascontiguousarray(out[13:17]); Bn = np.ascontiguousarray(out[18:18 + 2 * c])
    core.mvh_snp.argtypes = [C.c_int, C.c_int] + [_dp] * 7 + [C.c_int, C.c_double, _dp]
    o = np.zeros(8); n_nr = 0
    for mode in (1, 2, 3, 4):
        for q in range(pb["UtX"].shape[0]):
            x = np.ascontiguousarray(pb["UtX"][q])
            core.mvh_snp(n, c, _p(ev), _p(X), _p(x), _p(Y), _p(Vg), _p(Ve), _p(Bn), mode, nm["logl_mle_H0"], _p(o))
            beta, Vb, pw, pl, ps = MV.analyze_snp(ev, pb["UtW"], pb["UtY"], x, nm, mode)
            ref = np.array([beta[0], beta[1], Vb[0, 0], Vb[0, 1], Vb[1, 1], pw, pl, ps])
            assert np.allclose(o, ref, rtol=2e-6, atol=1e-300), (mode, q, o, ref)
            n_nr += (mode == 1) and pw < MV.P_NR
    assert n_nr >= 1


def test_moment_form_is_generic_in_the_number_of_phenotypes(core):
    """Three phenotypes on the host instantiation (the kernels of this round instantiate two): null fits, Wald (REML) and score
    tests agree with the restatement to rounding.  The likelihood-ratio test cannot be held to that bar by anyone: the reference's ML
    EM carries U_l^T V_e^-1/2 B from one iteration into the next iteration's rotated basis without re-expressing it (UltVehiBX in
    MphEM, src/mvlmm.cpp:673-690), and with d >= 3 the eigenvector signs of consecutive iterations are not a continuous function of
    (V_g, V_e), so its p_lrt jumps between a few discrete outcomes under 1e-14 relative perturbations of the phenotypes
    (tests/test_oracle_vs_ref.py::test_mvlmm_three_phenotypes_pins_and_ml_em_conditioning shows it on the compiled reference; with
    d = 2 the same probe moves nothing).  So here: the core's p_lrt either equals the restatement's to rounding or differs by no
    more than the restatement's own spread under that probe."""
    n, c, d = 280, 2, 3
    rng = np.random.default_rng(11)
    pb = _problem(n, c, 11)
    y3 = pb["UtY"] @ np.array([0.5, -0.4]) + (pb["U"].T @ rng.standard_normal(n))
    UtY = np.column_stack([pb["UtY"], y3])
    ev = np.ascontiguousarray(pb["ev"]); X = np.ascontiguousarray(pb["UtW"].T); Y = np.ascontiguousarray(UtY.T)
    Vg0, Ve0, _ = MV.mph_initial(ev, X, Y)
    out = np.zeros(128)
    core.mvh3_null.argtypes = [C.c_int, C.c_int] + [_dp] * 6
    k = core.mvh3_null(n, c, _p(ev), _p(X), _p(Y), _p(np.ascontiguousarray(Vg0)), _p(np.ascontiguousarray(Ve0)), _p(out))
    assert k == 4 * 9 + 2 + d * c
    nm = MV.null_model(ev, pb["UtW"], UtY)
    assert np.allclose(out[0:9], nm["Vg_remle"].ravel(), rtol=1e-6, atol=1e-9) and np.allclose(out[9:18], nm["Ve_remle"].ravel(), rtol=1e-6, atol=1e-9)
    assert out[18] == pytest.approx(nm["logl_remle_H0"], rel=1e-9)
    assert np.allclose(out[19:28], nm["Vg_mle"].ravel(), rtol=1e-6, atol=1e-9) and out[37] == pytest.approx(nm["logl_mle_H0"], rel=1e-9)
    Vg = np.ascontiguousarray(out[19:28]); Ve = np.ascontiguousarray(out[28:37]); Bn = np.ascontiguousarray(out[38:38 + d * c])
    core.mvh3_snp.argtypes = [C.c_int, C.c_int] + [_dp] * 7 + [C.c_int, C.c_double, _dp]
    o = np.zeros(d + d * (d + 1) // 2 + 3)
    prng = np.random.default_rng(0)
    perts = [1.0 + 1e-14 * prng.standard_normal(UtY.shape) for _ in range(6)]
    n_same = 0
    for q in range(8):
        x = np.ascontiguousarray(pb["UtX"][q])
        for mode in (1, 3, 2):
            core.mvh3_snp(n, c, _p(ev), _p(X), _p(x), _p(Y), _p(Vg), _p(Ve), _p(Bn), mode, nm["logl_mle_H0"], _p(o))
            beta, Vb, pw, pl, ps = MV.analyze_snp(ev, pb["UtW"], UtY, x, nm, mode)
            ref = np.concatenate([beta, Vb[np.triu_indices(d)], [pw, pl, ps]])
            if mode != 2:
                assert np.allclose(o, ref, rtol=5e-6, atol=1e-300), (mode, q, o, ref)
                continue
            spread = max(abs(MV.analyze_snp(ev, pb["UtW"], UtY * p_, x, nm, 2)[3] - pl) for p_ in perts) / pl
            gap = abs(o[10] - pl) / pl
            assert gap < 1e-8 or gap <= 1.5 * spread, (q, gap, spread)
            n_same += gap < 1e-8
    assert n_same >= 3                        # several SNPs land on the restatement's own outcome to rounding
