"""GPU parity tests: every C-ABI entry point against the CPU oracle on the same seeded
inputs.  Tolerances follow BASELINE.json's north_star: SNP ids/counts bit-exact, beta / se /
p-values within 1e-6 relative.  All calls go through the C ABI (gemma_b200.api)."""
import json
import os

import numpy as np
import pytest
import scipy.linalg

import gemma_b200
from gemma_b200 import synth
from oracle import oracle as O
from oracle import refpipe as R

pytestmark = pytest.mark.gpu

REL = 1e-6          # north_star tolerance on beta / se / p-values
EXP = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "expected.json")))


def rel_err(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    nan_ok = np.isnan(a) == np.isnan(b)
    assert nan_ok.all(), "NaN pattern differs"
    m = ~np.isnan(b)
    if not m.any():
        return 0.0
    return float(np.max(np.abs(a[m] - b[m]) / np.maximum(np.abs(b[m]), 1e-300)))


def check_sumstat(got, ref, a_mode, lam_rel=5e-5):
    fields = {1: ("beta", "se", "p_wald"), 2: ("p_lrt",), 3: ("beta", "se", "p_score"),
              4: ("beta", "se", "p_wald", "p_lrt", "p_score"), 9: ("beta", "se", "p_lrt", "p_score")}[a_mode]
    for k in fields:
        assert rel_err(got[k], ref[k]) < REL, (a_mode, k, rel_err(got[k], ref[k]))
    # lambda is only reproducible to ~1e-5 across summation orders; logl to 1e-8 relative
    for k in ("lambda_remle", "lambda_mle"):
        assert rel_err(got[k], ref[k]) < lam_rel, (a_mode, k, rel_err(got[k], ref[k]))
    assert rel_err(got["logl_H1"], ref["logl_H1"]) < 1e-8
    # untouched fields stay exactly 0 like the reference's initialisers (src/lmm.cpp:1536-1538)
    for k in set(got.dtype.names) - set(fields) - {"lambda_remle", "lambda_mle", "logl_H1"}:
        assert np.array_equal(got[k], ref[k]), k


def random_problem(n, c, l, seed, causal=True):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((n, 3 * n))
    K = O.center_matrix(A @ A.T / (3 * n))
    ev, U = scipy.linalg.eigh(K)
    ev, _ = O.zero_small_eval(ev)
    W = np.ones((n, c))
    if c > 1:
        W[:, :c - 1] = rng.standard_normal((n, c - 1))
    f = rng.uniform(0.05, 0.5, l)
    X = rng.binomial(2, f[None, :], size=(n, l)).astype(np.float64)
    g = U @ (np.sqrt(ev) * rng.standard_normal(n))
    y = 0.8 * g + rng.standard_normal(n) + W[:, :1].sum(axis=1) * 0.3
    if causal:
        y = y + 0.6 * (X[:, 0] - X[:, 0].mean()) + 0.25 * (X[:, 1] - X[:, 1].mean())
    return dict(U=U, ev=ev, W=W, y=y, X=X, UtW=U.T @ W, Uty=U.T @ y, trace_G=float(np.mean(ev)))


# ----------------------------------------------------------------------------------------
def test_dgemm_seam_matches_numpy_and_rejects_bad_shapes(ctx):
    rng = np.random.default_rng(1)
    A = rng.standard_normal((150, 70)); B = rng.standard_normal((70, 90))
    for ta, tb in (("N", "N"), ("T", "N"), ("N", "T"), ("T", "T")):
        a = A.T.copy() if ta == "T" else A
        b = B.T.copy() if tb == "T" else B
        C0 = rng.standard_normal((150, 90))
        C = ctx.dgemm(ta, tb, 1.5, a, b, -0.5, C0.copy())
        assert np.allclose(C, 1.5 * A @ B - 0.5 * C0, rtol=1e-12, atol=1e-12)
    # the reference's unit test: integer-valued 2000x200 * 200x1000 with known entries
    # (test/src/unittests-math.cpp:74-178 style): exact in FP64
    Ai = rng.integers(-5, 6, (300, 129)).astype(float); Bi = rng.integers(-5, 6, (129, 257)).astype(float)
    C = ctx.dgemm("N", "N", 1.0, Ai, Bi, 0.0, np.zeros((300, 257)))
    assert np.array_equal(C, Ai @ Bi)
    with pytest.raises(gemma_b200.GB200Error):       # "Range error in dgemm" (fastblas.cpp:207)
        ctx.dgemm("N", "N", 1.0, A, B.T.copy(), 0.0, np.zeros((150, 90)))
    with pytest.raises(gemma_b200.GB200Error):       # enforce(N>0) (fastblas.cpp:193-195)
        ctx.dgemm("N", "N", 1.0, np.zeros((4, 0)), np.zeros((0, 3)), 0.0, np.zeros((4, 3)))


@pytest.mark.parametrize("n,c,seed", [(257, 1, 3), (300, 3, 4), (64, 2, 5)])
def test_assoc_kernel_all_modes_vs_oracle(ctx, n, c, seed):
    pb = random_problem(n, c, 96, seed)
    ctx.lmm_setup_rotated(pb["U"], pb["ev"], pb["UtW"], pb["Uty"])
    l_mle, logl = O.calc_lambda_null("L", pb["ev"], pb["UtW"], pb["Uty"])
    UtX = pb["U"].T @ pb["X"]
    for mode in (1, 2, 3, 4, 9):
        ref = O.lmm_analyze_utx(pb["ev"], pb["UtW"], pb["Uty"], UtX, mode, l_mle_null=l_mle, logl_mle_H0=logl)
        for kern in (1, 2):                   # 1 = warp-per-SNP kernel, 2 = lockstep-CTA pipeline kernel
            ctx.set_option("lmm_kernel", kern)
            ctx.lmm_params(mode, l_mle_null=l_mle, logl_mle_H0=logl)
            got = ctx.lmm_assoc_utx(np.ascontiguousarray(UtX.T))
            check_sumstat(got, ref, mode)
    ctx.set_option("lmm_kernel", 0)


@pytest.mark.parametrize("n,c,seed", [(310, 4, 21), (280, 6, 22), (333, 7, 23), (401, 12, 24), (520, 20, 25), (700, 32, 26)])
def test_assoc_many_covariates_vs_oracle(ctx, n, c, seed):
    """c = 4..6: register-table kernel; c >= 7: the shared-memory-table kernel (GWAS runs carry age / sex / 10-20 PCs)."""
    pb = random_problem(n, c, 48, seed)
    UtW, Uty = ctx.lmm_setup(pb["U"], pb["ev"], pb["W"], pb["y"])
    nm = ctx.lmm_null(pb["trace_G"])
    l_mle, logl = O.calc_lambda_null("L", pb["ev"], pb["UtW"], pb["Uty"])
    l_re, logl_re = O.calc_lambda_null("R", pb["ev"], pb["UtW"], pb["Uty"])
    assert nm["l_mle_null"] == pytest.approx(l_mle, rel=5e-5) and nm["logl_mle_H0"] == pytest.approx(logl, rel=1e-9)
    assert nm["l_remle_null"] == pytest.approx(l_re, rel=5e-5) and nm["logl_remle_H0"] == pytest.approx(logl_re, rel=1e-9)
    pve, pve_se = O.calc_pve(pb["ev"], pb["UtW"], pb["Uty"], l_re, pb["trace_G"])
    assert nm["pve_null"] == pytest.approx(pve, rel=1e-5) and nm["pve_se_null"] == pytest.approx(pve_se, rel=1e-4)
    vg, ve, beta, se = O.calc_vgvebeta(pb["ev"], pb["UtW"], pb["Uty"], l_re)
    assert np.allclose(nm["beta_remle"], beta, rtol=1e-5, atol=1e-9) and np.allclose(nm["se_beta_remle"], se, rtol=1e-5)
    UtX = pb["U"].T @ pb["X"]
    for mode in (1, 4, 9):
        ref = O.lmm_analyze_utx(pb["ev"], pb["UtW"], pb["Uty"], UtX, mode, l_mle_null=l_mle, logl_mle_H0=logl)
        ctx.lmm_params(mode, l_mle_null=l_mle, logl_mle_H0=logl)
        got = ctx.lmm_assoc_utx(np.ascontiguousarray(UtX.T))
        check_sumstat(got, ref, mode)


def test_generic_covariate_kernel_is_bitwise_equal_to_register_kernel(ctx):
    """lmm_kernel = 3 forces the any-c kernel: same sums in the same order -> identical bits for c <= 6."""
    for n, c, seed in ((257, 1, 31), (300, 3, 32), (290, 5, 33)):
        pb = random_problem(n, c, 64, seed)
        ctx.lmm_setup_rotated(pb["U"], pb["ev"], pb["UtW"], pb["Uty"])
        l_mle, logl = O.calc_lambda_null("L", pb["ev"], pb["UtW"], pb["Uty"])
        UtXt = np.ascontiguousarray((pb["U"].T @ pb["X"]).T)
        ctx.lmm_params(4, l_mle_null=l_mle, logl_mle_H0=logl)
        ctx.set_option("lmm_kernel", 1); a = ctx.lmm_assoc_utx(UtXt); nm1 = ctx.lmm_null(pb["trace_G"])
        ctx.set_option("lmm_kernel", 3); b = ctx.lmm_assoc_utx(UtXt); nm3 = ctx.lmm_null(pb["trace_G"])
        ctx.set_option("lmm_kernel", 0)
        for k in a.dtype.names:
            assert np.array_equal(a[k], b[k], equal_nan=True), (c, k)
        for k in ("l_mle_null", "l_remle_null", "logl_mle_H0", "logl_remle_H0", "pve_null", "pve_se_null"):
            assert nm1[k] == nm3[k], (c, k)


def test_cuda_path_against_the_compiled_reference_itself(ctx):
    """The CUDA path vs the REFERENCE's own LMM::Analyze / null-model code (oracle/_ref/libgemma_ref.so: src/lmm.cpp compiled in
    place against the GSL API shim; prebuilt in the authoring container, shipped with the snapshot)."""
    from oracle import ref as REF
    if not REF.available():
        pytest.skip("oracle/_ref not shipped")
    rng = np.random.default_rng(17)
    for n, c, seed in ((260, 1, 41), (311, 3, 42), (290, 8, 43)):
        pb = random_problem(n, c, 72, seed)
        G = pb["X"].T.copy()
        G[rng.random(G.shape) < 0.02] = np.nan
        ctx.lmm_setup(pb["U"], pb["ev"], pb["W"], pb["y"])
        nm = ctx.lmm_null(pb["trace_G"])
        rnm = REF.null_model(pb["ev"], pb["UtW"], pb["Uty"], pb["trace_G"])
        for k in ("l_mle_null", "l_remle_null"):
            assert nm[k] == pytest.approx(rnm[k], rel=5e-5)
        for k in ("logl_mle_H0", "logl_remle_H0", "pve_null"):
            assert nm[k] == pytest.approx(rnm[k], rel=1e-6)
        assert np.allclose(nm["beta_remle"], rnm["beta_remle"], rtol=1e-5, atol=1e-9)
        for mode in (1, 2, 3, 4, 9):
            ref = REF.lmm_analyze(np.ones(n, dtype=np.int32), pb["U"], pb["ev"], pb["UtW"], pb["Uty"], pb["W"], pb["y"], G, mode,
                                  l_mle_null=rnm["l_mle_null"], logl_mle_H0=rnm["logl_mle_H0"])
            ctx.lmm_params(mode, l_mle_null=rnm["l_mle_null"], logl_mle_H0=rnm["logl_mle_H0"])
            got = ctx.lmm_batch_geno(G)
            check_sumstat(got, ref, mode)


def test_assoc_nondefault_search_grid_and_boundaries(ctx):
    # narrow / shifted lambda ranges force the "no sign change" and clamp branches (lmm.cpp:1985-2000)
    pb = random_problem(200, 1, 40, 9, causal=False)
    ctx.lmm_setup_rotated(pb["U"], pb["ev"], pb["UtW"], pb["Uty"])
    UtX = pb["U"].T @ pb["X"]
    for (lo, hi, nr) in ((1e-5, 1e5, 10), (1e-2, 1e-1, 3), (50.0, 5e4, 7), (1e-5, 1e5, 1), (1e-5, 1e5, 23)):
        l_mle, logl = O.calc_lambda_null("L", pb["ev"], pb["UtW"], pb["Uty"], lo, hi, nr)
        ref = O.lmm_analyze_utx(pb["ev"], pb["UtW"], pb["Uty"], UtX, 4, lo, hi, nr, l_mle, logl)
        for kern in (1, 2):
            ctx.set_option("lmm_kernel", kern)
            ctx.lmm_params(4, lo, hi, nr, l_mle, logl)
            got = ctx.lmm_assoc_utx(np.ascontiguousarray(UtX.T))
            check_sumstat(got, ref, 4)
    ctx.set_option("lmm_kernel", 0)


def test_null_model_vs_oracle(ctx):
    for n, c, seed in ((257, 1, 11), (180, 3, 12)):
        pb = random_problem(n, c, 4, seed)
        UtW, Uty = ctx.lmm_setup(pb["U"], pb["ev"], pb["W"], pb["y"])
        assert np.allclose(UtW, pb["UtW"], atol=1e-11) and np.allclose(Uty, pb["Uty"], atol=1e-11)
        nm = ctx.lmm_null(pb["trace_G"])
        l_mle, logl_mle = O.calc_lambda_null("L", pb["ev"], pb["UtW"], pb["Uty"])
        l_re, logl_re = O.calc_lambda_null("R", pb["ev"], pb["UtW"], pb["Uty"])
        assert nm["l_mle_null"] == pytest.approx(l_mle, rel=5e-5) and nm["logl_mle_H0"] == pytest.approx(logl_mle, rel=1e-9)
        assert nm["l_remle_null"] == pytest.approx(l_re, rel=5e-5) and nm["logl_remle_H0"] == pytest.approx(logl_re, rel=1e-9)
        pve, pve_se = O.calc_pve(pb["ev"], pb["UtW"], pb["Uty"], l_re, pb["trace_G"])
        assert nm["pve_null"] == pytest.approx(pve, rel=1e-5) and nm["pve_se_null"] == pytest.approx(pve_se, rel=1e-4)
        for tag, lam in (("mle", l_mle), ("remle", l_re)):
            vg, ve, beta, se = O.calc_vgvebeta(pb["ev"], pb["UtW"], pb["Uty"], lam)
            assert nm["vg_" + tag] == pytest.approx(vg, rel=1e-4) and nm["ve_" + tag] == pytest.approx(ve, rel=1e-5)
            assert np.allclose(nm["beta_" + tag], beta, rtol=1e-5, atol=1e-9)
            assert np.allclose(nm["se_beta_" + tag], se, rtol=1e-5)


def test_eigh_matches_lapack_semantics(ctx):
    rng = np.random.default_rng(21)
    n = 193
    A = rng.standard_normal((n, 2 * n)); K = A @ A.T / (2 * n)
    U, ev, tr, nz = ctx.eigh(K, center=True)
    Kc = O.center_matrix(K)
    ev_ref = scipy.linalg.eigh(Kc, eigvals_only=True)
    ev_ref, tr_ref = O.zero_small_eval(ev_ref)
    assert np.allclose(ev, ev_ref, atol=1e-11) and tr == pytest.approx(tr_ref, rel=1e-12)
    assert nz == 1 and ev[0] == 0.0                     # the centred matrix always has one (lapack.cpp:268)
    assert np.all(np.diff(ev) >= 0)                     # ascending
    assert np.allclose(U.T @ U, np.eye(n), atol=1e-11)  # eigenvectors in COLUMNS of row-major U
    ev_raw = scipy.linalg.eigh(Kc, eigvals_only=True)
    assert np.allclose((U * ev_raw) @ U.T, Kc, atol=1e-10)
    U2, ev2, _, _ = ctx.eigh(Kc, center=False)
    assert np.allclose(ev2, ev, atol=1e-11)


@pytest.mark.parametrize("n", [193, 1100])
def test_eigh_large_n_solver_agrees_with_the_default_one(ctx, n):
    """Beyond n = 32768 cusolverDnXsyevd refuses the problem and gb200_eigh switches to cusolverMgSyevd on the same device
    (csrc/eigh.cu); forced here at small n: same eigenvalues, orthonormal eigenvectors in columns, same spectral reconstruction."""
    rng = np.random.default_rng(n)
    A = rng.standard_normal((n, 2 * n)); K = A @ A.T / (2 * n)
    U1, ev1, tr1, nz1 = ctx.eigh(K, center=True)
    ctx.set_option("eigh_path", 2)
    try:
        U2, ev2, tr2, nz2 = ctx.eigh(K, center=True)
    finally:
        ctx.set_option("eigh_path", 0)
    assert np.allclose(ev2, ev1, atol=1e-11) and tr2 == pytest.approx(tr1, rel=1e-12) and nz2 == nz1 == 1
    assert np.allclose(U2.T @ U2, np.eye(n), atol=1e-11)
    Kc = O.center_matrix(K)
    assert np.allclose((U2 * ev2) @ U2.T, Kc, atol=1e-10)


@pytest.mark.parametrize("k_mode", [1, 2])
def test_kinship_geno_bed_and_precentred_paths(ctx, k_mode):
    n, l = 211, 500
    bed, G = synth.make_bed(n, l, seed=31, miss_rate=0.03)
    Gn = np.where(G < 0, np.nan, G)
    Gd = Gn + np.where(np.isnan(Gn), 0, np.random.default_rng(3).uniform(0, 0.2, Gn.shape))   # dosages
    for src in ("geno", "bed", "dosage", "precentred"):
        data = Gd if src == "dosage" else Gn
        Xc = O.kin_transform(data, k_mode)
        Kref = Xc @ Xc.T / l
        ctx.kin_begin(n, k_mode)
        if src in ("geno", "dosage"):
            ctx.kin_add_geno(data[:300]); ctx.kin_add_geno(data[300:])     # two batches
        elif src == "bed":
            ctx.kin_add_bed(bed[:123]); ctx.kin_add_bed(bed[123:])
        else:
            ctx.kin_add(Xc[:, :256]); ctx.kin_add(Xc[:, 256:])
        K, ns = ctx.kin_finish()
        assert ns == l
        assert np.allclose(K, Kref, rtol=1e-11, atol=1e-12), src
        assert np.array_equal(K, K.T)


def test_lmm_batch_entry_points_agree_with_oracle(ctx):
    n_total, l = 260, 120
    rng = np.random.default_rng(41)
    mask = np.ones(n_total, dtype=np.uint8); mask[rng.choice(n_total, 23, replace=False)] = 0
    n = int(mask.sum())
    pb = random_problem(n, 2, 4, 42)
    bed, G = synth.make_bed(n_total, l, seed=43, miss_rate=0.04)
    Gn = np.where(G < 0, np.nan, G)[:, mask == 1]
    X = O.lmm_impute(Gn)
    ctx.lmm_setup(pb["U"], pb["ev"], pb["W"], pb["y"])
    nm = ctx.lmm_null(pb["trace_G"])
    ctx.lmm_params(4, l_mle_null=nm["l_mle_null"], logl_mle_H0=nm["logl_mle_H0"])
    ref = O.lmm_analyze_utx(pb["ev"], pb["UtW"], pb["Uty"], pb["U"].T @ X, 4, l_mle_null=nm["l_mle_null"],
                            logl_mle_H0=nm["logl_mle_H0"])
    check_sumstat(ctx.lmm_batch(X), ref, 4)                                  # Xlarge layout, pre-imputed
    check_sumstat(ctx.lmm_batch_geno(Gn), ref, 4)                            # SNP-major with NaN
    check_sumstat(ctx.lmm_batch_bed(bed, n_total, mask), ref, 4)             # PLINK 2-bit + indicator_idv
    assert np.allclose(ctx.lmm_project(X), (pb["U"].T @ X).T, atol=1e-10)
    # empty batch is a no-op (the reference aborts: SURVEY appendix C.1)
    assert len(ctx.lmm_batch(np.zeros((n, 0)))) == 0
    # out-of-order / bad arguments fail loudly
    with pytest.raises(gemma_b200.GB200Error):
        ctx.lmm_batch_bed(bed, n_total, np.ones(n_total, dtype=np.uint8))    # mask selects != n individuals


def test_state_errors():
    c = gemma_b200.Context(0)
    with pytest.raises(gemma_b200.GB200Error):
        c.lmm_batch(np.zeros((8, 2)))
    with pytest.raises(gemma_b200.GB200Error):
        c.kin_add(np.zeros((8, 2)))
    with pytest.raises(gemma_b200.GB200Error):
        c.lmm_setup(np.eye(40), np.ones(40), np.ones((40, 33)), np.ones(40))  # n_cvt > GB200_MAX_CVT
    c.close()


# ---- golden fixtures through the GPU path ------------------------------------------------
@pytest.fixture(scope="module")
def mouse(golden_dir):
    d = os.path.join(golden_dir, "mouse_hs1940")
    bb = R.Bimbam(os.path.join(d, "mouse_hs1940.geno.txt.gz"))
    ph, ind = R.read_pheno(os.path.join(d, "mouse_hs1940.pheno.txt"), (1,))
    idv, W = R.process_cvt_phen(ind)
    isnp, n_miss, maf = R.qc_bimbam(bb, idv)
    return dict(bb=bb, ph=ph, idv=idv, W=W, isnp=isnp)


def test_mouse_hs1940_gk_then_lmm_matches_demo_txt(ctx, mouse):
    """BASELINE config 1: -gk then -lmm on example/mouse_hs1940, against example/demo.txt."""
    bb, idv = mouse["bb"], mouse["idv"]
    sel = np.nonzero(mouse["isnp"])[0]
    assert len(sel) == EXP["mouse_counts"]["ns_test"]
    ctx.kin_begin(bb.G.shape[1], 1)
    for s in range(0, len(sel), 4000):
        ctx.kin_add_geno(bb.G[sel[s:s + 4000]])
    K, ns = ctx.kin_finish()
    assert ns == len(sel)
    assert [[float("%.6g" % K[i, j]) for j in range(3)] for i in range(3)] == EXP["mouse_K3"]
    Kref = R.kinship_bimbam(bb, mouse["isnp"], 1)
    assert np.allclose(K, Kref, rtol=1e-10, atol=1e-12)
    keep = idv == 1
    Kt = R.text_roundtrip(K)[np.ix_(keep, keep)]              # the 10-digit .cXX.txt round trip
    U, ev, trace_G, _ = ctx.eigh(Kt, center=True)
    y = mouse["ph"][keep, 0]; W = mouse["W"][keep]
    ctx.lmm_setup(U, ev, W, y)
    nm = ctx.lmm_null(trace_G)
    assert "%.6f" % nm["pve_null"] == "%.6f" % EXP["mouse_pve"]
    assert "%.6f" % nm["pve_se_null"] == "%.6f" % EXP["mouse_pve_se"]
    ctx.lmm_params(1, l_mle_null=nm["l_mle_null"], logl_mle_H0=nm["logl_mle_H0"])
    Gs = bb.G[np.ix_(sel, keep)]
    out = ctx.lmm_batch_geno(Gs)
    for r, e in zip(out[:5], EXP["mouse_lmm1_rows"]):
        assert "%.6e" % r["beta"] == e["beta"] and "%.6e" % r["se"] == e["se"]
        assert "%.6e" % r["lambda_remle"] == e["l_remle"] and "%.6e" % r["p_wald"] == e["p_wald"]
    # all 10768 SNPs against the oracle run on the GPU's own eigendecomposition
    UtW = U.T @ W; Uty = U.T @ y
    ref = O.lmm_analyze_utx(ev, UtW, Uty, U.T @ O.lmm_impute(Gs), 1)
    check_sumstat(out, ref, 1)


def test_bxd_covariates_pins(ctx, golden_dir):
    d = os.path.join(golden_dir, "BXD")
    bb = R.Bimbam(os.path.join(d, "BXD_geno.txt.gz"))
    ph, ind = R.read_pheno(os.path.join(d, "BXD_pheno.txt"), (1,))
    rows, icvt = R.read_cvt(os.path.join(d, "BXD_covariates2.txt"))
    idv, W = R.process_cvt_phen(ind, rows, icvt)
    isnp_gk, _, _ = R.qc_bimbam(bb, idv, W)
    ctx.kin_begin(bb.G.shape[1], 1)
    ctx.kin_add_geno(bb.G[isnp_gk == 1])
    K, _ = ctx.kin_finish()
    keep = idv == 1
    isnp, _, _ = R.qc_bimbam(bb, idv, W, maf_level=0.1)
    U, ev, trace_G, _ = ctx.eigh(R.text_roundtrip(K)[np.ix_(keep, keep)], center=True)
    ctx.lmm_setup(U, ev, W[keep], ph[keep, 0])
    nm = ctx.lmm_null(trace_G)
    Gs = bb.G[np.ix_(isnp == 1, keep)]
    ctx.lmm_params(2, l_mle_null=nm["l_mle_null"], logl_mle_H0=nm["logl_mle_H0"])
    o2 = ctx.lmm_batch_geno(Gs)
    assert o2["p_lrt"][1] == pytest.approx(EXP["bxd_lmm2_row2_p_lrt"], abs=5e-7)     # lines[2] of the assoc file
    assert o2["p_lrt"].max() == pytest.approx(EXP["bxd_max_p_lrt"], abs=5e-7)
    ctx.lmm_params(9, l_mle_null=nm["l_mle_null"], logl_mle_H0=nm["logl_mle_H0"])
    o9 = ctx.lmm_batch_geno(Gs)
    assert o9["lambda_mle"].max() == pytest.approx(EXP["bxd_lmm9_max_l_mle"], abs=2e-6)


# ---- size-independent properties at a larger size ------------------------------------------
def test_properties_at_scale(ctx):
    n, l = 4096, 1536
    rng = np.random.default_rng(77)
    bed, G = synth.make_bed(n, l, seed=78)
    # cheap orthogonal U: Householder reflection; kinship-like spectrum
    v = rng.standard_normal(n); v /= np.linalg.norm(v)
    U = np.eye(n) - 2.0 * np.outer(v, v)
    ev = synth.spectrum_like_kinship(n, 79)
    y = rng.standard_normal(n) + 0.2 * G[5]
    W = np.ones((n, 1))
    ctx.lmm_setup(U, ev, W, y)
    nm = ctx.lmm_null(float(ev.mean()))
    ctx.lmm_params(4, l_mle_null=nm["l_mle_null"], logl_mle_H0=nm["logl_mle_H0"])
    o4 = ctx.lmm_batch_bed(bed, n)
    # (1) the combined mode reports exactly what the single-test modes report
    ctx.lmm_params(1, l_mle_null=nm["l_mle_null"], logl_mle_H0=nm["logl_mle_H0"])
    o1 = ctx.lmm_batch_bed(bed, n)
    for k in ("beta", "se", "p_wald", "lambda_remle"):
        assert np.array_equal(o1[k], o4[k]), k
    # (2) SNP order does not matter (warps pull SNPs from a ticket counter)
    perm = rng.permutation(l)
    ctx.lmm_params(4, l_mle_null=nm["l_mle_null"], logl_mle_H0=nm["logl_mle_H0"])
    op = ctx.lmm_batch_bed(bed[perm], n)
    for k in o4.dtype.names:
        assert np.array_equal(op[k], o4[k][perm]), k
    # (3) allele flip x -> 2 - x: beta changes sign, se and all p-values are invariant
    of = ctx.lmm_batch_geno(2.0 - G)
    assert rel_err(of["beta"], -o4["beta"]) < 1e-6 and rel_err(of["se"], o4["se"]) < 1e-6
    for k in ("p_wald", "p_lrt", "p_score"):
        assert rel_err(of[k], o4[k]) < 1e-6
    # (4) a sub-sample against the oracle at this size
    idx = np.arange(0, l, 97)
    UtW = U.T @ W; Uty = U.T @ y
    ref = O.lmm_analyze_utx(ev, UtW, Uty, U.T @ G[idx].T, 4, l_mle_null=nm["l_mle_null"],
                            logl_mle_H0=nm["logl_mle_H0"])
    check_sumstat(o4[idx], ref, 4)


# ---- int8 tensor-core projection (tcgen05) vs the FP64 projection ---------------------------
@pytest.mark.parametrize("n_total,n_drop,l,miss", [(300, 17, 200, 0.02), (1100, 0, 130, 0.0), (2500, 3, 517, 0.01)])
def test_i8_tensor_core_projection_matches_fp64(n_total, n_drop, l, miss):
    c = gemma_b200.Context(0)
    rng = np.random.default_rng(n_total)
    mask = np.ones(n_total, dtype=np.uint8)
    if n_drop:
        mask[rng.choice(n_total, n_drop, replace=False)] = 0
    n = int(mask.sum())
    Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    ev = synth.spectrum_like_kinship(n, 5)
    c.lmm_setup(Q, ev, np.ones((n, 1)), rng.standard_normal(n))
    bed, G = synth.make_bed(n_total, l, seed=n_total + 1, miss_rate=miss)
    X = O.lmm_impute(np.where(G < 0, np.nan, G)[:, mask == 1])          # n x l, mean-imputed
    ref = (Q.T @ X).T
    c.set_option("utx_path", 1)
    fp = c.lmm_project_bed(bed, n_total, mask if n_drop else None)
    assert np.allclose(fp, ref, rtol=0, atol=1e-11)
    scale = np.abs(ref).max()
    for T, tol in ((6, 1e-11), (8, 1e-11), (7, 1e-11), (5, 1e-9), (4, 1e-6)):
        c.set_option("utx_path", 2)
        c.set_option("n_slices", T)
        got = c.lmm_project_bed(bed, n_total, mask if n_drop else None)
        err = np.abs(got - ref).max() / scale
        assert err < tol, (T, err)
    # the count chosen from the column maxima of this U: cohorts below n = 8192 never go below 5 planes
    c.set_option("n_slices", 0)
    assert c.get_option("n_slices") == 5
    got = c.lmm_project_bed(bed, n_total, mask if n_drop else None)
    assert np.abs(got - ref).max() < 1e-10 * np.sqrt((X * X).mean()), np.abs(got - ref).max()
    c.close()


def test_i8_path_end_to_end_parity(ctx):
    """-lmm 4 through gb200_lmm_batch_bed with the tensor-core projection forced, vs the oracle."""
    n, l = 1536, 300
    rng = np.random.default_rng(5)
    pb = random_problem(n, 1, 4, 91)
    bed, G = synth.make_bed(n, l, seed=92, miss_rate=0.01)
    X = O.lmm_impute(np.where(G < 0, np.nan, G))
    y = pb["y"] + 0.5 * (X[:, 7] - X[:, 7].mean())
    ctx.lmm_setup(pb["U"], pb["ev"], pb["W"], y)
    nm = ctx.lmm_null(pb["trace_G"])
    ctx.lmm_params(4, l_mle_null=nm["l_mle_null"], logl_mle_H0=nm["logl_mle_H0"])
    ctx.set_option("utx_path", 2); ctx.set_option("n_slices", 6)
    got = ctx.lmm_batch_bed(bed, n)
    ctx.set_option("utx_path", 0); ctx.set_option("n_slices", 0)
    ref = O.lmm_analyze_utx(pb["ev"], pb["U"].T @ pb["W"], pb["U"].T @ y, pb["U"].T @ X, 4,
                            l_mle_null=nm["l_mle_null"], logl_mle_H0=nm["logl_mle_H0"])
    check_sumstat(got, ref, 4)


# ---- the GEMMA-compatible CLI end to end (BASELINE config 1 through the drop-in surface) -------
def test_cli_mouse_gk_then_lmm_matches_demo_txt(golden_dir, tmp_path):
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cli = os.path.join(root, "gemma_b200", "host", "gemma-b200")
    subprocess.check_call(["make", "-s", "-C", os.path.dirname(cli)])     # always: a stale binary must not pass for the source
    d = os.path.join(golden_dir, "mouse_hs1940")
    base = ["-g", d + "/mouse_hs1940.geno.txt.gz", "-p", d + "/mouse_hs1940.pheno.txt", "-a", d + "/mouse_hs1940.anno.txt",
            "-outdir", str(tmp_path)]
    r = subprocess.run([cli] + base + ["-gk", "-o", "mouse"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    k = open(tmp_path / "mouse.cXX.txt").read().splitlines()
    assert len(k) == 1940 and len(k[0].split("\t")) == 1940                  # test/test_suite.sh:119-123
    assert [[float("%.6g" % float(x)) for x in k[i].split("\t")[:3]] for i in range(3)] == EXP["mouse_K3"]
    r = subprocess.run([cli] + base + ["-n", "1", "-k", str(tmp_path / "mouse.cXX.txt"), "-lmm", "-o", "lmm"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "pve estimate =0.608801" in r.stdout and "se(pve) =0.032774" in r.stdout       # demo.txt:41-42
    lines = open(tmp_path / "lmm.assoc.txt").read().splitlines()
    assert len(lines) == 10769                                              # header + 10768 SNPs
    assert lines[0].split("\t") == ["chr", "rs", "ps", "n_miss", "allele1", "allele0", "af", "beta", "se", "logl_H1",
                                    "l_remle", "p_wald"]
    assert sum(len(l.split("\t")) for l in lines) == 129228                 # test/test_suite.sh:138 word count
    for line, e in zip(lines[1:6], EXP["mouse_lmm1_rows"]):                 # demo.txt:32-36, text for text
        f = line.split("\t")
        assert f[:7] == [e["chr"], e["rs"], e["ps"], e["n_miss"], e["allele1"], e["allele0"], e["af"]]
        assert [f[7], f[8], f[10], f[11]] == [e["beta"], e["se"], e["l_remle"], e["p_wald"]]
    # -lmm 4 writes the 15-column layout
    r = subprocess.run([cli] + base + ["-k", str(tmp_path / "mouse.cXX.txt"), "-lmm", "4", "-o", "lmm4", "-snps",
                                       d + "/../mouse_snps_first200.txt"], capture_output=True, text=True)
    if r.returncode == 0:
        h = open(tmp_path / "lmm4.assoc.txt").readline().rstrip("\n").split("\t")
        assert h[7:] == ["beta", "se", "logl_H1", "l_remle", "l_mle", "p_wald", "p_lrt", "p_score"]


def test_cli_mouse_loco_nind_matches_reference_pins(golden_dir, tmp_path):
    """test/dev_tests.rb:57-77 and test/dev_test_suite.sh:121-153, text for text."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cli = os.path.join(root, "gemma_b200", "host", "gemma-b200")
    subprocess.check_call(["make", "-s", "-C", os.path.dirname(cli)])     # always: a stale binary must not pass for the source
    d = os.path.join(golden_dir, "mouse_hs1940")
    e = EXP["mouse_loco"]
    base = ["-g", d + "/mouse_hs1940.geno.txt.gz", "-p", d + "/mouse_hs1940.pheno.txt", "-a", d + "/mouse_hs1940.anno.txt",
            "-snps", d + "/mouse_hs1940_snps.txt", "-nind", "400", "-loco", "1", "-outdir", str(tmp_path)]
    r = subprocess.run([cli] + base + ["-gk", "-o", "loco"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    txt = open(tmp_path / "loco.cXX.txt").read()
    assert len(txt.splitlines()) == e["cxx_lines"] and txt[:5] == e["cxx_head5"]
    assert "%.2f" % sum(float("%.2f" % float(w[:6])) for w in txt.split()) == e["cxx_sum2"]
    r = subprocess.run([cli] + base + ["-n", "1", "-k", str(tmp_path / "loco.cXX.txt"), "-lmm", "-no-check", "-o", "loco"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = open(tmp_path / "loco.assoc.txt").read().splitlines()
    assert len(lines) == e["assoc_lines"]
    assert lines[2].split("\t")[9] == e["row2_logl_H1"]
    assert "%.6e" % max(float(l.split("\t")[11]) for l in lines[1:]) == e["max_p_wald"]
    assert all(l.split("\t")[0] == "1" for l in lines[1:])            # only chromosome-1 SNPs are tested
    # binary side channel: -bin writes <file>.bin next to the text; feeding the exact K changes results only at the
    # level of the 10-digit rounding of the text file
    r = subprocess.run([cli] + base + ["-gk", "-bin", "-o", "locob"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert open(tmp_path / "locob.cXX.txt").read() == txt
    r = subprocess.run([cli] + base + ["-n", "1", "-k", str(tmp_path / "locob.cXX.txt.bin"), "-lmm", "-o", "locob"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    lb = open(tmp_path / "locob.assoc.txt").read().splitlines()
    assert len(lb) == len(lines)
    a = np.array([[float(x) for x in l.split("\t")[7:]] for l in lines[1:]])
    b = np.array([[float(x) for x in l.split("\t")[7:]] for l in lb[1:]])
    assert np.allclose(a, b, rtol=1e-5, atol=1e-8) and [l.split("\t")[:7] for l in lb] == [l.split("\t")[:7] for l in lines]


def test_reference_cli_with_the_plugin_reproduces_demo_txt(golden_dir, tmp_path):
    """The drop-in boundary, BUILT: oracle/_ref/gemma_ref_b200 is the reference's own CLI (every src/*.cpp compiled in place,
    unmodified) whose four hot-path seams -- fast_dgemm, EigenDecomp_Zeroed, BimbamKin / PlinkKin, LMM::AnalyzeBimbam /
    AnalyzePlink -- are bound to libgemma_b200.so by ONE extra translation unit (gemma_b200/host/gemma_seams.cpp; recipe
    `make -C oracle ref_b200`).  Its flag parsing, readers, QC, null model and writers are the reference's.  It must reproduce
    example/demo.txt:10-12 (K), :32-36 (first five -lmm 1 rows) and :41-42 (pve, se(pve)), and the PLINK seams must agree with
    the unbound reference CLI on a synthetic PLINK set."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "oracle", "_ref", "gemma_ref_b200")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/gemma_ref_b200 not shipped (built where /root/reference is present)")
    cwd = str(tmp_path); out = os.path.join(cwd, "output")
    d = os.path.join(golden_dir, "mouse_hs1940")
    base = ["-g", d + "/mouse_hs1940.geno.txt.gz", "-p", d + "/mouse_hs1940.pheno.txt", "-a", d + "/mouse_hs1940.anno.txt"]

    def run(args):
        r = subprocess.run([exe] + args, capture_output=True, text=True, cwd=cwd)
        assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
        return r.stdout + r.stderr

    run(base + ["-gk", "-o", "k"])
    k = open(os.path.join(out, "k.cXX.txt")).read().splitlines()
    assert len(k) == 1940
    assert [[float("%.6g" % float(x)) for x in k[i].split("\t")[:3]] for i in range(3)] == EXP["mouse_K3"]          # demo.txt:10-12
    run(base + ["-n", "1", "-k", "output/k.cXX.txt", "-lmm", "-o", "lmm"])
    lines = open(os.path.join(out, "lmm.assoc.txt")).read().splitlines()
    assert len(lines) == 1 + EXP["mouse_counts"]["ns_test"]
    hdr = lines[0].split("\t")
    for ln, e in zip(lines[1:6], EXP["mouse_lmm1_rows"]):                                                        # demo.txt:32-36
        f = dict(zip(hdr, ln.split("\t")))
        for key, val in e.items():
            assert f[key] == val, (key, f[key], val)
    log = open(os.path.join(out, "lmm.log.txt")).read()
    assert "## pve estimate in the null model = %s" % EXP["mouse_pve"] in log                                    # demo.txt:41
    assert "## se(pve) in the null model = %s" % EXP["mouse_pve_se"] in log                                      # demo.txt:42
    # PLINK seams (PlinkKin, AnalyzePlink) against the unbound reference CLI, whole files
    from oracle import ref as REF
    if os.path.exists(REF.EXE):
        n, l = 400, 300
        rng = np.random.default_rng(21)
        bed, G = synth.make_bed(n, l, seed=777, miss_rate=0.02)
        y = rng.standard_normal(n) + 0.5 * np.where(G[5] < 0, 0, G[5])
        y[rng.choice(n, 9, replace=False)] = np.nan
        _write_plink(os.path.join(cwd, "syn"), bed, y)
        REF.run_cli(["-bfile", "syn", "-gk", "1", "-o", "rk"], cwd)
        run(["-bfile", "syn", "-gk", "1", "-o", "pk"])
        Kr = np.loadtxt(os.path.join(out, "rk.cXX.txt")); Kp = np.loadtxt(os.path.join(out, "pk.cXX.txt"))
        assert np.allclose(Kr, Kp, rtol=1e-8, atol=1e-9)
        REF.run_cli(["-bfile", "syn", "-k", "output/rk.cXX.txt", "-lmm", "4", "-o", "ra"], cwd)
        run(["-bfile", "syn", "-k", "output/rk.cXX.txt", "-lmm", "4", "-o", "pa"])
        ha, na, xa = _assoc_table(os.path.join(out, "pa.assoc.txt")); hb, nb, xb = _assoc_table(os.path.join(out, "ra.assoc.txt"))
        assert ha == hb and na == nb
        bad, rel = _record_cli_deviations("plugin:gemma_ref_b200 vs gemma_ref, synthetic PLINK -lmm 4", ha, na, xa, xb, 2e-6,
                                          [i for i, h in enumerate(ha[7:]) if h.startswith("l_")])
        assert bad.any(axis=1).mean() <= 0.01 and rel.max() < 1e-4, (int(bad.any(axis=1).sum()), float(rel.max()))


_REPORT = {}


def _record_cli_deviations(name, header, ids, xa, xb, rtol, lam_cols=()):
    """Whole-file comparison of two CLIs' statistics columns: returns the mask of rows with a cell beyond its tolerance and
    RECORDS them -- count, worst relative deviation per column and the SNPs concerned -- in gpurun_out/cli_parity_report.json
    (copied to profiles/ after a GPU run), so that an allowance in an assert is never a blind one."""
    cols = header[7:]
    ok = np.isfinite(xb)
    assert np.array_equal(np.isfinite(xa), ok), "NaN pattern differs"
    with np.errstate(divide="ignore", invalid="ignore"):
        rel = np.where(ok, np.abs(xa - xb) / np.maximum(np.abs(xb), 1e-300), 0.0)
    tol = np.array([2e-5 if j in lam_cols else rtol for j in range(len(cols))])      # lambda: the Newton stopping tolerance is 1e-5
    bad = rel > tol[None, :]
    rows = np.nonzero(bad.any(axis=1))[0]
    entry = {"rows_compared": int(xa.shape[0]), "rows_beyond_tolerance": int(len(rows)), "tolerance": {c: float(t) for c, t in zip(cols, tol)},
             "max_rel_dev_all_rows": {c: float(rel[:, j].max()) for j, c in enumerate(cols)},
             "cells_beyond_tolerance": {c: int(bad[:, j].sum()) for j, c in enumerate(cols)},
             "snps": [{"rs": ids[r][1], "cols": {cols[j]: [float(xa[r, j]), float(xb[r, j])] for j in np.nonzero(bad[r])[0]},
                       "lambda_rel_dev": {cols[j]: float(rel[r, j]) for j in lam_cols}} for r in rows[:40]]}
    _REPORT[name] = entry
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    os.makedirs(os.path.join(root, "gpurun_out"), exist_ok=True)
    path = os.path.join(root, "gpurun_out", "cli_parity_report.json")
    try:
        prev = json.load(open(path))
    except Exception:
        prev = {}
    prev.update(_REPORT)
    json.dump(prev, open(path, "w"), indent=1)
    return bad, rel


def _assoc_table(path):
    lines = open(path).read().splitlines()
    hdr = lines[0].split("\t")
    rows = [ln.split("\t") for ln in lines[1:]]
    nonnum = [r[:7] for r in rows]
    num = np.array([[float(x) for x in r[7:]] for r in rows])
    return hdr, nonnum, num


def test_cli_against_the_reference_cli_end_to_end(golden_dir, tmp_path):
    """gemma-b200 vs the reference's own CLI (oracle/_ref/gemma_ref: all src/*.cpp compiled in place against the GSL API shim),
    same command lines, whole output files: mouse example (-gk, -lmm 4 over all 10 768 SNPs, LOCO) and a synthetic PLINK set with
    covariates (-gk 2 like the reference's HLC test, -lmm 4)."""
    import subprocess
    from oracle import ref as REF
    if not os.path.exists(REF.EXE):
        pytest.skip("reference CLI not shipped")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cli = os.path.join(root, "gemma_b200", "host", "gemma-b200")
    subprocess.check_call(["make", "-s", "-C", os.path.dirname(cli)])     # always: a stale binary must not pass for the source
    cwd = str(tmp_path); out = os.path.join(cwd, "output")

    def mine(args):
        r = subprocess.run([cli] + args + ["-outdir", out], capture_output=True, text=True, cwd=cwd)
        assert r.returncode == 0, r.stdout + r.stderr
        return r.stdout

    def compare(a, b, rtol=2e-6):
        ha, na, xa = _assoc_table(os.path.join(out, a + ".assoc.txt")); hb, nb, xb = _assoc_table(os.path.join(out, b + ".assoc.txt"))
        assert ha == hb and na == nb                                   # ids, positions, n_miss, alleles, af: text for text
        lam = [i for i, h in enumerate(ha[7:]) if h.startswith("l_")]
        bad, rel = _record_cli_deviations("lmm:%s_vs_%s" % (a, b), ha, na, xa, xb, rtol, lam)     # 7 printed digits -> 2e-6
        # The only rows allowed beyond the tolerance are those where the optimiser itself stopped on another iterate: CalcLambda
        # reports the Newton iterate BEFORE the one that met |dl| < 1e-5 |l| (src/lmm.cpp:2096), so a last-bit difference in a
        # borderline convergence test moves lambda by ~1e-5 and the lambda-dependent columns with it.  Such a row must show the
        # lambda shift, stay within 1e-4 everywhere, and the rows are counted and named in the report.
        for r in np.nonzero(bad.any(axis=1))[0]:
            assert lam and max(rel[r, j] for j in lam) > 1e-6, ("deviation without a lambda shift", na[r][1], rel[r].tolist())
            assert rel[r].max() < 1e-4, (na[r][1], rel[r].tolist())
        assert bad.any(axis=1).mean() <= 0.001, int(bad.any(axis=1).sum())
        return xa.shape[0]

    d = os.path.join(golden_dir, "mouse_hs1940")
    base = ["-g", d + "/mouse_hs1940.geno.txt.gz", "-p", d + "/mouse_hs1940.pheno.txt", "-a", d + "/mouse_hs1940.anno.txt"]
    REF.run_cli(base + ["-gk", "-o", "rk"], cwd)
    mine(base + ["-gk", "-o", "mk"])
    Kr = np.loadtxt(os.path.join(out, "rk.cXX.txt")); Km = np.loadtxt(os.path.join(out, "mk.cXX.txt"))
    assert np.allclose(Kr, Km, rtol=1e-8, atol=1e-9)
    REF.run_cli(base + ["-k", "output/rk.cXX.txt", "-lmm", "4", "-o", "r4"], cwd)
    mine(base + ["-k", os.path.join(out, "rk.cXX.txt"), "-lmm", "4", "-o", "m4"])
    assert compare("m4", "r4") == 10768
    loco = ["-snps", d + "/mouse_hs1940_snps.txt", "-nind", "400", "-loco", "1"]
    REF.run_cli(base + loco + ["-gk", "-o", "rl"], cwd)
    REF.run_cli(base + loco + ["-k", "output/rl.cXX.txt", "-lmm", "-no-check", "-o", "rl"], cwd)
    mine(base + loco + ["-k", os.path.join(out, "rl.cXX.txt"), "-lmm", "-o", "ml"])
    assert compare("ml", "rl") == 67
    # synthetic PLINK with two covariates + intercept, standardised kinship
    n, l = 700, 500
    rng = np.random.default_rng(11)
    bed, G = synth.make_bed(n, l, seed=555, miss_rate=0.01)
    y = rng.standard_normal(n) + 0.5 * np.where(G[2] < 0, 0, G[2])
    y[rng.choice(n, 19, replace=False)] = np.nan
    prefix = os.path.join(cwd, "syn")
    _write_plink(prefix, bed, y)
    cov = os.path.join(cwd, "cov.txt")
    with open(cov, "w") as f:
        for i in range(n):
            f.write("1 %.6f %.6f\n" % (rng.standard_normal(), rng.uniform(20, 70)))
    REF.run_cli(["-bfile", prefix, "-gk", "2", "-o", "rs"], cwd)
    mine(["-bfile", prefix, "-gk", "2", "-o", "ms"])
    Kr = np.loadtxt(os.path.join(out, "rs.sXX.txt")); Km = np.loadtxt(os.path.join(out, "ms.sXX.txt"))
    assert np.allclose(Kr, Km, rtol=1e-8, atol=1e-9)
    REF.run_cli(["-bfile", prefix, "-c", cov, "-k", "output/rs.sXX.txt", "-lmm", "4", "-o", "rp"], cwd)
    mine(["-bfile", prefix, "-c", cov, "-k", os.path.join(out, "rs.sXX.txt"), "-lmm", "4", "-o", "mp"])
    compare("mp", "rp")


def test_gxe_entry_points_and_cli_match_oracle_and_reference_cli(ctx, tmp_path):
    """G x E: gb200_lmm_gxe_batch_bed / _geno vs the oracle composition (itself checked against the reference
    CLI in tests/test_oracle_vs_ref.py), and gemma-b200 -gxe vs the reference CLI on whole assoc files."""
    import subprocess
    from oracle import ref as REF
    from test_oracle_vs_ref import _plink_gxe_case
    prefix, gxe_file, bed, G, env, y = _plink_gxe_case(tmp_path, n=310, l=120, seed=79)
    pl = R.Plink(prefix)
    ind_gxe = np.ones(len(y), dtype=np.int32); ind_gxe[[5, 17]] = 0
    idv, W = R.process_cvt_phen(pl.ind_pheno)
    idv2 = idv * ind_gxe
    isnp, _, _ = R.qc_plink(pl, idv2)
    K = R.kinship_plink(pl, R.qc_plink(pl, idv)[0], 1)
    prep = R.lmm_prepare(K, idv2, pl.pheno[:, 0], W)
    keep = idv2 == 1
    sel = np.nonzero(isnp)[0]
    Gs = np.where(pl.G[np.ix_(sel, keep)] < 0, np.nan, pl.G[np.ix_(sel, keep)])
    ctx.lmm_setup(prep["U"], prep["eval"], prep["W"], prep["y"])
    nm = ctx.lmm_null(prep["trace_G"])
    ctx.lmm_gxe_setup(env[keep])
    for mode in (1, 2, 3, 4):
        ref = R.lmm_gxe(prep, Gs, env[keep], mode, l_mle_null=prep["l_mle_null"])
        ctx.lmm_params(mode, l_mle_null=nm["l_mle_null"], logl_mle_H0=nm["logl_mle_H0"])
        got_bed = ctx.lmm_gxe_batch_bed(pl.bed[sel], len(idv2), idv2.astype(np.uint8))
        got_geno = ctx.lmm_gxe_batch_geno(Gs)
        check_sumstat(got_bed, ref, mode)
        check_sumstat(got_geno, ref, mode)
    if not os.path.exists(REF.EXE):
        return
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cli = os.path.join(root, "gemma_b200", "host", "gemma-b200")
    cwd = str(tmp_path); out = os.path.join(cwd, "output")
    REF.run_cli(["-bfile", prefix, "-gk", "1", "-o", "k"], cwd)
    REF.run_cli(["-bfile", prefix, "-gxe", gxe_file, "-k", "output/k.cXX.txt", "-lmm", "4", "-o", "rg"], cwd)
    r = subprocess.run([cli, "-bfile", prefix, "-gxe", gxe_file, "-k", os.path.join(out, "k.cXX.txt"), "-lmm", "4", "-o", "mg", "-outdir", out],
                       capture_output=True, text=True, cwd=cwd)
    assert r.returncode == 0, r.stdout + r.stderr
    ha, na, xa = _assoc_table(os.path.join(out, "mg.assoc.txt")); hb, nb, xb = _assoc_table(os.path.join(out, "rg.assoc.txt"))
    assert ha == hb and na == nb
    for j, h in enumerate(ha[7:]):
        assert np.allclose(xa[:, j], xb[:, j], rtol=2e-5 if h.startswith("l_") else 2e-6, atol=0), h


def test_lm_entry_points_and_cli_match_oracle_and_reference_cli(ctx, tmp_path):
    """-lm 1..4 (src/lm.cpp): gb200_lm_batch_bed / _geno vs the restatement (checked against the reference CLI in
    tests/test_oracle_vs_ref.py), and gemma-b200 -lm vs the reference CLI on whole assoc files (PLINK with covariates, BIMBAM mouse)."""
    import subprocess
    from oracle import ref as REF
    from test_oracle_vs_ref import _plink_lm_case
    prefix, cov = _plink_lm_case(tmp_path, n=333, l=140, seed=89)
    pl = R.Plink(prefix)
    rows, icvt = R.read_cvt(cov)
    idv, W = R.process_cvt_phen(pl.ind_pheno, rows, icvt)
    isnp, _, _ = R.qc_plink(pl, idv, W)
    keep = idv == 1
    sel = np.nonzero(isnp)[0]
    Gs = np.where(pl.G[np.ix_(sel, keep)] < 0, np.nan, pl.G[np.ix_(sel, keep)])
    ctx.lm_setup(W[keep], pl.pheno[keep, 0])
    for mode in (1, 2, 3, 4):
        ref = R.lm_analyze(W[keep], pl.pheno[keep, 0], Gs, mode)
        for got in (ctx.lm_batch_bed(pl.bed[sel], len(idv), 50 + mode, idv.astype(np.uint8)), ctx.lm_batch_geno(Gs, mode)):
            for k in ("beta", "se", "p_wald", "p_lrt", "p_score"):
                assert rel_err(got[k], ref[k]) < REL, (mode, k)
            assert np.all(got["lambda_remle"] == 0) and np.all(got["lambda_mle"] == 0)
    if not os.path.exists(REF.EXE):
        return
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cli = os.path.join(root, "gemma_b200", "host", "gemma-b200")
    cwd = str(tmp_path); out = os.path.join(cwd, "output")
    d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mouse_hs1940")
    cases = [(["-bfile", prefix, "-c", cov], "p"), (["-g", d + "/mouse_hs1940.geno.txt.gz", "-p", d + "/mouse_hs1940.pheno.txt", "-a", d + "/mouse_hs1940.anno.txt"], "b")]
    for base, tag in cases:
        for mode in ("1", "4"):
            REF.run_cli(base + ["-lm", mode, "-o", "r" + tag + mode], cwd)
            r = subprocess.run([cli] + base + ["-lm", mode, "-o", "m" + tag + mode, "-outdir", out], capture_output=True, text=True, cwd=cwd)
            assert r.returncode == 0, r.stdout + r.stderr
            a = open(os.path.join(out, "m" + tag + mode + ".assoc.txt")).read().splitlines()
            b = open(os.path.join(out, "r" + tag + mode + ".assoc.txt")).read().splitlines()
            assert len(a) == len(b) and a[0] == b[0]
            fa = [x.split("\t") for x in a[1:]]; fb = [x.split("\t") for x in b[1:]]
            assert [x[:8] for x in fa] == [x[:8] for x in fb]              # chr rs ps n_mis n_obs alleles af: text for text
            xa = np.array([[float(v) for v in x[8:]] for x in fa]); xb = np.array([[float(v) for v in x[8:]] for x in fb])
            assert np.allclose(xa, xb, rtol=2e-6, atol=0), (tag, mode)


def test_mvlmm_entry_points_match_restatement_and_reference_cli(ctx, golden_dir, tmp_path):
    """Multivariate LMM, two phenotypes (BASELINE config 5): gb200_mvlmm_setup / _null / _batch_geno / _batch_bed vs
    oracle/mvlmm_oracle.py on random problems (c = 1..3, Newton-Raphson branch included) and vs the reference CLI's mouse run
    (-n 1 6: example/demo.txt:62-80)."""
    from oracle import mvlmm_oracle as MV
    from oracle import ref as REF
    from test_mvlmm_core import _problem
    for n, c, seed in ((240, 1, 1), (300, 3, 3)):
        pb = _problem(n, c, seed)
        rng = np.random.default_rng(seed)
        W = pb["U"] @ pb["UtW"]; Y = pb["U"] @ pb["UtY"]
        ctx.mvlmm_setup(pb["U"], pb["ev"], W, Y)
        nm = ctx.mvlmm_null()
        ref = MV.null_model(pb["ev"], pb["UtW"], pb["UtY"])
        for k in ("Vg_remle", "Ve_remle", "Vg_mle", "Ve_mle"):
            assert np.allclose(nm[k], ref[k], rtol=1e-6, atol=1e-9), k
        assert nm["logl_remle_H0"] == pytest.approx(ref["logl_remle_H0"], rel=1e-9) and nm["logl_mle_H0"] == pytest.approx(ref["logl_mle_H0"], rel=1e-9)
        G = (pb["U"] @ pb["UtX"].T).T                                  # genotypes back in the original basis (exact integers up to rounding)
        G = np.rint(G)
        for mode in (1, 4):
            got = ctx.mvlmm_batch_geno(G, mode)
            for q in range(G.shape[0]):
                beta, Vb, pw, pl, ps = MV.analyze_snp(pb["ev"], pb["UtW"], pb["UtY"], pb["U"].T @ G[q], ref, mode)
                exp = np.array([beta[0], beta[1], Vb[0, 0], Vb[0, 1], Vb[1, 1], pw, pl, ps])
                assert np.allclose(got[q], exp, rtol=2e-6, atol=1e-300), (c, mode, q, got[q], exp)
    # PLINK rows (int8 tensor-core projection, n >= 1024) give the same statistics as the dosage entry point
    n = 1100
    pbp = _problem(n, 2, 9)
    bedp, Gp = synth.make_bed(n, 96, seed=321, miss_rate=0.01)
    ctx.mvlmm_setup(pbp["U"], pbp["ev"], pbp["U"] @ pbp["UtW"], pbp["U"] @ pbp["UtY"])
    ctx.mvlmm_null()
    a = ctx.mvlmm_batch_bed(bedp, n, a_mode=4)
    b = ctx.mvlmm_batch_geno(np.where(Gp < 0, np.nan, Gp), 4)
    assert np.allclose(a, b, rtol=1e-6, atol=1e-300)
    # mouse example, two phenotypes, through the PLINK-free BIMBAM entry point; reference CLI rows as the expectation
    d = os.path.join(golden_dir, "mouse_hs1940")
    bb = R.Bimbam(d + "/mouse_hs1940.geno.txt.gz")
    ph, ind = R.read_pheno(d + "/mouse_hs1940.pheno.txt", (1, 6))
    idv, W = R.process_cvt_phen(ind)
    isnp, _, _ = R.qc_bimbam(bb, idv)
    keep = idv == 1
    K = R.text_roundtrip(R.kinship_bimbam(bb, R.qc_bimbam(bb, R.process_cvt_phen(R.read_pheno(d + "/mouse_hs1940.pheno.txt", (1,))[1])[0])[0], 1))
    U, ev, _ = R.eigen_decomp_zeroed(O.center_matrix(np.ascontiguousarray(K[np.ix_(keep, keep)])))
    ctx.mvlmm_setup(U, ev, W[keep], ph[keep])
    nm = ctx.mvlmm_null()
    assert np.allclose([nm["Vg_remle"][0, 0], nm["Vg_remle"][0, 1], nm["Vg_remle"][1, 1]], [1.39398, -0.226714, 2.08168], rtol=5e-6)   # demo.txt:70-72
    assert np.allclose([nm["Ve_remle"][0, 0], nm["Ve_remle"][0, 1], nm["Ve_remle"][1, 1]], [0.348882, 0.0490525, 0.414433], rtol=5e-6)  # demo.txt:76-78
    sel = np.nonzero(isnp)[0][:64]
    Gs = bb.G[np.ix_(sel, np.nonzero(keep)[0])]
    got = ctx.mvlmm_batch_geno(Gs)[:, :6]
    exp = np.array([[float(x) for x in row[7:]] for row in EXP["mouse_mvlmm_rows"]["rows"]])
    assert np.allclose(got[:5], exp, rtol=2e-6, atol=0)                                                   # demo.txt:62-66
    if os.path.exists(REF.EXE):
        cwd = str(tmp_path)
        base = ["-g", d + "/mouse_hs1940.geno.txt.gz", "-p", d + "/mouse_hs1940.pheno.txt", "-a", d + "/mouse_hs1940.anno.txt"]
        REF.run_cli(base + ["-gk", "-o", "mouse"], cwd)
        REF.run_cli(base + ["-n", "1", "6", "-k", "output/mouse.cXX.txt", "-lmm", "-o", "mv"], cwd)
        lines = open(os.path.join(cwd, "output", "mv.assoc.txt")).read().splitlines()
        refrows = np.array([[float(x) for x in ln.split("\t")[7:]] for ln in lines[1:65]])
        assert np.allclose(got, refrows, rtol=3e-6, atol=0)
        # whole file through the CLI
        import subprocess
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        cli = os.path.join(root, "gemma_b200", "host", "gemma-b200")
        out = os.path.join(cwd, "output")
        r = subprocess.run([cli] + base + ["-n", "1", "6", "-k", os.path.join(out, "mouse.cXX.txt"), "-lmm", "-o", "mymv", "-outdir", out],
                           capture_output=True, text=True, cwd=cwd)
        assert r.returncode == 0, r.stdout + r.stderr
        mine = open(os.path.join(out, "mymv.assoc.txt")).read().splitlines()
        assert len(mine) == len(lines) and mine[0] == lines[0]
        fa = [x.split("\t") for x in mine[1:]]; fb = [x.split("\t") for x in lines[1:]]
        assert [x[:7] for x in fa] == [x[:7] for x in fb]
        xa = np.array([[float(v) for v in x[7:]] for x in fa]); xb = np.array([[float(v) for v in x[7:]] for x in fb])
        bad, rel = _record_cli_deviations("mvlmm:-lmm 1 whole mouse file", mine[0].split("\t"), fa, xa, xb, 5e-6)
        # MphEM stops on |dlogl| < em_prec (src/mvlmm.cpp:1177-1181): a borderline step flips the iteration count; counted + named in the report
        assert bad.any(axis=1).mean() <= 0.002 and rel.max() < 1e-3, (int(bad.any(axis=1).sum()), float(rel.max()))
        # -lmm 4 (Wald + LRT + score) on every 12th SNP
        snps = os.path.join(cwd, "sub.txt")
        with open(snps, "w") as f:
            f.write("\n".join(bb.rs[::12]) + "\n")
        REF.run_cli(base + ["-n", "1", "6", "-snps", snps, "-k", "output/mouse.cXX.txt", "-lmm", "4", "-o", "mv4"], cwd)
        r = subprocess.run([cli] + base + ["-n", "1", "6", "-snps", snps, "-k", os.path.join(out, "mouse.cXX.txt"), "-lmm", "4", "-o", "mymv4", "-outdir", out],
                           capture_output=True, text=True, cwd=cwd)
        assert r.returncode == 0, r.stdout + r.stderr
        a4 = open(os.path.join(out, "mymv4.assoc.txt")).read().splitlines(); b4 = open(os.path.join(out, "mv4.assoc.txt")).read().splitlines()
        assert len(a4) == len(b4) and a4[0] == b4[0]
        xa = np.array([[float(v) for v in x.split("\t")[7:]] for x in a4[1:]]); xb = np.array([[float(v) for v in x.split("\t")[7:]] for x in b4[1:]])
        bad, rel = _record_cli_deviations("mvlmm:-lmm 4 every 12th mouse SNP", a4[0].split("\t"), [x.split("\t") for x in a4[1:]], xa, xb, 5e-6)
        assert bad.any(axis=1).mean() <= 0.005 and rel.max() < 1e-3, (int(bad.any(axis=1).sum()), float(rel.max()))


# ---- kinship on the int8 tensor pipe (exact Z Z^T + rank-one centring) vs the FP64 oracle -------------
def test_kinship_int8_tensor_core_path_matches_oracle(ctx):
    n = 1300
    bed1, G1 = synth.make_bed(n, 700, seed=61)                       # no missing genotypes -> int8 path
    bed2, G2 = synth.make_bed(n, 333, seed=62, miss_rate=0.02)       # missing -> int8 GEMM + sparse FP64 correction terms
    bed3, G3 = synth.make_bed(n, 129, seed=63)                       # int8 again (odd size: padded to 256 columns)
    G = np.vstack([G1, G2, G3]); Gn = np.where(G < 0, np.nan, G)
    Xc = O.kin_transform(Gn, 1)
    Kref = Xc @ Xc.T / G.shape[0]
    ctx.profile_enable(True); ctx.profile_reset()
    ctx.kin_begin(n, 1)
    ctx.kin_add_bed(bed1); ctx.kin_add_bed(bed2); ctx.kin_add_bed(bed3)
    K, ns = ctx.kin_finish()
    ctx.profile_enable(False)
    assert ns == G.shape[0]
    assert np.allclose(K, Kref, rtol=1e-10, atol=1e-12)
    assert np.array_equal(K, K.T)
    # forcing the FP64 path gives the same matrix
    ctx.set_option("kin_path", 1)
    ctx.kin_begin(n, 1)
    ctx.kin_add_bed(np.vstack([bed1, bed2, bed3]))
    K2, _ = ctx.kin_finish()
    ctx.set_option("kin_path", 0)
    assert np.allclose(K2, K, rtol=1e-10, atol=1e-12)


def test_kinship_missing_genotypes_sparse_terms_and_dense_fallback(ctx):
    """Missing genotypes (imputed to the SNP mean by the reference, src/gemma_io.cpp:1688-1706): the int8 GEMM sees 0 there and
    the per-individual sparse kernel adds the X + X^T + G3 - b terms; chunks above the missing-rate limit take the dense FP64
    path.  Both must reproduce the oracle, also with a whole individual / a whole-SNP-but-one missing and n not a multiple of 16."""
    n = 1237
    bed1, G1 = synth.make_bed(n, 517, seed=71, miss_rate=0.004)
    bed2, G2 = synth.make_bed(n, 300, seed=72, miss_rate=0.08)
    bed3, G3 = synth.make_bed(n, 260, seed=73, miss_rate=0.35)       # above the default 20% limit -> dense FP64 for this call
    # individual 5 missing everywhere in batch 1; SNP 9 of batch 2 observed in only 3 individuals
    def set_missing(bed, G, s, i):
        byte, sh = i >> 2, 2 * (i & 3)
        bed[s, byte] = (int(bed[s, byte]) & (0xFF ^ (3 << sh))) | (1 << sh)
        G[s, i] = -9
    for s_ in range(G1.shape[0]): set_missing(bed1, G1, s_, 5)
    for i_ in range(3, n): set_missing(bed2, G2, 9, i_)
    G = np.vstack([G1, G2, G3]); Gn = np.where(G < 0, np.nan, G)
    Xc = O.kin_transform(Gn, 1)
    Kref = Xc @ Xc.T / G.shape[0]
    ctx.profile_enable(True); ctx.profile_reset()
    ctx.kin_begin(n, 1)
    ctx.kin_add_bed(bed1); ctx.kin_add_bed(bed2); ctx.kin_add_bed(bed3)
    K, ns = ctx.kin_finish()
    prof = {k: ctx.profile_get(k) for k in ("kin", "fix")}
    ctx.profile_enable(False)
    assert ns == G.shape[0]
    assert prof["kin"][1] >= 2 and prof["fix"][1] == 2, prof          # int8 GEMM + dense FP64 chunk; 2 sparse passes
    assert np.allclose(K, Kref, rtol=1e-10, atol=1e-12)
    assert np.array_equal(K, K.T)
    ctx.set_option("kin_miss_max_permille", 0)                       # every chunk with a hole -> dense FP64
    ctx.kin_begin(n, 1)
    ctx.kin_add_bed(bed1); ctx.kin_add_bed(bed2); ctx.kin_add_bed(bed3)
    K2, _ = ctx.kin_finish()
    ctx.set_option("kin_miss_max_permille", 200)
    assert np.allclose(K2, K, rtol=1e-10, atol=1e-12)


# ---- CTA-pair (tcgen05 cta_group::2) variants of the tensor-core kernels ---------------------------------
def test_cta_pair_projection_and_kinship_match_single_cta():
    c = gemma_b200.Context(0)
    rng = np.random.default_rng(123)
    n, l = 1500, 700
    Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    c.lmm_setup(Q, synth.spectrum_like_kinship(n, 5), np.ones((n, 1)), rng.standard_normal(n))
    bed, G = synth.make_bed(n, l, seed=321, miss_rate=0.01)
    ref = (Q.T @ O.lmm_impute(np.where(G < 0, np.nan, G))).T
    c.set_option("utx_path", 2)
    for T in (6, 8, 4):
        c.set_option("n_slices", T)
        c.set_option("cta_pair", 0)
        single = c.lmm_project_bed(bed, n)
        c.set_option("cta_pair", 1)
        c.set_option("hole_gemm", 0)                                # holes by the gather kernel, like the single-CTA kernel
        pair = c.lmm_project_bed(bed, n)
        assert np.array_equal(single, pair), T                      # exact integer accumulation: bit-identical
        c.set_option("hole_gemm", 1)                                # holes by a second tensor-core pass when they are many (device-side switch):
        pair_h = c.lmm_project_bed(bed, n)                          # mean * U^T q then carries the plane rounding of U as well
        assert np.abs(pair_h - pair).max() / np.abs(ref).max() < (1e-6 if T == 4 else 1e-11), T
    assert np.abs(pair - ref).max() / np.abs(ref).max() < 1e-6
    bedk, Gk = synth.make_bed(n, 900, seed=77)
    Xc = O.kin_transform(Gk, 1)
    for kp in (0, 1):
        c.set_option("kin_cta_pair", kp)
        c.kin_begin(n, 1); c.kin_add_bed(bedk); K, _ = c.kin_finish()
        assert np.allclose(K, Xc @ Xc.T / 900, rtol=1e-10, atol=1e-12), kp
    c.close()


def test_plink_entry_point_follows_analyzeplink_nan_rule(ctx):
    """gb200_lmm_batch_bed mirrors LMM::AnalyzePlink (src/lmm.cpp:1866-1884): when the lambda search fails the Wald
    test is skipped and p_wald / p_lrt are NaN; the dense / geno entry points mirror LMM::Analyze."""
    n, l = 300, 64
    pb = random_problem(n, 1, 4, 17, causal=False)
    bed, G = synth.make_bed(n, l, seed=18)
    X = np.ascontiguousarray(G.T)
    ctx.lmm_setup(pb["U"], pb["ev"], pb["W"], pb["y"])
    UtX = pb["U"].T @ X
    n_nan = 0
    l_re, _ = O.calc_lambda_null("R", pb["ev"], pb["UtW"], pb["Uty"])
    # a lower bound just under the null REML estimate makes Newton step out of [l_min, l_max] for some SNPs; which
    # factor does it depends on the last bits of the eigendecomposition, so probe with the (cheap) oracle first
    ranges = [(1e-5, 1e5, 10), (0.2, 0.25, 3)]
    for fac in (0.9999, 0.9997, 0.9995, 0.999, 0.998, 0.995, 0.99, 0.98, 0.95):
        for nr in (9, 5, 3):
            lo, hi = l_re * fac, 1e5
            l_mle, logl = O.calc_lambda_null("L", pb["ev"], pb["UtW"], pb["Uty"], lo, hi, nr)
            probe = O.lmm_analyze_utx(pb["ev"], pb["UtW"], pb["Uty"], UtX, 4, lo, hi, nr, l_mle, logl, plink=True)
            if np.isnan(probe["p_wald"]).any() and len(ranges) < 5:
                ranges.append((lo, hi, nr))
    for (lo, hi, nr) in ranges:
        l_mle, logl = O.calc_lambda_null("L", pb["ev"], pb["UtW"], pb["Uty"], lo, hi, nr)
        ctx.lmm_params(4, lo, hi, nr, l_mle, logl)
        for kern in (1, 2):
            ctx.set_option("lmm_kernel", kern)
            got_p = ctx.lmm_batch_bed(bed, n)
            got_b = ctx.lmm_batch(X)
            ref_p = O.lmm_analyze_utx(pb["ev"], pb["UtW"], pb["Uty"], UtX, 4, lo, hi, nr, l_mle, logl, plink=True)
            ref_b = O.lmm_analyze_utx(pb["ev"], pb["UtW"], pb["Uty"], UtX, 4, lo, hi, nr, l_mle, logl, plink=False)
            for got, ref in ((got_p, ref_p), (got_b, ref_b)):
                # a SNP whose Newton step lands within rounding of the range boundary may flip between the NaN and the
                # finite branch with the summation order: tolerate one such SNP per range, demand parity on the rest
                flip = np.isnan(got["logl_H1"]) != np.isnan(ref["logl_H1"])
                assert flip.sum() <= 1
                check_sumstat(got[~flip], ref[~flip], 4)
        n_nan += int(np.isnan(ref_p["p_wald"]).sum())
    ctx.set_option("lmm_kernel", 0)
    if len(ranges) > 2:
        assert n_nan > 0          # at least one probed range exercised the NaN branch through both kernels


def test_subbatch_pipeline_is_bitwise_identical_to_serial(ctx):
    """Batches >= 4096 SNPs are software-pipelined (projection of sub-batch i+1 overlaps the tests of sub-batch i on a
    second stream): same kernels, same inputs -> identical bits, in SNP order."""
    n, l = 1280, 4500
    pb = random_problem(n, 1, 4, 33)
    bed, G = synth.make_bed(n, l, seed=34, miss_rate=0.005)
    ctx.lmm_setup(pb["U"], pb["ev"], pb["W"], pb["y"])
    nm = ctx.lmm_null(pb["trace_G"])
    ctx.lmm_params(4, l_mle_null=nm["l_mle_null"], logl_mle_H0=nm["logl_mle_H0"])
    ctx.set_option("x_exact", 0)       # the pipelined variant runs without the exact x-sums (one buffer of them per context): compare like with like
    ctx.set_option("overlap", 0)
    serial = ctx.lmm_batch_bed(bed, n)
    ctx.set_option("overlap", 1)
    piped = ctx.lmm_batch_bed(bed, n)
    ctx.set_option("overlap", 0); ctx.set_option("x_exact", 1)
    for k in serial.dtype.names:
        assert np.array_equal(serial[k], piped[k], equal_nan=True), k
    idx = np.arange(0, l, 211)
    X = O.lmm_impute(np.where(G[idx] < 0, np.nan, G[idx]))
    ref = O.lmm_analyze_utx(pb["ev"], pb["UtW"], pb["Uty"], pb["U"].T @ X, 4, l_mle_null=nm["l_mle_null"],
                            logl_mle_H0=nm["logl_mle_H0"], plink=True)
    check_sumstat(piped[idx], ref, 4)


# ---- PLINK trio through the CLI (AnalyzePlink / PlinkKin surface) against the oracle's PLINK restatement ---------
def _write_plink(prefix, bed, y, rs_prefix="snp"):
    l, nb = bed.shape
    n = len(y)
    with open(prefix + ".bed", "wb") as f:
        f.write(bytes([0x6C, 0x1B, 0x01])); f.write(bed.tobytes())
    with open(prefix + ".bim", "w") as f:
        for s in range(l):
            f.write("%d\t%s%d\t0\t%d\tA\tG\n" % (1 + s % 19, rs_prefix, s, 1000 + 10 * s))
    with open(prefix + ".fam", "w") as f:
        for i in range(n):
            f.write("F%d I%d 0 0 1 %s\n" % (i, i, "NA" if np.isnan(y[i]) else "%.10g" % y[i]))


def test_cli_plink_gk_and_lmm4_match_oracle(tmp_path):
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cli = os.path.join(root, "gemma_b200", "host", "gemma-b200")
    subprocess.check_call(["make", "-s", "-C", os.path.dirname(cli)])     # always: a stale binary must not pass for the source
    n, l = 1150, 900
    rng = np.random.default_rng(3)
    bed, G = synth.make_bed(n, l, seed=444, miss_rate=0.01)
    bed[5] = 0xFF                                     # a monomorphic SNP (all 0) -> dropped by QC
    y = rng.standard_normal(n) + 0.4 * np.where(G[7] < 0, 0, G[7])
    y[rng.choice(n, 37, replace=False)] = np.nan      # individuals without phenotype are excluded from the LMM
    prefix = str(tmp_path / "syn")
    _write_plink(prefix, bed, y)
    out = str(tmp_path / "out")
    r = subprocess.run([cli, "-bfile", prefix, "-gk", "1", "-o", "k", "-outdir", out], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    r = subprocess.run([cli, "-bfile", prefix, "-k", out + "/k.cXX.txt", "-lmm", "4", "-o", "a", "-outdir", out],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    # oracle side
    pl = R.Plink(prefix)
    idv, W = R.process_cvt_phen(pl.ind_pheno)
    isnp, n_miss, maf = R.qc_plink(pl, idv)
    assert "## number of analyzed individuals = %d" % int(idv.sum()) in r.stdout
    assert "## number of analyzed SNPs         = %8d" % int(isnp.sum()) in r.stdout and isnp[5] == 0
    K = np.loadtxt(out + "/k.cXX.txt")
    Kref = R.kinship_plink(pl, isnp, 1)
    assert np.allclose(K, Kref, rtol=1e-8, atol=1e-9)                      # 10 significant digits in the text file
    prep = R.lmm_prepare(K, idv, pl.pheno[:, 0], W)
    keep = idv == 1
    X = O.lmm_impute(pl.G[np.ix_(np.nonzero(isnp)[0], keep)])
    ref = O.lmm_analyze_utx(prep["eval"], prep["UtW"], prep["Uty"], prep["U"].T @ X, 4, l_mle_null=prep["l_mle_null"],
                            logl_mle_H0=prep["logl_mle_H0"], plink=True)
    lines = open(out + "/a.assoc.txt").read().splitlines()
    hdr = lines[0].split("\t")
    assert hdr == ["chr", "rs", "ps", "n_miss", "allele1", "allele0", "af", "beta", "se", "logl_H1", "l_remle", "l_mle",
                   "p_wald", "p_lrt", "p_score"]
    assert len(lines) == 1 + int(isnp.sum())
    sel = np.nonzero(isnp)[0]
    got = np.array([[float(x) for x in ln.split("\t")[7:]] for ln in lines[1:]])
    for ln, s in zip(lines[1:], sel):
        f = ln.split("\t")
        assert f[1] == "snp%d" % s and int(f[3]) == int(n_miss[s]) and f[6] == "%.3f" % maf[s]
    # text carries 7 significant digits; the eigendecomposition differs (cuSOLVER vs LAPACK) only in rounding
    for col, key in zip(range(8), ("beta", "se", "logl_H1", "lambda_remle", "lambda_mle", "p_wald", "p_lrt", "p_score")):
        tol = 2e-4 if key.startswith("lambda") else 2e-6
        assert np.allclose(got[:, col], ref[key], rtol=tol, atol=1e-12), key
    # -gwasnps with PLINK input: every row of the filtered run is, text for text, the row of the same SNP in the unfiltered run
    # (the reference's AnalyzePlink tests all SNPs while WriteFiles skips rows, which shifts statistics onto other SNPs)
    sub = [int(s) for s in sel[3::7]]
    gw = str(tmp_path / "gw.txt")
    with open(gw, "w") as f:
        f.write("".join("snp%d\n" % s for s in sub))
    r = subprocess.run([cli, "-bfile", prefix, "-gwasnps", gw, "-k", out + "/k.cXX.txt", "-lmm", "4", "-o", "g", "-outdir", out],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    full = {ln.split("\t")[1]: ln for ln in lines[1:]}
    glines = open(out + "/g.assoc.txt").read().splitlines()
    assert glines[0] == lines[0] and len(glines) == 1 + len(sub)
    assert glines[1:] == [full["snp%d" % s] for s in sub]


def test_qc_bed_statistics_match_host_restatement(ctx):
    """gb200_qc_bed (device counting pass of ReadFile_bed + r2 terms) vs the oracle's numpy restatement."""
    n_total, l = 777, 300
    rng = np.random.default_rng(8)
    bed, G = synth.make_bed(n_total, l, seed=9, miss_rate=0.03)
    mask = np.ones(n_total, dtype=np.uint8); mask[rng.choice(n_total, 40, replace=False)] = 0
    Gk = np.where(G < 0, np.nan, G)[:, mask == 1]
    W = np.column_stack([rng.standard_normal(int(mask.sum())), np.ones(int(mask.sum()))])
    st = ctx.qc_bed(bed, n_total, mask, W)
    miss = np.isnan(Gk)
    assert np.array_equal(st["n_miss"], miss.sum(axis=1))
    assert np.array_equal(st["n_0"], (Gk == 0).sum(axis=1)) and np.array_equal(st["n_1"], (Gk == 1).sum(axis=1))
    assert np.array_equal(st["n_2"], (Gk == 2).sum(axis=1))
    maf = np.nansum(Gk, axis=1) / (2.0 * (Gk.shape[1] - miss.sum(axis=1)))
    assert np.array_equal(st["maf"], maf)                                   # integer sums: bit-exact
    X = np.where(miss, (2.0 * maf)[:, None], Gk)
    WtWi = np.linalg.inv(W.T @ W)
    Wtx = X @ W
    assert np.allclose(st["v_x"], (X * X).sum(axis=1), rtol=1e-13)
    assert np.allclose(st["v_w"], np.einsum("sa,ab,sb->s", Wtx, WtWi, Wtx), rtol=1e-10)
    st2 = ctx.qc_bed(bed, n_total)                                           # no mask, no covariates
    assert np.array_equal(st2["n_miss"], (G < 0).sum(axis=1)) and np.all(st2["v_w"] == 0)


@pytest.mark.gpu
def test_device_tails_match_the_restated_gsl_tails_including_the_asymptotic_branch(ctx):
    """gsl_cdf_fdist_Q(x, 1, df) and gsl_cdf_chisq_Q(x, 1) as the kernels compute them (gb200_cdf_tails) against the oracle's
    restatement, over df on both sides of GSL's df/2 > 1e5 switch to the A&S 26.5.17 form (cdf/beta_inc.c; call sites
    src/lmm.cpp:1161,1206,1553)."""
    xs, dfs = [], []
    for df in (10.0, 194.0, 1408.0, 49998.0, 199998.0, 200004.0, 300000.0, 1000000.0):
        for x in (1e-8, 0.01, 0.5, 1.0, 3.0, 10.0, 30.0, 60.0, 300.0, 2000.0, df - 0.5, df + 0.5, 5.0 * df):
            xs.append(x); dfs.append(df)
    got = ctx.cdf_tails(np.array(xs), np.array(dfs))
    want = np.array([O.fdist_Q(x, 1.0, d) for x, d in zip(xs, dfs)])
    big = want > 1e-290
    assert np.all(np.abs(got[big] - want[big]) <= 1e-9 * want[big] + 1e-17), np.abs(got - want)[big].max()   # 1 - P leaves ~1e-16 absolute
    assert np.all(got[~big] <= 1e-290)
    assert np.all(got[(np.array(dfs) > 2e5) & (np.array(xs) == 300.0)] == 0.0)       # the reference's own 1 - P = 0
    xc = np.array([-1.0, 0.0, 1e-12, 0.3, 1.0, 10.0, 100.0, 1400.0])
    assert np.allclose(ctx.cdf_tails(xc), [O.chisq1_Q(x) for x in xc], rtol=1e-12, atol=0)


@pytest.mark.gpu
def test_dosage_rows_take_the_tensor_core_projection_as_exact_digit_rows(ctx):
    """BIMBAM mean genotypes (src/lmm.cpp:1590-1618, dgemm at :1521): values printed with up to 6 decimals are integers over a power
    of ten, so each SNP row goes through the int8 GEMM as 1-3 exact base-256 digit rows (i8gemm_sm100.cu, i8_project_geno).  Same
    statistics as the FP64 projection and as the oracle; a batch with a long decimal stays on the FP64 GEMM."""
    n, l = 1300, 96
    rng = np.random.default_rng(77)
    pb = random_problem(n, 1, 4, 78)
    f = rng.uniform(0.05, 0.5, l)
    G = rng.binomial(2, f[:, None], size=(l, n)).astype(np.float64)              # SNP-major, plain 0/1/2 rows ...
    for s, dec in ((3, 1), (7, 2), (8, 3), (20, 4), (21, 5), (40, 6), (41, 3)):      # ... and dosage rows with 1..6 decimals
        G[s] = np.round(np.clip(G[s] + rng.normal(0, 0.3, n), 0, 2), dec)
    G[50] = np.round(rng.uniform(0, 2, n), 2)
    G[rng.random(G.shape) < 0.01] = np.nan                                       # missing entries: mean-imputed (lmm.cpp:1611-1618)
    X = O.lmm_impute(G)
    y = pb["y"] + 0.5 * (X[:, 8] - X[:, 8].mean())
    ctx.lmm_setup(pb["U"], pb["ev"], pb["W"], y)
    nm = ctx.lmm_null(pb["trace_G"])
    ctx.lmm_params(4, l_mle_null=nm["l_mle_null"], logl_mle_H0=nm["logl_mle_H0"])
    ref = O.lmm_analyze_utx(pb["ev"], pb["U"].T @ pb["W"], pb["U"].T @ y, pb["U"].T @ X, 4,
                            l_mle_null=nm["l_mle_null"], logl_mle_H0=nm["logl_mle_H0"])
    ctx.set_option("utx_path", 1)
    fp = ctx.lmm_batch_geno(G)
    ctx.set_option("utx_path", 0)
    ctx.profile_enable(True); ctx.profile_reset()
    got = ctx.lmm_batch_geno(G)
    assert ctx.profile_get("fix")[1] >= 1                     # the digit-row path ran (its recombination pass is profiled as "fix")
    check_sumstat(got, ref, 4)
    check_sumstat(fp, ref, 4)
    for k in ("beta", "se", "p_wald", "p_lrt", "p_score"):
        m = ~np.isnan(fp[k])
        assert np.allclose(got[k][m], fp[k][m], rtol=1e-7, atol=1e-300), k
    # one value with nine decimals: the batch is not representable, the FP64 GEMM takes it (same answers)
    G2 = G.copy(); G2[5, 11] = 0.123456789
    ctx.profile_reset()
    got2 = ctx.lmm_batch_geno(G2)
    assert ctx.profile_get("fix")[1] == 0
    ctx.profile_enable(False)
    X2 = O.lmm_impute(G2)
    ref2 = O.lmm_analyze_utx(pb["ev"], pb["U"].T @ pb["W"], pb["U"].T @ y, pb["U"].T @ X2, 4,
                             l_mle_null=nm["l_mle_null"], logl_mle_H0=nm["logl_mle_H0"])
    check_sumstat(got2, ref2, 4)


@pytest.mark.gpu
@pytest.mark.parametrize("c", [1, 3])
def test_exact_x_sums_remove_the_plane_rounding_from_beta(ctx, c):
    """Four digit planes leave ~1e-9 of rounding noise on every projected value; beta = P_xy / P_xx inherits it divided by the
    SNP's |z-score|.  The bed entry points therefore also form x . v_q, v_q = U (h(l_mle_null) (.) q), in genotype space (an FP64
    dot product, no planes) and the per-SNP kernel replaces the projected order-1 x-sums of the score test / final Wald tables by
    exact + (interpolated - projected at l_mle_null) (LmmConst::xex): the noise reaches beta only in second order."""
    n, l = 1536, 400
    pb = random_problem(n, c, 4, 191 + c)
    bed, G = synth.make_bed(n, l, seed=192, miss_rate=0.01)
    X = O.lmm_impute(np.where(G < 0, np.nan, G))
    y = pb["y"] + 0.5 * (X[:, 7] - X[:, 7].mean())
    ctx.lmm_setup(pb["U"], pb["ev"], pb["W"], y)
    nm = ctx.lmm_null(pb["trace_G"])
    ctx.lmm_params(4, l_mle_null=nm["l_mle_null"], logl_mle_H0=nm["logl_mle_H0"])
    ref = O.lmm_analyze_utx(pb["ev"], pb["U"].T @ pb["W"], pb["U"].T @ y, pb["U"].T @ X, 4,
                            l_mle_null=nm["l_mle_null"], logl_mle_H0=nm["logl_mle_H0"])
    ctx.set_option("utx_path", 2); ctx.set_option("n_slices", 4)
    errs = {}
    try:
        for xe in (0, 1):
            ctx.set_option("x_exact", xe)
            got = ctx.lmm_batch_bed(bed, n)
            errs[xe] = {k: float(np.nanmax(np.abs(got[k] - ref[k]) / np.maximum(np.abs(ref[k]), 1e-300))) for k in ("beta", "se", "p_wald", "p_lrt", "p_score")}
    finally:
        ctx.set_option("x_exact", 1); ctx.set_option("utx_path", 0); ctx.set_option("n_slices", 0)
    assert errs[1]["beta"] < 1e-7 and errs[1]["beta"] < errs[0]["beta"] / 10, errs      # second order: the correction is taken at l_mle_null, the Wald table at lambda_hat
    for k in ("se", "p_wald", "p_lrt", "p_score"):
        assert errs[1][k] < 1e-7, errs
