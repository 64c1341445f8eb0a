"""ctypes binding of oracle/_ref/libgemma_ref.so: the REFERENCE's own src/lmm.cpp, mathfunc.cpp, gemma_io.cpp, ... compiled
in place against the GSL API shim (oracle/gsl_shim/, recipe: `make -C oracle ref`).

TEST INFRASTRUCTURE ONLY: used by tests/test_oracle_vs_ref.py to validate the restated oracle against the code it restates,
and by bench.py's reference arm.  Never imported by the product."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "libgemma_ref.so")
REF_SRC = "/root/reference/src"
_LIB = None
_dp = C.POINTER(C.c_double)


def available():
    return os.path.exists(SO) or os.path.isdir(REF_SRC)


def build():
    """(Re)build when the reference sources are present; otherwise use the prebuilt library shipped with the snapshot."""
    if os.path.isdir(REF_SRC):
        subprocess.check_call(["make", "-C", _HERE, "-s", "ref"])
    if not os.path.exists(SO):
        raise RuntimeError("oracle/_ref/libgemma_ref.so is missing and /root/reference is absent")
    return SO


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.ref_getab_index.restype = C.c_size_t
        L.ref_getab_index.argtypes = [C.c_size_t] * 3
        L.ref_eval_fn.restype = C.c_double
        L.ref_eval_fn.argtypes = [C.c_char, C.c_int, C.c_int, C.c_double, C.c_size_t, C.c_size_t, _dp, _dp, C.c_size_t, _dp, _dp]
        L.ref_null_model.restype = C.c_int
        L.ref_null_model.argtypes = [C.c_size_t, C.c_size_t, _dp, _dp, C.c_size_t, _dp, C.c_double, C.c_double, C.c_size_t, C.c_double,
                                     _dp, _dp, _dp, _dp, _dp, _dp]
        L.ref_lmm_analyze.restype = C.c_int
        L.ref_lmm_analyze.argtypes = [C.c_size_t, C.POINTER(C.c_int), C.c_size_t, C.c_size_t, _dp, _dp, _dp, _dp, _dp, _dp, _dp, C.c_size_t,
                                      C.c_int, C.c_double, C.c_double, C.c_size_t, C.c_double, C.c_double, _dp]
        L.ref_assoc_utx.restype = C.c_int
        L.ref_assoc_utx.argtypes = [C.c_size_t, C.c_size_t, _dp, _dp, _dp, _dp, C.c_size_t, C.c_size_t, C.c_int, C.c_double, C.c_double, C.c_size_t,
                                    C.c_double, C.c_double, _dp]
        L.ref_qc_bimbam.restype = C.c_long
        L.ref_qc_bimbam.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.c_size_t, _dp, C.c_size_t, C.c_size_t, C.c_double, C.c_double, C.c_double,
                                    C.c_double, C.POINTER(C.c_int), C.POINTER(C.c_long), _dp, C.c_size_t, C.POINTER(C.c_long)]
        L.ref_qc_plink.restype = C.c_long
        L.ref_qc_plink.argtypes = L.ref_qc_bimbam.argtypes
        L.ref_plink_kin.restype = C.c_int
        L.ref_plink_kin.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.c_size_t, C.c_int, C.c_size_t, _dp]
        L.ref_lmm_analyze_plink.restype = C.c_int
        L.ref_lmm_analyze_plink.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_size_t, C.c_size_t, C.c_size_t,
                                            _dp, _dp, _dp, _dp, _dp, _dp, C.c_int, C.c_double, C.c_double, C.c_size_t, C.c_double, C.c_double,
                                            _dp, C.c_size_t]
        L.ref_bimbam_kin.restype = C.c_int
        L.ref_bimbam_kin.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.c_size_t, C.c_int, C.c_size_t, _dp]
        L.ref_read_kin.restype = C.c_int
        L.ref_read_kin.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.c_size_t, C.POINTER(C.c_char_p), C.c_int, _dp, C.c_size_t]
        L.ref_center_matrix.restype = None
        L.ref_center_matrix.argtypes = [_dp, C.c_size_t]
        _LIB = L
    return _LIB


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return a.ctypes.data_as(_dp)


SUMSTAT = ("beta", "se", "lambda_remle", "lambda_mle", "p_wald", "p_lrt", "p_score", "logl_H1")


def getab_index(a, b, n_cvt):
    return lib().ref_getab_index(a, b, n_cvt)


def eval_fn(fn, which, calc_null, l, eval_, UtW, Uty, Utx=None):
    eval_, UtW, Uty = _f(eval_), _f(UtW), _f(Uty)
    n, c = UtW.shape
    x = _f(Utx) if Utx is not None else None
    return lib().ref_eval_fn(fn.encode(), which, int(calc_null), l, n, c, _p(eval_), _p(UtW), c, _p(Uty), _p(x) if x is not None else None)


def null_model(eval_, UtW, Uty, trace_G, l_min=1e-5, l_max=1e5, n_region=10):
    eval_, UtW, Uty = _f(eval_), _f(UtW), _f(Uty)
    n, c = UtW.shape
    out = np.zeros(8); br, sr, bm, sm, vv = np.zeros(c), np.zeros(c), np.zeros(c), np.zeros(c), np.zeros(2)
    lib().ref_null_model(n, c, _p(eval_), _p(UtW), c, _p(Uty), l_min, l_max, n_region, trace_G, _p(out), _p(br), _p(sr), _p(bm), _p(sm), _p(vv))
    return dict(l_mle_null=out[0], logl_mle_H0=out[1], l_remle_null=out[2], logl_remle_H0=out[3], pve_null=out[4], pve_se_null=out[5],
                vg_remle=out[6], ve_remle=out[7], vg_mle=vv[0], ve_mle=vv[1], beta_remle=br, se_beta_remle=sr, beta_mle=bm, se_beta_mle=sm)


def lmm_analyze(indicator_idv, U, eval_, UtW, Uty, W, y, G, a_mode, l_min=1e-5, l_max=1e5, n_region=10, l_mle_null=0.0, logl_mle_H0=0.0):
    """LMM::Analyze on G (l x ni_total SNP-major, NaN = missing).  Returns a structured array like oracle.SUMSTAT_DTYPE."""
    idv = np.ascontiguousarray(indicator_idv, dtype=np.int32)
    U, eval_, UtW, Uty, W, y, G = _f(U), _f(eval_), _f(UtW), _f(Uty), _f(W), _f(y), _f(G)
    n, c = UtW.shape
    l = G.shape[0]
    out = np.zeros((l, 8))
    got = lib().ref_lmm_analyze(len(idv), idv.ctypes.data_as(C.POINTER(C.c_int)), n, c, _p(U), _p(eval_), _p(UtW), _p(Uty), _p(W), _p(y), _p(G), l,
                                a_mode, l_min, l_max, n_region, l_mle_null, logl_mle_H0, _p(out))
    assert got == l, (got, l)
    r = np.zeros(l, dtype=[(k, "<f8") for k in SUMSTAT])
    for i, k in enumerate(SUMSTAT):
        r[k] = out[:, i]
    return r


def assoc_utx(eval_, UtW, Uty, UtX, a_mode, l_min=1e-5, l_max=1e5, n_region=10, l_mle_null=0.0, logl_mle_H0=0.0):
    """Per-SNP part of the reference's batch_compute on a given U^T X (n x l)."""
    eval_, UtW, Uty, UtX = _f(eval_), _f(UtW), _f(Uty), _f(UtX)
    n, c = UtW.shape
    l = UtX.shape[1]
    out = np.zeros((l, 8))
    lib().ref_assoc_utx(n, c, _p(eval_), _p(UtW), _p(Uty), _p(UtX), l, UtX.shape[1], a_mode, l_min, l_max, n_region, l_mle_null, logl_mle_H0, _p(out))
    r = np.zeros(l, dtype=[(k, "<f8") for k in SUMSTAT])
    for i, k in enumerate(SUMSTAT):
        r[k] = out[:, i]
    return r


def qc_bimbam(path, indicator_idv, W, maf_level=0.01, miss_level=0.05, hwe_level=0.0, r2_level=0.9999, cap=1 << 22):
    idv = np.ascontiguousarray(indicator_idv, dtype=np.int32)
    Wt = _f(W[idv == 1])
    isnp = np.zeros(cap, dtype=np.int32); n_miss = np.zeros(cap, dtype=np.int64); maf = np.zeros(cap); ns_test = C.c_long()
    tot = lib().ref_qc_bimbam(path.encode(), idv.ctypes.data_as(C.POINTER(C.c_int)), len(idv), _p(Wt), Wt.shape[0], Wt.shape[1], maf_level, miss_level,
                              hwe_level, r2_level, isnp.ctypes.data_as(C.POINTER(C.c_int)), n_miss.ctypes.data_as(C.POINTER(C.c_long)), _p(maf), cap,
                              C.byref(ns_test))
    assert tot >= 0
    return isnp[:tot].copy(), n_miss[:tot].copy(), maf[:tot].copy(), ns_test.value


def qc_plink(prefix, indicator_idv, W, maf_level=0.01, miss_level=0.05, hwe_level=0.0, r2_level=0.9999, cap=1 << 22):
    idv = np.ascontiguousarray(indicator_idv, dtype=np.int32)
    Wt = _f(W[idv == 1])
    isnp = np.zeros(cap, dtype=np.int32); n_miss = np.zeros(cap, dtype=np.int64); maf = np.zeros(cap); ns_test = C.c_long()
    tot = lib().ref_qc_plink(prefix.encode(), idv.ctypes.data_as(C.POINTER(C.c_int)), len(idv), _p(Wt), Wt.shape[0], Wt.shape[1], maf_level, miss_level,
                             hwe_level, r2_level, isnp.ctypes.data_as(C.POINTER(C.c_int)), n_miss.ctypes.data_as(C.POINTER(C.c_long)), _p(maf), cap,
                             C.byref(ns_test))
    assert tot >= 0
    return isnp[:tot].copy(), n_miss[:tot].copy(), maf[:tot].copy(), ns_test.value


def plink_kin(prefix, indicator_snp, k_mode, ni_total):
    isnp = np.ascontiguousarray(indicator_snp, dtype=np.int32)
    K = np.zeros((ni_total, ni_total))
    assert lib().ref_plink_kin(prefix.encode(), isnp.ctypes.data_as(C.POINTER(C.c_int)), len(isnp), k_mode, ni_total, _p(K)) == 0
    return K


def lmm_analyze_plink(prefix, indicator_idv, indicator_snp, U, eval_, UtW, Uty, W, y, a_mode, l_min=1e-5, l_max=1e5, n_region=10, l_mle_null=0.0,
                      logl_mle_H0=0.0):
    idv = np.ascontiguousarray(indicator_idv, dtype=np.int32); isnp = np.ascontiguousarray(indicator_snp, dtype=np.int32)
    U, eval_, UtW, Uty, W, y = _f(U), _f(eval_), _f(UtW), _f(Uty), _f(W), _f(y)
    n, c = UtW.shape
    cap = int(isnp.sum())
    out = np.zeros((cap, 8))
    got = lib().ref_lmm_analyze_plink(prefix.encode(), len(idv), idv.ctypes.data_as(C.POINTER(C.c_int)), isnp.ctypes.data_as(C.POINTER(C.c_int)), len(isnp),
                                      n, c, _p(U), _p(eval_), _p(UtW), _p(Uty), _p(W), _p(y), a_mode, l_min, l_max, n_region, l_mle_null, logl_mle_H0,
                                      _p(out), cap)
    assert got == cap, (got, cap)
    r = np.zeros(cap, dtype=[(k, "<f8") for k in SUMSTAT])
    for i, k in enumerate(SUMSTAT):
        r[k] = out[:, i]
    return r


def bimbam_kin(path, indicator_snp, k_mode, ni_total):
    isnp = np.ascontiguousarray(indicator_snp, dtype=np.int32)
    K = np.zeros((ni_total, ni_total))
    rc = lib().ref_bimbam_kin(path.encode(), isnp.ctypes.data_as(C.POINTER(C.c_int)), len(isnp), k_mode, ni_total, _p(K))
    assert rc == 0
    return K


def read_kin(path, indicator_idv, k_mode=1, ids=None):
    """ReadFile_kin: the kinship matrix of the analysed individuals as the reference reads it from a -km 1 / -km 2 file."""
    idv = np.ascontiguousarray(indicator_idv, dtype=np.int32)
    n = int(idv.sum())
    G = np.zeros((n, n))
    arr = None
    if ids is not None:
        arr = (C.c_char_p * len(ids))(*[s.encode() for s in ids])
    rc = lib().ref_read_kin(path.encode(), idv.ctypes.data_as(C.POINTER(C.c_int)), len(idv), arr, k_mode, _p(G), n)
    assert rc == 0
    return G


def center_matrix(G):
    G = _f(G).copy()
    lib().ref_center_matrix(_p(G), G.shape[0])
    return G


EXE = os.path.join(_HERE, "_ref", "gemma_ref")


def run_cli(args, cwd):
    """Run the reference's whole CLI (oracle/_ref/gemma_ref: every src/*.cpp compiled in place against the GSL API shim, BLAS /
    LAPACK from the OpenBLAS bundled with scipy).  Returns stdout+stderr; outputs land in <cwd>/output like the reference's."""
    build()
    if not os.path.exists(EXE):
        raise RuntimeError("oracle/_ref/gemma_ref is missing")
    r = subprocess.run([EXE] + list(args), cwd=cwd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("reference CLI failed (%d): %s" % (r.returncode, (r.stdout + r.stderr)[-2000:]))
    return r.stdout + r.stderr
