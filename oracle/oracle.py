"""ctypes binding of the CPU oracle (oracle/gemma_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product (gemma_b200)
never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_double_p = C.POINTER(C.c_double)


class SumStat(C.Structure):
    """SUMSTAT, src/param.h:54-66."""
    _fields_ = [(k, C.c_double) for k in
                ("beta", "se", "lambda_remle", "lambda_mle", "p_wald", "p_lrt", "p_score", "logl_H1")]


SUMSTAT_DTYPE = np.dtype([(k, "<f8") for k in
                          ("beta", "se", "lambda_remle", "lambda_mle", "p_wald", "p_lrt", "p_score", "logl_H1")])


def build(force=False):
    so = os.path.join(_HERE, "libgemma_oracle.so")
    src = os.path.join(_HERE, "gemma_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libgemma_oracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.go_getab_index.restype = C.c_size_t
        L.go_getab_index.argtypes = [C.c_size_t] * 3
        L.go_cdf_fdist_Q.restype = C.c_double
        L.go_cdf_fdist_Q.argtypes = [C.c_double] * 3
        L.go_cdf_chisq1_Q.restype = C.c_double
        L.go_cdf_chisq1_Q.argtypes = [C.c_double]
        L.go_lmm_analyze_utx.restype = C.c_int
        L.go_lmm_analyze_utx.argtypes = [C.c_size_t, C.c_size_t, c_double_p, c_double_p, C.c_size_t,
                                         c_double_p, c_double_p, C.c_size_t, C.c_size_t, C.c_int,
                                         C.c_double, C.c_double, C.c_size_t, C.c_double, C.c_double,
                                         C.c_void_p, C.POINTER(C.c_long)]
        L.go_lmm_analyze_utx_plink.restype = C.c_int
        L.go_lmm_analyze_utx_plink.argtypes = L.go_lmm_analyze_utx.argtypes
        L.go_calc_lambda_null.restype = C.c_int
        L.go_calc_lambda_null.argtypes = [C.c_char, C.c_size_t, C.c_size_t, c_double_p, c_double_p,
                                          C.c_size_t, c_double_p, C.c_double, C.c_double, C.c_size_t,
                                          c_double_p, c_double_p]
        L.go_calc_pve.restype = C.c_int
        L.go_calc_pve.argtypes = [C.c_size_t, C.c_size_t, c_double_p, c_double_p, C.c_size_t, c_double_p,
                                  C.c_double, C.c_double, c_double_p, c_double_p]
        L.go_calc_vgvebeta.restype = C.c_int
        L.go_calc_vgvebeta.argtypes = [C.c_size_t, C.c_size_t, c_double_p, c_double_p, C.c_size_t,
                                       c_double_p, C.c_double, c_double_p, c_double_p, c_double_p,
                                       c_double_p]
        L.go_eval_fn.restype = C.c_double
        L.go_eval_fn.argtypes = [C.c_char, C.c_int, C.c_int, C.c_double, C.c_size_t, C.c_size_t,
                                 c_double_p, c_double_p, C.c_size_t, c_double_p, c_double_p]
        L.go_kin_transform.restype = C.c_int
        L.go_kin_transform.argtypes = [c_double_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int,
                                       c_double_p, C.c_size_t]
        L.go_kin_accumulate.restype = C.c_int
        L.go_kin_accumulate.argtypes = [c_double_p, C.c_size_t, C.c_size_t, C.c_size_t, c_double_p,
                                        C.c_size_t]
        L.go_center_matrix.restype = C.c_int
        L.go_center_matrix.argtypes = [c_double_p, C.c_size_t, C.c_size_t]
        L.go_zero_small_eval.restype = C.c_double
        L.go_zero_small_eval.argtypes = [c_double_p, C.c_size_t]
        L.go_bed_decode.restype = None
        L.go_bed_decode.argtypes = [C.c_char_p, C.c_size_t, c_double_p]
        L.go_lmm_impute.restype = None
        L.go_lmm_impute.argtypes = [c_double_p, C.c_size_t, C.c_size_t, C.c_size_t, c_double_p, C.c_size_t]
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(c_double_p)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def getab_index(a, b, n_cvt):
    return lib().go_getab_index(a, b, n_cvt)


def fdist_Q(x, nu1, nu2):
    return lib().go_cdf_fdist_Q(x, nu1, nu2)


def chisq1_Q(x):
    return lib().go_cdf_chisq1_Q(x)


def lmm_analyze_utx(eval_, UtW, Uty, UtX, a_mode, l_min=1e-5, l_max=1e5, n_region=10,
                    l_mle_null=0.0, logl_mle_H0=0.0, return_evals=False, plink=False):
    """UtX: n x l (SNP per column, like the reference's UtXlarge). Returns SUMSTAT array."""
    eval_, UtW, Uty, UtX = _f64(eval_), _f64(UtW), _f64(Uty), _f64(UtX)
    n, c = UtW.shape
    l = UtX.shape[1]
    out = np.zeros(l, dtype=SUMSTAT_DTYPE)
    nev = C.c_long(0)
    fn = lib().go_lmm_analyze_utx_plink if plink else lib().go_lmm_analyze_utx
    rc = fn(n, c, _p(eval_), _p(UtW), c, _p(Uty), _p(UtX), l, UtX.shape[1],
                                  a_mode, l_min, l_max, n_region, l_mle_null, logl_mle_H0,
                                  out.ctypes.data_as(C.c_void_p), C.byref(nev))
    assert rc == 0
    return (out, nev.value) if return_evals else out


def calc_lambda_null(func, eval_, UtW, Uty, l_min=1e-5, l_max=1e5, n_region=10):
    eval_, UtW, Uty = _f64(eval_), _f64(UtW), _f64(Uty)
    n, c = UtW.shape
    lam, logl = C.c_double(), C.c_double()
    lib().go_calc_lambda_null(func.encode(), n, c, _p(eval_), _p(UtW), c, _p(Uty), l_min, l_max,
                              n_region, C.byref(lam), C.byref(logl))
    return lam.value, logl.value


def calc_pve(eval_, UtW, Uty, lam, trace_G):
    eval_, UtW, Uty = _f64(eval_), _f64(UtW), _f64(Uty)
    n, c = UtW.shape
    pve, se = C.c_double(), C.c_double()
    lib().go_calc_pve(n, c, _p(eval_), _p(UtW), c, _p(Uty), lam, trace_G, C.byref(pve), C.byref(se))
    return pve.value, se.value


def calc_vgvebeta(eval_, UtW, Uty, lam):
    eval_, UtW, Uty = _f64(eval_), _f64(UtW), _f64(Uty)
    n, c = UtW.shape
    vg, ve = C.c_double(), C.c_double()
    beta, se = np.zeros(c), np.zeros(c)
    lib().go_calc_vgvebeta(n, c, _p(eval_), _p(UtW), c, _p(Uty), lam, C.byref(vg), C.byref(ve),
                           _p(beta), _p(se))
    return vg.value, ve.value, beta, se


def eval_fn(fn, which, calc_null, l, eval_, UtW, Uty, Utx=None):
    eval_, UtW, Uty = _f64(eval_), _f64(UtW), _f64(Uty)
    n, c = UtW.shape
    px = None
    if Utx is not None:
        Utx = _f64(Utx)
        px = _p(Utx)
    return lib().go_eval_fn(fn.encode(), which, int(calc_null), l, n, c, _p(eval_), _p(UtW), c,
                            _p(Uty), px)


def kin_transform(G, k_mode):
    """G: l x n SNP-major with NaN missing -> Xc n x l (centred / standardised)."""
    G = _f64(G)
    l, n = G.shape
    Xc = np.zeros((n, l))
    lib().go_kin_transform(_p(G), l, n, n, k_mode, _p(Xc), l)
    return Xc


def kin_accumulate(Xc, K):
    Xc = _f64(Xc)
    n, l = Xc.shape
    assert K.flags.c_contiguous and K.dtype == np.float64
    lib().go_kin_accumulate(_p(Xc), n, l, l, _p(K), K.shape[1])
    return K


def center_matrix(G):
    G = _f64(G).copy()
    lib().go_center_matrix(_p(G), G.shape[0], G.shape[1])
    return G


def zero_small_eval(ev):
    ev = _f64(ev).copy()
    tr = lib().go_zero_small_eval(_p(ev), ev.size)
    return ev, tr


def bed_decode(row_bytes, n):
    g = np.zeros(n)
    lib().go_bed_decode(bytes(row_bytes), n, _p(g))
    return g


def lmm_impute(G):
    G = _f64(G)
    l, n = G.shape
    X = np.zeros((n, l))
    lib().go_lmm_impute(_p(G), l, n, n, _p(X), l)
    return X
