"""ctypes binding of libgemma_b200.so (include/gemma_b200.h).

Host-side mirror of the reference's seams for Python callers (tests, bench): names and
argument meaning follow the C ABI one to one; numpy arrays are passed as caller-owned
host buffers.  There is no fallback: if the CUDA library is missing, import fails loudly.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GB200_LIB", os.path.join(_HERE, "csrc", "libgemma_b200.so"))   # override for A/B builds only

SUMSTAT_DTYPE = np.dtype([(k, "<f8") for k in
                          ("beta", "se", "lambda_remle", "lambda_mle", "p_wald", "p_lrt", "p_score", "logl_H1")])

_dp = C.POINTER(C.c_double)
_vp = C.c_void_p
_sz = C.c_size_t

STATUS = {0: "GB200_OK", 1: "GB200_ERR_ARG", 2: "GB200_ERR_CUDA", 3: "GB200_ERR_STATE",
          4: "GB200_ERR_UNSUPPORTED", 5: "GB200_ERR_NUMERIC", 6: "GB200_ERR_NOMEM"}


class NullModel(C.Structure):
    _fields_ = [(k, C.c_double) for k in
                ("l_mle_null", "logl_mle_H0", "l_remle_null", "logl_remle_H0", "pve_null", "pve_se_null",
                 "vg_mle", "ve_mle", "vg_remle", "ve_remle")]


class SnpQC(C.Structure):
    _fields_ = [("n_miss", C.c_int), ("n_0", C.c_int), ("n_1", C.c_int), ("n_2", C.c_int),
                ("maf", C.c_double), ("v_x", C.c_double), ("v_w", C.c_double)]


SNPQC_DTYPE = np.dtype([("n_miss", "<i4"), ("n_0", "<i4"), ("n_1", "<i4"), ("n_2", "<i4"), ("maf", "<f8"), ("v_x", "<f8"),
                        ("v_w", "<f8")])


class GB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("%s: %s" % (STATUS.get(code, code), msg))
        self.code = code


# every exported symbol with its signature; tests check the library exports exactly these
SIGNATURES = {
    "gb200_abi_version": (C.c_int, []),
    "gb200_create": (C.c_int, [C.POINTER(_vp), C.c_int, _vp]),
    "gb200_destroy": (None, [_vp]),
    "gb200_last_error": (C.c_char_p, [_vp]),
    "gb200_stream": (_vp, [_vp]),
    "gb200_synchronize": (C.c_int, [_vp]),
    "gb200_profile_enable": (C.c_int, [_vp, C.c_int]),
    "gb200_profile_reset": (C.c_int, [_vp]),
    "gb200_profile_get": (C.c_int, [_vp, C.c_char_p, _dp, C.POINTER(C.c_long)]),
    "gb200_measure_fp64_fma": (C.c_int, [_vp, C.c_double, _dp, _dp]),
    "gb200_lmm_counters": (C.c_int, [_vp, C.POINTER(C.c_ulonglong), C.c_int]),
    "gb200_cdf_tails": (C.c_int, [_vp, C.c_int, _vp, C.c_double, _vp, _vp, _sz]),
    "gb200_dgemm": (C.c_int, [_vp, C.c_char_p, C.c_char_p, C.c_double, _vp, _sz, _sz, _sz, _vp, _sz, _sz, _sz,
                              C.c_double, _vp, _sz, _sz, _sz]),
    "gb200_kin_begin": (C.c_int, [_vp, _sz, C.c_int]),
    "gb200_kin_add": (C.c_int, [_vp, _vp, _sz, _sz, _sz]),
    "gb200_kin_add_geno": (C.c_int, [_vp, _vp, _sz, _sz, _sz]),
    "gb200_kin_add_bed": (C.c_int, [_vp, _vp, _sz, _sz]),
    "gb200_kin_add_bed_dev": (C.c_int, [_vp, _vp, _sz, _sz]),
    "gb200_kin_finish": (C.c_int, [_vp, _vp, _sz, C.POINTER(_sz)]),
    "gb200_kin_finish_dev": (C.c_int, [_vp, C.POINTER(_vp), C.POINTER(_sz)]),
    "gb200_eigh": (C.c_int, [_vp, _vp, _sz, _sz, C.c_int, _vp, _sz, _vp, _dp, C.POINTER(C.c_int),
                             C.POINTER(C.c_int)]),
    "gb200_eigh_dev": (C.c_int, [_vp, _vp, _sz, C.c_int, _vp, _vp, _dp, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "gb200_qc_bed": (C.c_int, [_vp, _vp, _vp, _sz, _sz, _sz, _vp, _vp, _sz, _vp]),
    "gb200_lmm_setup": (C.c_int, [_vp, _sz, _sz, _vp, _sz, _vp, _vp, _sz, _vp, _vp, _vp]),
    "gb200_lmm_setup_rotated": (C.c_int, [_vp, _sz, _sz, _vp, _sz, _vp, _vp, _sz, _vp]),
    "gb200_lmm_setup_rotated_dev": (C.c_int, [_vp, _sz, _sz, _vp, _vp, _vp, _vp]),
    "gb200_lmm_null": (C.c_int, [_vp, C.c_double, C.c_double, _sz, C.c_double, C.POINTER(NullModel), _vp, _vp,
                                 _vp, _vp]),
    "gb200_lmm_params": (C.c_int, [_vp, C.c_int, C.c_double, C.c_double, _sz, C.c_double, C.c_double]),
    "gb200_lmm_batch": (C.c_int, [_vp, _vp, _sz, _sz, _vp]),
    "gb200_lmm_batch_geno": (C.c_int, [_vp, _vp, _sz, _sz, _vp]),
    "gb200_lmm_batch_bed": (C.c_int, [_vp, _vp, _vp, _sz, _sz, _sz, _vp]),
    "gb200_lmm_batch_bed_dev": (C.c_int, [_vp, _vp, _vp, _sz, _sz, _sz, _vp]),
    "gb200_mvlmm_setup": (C.c_int, [_vp, _sz, _sz, _sz, _vp, _sz, _vp, _vp, _sz, _vp, _sz]),
    "gb200_mvlmm_null": (C.c_int, [_vp] * 9),
    "gb200_mvlmm_batch_geno": (C.c_int, [_vp, _vp, _sz, _sz, C.c_int, _vp]),
    "gb200_mvlmm_batch_bed": (C.c_int, [_vp, _vp, _vp, _sz, _sz, _sz, C.c_int, _vp]),
    "gb200_lm_setup": (C.c_int, [_vp, _sz, _sz, _vp, _sz, _vp]),
    "gb200_lm_batch_geno": (C.c_int, [_vp, _vp, _sz, _sz, C.c_int, _vp]),
    "gb200_lm_batch_bed": (C.c_int, [_vp, _vp, _vp, _sz, _sz, _sz, C.c_int, _vp]),
    "gb200_lmm_gxe_setup": (C.c_int, [_vp, _vp]),
    "gb200_lmm_gxe_batch_geno": (C.c_int, [_vp, _vp, _sz, _sz, _vp]),
    "gb200_lmm_gxe_batch_bed": (C.c_int, [_vp, _vp, _vp, _sz, _sz, _sz, _vp]),
    "gb200_lmm_assoc_utx": (C.c_int, [_vp, _vp, _sz, _sz, _vp]),
    "gb200_lmm_project": (C.c_int, [_vp, _vp, _sz, _sz, _vp]),
    "gb200_lmm_project_bed": (C.c_int, [_vp, _vp, _vp, _sz, _sz, _sz, _vp]),
    "gb200_set_option": (C.c_int, [_vp, C.c_char_p, C.c_long]),
    "gb200_get_option": (C.c_int, [_vp, C.c_char_p, C.POINTER(C.c_long)]),
}

_LIB = None


def load_library():
    """Load libgemma_b200.so (built in-tree by __graft_entry__.build()). No fallback."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("libgemma_b200.so is missing at %s -- run `python -c 'import __graft_entry__ as g; "
                              "g.build()'` (there is no CPU fallback)" % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)            # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _LIB = lib
    return _LIB


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_vp)


class Context:
    """One context per GPU / host thread (gb200_create)."""

    def __init__(self, device=-1, stream=None):
        self.lib = load_library()
        h = _vp()
        rc = self.lib.gb200_create(C.byref(h), device, stream)
        if rc != 0:
            raise GB200Error(rc, "gb200_create failed (no CUDA device? there is no CPU fallback)")
        self.h = h
        self.n = 0
        self.n_cvt = 0

    def close(self):
        if getattr(self, "h", None):
            self.lib.gb200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc != 0:
            raise GB200Error(rc, self.lib.gb200_last_error(self.h).decode())

    # ---- misc
    def stream(self):
        return self.lib.gb200_stream(self.h)

    def synchronize(self):
        self._chk(self.lib.gb200_synchronize(self.h))

    def set_option(self, name, value):
        self._chk(self.lib.gb200_set_option(self.h, name.encode(), int(value)))

    def get_option(self, name):
        v = C.c_long()
        self._chk(self.lib.gb200_get_option(self.h, name.encode(), C.byref(v)))
        return v.value

    def profile_enable(self, on=True):
        self._chk(self.lib.gb200_profile_enable(self.h, int(on)))

    def profile_reset(self):
        self._chk(self.lib.gb200_profile_reset(self.h))

    def profile_get(self, name):
        ms, n = C.c_double(), C.c_long()
        self._chk(self.lib.gb200_profile_get(self.h, name.encode(), C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def measure_fp64_fma(self, seconds=0.5):
        t, ms = C.c_double(), C.c_double()
        self._chk(self.lib.gb200_measure_fp64_fma(self.h, seconds, C.byref(t), C.byref(ms)))
        return t.value, ms.value

    def lmm_counters(self, reset=False):
        a = (C.c_ulonglong * 6)()
        self._chk(self.lib.gb200_lmm_counters(self.h, a, int(reset)))
        return dict(zip(("common_slots", "two_power_passes", "three_power_passes", "with_logdet", "unused", "snps"), [int(x) for x in a]))

    def cdf_tails(self, x, nu2=None, nu1=1.0):
        """Device restatement of gsl_cdf_fdist_Q(x, nu1, nu2) (nu2 given) or gsl_cdf_chisq_Q(x, 1) (nu2 None)."""
        x = _f64(np.atleast_1d(x)); out = np.empty_like(x)
        if nu2 is None:
            self._chk(self.lib.gb200_cdf_tails(self.h, 1, _ptr(x), 1.0, None, _ptr(out), x.size))
        else:
            nu2 = _f64(np.broadcast_to(np.asarray(nu2, dtype=np.float64), x.shape))
            self._chk(self.lib.gb200_cdf_tails(self.h, 0, _ptr(x), float(nu1), _ptr(nu2), _ptr(out), x.size))
        return out

    # ---- fast_dgemm seam
    def dgemm(self, TransA, TransB, alpha, A, B, beta, Cm):
        A, B = _f64(A), _f64(B)
        assert Cm.dtype == np.float64 and Cm.flags.c_contiguous
        self._chk(self.lib.gb200_dgemm(self.h, TransA.encode(), TransB.encode(), alpha, _ptr(A), A.shape[0],
                                       A.shape[1], A.shape[1], _ptr(B), B.shape[0], B.shape[1], B.shape[1],
                                       beta, _ptr(Cm), Cm.shape[0], Cm.shape[1], Cm.shape[1]))
        return Cm

    # ---- -gk
    def kin_begin(self, n, k_mode=1):
        self._chk(self.lib.gb200_kin_begin(self.h, n, k_mode))
        self._kin_n = n

    def kin_add(self, Xb):
        Xb = _f64(Xb)
        self._chk(self.lib.gb200_kin_add(self.h, _ptr(Xb), Xb.shape[0], Xb.shape[1], Xb.shape[1]))

    def kin_add_geno(self, G):
        G = _f64(G)
        self._chk(self.lib.gb200_kin_add_geno(self.h, _ptr(G), G.shape[0], G.shape[1], G.shape[1]))

    def kin_add_bed(self, bed):
        bed = np.ascontiguousarray(bed, dtype=np.uint8)
        self._chk(self.lib.gb200_kin_add_bed(self.h, _ptr(bed), bed.shape[0], bed.shape[1]))

    def kin_add_bed_dev(self, dev_ptr, l, bytes_per_snp):
        self._chk(self.lib.gb200_kin_add_bed_dev(self.h, dev_ptr, l, bytes_per_snp))

    def kin_finish(self, out=None):
        """out: optional preallocated n x n float64 C-contiguous array (e.g. a view of pinned memory) that receives K."""
        n = self._kin_n
        K = np.empty((n, n)) if out is None else out
        assert K.shape == (n, n) and K.dtype == np.float64 and K.flags.c_contiguous
        ns = _sz()
        self._chk(self.lib.gb200_kin_finish(self.h, _ptr(K), n, C.byref(ns)))
        return K, ns.value

    def kin_finish_dev(self):
        p, ns = _vp(), _sz()
        self._chk(self.lib.gb200_kin_finish_dev(self.h, C.byref(p), C.byref(ns)))
        return p.value, ns.value

    # ---- eigen
    def eigh(self, G, center=True):
        G = _f64(G).copy()
        n = G.shape[0]
        U = np.empty((n, n)); ev = np.empty(n)
        tr = C.c_double(); nz = C.c_int(); nn = C.c_int()
        self._chk(self.lib.gb200_eigh(self.h, _ptr(G), n, n, int(center), _ptr(U), n, _ptr(ev), C.byref(tr),
                                      C.byref(nz), C.byref(nn)))
        return U, ev, tr.value, nz.value

    def eigh_dev(self, G_dev, n, U_dev, eval_dev, center=True):
        """Device pointers (ints): G_dev is destroyed; returns (trace_G, n_zero, n_negative)."""
        tr = C.c_double(); nz = C.c_int(); nn = C.c_int()
        self._chk(self.lib.gb200_eigh_dev(self.h, G_dev, n, int(center), U_dev, eval_dev, C.byref(tr), C.byref(nz), C.byref(nn)))
        return tr.value, nz.value, nn.value

    # ---- SNP QC statistics (PLINK)
    def qc_bed(self, bed, ni_total, idv_mask=None, W=None):
        bed = np.ascontiguousarray(bed, dtype=np.uint8)
        m = None if idv_mask is None else np.ascontiguousarray(idv_mask, dtype=np.uint8)
        out = np.zeros(bed.shape[0], dtype=SNPQC_DTYPE)
        Wc = WtWi = None; c = 0
        if W is not None:
            Wc = _f64(W); c = Wc.shape[1]; WtWi = _f64(np.linalg.inv(Wc.T @ Wc))
        self._chk(self.lib.gb200_qc_bed(self.h, _ptr(bed), _ptr(m), ni_total, bed.shape[0], bed.shape[1], _ptr(Wc), _ptr(WtWi),
                                        c, _ptr(out)))
        return out

    # ---- -lmm
    def lmm_setup(self, U, eval_, W, y):
        U, eval_, W, y = _f64(U), _f64(eval_), _f64(W), _f64(y)
        n, c = W.shape
        UtW = np.empty((n, c)); Uty = np.empty(n)
        self._chk(self.lib.gb200_lmm_setup(self.h, n, c, _ptr(U), n, _ptr(eval_), _ptr(W), c, _ptr(y), _ptr(UtW),
                                           _ptr(Uty)))
        self.n, self.n_cvt = n, c
        return UtW, Uty

    def lmm_setup_rotated(self, U, eval_, UtW, Uty):
        U, eval_, UtW, Uty = _f64(U), _f64(eval_), _f64(UtW), _f64(Uty)
        n, c = UtW.shape
        self._chk(self.lib.gb200_lmm_setup_rotated(self.h, n, c, _ptr(U), n, _ptr(eval_), _ptr(UtW), c, _ptr(Uty)))
        self.n, self.n_cvt = n, c

    def lmm_setup_rotated_dev(self, n, n_cvt, U_dev, eval_dev, UtWt_dev, Uty_dev):
        """Device pointers (ints); U_dev is borrowed."""
        self._chk(self.lib.gb200_lmm_setup_rotated_dev(self.h, n, n_cvt, U_dev, eval_dev, UtWt_dev, Uty_dev))
        self.n, self.n_cvt = n, n_cvt

    def lmm_null(self, trace_G, l_min=1e-5, l_max=1e5, n_region=10):
        nm = NullModel()
        c = self.n_cvt
        b1, s1, b2, s2 = (np.zeros(c) for _ in range(4))
        self._chk(self.lib.gb200_lmm_null(self.h, l_min, l_max, n_region, trace_G, C.byref(nm), _ptr(b1), _ptr(s1),
                                          _ptr(b2), _ptr(s2)))
        d = {k: getattr(nm, k) for k, _ in NullModel._fields_}
        d.update(beta_mle=b1, se_beta_mle=s1, beta_remle=b2, se_beta_remle=s2)
        return d

    def lmm_params(self, a_mode, l_min=1e-5, l_max=1e5, n_region=10, l_mle_null=0.0, logl_mle_H0=0.0):
        self._chk(self.lib.gb200_lmm_params(self.h, a_mode, l_min, l_max, n_region, l_mle_null, logl_mle_H0))

    def lmm_batch(self, Xb):
        """Xb: n x l (reference Xlarge layout)."""
        Xb = _f64(Xb)
        out = np.zeros(Xb.shape[1], dtype=SUMSTAT_DTYPE)
        self._chk(self.lib.gb200_lmm_batch(self.h, _ptr(Xb), Xb.shape[1], Xb.shape[1], _ptr(out)))
        return out

    def lmm_batch_geno(self, G):
        """G: l x n SNP-major, NaN = missing."""
        G = _f64(G)
        out = np.zeros(G.shape[0], dtype=SUMSTAT_DTYPE)
        self._chk(self.lib.gb200_lmm_batch_geno(self.h, _ptr(G), G.shape[0], G.shape[1], _ptr(out)))
        return out

    def lmm_batch_bed(self, bed, ni_total, idv_mask=None):
        bed = np.ascontiguousarray(bed, dtype=np.uint8)
        m = None if idv_mask is None else np.ascontiguousarray(idv_mask, dtype=np.uint8)
        out = np.zeros(bed.shape[0], dtype=SUMSTAT_DTYPE)
        self._chk(self.lib.gb200_lmm_batch_bed(self.h, _ptr(bed), _ptr(m), ni_total, bed.shape[0], bed.shape[1],
                                               _ptr(out)))
        return out

    # ---- multivariate LMM (two phenotypes, src/mvlmm.cpp)
    def mvlmm_setup(self, U, eval_, W, Y):
        U, eval_, W, Y = _f64(U), _f64(eval_), _f64(W), _f64(Y)
        n, c = W.shape
        self._chk(self.lib.gb200_mvlmm_setup(self.h, n, c, Y.shape[1], _ptr(U), n, _ptr(eval_), _ptr(W), c, _ptr(Y), Y.shape[1]))
        self.n, self.n_cvt = n, c

    def mvlmm_null(self):
        c = self.n_cvt
        a = [np.zeros(4), np.zeros(4), np.zeros(2 * c), C.c_double(), np.zeros(4), np.zeros(4), np.zeros(2 * c), C.c_double()]
        self._chk(self.lib.gb200_mvlmm_null(self.h, _ptr(a[0]), _ptr(a[1]), _ptr(a[2]), C.byref(a[3]), _ptr(a[4]), _ptr(a[5]), _ptr(a[6]),
                                            C.byref(a[7])))
        return dict(Vg_remle=a[0].reshape(2, 2), Ve_remle=a[1].reshape(2, 2), B_remle=a[2].reshape(2, c), logl_remle_H0=a[3].value,
                    Vg_mle=a[4].reshape(2, 2), Ve_mle=a[5].reshape(2, 2), B_mle=a[6].reshape(2, c), logl_mle_H0=a[7].value)

    def mvlmm_batch_geno(self, G, a_mode=1):
        """rows: beta_1, beta_2, Vbeta_11, Vbeta_12, Vbeta_22, p_wald, p_lrt, p_score"""
        G = _f64(G)
        out = np.zeros((G.shape[0], 8))
        self._chk(self.lib.gb200_mvlmm_batch_geno(self.h, _ptr(G), G.shape[0], G.shape[1], int(a_mode), _ptr(out)))
        return out

    def mvlmm_batch_bed(self, bed, ni_total, idv_mask=None, a_mode=1):
        bed = np.ascontiguousarray(bed, dtype=np.uint8)
        m = None if idv_mask is None else np.ascontiguousarray(idv_mask, dtype=np.uint8)
        out = np.zeros((bed.shape[0], 8))
        self._chk(self.lib.gb200_mvlmm_batch_bed(self.h, _ptr(bed), _ptr(m), ni_total, bed.shape[0], bed.shape[1], int(a_mode), _ptr(out)))
        return out

    # ---- -lm (linear model, src/lm.cpp)
    def lm_setup(self, W, y):
        W, y = _f64(W), _f64(y)
        self._chk(self.lib.gb200_lm_setup(self.h, W.shape[0], W.shape[1], _ptr(W), W.shape[1], _ptr(y)))

    def lm_batch_geno(self, G, a_mode):
        G = _f64(G)
        out = np.zeros(G.shape[0], dtype=SUMSTAT_DTYPE)
        self._chk(self.lib.gb200_lm_batch_geno(self.h, _ptr(G), G.shape[0], G.shape[1], int(a_mode), _ptr(out)))
        return out

    def lm_batch_bed(self, bed, ni_total, a_mode, idv_mask=None):
        bed = np.ascontiguousarray(bed, dtype=np.uint8)
        m = None if idv_mask is None else np.ascontiguousarray(idv_mask, dtype=np.uint8)
        out = np.zeros(bed.shape[0], dtype=SUMSTAT_DTYPE)
        self._chk(self.lib.gb200_lm_batch_bed(self.h, _ptr(bed), _ptr(m), ni_total, bed.shape[0], bed.shape[1], int(a_mode), _ptr(out)))
        return out

    # ---- G x E (AnalyzePlinkGXE / AnalyzeBimbamGXE)
    def lmm_gxe_setup(self, env):
        env = _f64(env)
        assert env.shape == (self.n,)
        self._chk(self.lib.gb200_lmm_gxe_setup(self.h, _ptr(env)))

    def lmm_gxe_batch_geno(self, G):
        G = _f64(G)
        out = np.zeros(G.shape[0], dtype=SUMSTAT_DTYPE)
        self._chk(self.lib.gb200_lmm_gxe_batch_geno(self.h, _ptr(G), G.shape[0], G.shape[1], _ptr(out)))
        return out

    def lmm_gxe_batch_bed(self, bed, ni_total, idv_mask=None):
        bed = np.ascontiguousarray(bed, dtype=np.uint8)
        m = None if idv_mask is None else np.ascontiguousarray(idv_mask, dtype=np.uint8)
        out = np.zeros(bed.shape[0], dtype=SUMSTAT_DTYPE)
        self._chk(self.lib.gb200_lmm_gxe_batch_bed(self.h, _ptr(bed), _ptr(m), ni_total, bed.shape[0], bed.shape[1], _ptr(out)))
        return out

    def lmm_batch_bed_dev(self, bed_dev, mask_dev, ni_total, l, bytes_per_snp, out_dev):
        self._chk(self.lib.gb200_lmm_batch_bed_dev(self.h, bed_dev, mask_dev, ni_total, l, bytes_per_snp, out_dev))

    def lmm_assoc_utx(self, UtXt):
        """UtXt: l x n (U^T x contiguous per SNP)."""
        UtXt = _f64(UtXt)
        out = np.zeros(UtXt.shape[0], dtype=SUMSTAT_DTYPE)
        self._chk(self.lib.gb200_lmm_assoc_utx(self.h, _ptr(UtXt), UtXt.shape[0], UtXt.shape[1], _ptr(out)))
        return out

    def lmm_project_bed(self, bed, ni_total, idv_mask=None):
        bed = np.ascontiguousarray(bed, dtype=np.uint8)
        m = None if idv_mask is None else np.ascontiguousarray(idv_mask, dtype=np.uint8)
        out = np.empty((bed.shape[0], self.n))
        self._chk(self.lib.gb200_lmm_project_bed(self.h, _ptr(bed), _ptr(m), ni_total, bed.shape[0], bed.shape[1],
                                                 _ptr(out)))
        return out

    def lmm_project(self, Xb):
        Xb = _f64(Xb)
        out = np.empty((Xb.shape[1], Xb.shape[0]))
        self._chk(self.lib.gb200_lmm_project(self.h, _ptr(Xb), Xb.shape[1], Xb.shape[1], _ptr(out)))
        return out
