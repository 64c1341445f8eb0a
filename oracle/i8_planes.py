"""CPU restatement of the int8 digit-plane projection of U^T x (gemma_b200/csrc/i8gemm_sm100.cu) -- TEST INFRASTRUCTURE ONLY.

Not part of the reference (which calls cblas_dgemm, src/lmm.cpp:1521): this restates OUR representation so that its error
bound can be checked without a GPU.  Every eigenvector (column i of U) is divided by s_i = max_j |U_ji| / (127.4 * 256^(T-1)) and
rounded, Q_ji = rint(U_ji / s_i); Q is written in balanced base-256 digits (T - 1 digits in [-128, 127] and a top digit with
|d| <= 127: the whole int8 range), one int8 plane per digit.  With integer genotypes the T plane products are exact integers
(the tensor cores accumulate in int32), and the recombination sum_t 256^(T-1-t) P_t is exact in int64 / FP64 up to 2^53; the
only error is the rounding of U, at most n s_i max|x| / 2 and about sqrt(n / 12) s_i rms(x) in practice."""
import math

import numpy as np

TOP = 127.4


def default_planes(n):
    """i8_default_planes: U-independent worst case (column maximum 1): smallest T with
    sqrt(n) / (sqrt(12) 127.4 256^(T-1)) <= 2^-30, clamped to 4..8."""
    need = math.sqrt(n if n > 1 else 2) / (math.sqrt(12.0) * TOP) * 2.0 ** 30
    return min(8, max(4, 1 + int(math.ceil(math.log2(need) / 8.0))))


def choose_planes(colmax_max, n, linear_sums_exact=False):
    """i8_choose_planes: the same bound with the measured largest column maximum, target 2^-29; at least 5 planes below n = 8192.
    With the exact linear x-sums of the side GEMM (LmmConst::xsum) the projected values only feed sums quadratic in x: target 2^-21,
    at least 3 planes."""
    if not (colmax_max > 0 and math.isfinite(colmax_max)):
        return 4
    need = colmax_max * math.sqrt(n if n > 1 else 2) / (math.sqrt(12.0) * TOP) * 2.0 ** (21 if linear_sums_exact else 29)
    T = min(8, max(3 if linear_sums_exact else 4, 1 + int(math.ceil(math.log2(need) / 8.0))))
    return max(T, 5) if n < 8192 else T


def slice_planes(U, T):
    """Returns planes (T, n_eig, n_ind) int8 with plane 0 the most significant digit, and scale (n_eig,) = s_i."""
    n = U.shape[0]
    colmax = np.abs(U).max(axis=0)
    top = TOP * 256.0 ** (T - 1)
    ok = colmax > 0
    mult = np.where(ok, top / np.where(ok, colmax, 1.0), 0.0)
    scale = np.where(ok, colmax / top, 0.0)
    Q = np.rint(U * mult[None, :]).astype(np.int64)                          # Q[j, i]
    planes = np.zeros((T, n, n), dtype=np.int8)
    for t in range(T - 1, 0, -1):
        d = ((Q + 128) & 255) - 128
        Q = (Q - d) >> 8
        planes[t] = d.T.astype(np.int8)
    assert np.abs(Q).max() <= 127
    planes[0] = Q.T.astype(np.int8)
    return planes, scale


def project(planes, scale, X):
    """X: (n_ind, l) integer genotypes.  Exact integer plane products, recombined most significant digit first."""
    T = planes.shape[0]
    acc = np.zeros((planes.shape[1], X.shape[1]), dtype=np.int64)
    Xi = X.astype(np.int64)
    for t in range(T):
        acc = acc * 256 + planes[t].astype(np.int64) @ Xi
    return acc.astype(np.float64) * scale[:, None]
