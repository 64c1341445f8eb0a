"""CPU restatement of the int8 digit-plane projection of U^T x (gemma_b200/csrc/i8gemm_sm100.cu) -- TEST INFRASTRUCTURE ONLY.

Not part of the reference (which calls cblas_dgemm, src/lmm.cpp:1521): this restates OUR representation so that its error
bound can be checked without a GPU.  Every eigenvector (column i of U) is scaled by a power of two sigma_i = 2^e > max_j |U_ji|
and rounded to B = 6 + 8 (T - 1) fractional bits, Q_ji = rint(U_ji 2^(B - e)); Q is written in balanced base-256 digits
(T - 1 digits in [-128, 127] and a top digit with |d| <= 65), one int8 plane per digit.  With integer genotypes the T plane
products are exact integers (the tensor cores accumulate in int32), and the recombination sum_t 256^(T-1-t) P_t is exact in
int64 / FP64 up to 2^53; the only error is the rounding of U, at most n 2^-(B+1) sigma_i max|x| and about sqrt(n) 2^-B sigma_i |x|
in practice."""
import math

import numpy as np


def default_planes(n):
    """i8_default_planes: smallest T with sqrt(n) 2^-(6 + 8 (T - 1)) <= 2^-30, clamped to 4..8."""
    need = 30.0 + 0.5 * math.log2(n if n > 1 else 2)
    return min(8, max(4, int(math.ceil((need - 6.0) / 8.0)) + 1))


def slice_planes(U, T):
    """Returns planes (T, n_eig, n_ind) int8 with plane 0 the most significant digit, and scale (n_eig,) = 2^(e - B)."""
    n = U.shape[0]
    B = 6 + 8 * (T - 1)
    colmax = np.abs(U).max(axis=0)
    e = np.where(colmax > 0, np.frexp(colmax)[1], 0).astype(np.int64)        # colmax = f 2^e, f in [0.5, 1)
    Q = np.rint(np.ldexp(U, (B - e)[None, :].astype(np.int32))).astype(np.int64)   # Q[j, i]
    planes = np.zeros((T, n, n), dtype=np.int8)
    for t in range(T - 1, 0, -1):
        d = ((Q + 128) & 255) - 128
        Q = (Q - d) >> 8
        planes[t] = d.T.astype(np.int8)
    assert np.abs(Q).max() <= 65
    planes[0] = Q.T.astype(np.int8)
    return planes, np.ldexp(1.0, (e - B).astype(np.int32))


def project(planes, scale, X):
    """X: (n_ind, l) integer genotypes.  Exact integer plane products, recombined most significant digit first."""
    T = planes.shape[0]
    acc = np.zeros((planes.shape[1], X.shape[1]), dtype=np.int64)
    Xi = X.astype(np.int64)
    for t in range(T):
        acc = acc * 256 + planes[t].astype(np.int64) @ Xi
    return acc.astype(np.float64) * scale[:, None]
