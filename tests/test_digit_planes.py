"""The int8 digit-plane representation behind the tensor-core projection (oracle/i8_planes.py restates
gemma_b200/csrc/i8gemm_sm100.cu: col_scale_kernel, slice_kernel, the plane recombination): reconstruction, digit ranges,
the error bound that fixes the default plane count.  CPU only."""
import numpy as np
import pytest

from oracle import i8_planes as P


def _orthogonal(n, seed):
    q, _ = np.linalg.qr(np.random.default_rng(seed).standard_normal((n, n)))
    return q


def test_plane_count_rules():
    # U-independent worst case (an eigenvector concentrated on one individual): sqrt(n) / (sqrt(12) 127.4 256^(T-1)) <= 2^-30
    assert [P.default_planes(n) for n in (1024, 10000, 50000, 65536, 250000)] == [5] * 5 and P.default_planes(2) == 4
    assert P.default_planes(3 * 10 ** 6) == 5 and P.default_planes(4 * 10 ** 6) == 6
    for n in (1000, 50000, 65536, 200000, 10 ** 6, 10 ** 7):
        T = P.default_planes(n)
        bound = lambda t: np.sqrt(n) / (np.sqrt(12.0) * 127.4 * 256.0 ** (t - 1))
        assert bound(T) <= 2.0 ** -30 < bound(T - 1)
    # measured rule (target 2^-29): delocalised eigenvectors (column maximum ~ sqrt(4 ln n / n)) need 4 planes, concentrated ones more;
    # cohorts below n = 8192 never go below 5
    for n in (10000, 50000):
        assert P.choose_planes(np.sqrt(4 * np.log(n) / n), n) == 4
        assert P.choose_planes(13.5 / np.sqrt(n), n) == 4 and P.choose_planes(14.2 / np.sqrt(n), n) == 5
        assert P.choose_planes(1.0, n) == 5
    # with the exact linear x-sums (side GEMM) the projection only feeds sums quadratic in x: 3 planes for delocalised eigenvectors
    for n in (10000, 50000):
        assert P.choose_planes(np.sqrt(4 * np.log(n) / n), n, linear_sums_exact=True) == 3
        assert P.choose_planes(1.0, n, linear_sums_exact=True) == 4
    assert P.choose_planes(0.05, 4000, linear_sums_exact=True) == 5
    q = _orthogonal(400, 1)
    assert P.choose_planes(np.abs(q).max(), 400) == 5 and P.choose_planes(np.abs(q).max() * np.sqrt(400 / 8192), 8192) == 4
    assert P.choose_planes(0.0, 400) == 4


@pytest.mark.parametrize("n,T", [(96, 5), (257, 4), (257, 6), (300, 7)])
def test_planes_reconstruct_the_rounded_matrix(n, T):
    U = _orthogonal(n, n)
    U[:, 3] = 0.0; U[5, 7] = 0.5; U[:, 7] = np.clip(U[:, 7], -0.5, 0.5)        # an empty column; a column whose maximum is a power of two
    planes, scale = P.slice_planes(U, T)
    assert planes.dtype == np.int8 and np.abs(planes[0].astype(int)).max() <= 127
    assert np.abs(planes[0].astype(int)).max(axis=1)[[0, 1, 2, 7]].min() >= 126       # the top digit uses the whole int8 range
    assert not planes[:, 3, :].any() and scale[3] == 0.0
    Q = np.zeros((n, n), dtype=np.int64)
    for t in range(T):
        Q = Q * 256 + planes[t]
    back = Q.T * scale[None, :]
    colmax = np.abs(U).max(axis=0)
    half_unit = colmax / (127.4 * 256.0 ** (T - 1)) / 2.0
    assert np.all(np.abs(back - U) <= half_unit[None, :] * (1 + 1e-9) + 4e-16)   # half a unit of the last kept digit, per column


@pytest.mark.parametrize("n,miss", [(384, False), (500, True)])
def test_projection_error_bound(n, miss):
    rng = np.random.default_rng(n)
    U = _orthogonal(n, 7)
    X = rng.binomial(2, rng.uniform(0.05, 0.5, 40)[None, :], size=(n, 40)).astype(np.int64)
    if miss:
        X[rng.random(X.shape) < 0.02] = 0                    # holes enter the int8 operand as 0 (the mean term is added in FP64 afterwards)
    exact = U.T @ X.astype(np.float64)
    cm = np.abs(U).max()
    for T in (3, 4, 5):
        unit = cm / (127.4 * 256.0 ** (T - 1))
        planes, scale = P.slice_planes(U, T)
        err = np.abs(P.project(planes, scale, X) - exact).max()
        worst = n * unit / 2 * 2                              # every entry off by half a unit, |x| <= 2
        typical = np.sqrt(n / 12.0) * unit * 1.0             # independent rounding noise, rms(x) ~ 1
        assert err <= worst and err <= 6 * typical + 64 * np.finfo(float).eps * np.abs(exact).max(), (T, err, typical)
    # at the chosen plane count the projection is indistinguishable from the FP64 product at the 1e-6 parity bar:
    # relative to the typical size of a projected value (rms(x) ~ 1) the noise stays below the 2^-29 design target
    T = P.choose_planes(cm, n)
    planes, scale = P.slice_planes(U, T)
    assert T == 5 and np.abs(P.project(planes, scale, X) - exact).max() < 6 * 2.0 ** -29
    planes, scale = P.slice_planes(U, 4)                                  # what a large cohort with such delocalised eigenvectors uses
    assert np.abs(P.project(planes, scale, X) - exact).max() < 6 * 2.0 ** -29


def test_kinship_missing_genotype_identity():
    """The algebra of the int8 kinship path with holes (i8gemm_sm100.cu, kin_miss_fix_kernel): with z = 0 at missing entries,
    q the missing indicator and m_s the SNP mean over the observed entries, the reference's mean-imputed centred product
    (src/gemma_io.cpp:1688-1706) equals Z Z^T - a 1^T - 1 a^T + beta + (Y + Y^T) - b 1^T - 1 b^T with
    a_i = sum_s m_s z_si, beta = sum_s m_s^2, b_i = sum_s m_s^2 q_si, Y_ji = sum_s q_sj (m_s z_si + m_s^2 q_si / 2):
    the dense part is exact integer arithmetic, everything else is O(#missing x n)."""
    rng = np.random.default_rng(3)
    n, l = 60, 400
    G = rng.binomial(2, rng.uniform(0.05, 0.5, l)[:, None], size=(l, n)).astype(float)
    q = rng.random((l, n)) < 0.03
    z = np.where(q, 0.0, G)
    m = z.sum(1) / (n - q.sum(1))
    Xc = np.where(q, 0.0, G - m[:, None])                     # what the reference accumulates: imputed entries centre to 0
    ref = Xc.T @ Xc
    a = z.T @ m
    beta = float(m @ m)
    b = q.T.astype(float) @ (m * m)
    Y = q.T.astype(float) @ (m[:, None] * z + 0.5 * (m * m)[:, None] * q)      # row j: sum over the SNPs where j is missing
    got = z.T @ z - a[:, None] - a[None, :] + beta + Y + Y.T - b[:, None] - b[None, :]
    assert np.allclose(got, ref, rtol=0, atol=1e-9)
    ZZ = z.T.astype(np.int64) @ z.astype(np.int64)            # the tensor-pipe part is integer-exact
    assert np.array_equal(ZZ, (z.T @ z).astype(np.int64))
