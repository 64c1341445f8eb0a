"""Deterministic synthetic genotypes for tests and bench.

Counter-based generator keyed on (seed, snp j, individual i): allele frequency
f_j ~ U(0.05, 0.5), genotype g_ij ~ Binomial(2, f_j), optional missingness.  The same
integer hash is evaluated with numpy on the host and with torch int64 ops on the device,
so the CPU baseline and the GPU see bit-identical PLINK .bed bytes without storing them.

PLINK 2-bit codes follow the reference's decoder (src/gemma_io.cpp:1665-1682): per sample
(low bit, high bit): (0,0) -> 2 minor alleles, (0,1) -> 1, (1,1) -> 0, (1,0) -> missing;
four samples per byte, sample 0 in the two lowest bits, SNP-major rows of ceil(n/4) bytes.
"""
import numpy as np

SEED = 20260923
_M1 = 0xBF58476D1CE4E5B9
_M2 = 0x94D049BB133111EB
_G1 = 0x9E3779B97F4A7C15
_G2 = 0xC2B2AE3D27D4EB4F
_MASK = (1 << 64) - 1


def _mix_np(z):
    z = z.astype(np.uint64)
    z = (z ^ (z >> np.uint64(30))) * np.uint64(_M1)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(_M2)
    return z ^ (z >> np.uint64(31))


def _u01_np(h):
    return (h >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def snp_freq(seed, j):
    """f_j in [0.05, 0.5) for SNP indices j (numpy int array)."""
    with np.errstate(over="ignore"):
        h = _mix_np(np.uint64(seed) * np.uint64(_G2) + np.asarray(j, dtype=np.uint64) * np.uint64(_G1) + np.uint64(12345))
    return 0.05 + 0.45 * _u01_np(h)


def genotypes(n, l, seed=SEED, snp_offset=0, miss_rate=0.0):
    """Returns int8 array [l, n] with values 0/1/2 and -9 for missing."""
    j = (np.arange(l, dtype=np.uint64) + np.uint64(snp_offset))[:, None]
    i = np.arange(n, dtype=np.uint64)[None, :]
    with np.errstate(over="ignore"):
        base = np.uint64(seed) + j * np.uint64(_G1) + i * np.uint64(_G2)
        h1 = _mix_np(base)
        h2 = _mix_np(base + np.uint64(0x632BE59BD9B4E019))
        f = snp_freq(seed, j)
        g = (_u01_np(h1) < f).astype(np.int8) + (_u01_np(h2) < f).astype(np.int8)
        if miss_rate > 0:
            h3 = _mix_np(base + np.uint64(0x1F83D9ABFB41BD6B))
            g = np.where(_u01_np(h3) < miss_rate, np.int8(-9), g)
    return g


_CODE = np.array([3, 2, 0], dtype=np.uint8)   # g=0 -> 0b11, g=1 -> 0b10, g=2 -> 0b00


def pack_bed(g):
    """int8 [l, n] (0/1/2/-9) -> uint8 [l, ceil(n/4)] PLINK SNP-major rows."""
    l, n = g.shape
    code = np.where(g < 0, np.uint8(1), _CODE[np.clip(g, 0, 2)])
    nb = (n + 3) // 4
    pad = np.zeros((l, nb * 4), dtype=np.uint8)
    pad[:, :n] = code
    pad = pad.reshape(l, nb, 4)
    return (pad[:, :, 0] | (pad[:, :, 1] << 2) | (pad[:, :, 2] << 4) | (pad[:, :, 3] << 6)).astype(np.uint8)


def make_bed(n, l, seed=SEED, snp_offset=0, miss_rate=0.0):
    """(bed uint8 [l, ceil(n/4)], G float64 [l, n] with -9 for missing)."""
    g = genotypes(n, l, seed, snp_offset, miss_rate)
    return pack_bed(g), g.astype(np.float64)


# ---- device twin (torch int64 arithmetic wraps like uint64) --------------------------------
def _s64(v):
    v &= _MASK
    return v - (1 << 64) if v >= (1 << 63) else v


def _lsr(x, s):
    return (x >> s) & ((1 << (64 - s)) - 1)


def _mix_t(z):
    z = (z ^ _lsr(z, 30)) * _s64(_M1)
    z = (z ^ _lsr(z, 27)) * _s64(_M2)
    return z ^ _lsr(z, 31)


def _u01_t(h):
    import torch
    return _lsr(h, 11).to(torch.float64) * (1.0 / 9007199254740992.0)


def make_bed_torch(n, l, device, seed=SEED, snp_offset=0, miss_rate=0.0, chunk=2048):
    """uint8 CUDA tensor [l, ceil(n/4)], bit-identical to make_bed(...)[0]."""
    import torch
    nb = (n + 3) // 4
    out = torch.empty((l, nb), dtype=torch.uint8, device=device)
    i = torch.arange(nb * 4, dtype=torch.int64, device=device)[None, :]
    valid = (i < n)
    for s0 in range(0, l, chunk):
        lc = min(chunk, l - s0)
        j = (torch.arange(lc, dtype=torch.int64, device=device) + (snp_offset + s0))[:, None]
        base = _s64(seed) + j * _s64(_G1) + i * _s64(_G2)
        hf = _mix_t(_s64(seed * _G2) + j * _s64(_G1) + 12345)
        f = 0.05 + 0.45 * _u01_t(hf)
        g = (_u01_t(_mix_t(base)) < f).to(torch.int64) + (_u01_t(_mix_t(base + _s64(0x632BE59BD9B4E019))) < f).to(torch.int64)
        code = torch.where(g == 0, 3, torch.where(g == 1, 2, 0))
        if miss_rate > 0:
            m = _u01_t(_mix_t(base + _s64(0x1F83D9ABFB41BD6B))) < miss_rate
            code = torch.where(m, 1, code)
        code = torch.where(valid, code, 0).view(lc, nb, 4)
        out[s0:s0 + lc] = (code[:, :, 0] | (code[:, :, 1] << 2) | (code[:, :, 2] << 4) | (code[:, :, 3] << 6)).to(torch.uint8)
    return out


def phenotype(n, G_causal, seed=SEED, h2=0.5):
    """y = sum_j g_ij b_j + e with b_j ~ N(0, h2/(m 2f(1-f))) over the m causal rows of G_causal."""
    rng = np.random.default_rng(seed)
    m = G_causal.shape[0]
    f = np.clip(G_causal.mean(axis=1) / 2.0, 0.01, 0.99)
    b = rng.standard_normal(m) * np.sqrt(h2 / (m * 2 * f * (1 - f)))
    e = rng.standard_normal(n) * np.sqrt(1 - h2)
    return G_causal.T @ b + e


def spectrum_like_kinship(n, seed=SEED):
    """A cheap orthogonal U (product of Householder reflections + permutation) and a kinship-like
    eigenvalue spectrum (one zero eigenvalue, the rest spread over [1e-3, ~10]) for configs where
    computing K and its eigendecomposition is not the thing being measured."""
    rng = np.random.default_rng(seed + 1)
    ev = np.sort(np.concatenate([[0.0], rng.gamma(0.6, 1.6, n - 1) + 1e-3]))
    return ev


def polygenic_rotated(ev, seed=SEED, h2=0.5):
    """U^T y of a phenotype drawn from the LMM itself, y ~ N(0, h2 K / mean(eval) + (1 - h2) I): in the eigenbasis the entries are
    independent N(0, h2 eval_i / mean(eval) + 1 - h2).  pve ~ h2, so the REML / ML roots are interior and every SNP runs the whole
    grid + Brent + Newton search (a phenotype unrelated to K puts lambda on the l_min boundary and skips it)."""
    ev = np.asarray(ev, dtype=np.float64)
    rng = np.random.default_rng(seed + 77)
    return np.sqrt(h2 * ev / ev.mean() + (1.0 - h2)) * rng.standard_normal(ev.shape[0])
