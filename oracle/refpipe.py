"""numpy restatement of the reference's host-side flow around the hot path
(readers, SNP QC, -gk, -eigen, null model, -lmm) driving the C oracle.

TEST INFRASTRUCTURE ONLY (see oracle/gemma_oracle.c header).  Each function
cites the reference file:line it follows.  Eigendecomposition uses LAPACK
dsyevr through scipy (driver='evr'), the routine the reference calls at
src/lapack.cpp:205,220.
"""
import gzip
import io

import numpy as np
import scipy.linalg

from . import oracle as O


def _open(path):
    if str(path).endswith(".gz"):
        return io.TextIOWrapper(gzip.open(path, "rb"))
    return open(path, "r")


def _tok(line):
    # strtok(line, " ,\t")  (src/gemma_io.cpp:706 etc.)
    return line.replace(",", " ").replace("\t", " ").split()


def read_pheno(path, p_column=(1,)):
    """src/gemma_io.cpp:386-444 ReadFile_pheno. Returns (pheno[n,d], indicator[n,d])."""
    ph, ind = [], []
    with _open(path) as f:
        for line in f:
            t = _tok(line.rstrip("\r\n"))
            if not t and line.strip() == "":
                # the reference still pushes a row for an empty line only if strtok fails -> enforce; skip
                continue
            row, irow = [], []
            for c in p_column:
                s = t[c - 1]
                if s == "NA":
                    row.append(-9.0); irow.append(0)
                else:
                    row.append(float(s)); irow.append(1)
            ph.append(row); ind.append(irow)
    return np.array(ph, dtype=np.float64), np.array(ind, dtype=np.int32)


def read_cvt(path):
    """src/gemma_io.cpp:446-511 ReadFile_cvt. Returns (cvt rows list, indicator_cvt)."""
    rows, ind = [], []
    with _open(path) as f:
        for line in f:
            t = _tok(line.rstrip("\r\n"))
            na = any(s == "NA" for s in t)
            rows.append([(-9.0 if s == "NA" else float(s)) for s in t])
            ind.append(0 if na else 1)
    return rows, np.array(ind, dtype=np.int32)


def read_anno(path):
    """src/gemma_io.cpp:280-341 ReadFile_anno: rs, bp, chr[, cM]."""
    m = {}
    with _open(path) as f:
        for line in f:
            t = _tok(line.rstrip("\r\n"))
            if not t:
                continue
            rs = t[0]
            bp = -9 if t[1] == "NA" else int(float(t[1]))
            ch = "-9" if (len(t) < 3 or t[2] == "NA") else t[2]
            cm = -9.0 if (len(t) < 4 or t[3] == "NA") else float(t[3])
            m[rs] = (ch, bp, cm)
    return m


def process_cvt_phen(ind_pheno, cvt_rows=None, ind_cvt=None):
    """src/param.cpp:1993-2098 ProcessCvtPhen + :1937-1990 CheckCvt.
    Returns (indicator_idv, W_full[n_total, c]) where W rows of excluded individuals are unused."""
    indicator_idv = np.all(ind_pheno == 1, axis=1).astype(np.int32)
    if ind_cvt is not None and len(ind_cvt):
        indicator_idv = indicator_idv * ind_cvt
    n_total = len(indicator_idv)
    if cvt_rows is None or ind_cvt is None or len(ind_cvt) == 0:
        return indicator_idv, np.ones((n_total, 1))
    sel = [i for i in range(n_total) if indicator_idv[i] == 1 and ind_cvt[i] == 1]
    W = np.array([cvt_rows[i] for i in sel], dtype=np.float64)
    const = [j for j in range(W.shape[1]) if W[:, j].min() == W[:, j].max()]
    if len(const) == W.shape[1]:
        return indicator_idv, np.ones((n_total, 1))       # covariates dropped, intercept only
    Wfull = np.zeros((n_total, W.shape[1] + (0 if const else 1)))
    for k, i in enumerate(sel):
        Wfull[i, :W.shape[1]] = W[k]
        if not const:
            Wfull[i, -1] = 1.0                              # intercept appended as LAST column
    return indicator_idv, Wfull


class Bimbam:
    """Parsed BIMBAM mean-genotype file: rs ids, alleles, G[p_total, n_total] with NaN = NA."""

    def __init__(self, path):
        rs, a1, a0, rows = [], [], [], []
        with _open(path) as f:
            for line in f:
                t = _tok(line.rstrip("\r\n"))
                if not t:
                    continue
                rs.append(t[0]); a1.append(t[1]); a0.append(t[2])
                rows.append(np.array([np.nan if s == "NA" else float(s) for s in t[3:]]))
        self.rs, self.a1, self.a0 = rs, a1, a0
        self.G = np.vstack(rows)


def qc_bimbam(bb, indicator_idv, W=None, miss_level=0.05, maf_level=0.01, r2_level=0.9999,
              snps=None):
    """src/gemma_io.cpp:639-873 ReadFile_geno QC pass. Returns (indicator_snp, n_miss, maf)."""
    keep = indicator_idv == 1
    G = bb.G[:, :len(keep)][:, keep]          # only the first len(indicator) columns are parsed (:702)
    ni_test = int(keep.sum())
    p = G.shape[0]
    ind = np.zeros(p, dtype=np.int32)
    n_miss_a = np.zeros(p, dtype=np.int64)
    maf_a = np.zeros(p)
    Wt = None
    if W is not None and W.shape[1] != 1:
        Wt = W[keep]
        WtWi = np.linalg.inv(Wt.T @ Wt)
    for t in range(p):
        if snps is not None and bb.rs[t] not in snps:
            continue
        g = G[t]
        miss = np.isnan(g)
        n_miss = int(miss.sum())
        gv = g[~miss]
        maf = float(np.cumsum(gv)[-1]) if gv.size else 0.0   # strictly sequential left fold, as the reference accumulates
        # (Python >= 3.12 sum() is Neumaier-compensated and differs from the plain loop by an ulp)
        maf /= 2.0 * (ni_test - n_miss)
        n_miss_a[t] = n_miss; maf_a[t] = maf
        if n_miss / ni_test > miss_level:
            continue
        if (maf < maf_level or maf > 1.0 - maf_level) and maf_level != -1:
            continue
        if gv.size == 0 or np.all(gv == gv[0]):
            continue
        if Wt is not None:
            x = g.copy(); x[miss] = maf * 2.0
            Wtx = Wt.T @ x
            v_w = Wtx @ (WtWi @ Wtx); v_x = x @ x
            if v_w / v_x > r2_level:
                continue
        ind[t] = 1
    return ind, n_miss_a, maf_a


def trim_individuals(indicator, ni_max):
    """src/param.cpp:74-90 (-nind, a test-speed switch): the vector is resized to the NUMBER of set flags
    seen before the scan stops (i.e. min(#set, ni_max)), not to the position of the ni_max-th set flag."""
    if not ni_max:
        return indicator
    count = 0
    for v in indicator:
        if v:
            count += 1
        if count >= ni_max:
            break
    return indicator[:count] if count != len(indicator) else indicator


def loco_sets(anno, loco):
    """src/param.cpp:52-66 LOCO_set_Snps: (ksnps, gwasnps) = annotated SNPs off / on chromosome `loco`."""
    ks = {rs for rs, v in anno.items() if v[0] != loco}
    gw = {rs for rs, v in anno.items() if v[0] == loco}
    return ks, gw


def kinship_bimbam(bb, indicator_snp, k_mode=1, batch=20000, ni_total=None, ksnps=None):
    """src/gemma_io.cpp:1418-1597 BimbamKin: all ni_total individuals (the first ni_total columns under -nind),
    SNPs with indicator 1 that are also in ksnps when that set is non-empty (:1478-1480)."""
    sel = np.nonzero(indicator_snp)[0]
    if ksnps:
        sel = np.array([t for t in sel if bb.rs[t] in ksnps], dtype=np.int64)
    n = bb.G.shape[1] if ni_total is None else ni_total
    K = np.zeros((n, n))
    for s in range(0, len(sel), batch):
        Xc = O.kin_transform(bb.G[sel[s:s + batch], :n], k_mode)
        K += Xc @ Xc.T            # the dgemm of :1554 (BLAS instead of the oracle's O(n^2 l) loop)
    K *= 1.0 / len(sel)
    return K


def text_roundtrip(M):
    """WriteMatrix precision(10) general format (src/param.cpp:1899-1906) then atof."""
    flat = np.array([float("%.10g" % v) for v in np.asarray(M).ravel()])
    return flat.reshape(np.shape(M))


def eigen_decomp_zeroed(G):
    """src/lapack.cpp:260-291 (dsyevr 'V','A','L', abstol 1e-7; eval<1e-10 -> 0). Returns U, eval, trace_G."""
    ev, U = scipy.linalg.eigh(G, lower=True, driver="evr")
    ev, tr = O.zero_small_eval(ev)
    return U, ev, tr


def lmm_prepare(K_total, indicator_idv, y_total, W_total):
    """LMM branch of BatchRun up to the null model (src/gemma.cpp:2556-2760)."""
    keep = indicator_idv == 1
    G = np.ascontiguousarray(K_total[np.ix_(keep, keep)])
    G = O.center_matrix(G)
    U, ev, trace_G = eigen_decomp_zeroed(G)
    W = np.ascontiguousarray(W_total[keep]); y = np.ascontiguousarray(y_total[keep])
    UtW = U.T @ W; Uty = U.T @ y
    l_mle, logl_mle = O.calc_lambda_null("L", ev, UtW, Uty)
    l_remle, logl_remle = O.calc_lambda_null("R", ev, UtW, Uty)
    pve, pve_se = O.calc_pve(ev, UtW, Uty, l_remle, trace_G)
    return dict(U=U, eval=ev, trace_G=trace_G, UtW=UtW, Uty=Uty, W=W, y=y, l_mle_null=l_mle,
                logl_mle_H0=logl_mle, l_remle_null=l_remle, logl_remle_H0=logl_remle, pve=pve,
                pve_se=pve_se)


def lmm_genotypes_bimbam(bb, indicator_snp, indicator_idv, sel=None):
    """src/lmm.cpp:1590-1618: analysed individuals only, mean-imputed; returns X n x l."""
    idx = np.nonzero(indicator_snp)[0] if sel is None else sel
    keep = np.nonzero(indicator_idv == 1)[0]
    return O.lmm_impute(bb.G[np.ix_(idx, keep)])


def lmm_analyze(prep, X, a_mode, **kw):
    UtX = prep["U"].T @ X          # src/lmm.cpp:1521
    return O.lmm_analyze_utx(prep["eval"], prep["UtW"], prep["Uty"], UtX, a_mode,
                             l_mle_null=prep["l_mle_null"], logl_mle_H0=prep["logl_mle_H0"], **kw)


def lm_analyze(W, y, G, a_mode):
    """-lm restatement (LM::AnalyzeBimbam / AnalyzePlink + CalcvPv + LmCalcP, src/lm.cpp:224-288, 382-640).  W: n x c, y: n, G: l x n
    genotypes of the analysed individuals (NaN = missing, mean-imputed per SNP).  Returns SUMSTAT rows {beta, se, 0, 0, p_wald, p_lrt,
    p_score, -0}; se is the score-test one in mode 3 (:281-285)."""
    n, c = W.shape
    test_mode = a_mode - 50 if a_mode > 50 else a_mode
    WtWi = np.linalg.inv(W.T @ W)
    Wty = W.T @ y
    yPwy = float(y @ y - (WtWi @ Wty) @ Wty)
    df = float(n) - float(c) - 1.0
    out = np.zeros(G.shape[0], dtype=O.SUMSTAT_DTYPE)
    for t in range(G.shape[0]):
        x = G[t].copy()
        miss = np.isnan(x)
        x[miss] = float(np.cumsum(x[~miss])[-1]) / float((~miss).sum())
        Wtx = W.T @ x
        WtWiWtx = WtWi @ Wtx
        xPwx = float(x @ x - WtWiWtx @ Wtx)
        xPwy = float(x @ y - WtWiWtx @ Wty)
        yPxy = yPwy - xPwy * xPwy / xPwx
        beta = xPwy / xPwx
        se_wald = np.sqrt(yPxy / (df * xPwx)); se_score = np.sqrt(yPwy / (float(n) * xPwx))
        out["beta"][t] = beta; out["se"][t] = se_score if test_mode == 3 else se_wald
        out["p_wald"][t] = O.fdist_Q(beta * beta / (se_wald * se_wald), 1.0, df)
        out["p_score"][t] = O.fdist_Q(beta * beta / (se_score * se_score), 1.0, df)
        out["p_lrt"][t] = O.chisq1_Q(float(n) * (np.log(yPwy) - np.log(yPxy)))
        out["logl_H1"][t] = -0.0
    return out


def lmm_gxe(prep, G, env, a_mode, l_min=1e-5, l_max=1e5, n_region=10, l_mle_null=0.0):
    """G x E restatement (LMM::AnalyzePlinkGXE / AnalyzeBimbamGXE, src/lmm.cpp:2283-2608) composed from the oracle's univariate
    pieces.  G: l x n genotypes of the analysed individuals (NaN = missing); env: n.  Per SNP: mean-impute, flip to 2 - x when
    the mean exceeds 1 (:2537-2540), covariates [W, env, x], tested variable x * env; -lmm 2/4 compare with the per-SNP null that
    contains x (param0 = calc_null with n_cvt + 2 covariates, :2560-2563); beta changes sign for flipped SNPs (:2587-2589).
    logl_H0 stays 0 in the other modes, as the reference's local does."""
    U, ev, Uty = prep["U"], prep["eval"], prep["Uty"]
    l = G.shape[0]
    out = np.zeros(l, dtype=O.SUMSTAT_DTYPE)
    Ute = U.T @ env
    for t in range(l):
        x = G[t].copy()
        miss = np.isnan(x)
        x_mean = float(np.cumsum(x[~miss])[-1]) / float((~miss).sum())
        x[miss] = x_mean
        flip = x_mean > 1
        if flip:
            x = 2 - x
        UtW_e = np.column_stack([prep["UtW"], Ute, U.T @ x])
        Utx = U.T @ (x * env)
        logl_H0 = 0.0
        if a_mode in (2, 4):
            _, logl_H0 = O.calc_lambda_null("L", ev, UtW_e, Uty, l_min, l_max, n_region)
        r = O.lmm_analyze_utx(ev, UtW_e, Uty, Utx[:, None], a_mode, l_min, l_max, n_region, l_mle_null, logl_H0)[0]
        out[t] = r
        if flip:
            out["beta"][t] = -out["beta"][t]
    return out


# ---- PLINK (test infrastructure; src/gemma_io.cpp:514-636, 876-1064, 1599-1738) -------------------------
class Plink:
    """.bim/.fam/.bed trio: rs ids, alleles, phenotypes (column 6+, -9/NA missing) and G[p, n] with NaN = missing."""

    def __init__(self, prefix, p_column=(1,)):
        self.chr, self.rs, self.cM, self.bp, self.a1, self.a0 = [], [], [], [], [], []
        for line in open(prefix + ".bim"):
            t = line.split()
            if not t:
                continue
            self.chr.append(t[0]); self.rs.append(t[1]); self.cM.append(float(t[2])); self.bp.append(int(t[3]))
            self.a1.append(t[4]); self.a0.append(t[5])
        ph, ind = [], []
        for line in open(prefix + ".fam"):
            t = _tok(line.rstrip("\r\n"))
            if not t:
                continue
            row, irow = [], []
            for c in p_column:
                s = t[5 + c - 1]
                if s == "NA" or float(s) == -9:
                    row.append(-9.0); irow.append(0)
                else:
                    row.append(float(s)); irow.append(1)
            ph.append(row); ind.append(irow)
        self.pheno = np.array(ph); self.ind_pheno = np.array(ind, dtype=np.int32)
        n = len(ph)
        raw = np.fromfile(prefix + ".bed", dtype=np.uint8)[3:]
        nb = (n + 3) // 4
        self.bed = raw.reshape(-1, nb)
        code = np.stack([(self.bed >> (2 * q)) & 3 for q in range(4)], axis=2).reshape(len(self.bed), nb * 4)[:, :n]
        lut = np.array([2.0, np.nan, 1.0, 0.0])          # bits (hi,lo): 00->2, 01 (lo=1,hi=0)->missing, 10->1, 11->0
        self.G = lut[code]


def qc_plink(pl, indicator_idv, W=None, miss_level=0.05, maf_level=0.01, r2_level=0.9999):
    """src/gemma_io.cpp:876-1064 ReadFile_bed QC pass."""
    keep = indicator_idv == 1
    G = pl.G[:, keep]
    ni_test = int(keep.sum())
    miss = np.isnan(G)
    n_miss = miss.sum(axis=1)
    s = np.nansum(G, axis=1)
    maf = s / (2.0 * (ni_test - n_miss))
    n0 = (G == 0).sum(axis=1); n1 = (G == 1).sum(axis=1); n2 = (G == 2).sum(axis=1)
    ind = np.ones(len(G), dtype=np.int32)
    ind[n_miss / ni_test > miss_level] = 0
    if maf_level != -1:
        ind[(maf < maf_level) | (maf > 1.0 - maf_level)] = 0
    ind[((n0 + n1) == 0) | ((n1 + n2) == 0) | ((n2 + n0) == 0)] = 0
    if W is not None and W.shape[1] != 1:
        Wt = W[keep]; WtWi = np.linalg.inv(Wt.T @ Wt)
        for t in np.nonzero(ind)[0]:
            x = G[t].copy(); x[miss[t]] = maf[t] * 2.0
            Wtx = Wt.T @ x
            if (Wtx @ (WtWi @ Wtx)) / (x @ x) > r2_level:
                ind[t] = 0
    return ind, n_miss, maf


def kinship_plink(pl, indicator_snp, k_mode=1, batch=20000):
    """src/gemma_io.cpp:1599-1738 PlinkKin (all ni_total individuals)."""
    sel = np.nonzero(indicator_snp)[0]
    n = pl.G.shape[1]
    K = np.zeros((n, n))
    for s in range(0, len(sel), batch):
        Xc = O.kin_transform(pl.G[sel[s:s + batch]], k_mode)
        K += Xc @ Xc.T
    return K / len(sel)
