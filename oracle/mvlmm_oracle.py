"""CPU restatement of the reference's multivariate LMM (src/mvlmm.cpp) -- TEST INFRASTRUCTURE ONLY.

The "next" row 2 of SURVEY 8(f) (BASELINE config 5).  Pinned against the compiled reference itself
(oracle/_ref/gemma_ref, tests/test_oracle_vs_ref.py::test_mvlmm_*), which reproduces example/demo.txt:62-66.
numpy, small d (number of phenotypes <= 4: the pairwise initialisation of MphInitial for d > 4, mvlmm.cpp:2804-2880, is
not restated).  Every function cites the reference lines it follows.  Layout as in the reference: Y is d x n, X is c x n
(rows = covariates, then the SNP), eval is n; everything is already rotated by U^T.

The Newton-Raphson derivatives are written in their closed block form instead of the reference's ~1000 lines of index
loops (Calc_* helpers, mvlmm.cpp:1015-2050): with H_k = delta_k V_g + V_e, P = H^-1 - H^-1 X (X' H^-1 X)^-1 X' H^-1 and
D_t the derivative of H with respect to one free element t of V_g or V_e (E_ij + E_ji off the diagonal, times delta_k for
V_g: the factor 2 of mvlmm.cpp:1125-1131),
    gradient_t    = -1/2 tr(P D_t) + 1/2 y' P D_t P y            (REML; tr(H^-1 D_t) instead of tr(P D_t) for ML)   :2434-2447
    Hessian_{t,u} =  1/2 tr(P D_t P D_u) - y' P D_t P D_u P y    (REML; tr(H^-1 D_t H^-1 D_u) for ML)                :2470-2488
which are the quantities CalcDev assembles (Calc_tracePD, Calc_yPDPy, Calc_tracePDPD, Calc_yPDPDPy, Calc_traceHiD, ...)."""
import numpy as np
import scipy.linalg
import scipy.stats

from . import oracle as O

EM_ITER, EM_PREC, NR_ITER, NR_PREC, P_NR = 10000, 1e-4, 100, 1e-4, 0.001      # src/param.cpp:98-99


def _eigh(A):
    """EigenDecomp(.., 0) = dsyevr, ascending (src/lapack.cpp:240-254)."""
    return scipy.linalg.eigh(A, lower=True, driver="evr")


def eigen_proc(V_g, V_e):
    """mvlmm.cpp:213-282.  Returns D_l, UltVeh, UltVehi, logdet_Ve."""
    d_e, U = _eigh(V_e)
    d = V_g.shape[0]
    Veh = np.zeros((d, d)); Vehi = np.zeros((d, d)); logdet = 0.0
    for i in range(d):
        if d_e[i] <= 0:
            continue
        logdet += np.log(d_e[i])
        s = np.sqrt(d_e[i])
        Veh += s * np.outer(U[:, i], U[:, i]); Vehi += (1.0 / s) * np.outer(U[:, i], U[:, i])
    Lam = Vehi @ (V_g @ Vehi)
    D_l, U_l = _eigh(Lam)
    D_l = np.where(D_l < 0, 0.0, D_l)
    return D_l, U_l.T @ Veh, U_l.T @ Vehi, logdet


def calc_qi(ev, D_l, X):
    """mvlmm.cpp:285-329: Q = sum_k x_k x_k' (x) diag_l 1/(D_l delta_k + 1); returns Qi, log|det Q|."""
    c, n = X.shape; d = len(D_l)
    w = 1.0 / (np.outer(D_l, ev) + 1.0)                      # d x n
    S = np.einsum("ik,jk,lk->ijl", X, X, w)                  # c x c x d
    Q = np.zeros((c * d, c * d))
    for i in range(c):
        for j in range(c):
            for l in range(d):
                Q[i * d + l, j * d + l] = S[i, j, l]
    sign, logdet = np.linalg.slogdet(Q)
    return np.linalg.inv(Q), logdet


def calc_xHiy(ev, D_l, X, UltVehiY):
    """mvlmm.cpp:334-359: xHiy[j d + i] = sum_k x_jk y_ik / (delta_k D_i + 1)."""
    w = 1.0 / (np.outer(D_l, ev) + 1.0)
    return np.einsum("jk,ik,ik->ji", X, UltVehiY, w).reshape(-1)


def mph_calc_logl(ev, xHiy, D_l, UltVehiY, Qi):
    """mvlmm.cpp:565-594."""
    v = np.outer(D_l, ev) + 1.0
    logl = float(np.sum(UltVehiY ** 2 / v + np.log(v)))
    logl -= float(xHiy @ (Qi @ xHiy))
    return -0.5 * logl


def _logl_const(fn, X, n, d):
    """mvlmm.cpp:641-648 / 2665-2672."""
    c = X.shape[0]
    if fn == "R":
        sign, ld = np.linalg.slogdet(X @ X.T)
        return -0.5 * (n - c) * d * np.log(2.0 * np.pi) + 0.5 * d * ld
    return -0.5 * n * d * np.log(2.0 * np.pi)


def mph_em(fn, max_iter, max_prec, ev, X, Y, V_g, V_e, B):
    """MphEM, mvlmm.cpp:599-724.  Updates V_g, V_e, B in place; returns the last log-likelihood."""
    c, n = X.shape; d = Y.shape[0]
    XXti = np.linalg.inv(X @ X.T)
    const = _logl_const(fn, X, n, d)
    logl_old = logl_new = 0.0
    UltVehiBX = np.zeros((d, n)); UltVehiB = np.zeros((d, c))
    for t in range(max_iter):
        D_l, UltVeh, UltVehi, logdet_Ve = eigen_proc(V_g, V_e)
        Qi, logdet_Q = calc_qi(ev, D_l, X)
        UltVehiY = UltVehi @ Y
        xHiy = calc_xHiy(ev, D_l, X, UltVehiY)
        logl_new = const + mph_calc_logl(ev, xHiy, D_l, UltVehiY, Qi) - 0.5 * n * logdet_Ve
        if fn == "R":
            logl_new += -0.5 * (logdet_Q - c * logdet_Ve)
        if t != 0 and abs(logl_new - logl_old) < max_prec:
            break
        logl_old = logl_new
        OmegaU = D_l[:, None] / (np.outer(D_l, ev) + 1.0)                 # CalcOmega :363-382
        OmegaE = ev[None, :] * OmegaU
        if fn == "R":
            UltVehiB = (Qi @ xHiy).reshape(c, d).T                           # UpdateRL_B :420-441
            UltVehiBX = UltVehiB @ X
        elif t == 0:
            UltVehiB = UltVehi @ B
            UltVehiBX = UltVehiB @ X
        UltVehiU = (UltVehiY - UltVehiBX) * OmegaE                           # UpdateU :384-391
        if fn == "L":
            UltVehiB = ((UltVehiY - UltVehiU) @ X.T) @ XXti                  # UpdateL_B :402-418
            UltVehiBX = UltVehiB @ X
        UltVehiE = UltVehiY - UltVehiBX - UltVehiU                           # UpdateE :393-400
        U_hat = UltVeh.T @ UltVehiU; E_hat = UltVeh.T @ UltVehiE
        B[:, :] = UltVeh.T @ UltVehiB
        # CalcSigma :485-560
        S_uu = np.diag(OmegaU.sum(axis=1)); S_ee = np.diag(OmegaE.sum(axis=1))
        if fn == "R":
            Qi4 = Qi.reshape(c, d, c, d)
            T = np.einsum("jk,jalb,lk->abk", X, Qi4, X)                      # x_k' Qi_[a,b] x_k
            we = 1.0 / (np.outer(D_l, ev) + 1.0); wu = we * D_l[:, None]
            S_uu = S_uu + np.einsum("k,ak,bk,abk->ab", ev, wu, wu, T)
            S_ee = S_ee + np.einsum("ak,bk,abk->ab", we, we, T)
        S_uu = UltVeh.T @ S_uu @ UltVeh; S_ee = UltVeh.T @ S_ee @ UltVeh
        nz = ev != 0                                                          # UpdateV :443-483
        V_g[:, :] = ((U_hat[:, nz] / ev[nz]) @ U_hat[:, nz].T + S_uu) / n
        V_e[:, :] = (E_hat @ E_hat.T + S_ee) / n
    return logl_new


def mph_calc_p(ev, x, W, Y, V_g, V_e):
    """MphCalcP, mvlmm.cpp:727-831.  Returns p, beta (d), Vbeta (d x d)."""
    c, n = W.shape; d = Y.shape[0]
    D_l, UltVeh, UltVehi, _ = eigen_proc(V_g, V_e)
    Qi, _ = calc_qi(ev, D_l, W)
    UltVehiY = UltVehi @ Y
    w = 1.0 / (np.outer(D_l, ev) + 1.0)
    xPy = np.einsum("k,ik,ik->i", x, UltVehiY, w)
    xPx = np.diag(np.einsum("k,k,ik->i", x, x, w))
    WHix = np.zeros((c * d, d)); WHiy = np.zeros(c * d)
    a1 = np.einsum("k,jk,ik->ji", x, W, w); a2 = np.einsum("ik,jk,ik->ji", UltVehiY, W, w)
    for i in range(d):
        for j in range(c):
            WHix[j * d + i, i] = a1[j, i]; WHiy[j * d + i] = a2[j, i]
    QiWHix = Qi @ WHix
    xPx = xPx - WHix.T @ QiWHix
    xPy = xPy - QiWHix.T @ WHiy
    b = np.linalg.solve(xPx, xPy)
    Vb = np.linalg.inv(xPx)
    beta = UltVeh.T @ b
    Vbeta = UltVeh.T @ (Vb @ UltVeh)
    stat = float(b @ xPy)
    return float(scipy.stats.chi2.sf(stat, d)), beta, Vbeta


def mph_calc_beta(ev, W, Y, V_g, V_e):
    """MphCalcBeta, mvlmm.cpp:835-937: B (d x c)."""
    c, n = W.shape; d = Y.shape[0]
    D_l, UltVeh, UltVehi, _ = eigen_proc(V_g, V_e)
    Qi, _ = calc_qi(ev, D_l, W)
    WHiy = calc_xHiy(ev, D_l, W, UltVehi @ Y)
    q = (Qi @ WHiy).reshape(c, d)
    return (UltVeh.T @ q.T)


def _free_elements(d):
    return [(i, j) for i in range(d) for j in range(i, d)]             # GetIndex order, mvlmm.cpp:1093-1109


def _nr_quantities(fn, ev, X, Y, V_g, V_e):
    """CalcHiQi + Calc_Hiy_all + Calc_xHi_all + Calc_xHiy + CalcDev (mvlmm.cpp:942-1090, 2360-2556) in block form.
    Returns logdet_H, logdet_Q, yPy, gradient (2v), Hessian (2v x 2v)."""
    c, n = X.shape; d = Y.shape[0]
    H = ev[:, None, None] * V_g[None] + V_e[None]                      # n x d x d
    Hi = np.linalg.inv(H)
    logdet_H = float(np.sum(np.linalg.slogdet(H)[1]))
    Q = np.einsum("ik,jk,kab->iajb", X, X, Hi).reshape(c * d, c * d)    # (i a),(j b)
    Qi = np.linalg.inv(Q)
    logdet_Q = float(np.linalg.slogdet(Q)[1])
    Hiy = np.einsum("kab,bk->ak", Hi, Y)                                # d x n
    xHiy = np.einsum("ik,ak->ia", X, Hiy).reshape(-1)
    QixHiy = Qi @ xHiy
    yPy = float(np.sum(Y * Hiy) - QixHiy @ xHiy)
    Bh = QixHiy.reshape(c, d).T                                         # d x c
    g = np.einsum("kab,bk->ak", Hi, Y - Bh @ X)                         # (P y)_k, d x n
    el = _free_elements(d); v = len(el)
    # D_t for every free element: list of (E (d x d), scale_k)
    Ds = []
    for (i, j) in el:
        E = np.zeros((d, d)); E[i, j] = 1.0; E[j, i] = 1.0
        Ds.append((E, ev))                                              # V_g element
    for (i, j) in el:
        E = np.zeros((d, d)); E[i, j] = 1.0; E[j, i] = 1.0
        Ds.append((E, np.ones(n)))                                      # V_e element
    HiDHi = [np.einsum("k,kab,bc,kcd->kad", s, Hi, E, Hi) for (E, s) in Ds]          # n x d x d each
    A = [np.einsum("ik,jk,kab->iajb", X, X, M).reshape(c * d, c * d) for M in HiDHi]   # sum_k X_k Hi D Hi X_k'
    u = [np.einsum("k,ab,bk->ak", s, E, g) for (E, s) in Ds]                         # D_t (P y), d x n
    Hu = [np.einsum("kab,bk->ak", Hi, uu) for uu in u]
    XHu = [np.einsum("ik,ak->ia", X, h).reshape(-1) for h in Hu]
    grad = np.zeros(2 * v); Hess = np.zeros((2 * v, 2 * v))
    for t in range(2 * v):
        E, s = Ds[t]
        trHiD = float(np.einsum("k,kab,ba->", s, Hi, E))
        yPDPy = float(np.sum(g * u[t]))
        if fn == "R":
            grad[t] = -0.5 * (trHiD - float(np.trace(Qi @ A[t]))) + 0.5 * yPDPy
        else:
            grad[t] = -0.5 * trHiD + 0.5 * yPDPy
    for t in range(2 * v):
        for r in range(t, 2 * v):
            E2, s2 = Ds[r]
            yPDPDPy = float(np.sum(u[t] * Hu[r]) - XHu[t] @ (Qi @ XHu[r]))
            trHH = float(np.einsum("kab,k,ba->", HiDHi[t], s2, E2))                   # tr(Hi D_t Hi D_r)
            if fn == "R":
                M3 = np.einsum("kab,k,bc,kcd->kad", HiDHi[t], s2, E2, Hi)             # Hi D_t Hi D_r Hi
                A3 = np.einsum("ik,jk,kab->iajb", X, X, M3).reshape(c * d, c * d)
                trPP = trHH - 2.0 * float(np.trace(Qi @ A3)) + float(np.trace(Qi @ A[t] @ Qi @ A[r]))
                h = 0.5 * trPP - yPDPDPy
            else:
                h = 0.5 * trHH - yPDPDPy
            Hess[t, r] = Hess[r, t] = h
    return logdet_H, logdet_Q, yPy, grad, Hess


def _is_pd(V):
    return bool(np.all(_eigh(V)[0] > 0))


def mph_nr(fn, max_iter, max_prec, ev, X, Y, V_g, V_e):
    """MphNR, mvlmm.cpp:2608-2760 (+ UpdateVgVe :2557-2606).  Updates V_g, V_e in place; returns logl, -Hessian^-1."""
    c, n = X.shape; d = Y.shape[0]
    el = _free_elements(d); v = len(el)
    const = _logl_const(fn, X, n, d)
    logl_old = logl_new = 0.0
    grad = np.zeros(2 * v); Hinv = np.zeros((2 * v, 2 * v)); Hess = np.zeros((2 * v, 2 * v))
    for t in range(max_iter):
        Vg_save, Ve_save = V_g.copy(), V_e.copy()
        step_scale, step_iter = 1.0, 0
        while True:
            V_g[:, :] = Vg_save; V_e[:, :] = Ve_save
            if t != 0:
                vec = np.array([V_g[i, j] for (i, j) in el] + [V_e[i, j] for (i, j) in el])
                vec = vec - step_scale * (Hinv @ grad)
                for q, (i, j) in enumerate(el):
                    V_g[i, j] = V_g[j, i] = vec[q]; V_e[i, j] = V_e[j, i] = vec[q + v]
            flag_pd = _is_pd(V_e) and _is_pd(V_g)
            if flag_pd:
                logdet_H, logdet_Q, yPy, g_new, H_new = _nr_quantities(fn, ev, X, Y, V_g, V_e)
                logl_new = const - 0.5 * logdet_H - 0.5 * yPy - (0.5 * logdet_Q if fn == "R" else 0.0)
            step_scale /= 2.0; step_iter += 1
            if not ((not flag_pd or logl_new < logl_old or logl_new - logl_old > 10) and step_iter < 10 and t != 0):
                break
        if t != 0:
            if logl_new < logl_old or not flag_pd:
                V_g[:, :] = Vg_save; V_e[:, :] = Ve_save
                break
            if logl_new - logl_old < max_prec:
                break
        logl_old = logl_new
        grad, Hess = g_new, H_new
        Hinv = np.linalg.inv(Hess)
    return logl_new, -Hinv


def mph_initial(ev, X, Y, l_min=1e-5, l_max=1e5, n_region=10):
    """MphInitial, mvlmm.cpp:2763-2948 (d <= 4): univariate REML per trait for the diagonals, GLS B."""
    d = Y.shape[0]; c = X.shape[0]
    if d > 4:
        raise NotImplementedError("pairwise initialisation for d > 4 (mvlmm.cpp:2804-2880) is not restated")
    V_g = np.zeros((d, d)); V_e = np.zeros((d, d))
    Xt = np.ascontiguousarray(X.T)
    for i in range(d):
        lam, _ = O.calc_lambda_null("R", ev, Xt, Y[i], l_min, l_max, n_region)
        vg, ve, _, _ = O.calc_vgvebeta(ev, Xt, Y[i], lam)
        V_g[i, i] = vg; V_e[i, i] = ve
    B = mph_calc_beta(ev, X, Y, V_g, V_e)
    return V_g, V_e, B


def null_model(ev, UtW, UtY):
    """Null part of MVLMM::AnalyzeBimbam, mvlmm.cpp:3056-3160.  UtW: n x c, UtY: n x d.  Returns the REML and ML estimates; the
    per-SNP fits start from the LAST ones (the ML estimates: V_g_null is copied after the 'L' block, :3205-3207)."""
    X = np.ascontiguousarray(UtW.T); Y = np.ascontiguousarray(UtY.T)
    V_g, V_e, B = mph_initial(ev, X, Y)
    mph_em("R", EM_ITER, EM_PREC, ev, X, Y, V_g, V_e, B)
    logl_remle, cov_r = mph_nr("R", NR_ITER, NR_PREC, ev, X, Y, V_g, V_e)
    B = mph_calc_beta(ev, X, Y, V_g, V_e)
    out = dict(Vg_remle=V_g.copy(), Ve_remle=V_e.copy(), logl_remle_H0=logl_remle, cov_remle=cov_r, B_remle=B.copy())
    mph_em("L", EM_ITER, EM_PREC, ev, X, Y, V_g, V_e, B)
    logl_mle, cov_m = mph_nr("L", NR_ITER, NR_PREC, ev, X, Y, V_g, V_e)
    B = mph_calc_beta(ev, X, Y, V_g, V_e)
    out.update(Vg_mle=V_g.copy(), Ve_mle=V_e.copy(), logl_mle_H0=logl_mle, B_mle=B.copy())
    return out


def analyze_snp_wald(ev, UtW, UtY, Utx, nm):
    """-lmm 1 branch of the per-SNP loop, mvlmm.cpp:3286-3347.  Returns beta (d), Vbeta (d x d), p_wald."""
    W = np.ascontiguousarray(UtW.T); Y = np.ascontiguousarray(UtY.T)
    X = np.vstack([W, Utx[None, :]])
    d = Y.shape[0]
    V_g, V_e = nm["Vg_mle"].copy(), nm["Ve_mle"].copy()
    B = np.hstack([nm["B_mle"], np.zeros((d, 1))])
    mph_em("R", EM_ITER // 10, EM_PREC * 10, ev, X, Y, V_g, V_e, B)
    p, beta, Vbeta = mph_calc_p(ev, Utx, W, Y, V_g, V_e)
    if p < P_NR:
        mph_nr("R", NR_ITER // 10, NR_PREC * 10, ev, X, Y, V_g, V_e)
        p, beta, Vbeta = mph_calc_p(ev, Utx, W, Y, V_g, V_e)
    return beta, Vbeta, p


def analyze_snp(ev, UtW, UtY, Utx, nm, a_mode):
    """Per-SNP body of MVLMM::AnalyzeBimbam for -lmm 1/2/3/4 (mvlmm.cpp:3286-3360; crt = 0).  V_g, V_e, B are reset to the null
    (ML) estimates once per SNP and then carried from the LRT block into the Wald block, as in the reference.
    Returns beta (d), Vbeta (d x d), p_wald, p_lrt, p_score."""
    W = np.ascontiguousarray(UtW.T); Y = np.ascontiguousarray(UtY.T)
    X = np.vstack([W, Utx[None, :]])
    d = Y.shape[0]
    V_g, V_e = nm["Vg_mle"].copy(), nm["Ve_mle"].copy()
    B = np.hstack([nm["B_mle"], np.zeros((d, 1))])
    beta = np.zeros(d); Vbeta = np.zeros((d, d)); p_wald = p_lrt = p_score = 0.0
    if a_mode in (3, 4):
        p_score, beta, Vbeta = mph_calc_p(ev, Utx, W, Y, nm["Vg_mle"], nm["Ve_mle"])
    if a_mode in (2, 4):
        logl_H1 = mph_em("L", EM_ITER // 10, EM_PREC * 10, ev, X, Y, V_g, V_e, B)
        _, beta, Vbeta = mph_calc_p(ev, Utx, W, Y, V_g, V_e)
        p_lrt = float(scipy.stats.chi2.sf(2.0 * (logl_H1 - nm["logl_mle_H0"]), d))
        if p_lrt < P_NR:
            logl_H1, _ = mph_nr("L", NR_ITER // 10, NR_PREC * 10, ev, X, Y, V_g, V_e)
            _, beta, Vbeta = mph_calc_p(ev, Utx, W, Y, V_g, V_e)
            p_lrt = float(scipy.stats.chi2.sf(2.0 * (logl_H1 - nm["logl_mle_H0"]), d))
    if a_mode in (1, 4):
        mph_em("R", EM_ITER // 10, EM_PREC * 10, ev, X, Y, V_g, V_e, B)
        p_wald, beta, Vbeta = mph_calc_p(ev, Utx, W, Y, V_g, V_e)
        if p_wald < P_NR:
            mph_nr("R", NR_ITER // 10, NR_PREC * 10, ev, X, Y, V_g, V_e)
            p_wald, beta, Vbeta = mph_calc_p(ev, Utx, W, Y, V_g, V_e)
    return beta, Vbeta, p_wald, p_lrt, p_score
