// mvlmm_core.cuh -- multivariate LMM (two phenotypes and up) in "moment form".
//
// What it replaces in the reference (src/mvlmm.cpp): EigenProc :213-282, CalcQi :285-329, CalcXHiY :334-359, CalcOmega :363-382,
// UpdateU/E/L_B/RL_B :384-441, UpdateV :443-483, CalcSigma :485-560, MphCalcLogL :565-594, MphEM :599-724, MphCalcP :727-831,
// MphCalcBeta :835-937, the Newton-Raphson machinery CalcHiQi ... CalcDev :942-2556, UpdateVgVe :2557-2606, MphNR :2608-2760.
//
// After the simultaneous diagonalisation of (V_g, V_e) -- A = U_l^T V_e^{-1/2}, H_k = A^-1 diag(delta_k D_l + 1) A^-T -- every
// quantity of the EM and of the Newton-Raphson step is a small algebraic expression in WEIGHTED MOMENTS of the rotated data
//      M[w][a,b] = sum_k w_k z_a,k z_b,k ,   z = (x_1 .. x_C1, y_1 .. y_D)   (covariates, SNP, phenotypes; all U^T-rotated)
// for the weights  w_l, w_l w_m, delta w_l w_m  (EM, gradient) and  delta^p w_l w_m w_o, p = 0..2  (Hessian), w_l = 1/(delta D_l + 1).
// The reference materialises d x n and (dc x dc) x n arrays per SNP and iteration instead.  One warp owns one SNP: each moment
// family is one lane-strided pass over the n individuals followed by a shuffle butterfly, and the small algebra (D x D eigen
// problems, C1 x C1 inverses per trait direction, the 2v x 2v Newton system) runs redundantly in every lane.
// The same code compiles for the host (GB_MV_HOST: one "lane"): the test harness tests/host/mvlmm_check.cpp checks this
// formulation against the restated reference algorithm without a GPU.
#pragma once
#include <math.h>
#include <float.h>

#ifdef GB_MV_HOST
#define GB_HD
#define GB_HD_BIG
#define GB_MV_LANE 0
#define GB_MV_NLANE 1
static inline double gb_mv_allsum(double v) { return v; }
#else
#define GB_HD __device__ __forceinline__
#define GB_HD_BIG __device__ __noinline__      /* the EM / NR / test routines are called, not inlined: keeps ptxas time and code size bounded */
#define GB_MV_LANE (threadIdx.x & 31)
#define GB_MV_NLANE 32
__device__ __forceinline__ double gb_mv_allsum(double v) {
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
  return v;
}
#endif

namespace gbmv {

constexpr int pair_index(int a, int b, int nz) { return a <= b ? (2 * nz - a + 1) * a / 2 + (b - a) : (2 * nz - b + 1) * b / 2 + (a - b); }

// ---- small dense helpers (N <= 8) ---------------------------------------------------------------------------------------
template <int N>
GB_HD_BIG bool inv_small(const double (&A)[N][N], double (&Ai)[N][N], double &logabsdet) {
  double M[N][2 * N];
  for (int i = 0; i < N; ++i) for (int j = 0; j < N; ++j) { M[i][j] = A[i][j]; M[i][N + j] = (i == j) ? 1.0 : 0.0; }
  logabsdet = 0.0;
  for (int k = 0; k < N; ++k) {
    int pr = k;
    for (int i = k + 1; i < N; ++i) if (fabs(M[i][k]) > fabs(M[pr][k])) pr = i;
    if (M[pr][k] == 0.0) return false;
    if (pr != k) for (int j = 0; j < 2 * N; ++j) { const double t = M[k][j]; M[k][j] = M[pr][j]; M[pr][j] = t; }
    const double piv = M[k][k];
    logabsdet += log(fabs(piv));
    for (int j = 0; j < 2 * N; ++j) M[k][j] /= piv;
    for (int i = 0; i < N; ++i) if (i != k) { const double f = M[i][k]; if (f != 0.0) for (int j = 0; j < 2 * N; ++j) M[i][j] -= f * M[k][j]; }
  }
  for (int i = 0; i < N; ++i) for (int j = 0; j < N; ++j) Ai[i][j] = M[i][N + j];
  return true;
}

// symmetric eigenproblem, cyclic Jacobi: A = V diag(ev) V^T (columns of V); order unspecified (nothing below depends on it)
template <int N>
GB_HD_BIG void sym_eig(const double (&A0)[N][N], double (&ev)[N], double (&V)[N][N]) {
  double A[N][N];
  for (int i = 0; i < N; ++i) for (int j = 0; j < N; ++j) { A[i][j] = A0[i][j]; V[i][j] = (i == j) ? 1.0 : 0.0; }
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0.0, dia = 0.0;
    for (int i = 0; i < N; ++i) { dia += fabs(A[i][i]); for (int j = i + 1; j < N; ++j) off += fabs(A[i][j]); }
    if (off <= 1e-300 || off <= DBL_EPSILON * 1e-3 * dia) break;
    for (int p = 0; p < N; ++p) for (int q = p + 1; q < N; ++q) {
      if (A[p][q] == 0.0) continue;
      const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
      const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
      const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
      for (int k = 0; k < N; ++k) { const double akp = A[k][p], akq = A[k][q]; A[k][p] = c * akp - s * akq; A[k][q] = s * akp + c * akq; }
      for (int k = 0; k < N; ++k) { const double apk = A[p][k], aqk = A[q][k]; A[p][k] = c * apk - s * aqk; A[q][k] = s * apk + c * aqk; }
      for (int k = 0; k < N; ++k) { const double vkp = V[k][p], vkq = V[k][q]; V[k][p] = c * vkp - s * vkq; V[k][q] = s * vkp + c * vkq; }
    }
  }
  for (int i = 0; i < N; ++i) ev[i] = A[i][i];
}

// ---- data and moments ------------------------------------------------------------------------------------------------------
template <int NZ>
struct MvData {
  int n;
  const double *delta;
  const double *z[NZ];          // z[a][k]: rows of covariates / SNP first, then the D phenotype rows
};

// NW weight families in one pass: out[w][pair] = sum_k W(w,k) z_a z_b ; scal[w] = sum_k W(w,k).  WF: functor (delta_k, double (&w)[NW])
template <int NZ, int NW, class WF>
GB_HD void moments(const MvData<NZ> &d, WF wf, double (&out)[NW][NZ * (NZ + 1) / 2], double (&scal)[NW]) {
  constexpr int NP = NZ * (NZ + 1) / 2;
  for (int w = 0; w < NW; ++w) { scal[w] = 0.0; for (int p = 0; p < NP; ++p) out[w][p] = 0.0; }
  for (int k = GB_MV_LANE; k < d.n; k += GB_MV_NLANE) {
    double z[NZ], wt[NW];
    for (int a = 0; a < NZ; ++a) z[a] = d.z[a][k];
    wf(d.delta[k], wt);
    double pr[NP];
    for (int a = 0; a < NZ; ++a) for (int b = a; b < NZ; ++b) pr[pair_index(a, b, NZ)] = z[a] * z[b];
    for (int w = 0; w < NW; ++w) {
      scal[w] += wt[w];
      for (int p = 0; p < NP; ++p) out[w][p] = fma(wt[w], pr[p], out[w][p]);
    }
  }
  for (int w = 0; w < NW; ++w) { scal[w] = gb_mv_allsum(scal[w]); for (int p = 0; p < NP; ++p) out[w][p] = gb_mv_allsum(out[w][p]); }
}

// ---- model state -------------------------------------------------------------------------------------------------------------
template <int D>
struct Basis {                    // EigenProc
  double Dl[D];
  double UltVeh[D][D], UltVehi[D][D];
  double logdet_Ve;
};

template <int D>
GB_HD_BIG void eigen_proc(const double (&V_g)[D][D], const double (&V_e)[D][D], Basis<D> &b) {
  double de[D], U[D][D], Veh[D][D], Vehi[D][D];
  sym_eig<D>(V_e, de, U);
  b.logdet_Ve = 0.0;
  for (int i = 0; i < D; ++i) for (int j = 0; j < D; ++j) { Veh[i][j] = 0.0; Vehi[i][j] = 0.0; }
  for (int q = 0; q < D; ++q) {
    if (de[q] <= 0) continue;
    b.logdet_Ve += log(de[q]);
    const double s = sqrt(de[q]);
    for (int i = 0; i < D; ++i) for (int j = 0; j < D; ++j) { Veh[i][j] += s * U[i][q] * U[j][q]; Vehi[i][j] += (1.0 / s) * U[i][q] * U[j][q]; }
  }
  double T[D][D], Lam[D][D];
  for (int i = 0; i < D; ++i) for (int j = 0; j < D; ++j) { double t = 0.0; for (int q = 0; q < D; ++q) t += V_g[i][q] * Vehi[q][j]; T[i][j] = t; }
  for (int i = 0; i < D; ++i) for (int j = 0; j < D; ++j) { double t = 0.0; for (int q = 0; q < D; ++q) t += Vehi[i][q] * T[q][j]; Lam[i][j] = t; }
  for (int i = 0; i < D; ++i) for (int j = i + 1; j < D; ++j) { const double m = 0.5 * (Lam[i][j] + Lam[j][i]); Lam[i][j] = Lam[j][i] = m; }
  double Ul[D][D];
  sym_eig<D>(Lam, b.Dl, Ul);
  for (int i = 0; i < D; ++i) if (b.Dl[i] < 0) b.Dl[i] = 0.0;
  for (int i = 0; i < D; ++i) for (int j = 0; j < D; ++j) {
    double t1 = 0.0, t2 = 0.0;
    for (int q = 0; q < D; ++q) { t1 += Ul[q][i] * Veh[q][j]; t2 += Ul[q][i] * Vehi[q][j]; }
    b.UltVeh[i][j] = t1; b.UltVehi[i][j] = t2;
  }
}

constexpr int dd_index(int l, int m, int D) { return l <= m ? (2 * D - l + 1) * l / 2 + (m - l) : (2 * D - m + 1) * m / 2 + (l - m); }

// quantities of one (V_g, V_e) evaluation that both EM and NR need
template <int D, int C1>
struct Eval {
  static constexpr int NZ = C1 + D, NP = NZ * (NZ + 1) / 2, DD = D * (D + 1) / 2;
  Basis<D> bs;
  double M1[D][NP], sw[D], logv[D];       // weights w_l ; sum_k w_l ; sum_k log(delta D_l + 1)
  double Qi[D][C1][C1];                   // inverse of Q_l = sum w_l x x'
  double logdetQ;                         // sum_l log|det Q_l|
  double xHiy[D][C1];                     // sum_k w_l x_j y'_l
  double Bp[D][C1];                       // Qi_l xHiy_l  (UltVehiB of the REML update)
  double yPy_t;                           // sum_l (y'y'_l - xHiy_l' Qi_l xHiy_l)
  double logl;                            // without the constant
};

// moments with weights w_l (+ log terms), then Q, xHiy, logl.  x rows used: the first CX of the C1 (CX = C1 normally; MphCalcP uses the covariates only)
template <int D, int C1>
GB_HD_BIG bool evaluate(const MvData<C1 + D> &dat, const double (&V_g)[D][D], const double (&V_e)[D][D], bool reml, Eval<D, C1> &e) {
  constexpr int NZ = C1 + D;
  eigen_proc<D>(V_g, V_e, e.bs);
  const Basis<D> &b = e.bs;
  double Dl[D];
  for (int l = 0; l < D; ++l) Dl[l] = b.Dl[l];
  moments<NZ, D>(dat, [&](double delta, double (&w)[D]) { for (int l = 0; l < D; ++l) w[l] = 1.0 / (delta * Dl[l] + 1.0); }, e.M1, e.sw);
  // log terms: a separate cheap pass
  for (int l = 0; l < D; ++l) e.logv[l] = 0.0;
  for (int k = GB_MV_LANE; k < dat.n; k += GB_MV_NLANE) for (int l = 0; l < D; ++l) e.logv[l] += log(dat.delta[k] * Dl[l] + 1.0);
  for (int l = 0; l < D; ++l) e.logv[l] = gb_mv_allsum(e.logv[l]);
  e.logdetQ = 0.0; e.yPy_t = 0.0;
  double sumlog = 0.0;
  for (int l = 0; l < D; ++l) {
    double Q[C1][C1], ld;
    for (int i = 0; i < C1; ++i) for (int j = 0; j < C1; ++j) Q[i][j] = e.M1[l][pair_index(i, j, NZ)];
    if (!inv_small<C1>(Q, e.Qi[l], ld)) return false;
    e.logdetQ += ld;
    double yy = 0.0;
    for (int s = 0; s < D; ++s) for (int t = 0; t < D; ++t) yy += b.UltVehi[l][s] * b.UltVehi[l][t] * e.M1[l][pair_index(C1 + s, C1 + t, NZ)];
    for (int j = 0; j < C1; ++j) { double v = 0.0; for (int s = 0; s < D; ++s) v += b.UltVehi[l][s] * e.M1[l][pair_index(j, C1 + s, NZ)]; e.xHiy[l][j] = v; }
    double q = 0.0;
    for (int i = 0; i < C1; ++i) { double v = 0.0; for (int j = 0; j < C1; ++j) v += e.Qi[l][i][j] * e.xHiy[l][j]; e.Bp[l][i] = v; q += v * e.xHiy[l][i]; }
    e.yPy_t += yy - q;
    sumlog += e.logv[l];
  }
  e.logl = -0.5 * (e.yPy_t + sumlog) - 0.5 * (double)dat.n * b.logdet_Ve;
  if (reml) e.logl += -0.5 * (e.logdetQ - (double)C1 * b.logdet_Ve);
  return true;
}

template <int D, int C1>
struct Fit { double V_g[D][D], V_e[D][D], B[D][C1]; double logl; };

// log-likelihood constant (mvlmm.cpp:641-648): needs X X' of the C1 x rows (unit-weight moments)
template <int D, int C1>
GB_HD_BIG double logl_const(const MvData<C1 + D> &dat, bool reml) {
  constexpr int NZ = C1 + D;
  const double l2pi = 1.8378770664093453;
  if (!reml) return -0.5 * (double)dat.n * (double)D * l2pi;
  double M0[1][NZ * (NZ + 1) / 2], s0[1];
  moments<NZ, 1>(dat, [](double, double (&w)[1]) { w[0] = 1.0; }, M0, s0);
  double XXt[C1][C1], XXti[C1][C1], ld = 0.0;
  for (int i = 0; i < C1; ++i) for (int j = 0; j < C1; ++j) XXt[i][j] = M0[0][pair_index(i, j, NZ)];
  inv_small<C1>(XXt, XXti, ld);
  return -0.5 * (double)(dat.n - C1) * (double)D * l2pi + 0.5 * (double)D * ld;
}

template <int NZ>
GB_HD double quad(const double *M, const double *u, const double *v) {
  double s = 0.0;
  for (int a = 0; a < NZ; ++a) for (int b = 0; b < NZ; ++b) s += u[a] * v[b] * M[pair_index(a, b, NZ)];
  return s;
}

// MphEM (mvlmm.cpp:599-724).  fit.V_g / V_e / B are the starting values and receive the result.
template <int D, int C1>
GB_HD_BIG double mph_em(bool reml, int max_iter, double max_prec, const MvData<C1 + D> &dat, Fit<D, C1> &fit) {
  constexpr int NZ = C1 + D, NP = NZ * (NZ + 1) / 2, DD = D * (D + 1) / 2;
  const double cst = logl_const<D, C1>(dat, reml);
  double M0[1][NP], s0[1];
  if (!reml) moments<NZ, 1>(dat, [](double, double (&w)[1]) { w[0] = 1.0; }, M0, s0);
  double XXti[C1][C1];
  if (!reml) { double XXt[C1][C1], ld; for (int i = 0; i < C1; ++i) for (int j = 0; j < C1; ++j) XXt[i][j] = M0[0][pair_index(i, j, NZ)]; inv_small<C1>(XXt, XXti, ld); }
  double logl_old = 0.0, logl_new = 0.0;
  double Bold[D][C1];                                     // UltVehiB carried from the previous iteration (ML branch)
  for (int l = 0; l < D; ++l) for (int j = 0; j < C1; ++j) Bold[l][j] = 0.0;
  Eval<D, C1> e;
  for (int t = 0; t < max_iter; ++t) {
    if (!evaluate<D, C1>(dat, fit.V_g, fit.V_e, reml, e)) break;
    logl_new = cst + e.logl;
    if (t != 0 && fabs(logl_new - logl_old) < max_prec) break;
    logl_old = logl_new;
    const Basis<D> &b = e.bs;
    double Dl[D];
    for (int l = 0; l < D; ++l) Dl[l] = b.Dl[l];
    // second-order moments: w_l w_m and delta w_l w_m
    double M2[DD][NP], M3[DD][NP], s2[DD], s3[DD];
    moments<NZ, DD>(dat, [&](double delta, double (&w)[DD]) {
      double wl[D]; for (int l = 0; l < D; ++l) wl[l] = 1.0 / (delta * Dl[l] + 1.0);
      for (int l = 0; l < D; ++l) for (int m = l; m < D; ++m) w[dd_index(l, m, D)] = wl[l] * wl[m]; }, M2, s2);
    moments<NZ, DD>(dat, [&](double delta, double (&w)[DD]) {
      double wl[D]; for (int l = 0; l < D; ++l) wl[l] = 1.0 / (delta * Dl[l] + 1.0);
      for (int l = 0; l < D; ++l) for (int m = l; m < D; ++m) w[dd_index(l, m, D)] = delta * wl[l] * wl[m]; }, M3, s3);
    // B' used by UpdateU ("old") and by UpdateE ("new")
    double Bo[D][C1], Bn[D][C1];
    if (reml) { for (int l = 0; l < D; ++l) for (int j = 0; j < C1; ++j) { Bo[l][j] = e.Bp[l][j]; Bn[l][j] = e.Bp[l][j]; } }
    else {
      if (t == 0) { for (int l = 0; l < D; ++l) for (int j = 0; j < C1; ++j) { double v = 0.0; for (int s = 0; s < D; ++s) v += b.UltVehi[l][s] * fit.B[s][j]; Bo[l][j] = v; } }
      else { for (int l = 0; l < D; ++l) for (int j = 0; j < C1; ++j) Bo[l][j] = Bold[l][j]; }
      // UpdateL_B: ((UltVehiY - UltVehiU) X') (X X')^-1 with UltVehiY - UltVehiU = w y' + (1 - w) Bo x
      for (int l = 0; l < D; ++l) {
        double rhs[C1];
        for (int j = 0; j < C1; ++j) {
          double v = 0.0;
          for (int s = 0; s < D; ++s) v += b.UltVehi[l][s] * e.M1[l][pair_index(j, C1 + s, NZ)];
          for (int i = 0; i < C1; ++i) v += Bo[l][i] * (M0[0][pair_index(i, j, NZ)] - e.M1[l][pair_index(i, j, NZ)]);
          rhs[j] = v;
        }
        for (int j = 0; j < C1; ++j) { double v = 0.0; for (int i = 0; i < C1; ++i) v += rhs[i] * XXti[i][j]; Bn[l][j] = v; }
      }
    }
    for (int l = 0; l < D; ++l) for (int j = 0; j < C1; ++j) Bold[l][j] = Bn[l][j];
    // coefficient vectors over z: r_l = y'_l - Bo_l x (used with weight w_l), q_l = (Bo_l - Bn_l) x (unit weight)
    double co[D][NZ], cq[D][NZ];
    for (int l = 0; l < D; ++l) {
      for (int j = 0; j < C1; ++j) { co[l][j] = -Bo[l][j]; cq[l][j] = Bo[l][j] - Bn[l][j]; }
      for (int s = 0; s < D; ++s) { co[l][C1 + s] = b.UltVehi[l][s]; cq[l][C1 + s] = 0.0; }
    }
    double UU[D][D], EE[D][D];
    for (int l = 0; l < D; ++l) for (int m = 0; m < D; ++m) {
      const int lm = dd_index(l, m, D);
      UU[l][m] = Dl[l] * Dl[m] * quad<NZ>(M3[lm], co[l], co[m]);
      double ee = quad<NZ>(M2[lm], co[l], co[m]);
      if (!reml) ee += quad<NZ>(e.M1[l], co[l], cq[m]) + quad<NZ>(e.M1[m], cq[l], co[m]) + quad<NZ>(M0[0], cq[l], cq[m]);
      EE[l][m] = ee;
    }
    // CalcSigma: diagonal in the rotated basis
    for (int a = 0; a < D; ++a) {
      double suu = Dl[a] * e.sw[a], see = (double)dat.n - e.sw[a];
      if (Dl[a] == 0.0) see = 0.0;                         // OmegaE = delta D w = 0 exactly when D_l = 0 (n - sum w would round to ~0 anyway)
      if (reml) {
        double qu = 0.0, qe = 0.0;
        const int aa = dd_index(a, a, D);
        for (int i = 0; i < C1; ++i) for (int j = 0; j < C1; ++j) { qu += e.Qi[a][i][j] * M3[aa][pair_index(i, j, NZ)]; qe += e.Qi[a][i][j] * M2[aa][pair_index(i, j, NZ)]; }
        suu += Dl[a] * Dl[a] * qu; see += qe;
      }
      UU[a][a] += suu; EE[a][a] += see;
    }
    // back to the original basis: V = UltVeh' (.) UltVeh / n ; B = UltVeh' Bn
    for (int i = 0; i < D; ++i) for (int j = 0; j < D; ++j) {
      double vg = 0.0, ve = 0.0;
      for (int l = 0; l < D; ++l) for (int m = 0; m < D; ++m) { vg += b.UltVeh[l][i] * UU[l][m] * b.UltVeh[m][j]; ve += b.UltVeh[l][i] * EE[l][m] * b.UltVeh[m][j]; }
      fit.V_g[i][j] = vg / (double)dat.n; fit.V_e[i][j] = ve / (double)dat.n;
    }
    for (int i = 0; i < D; ++i) for (int j = 0; j < C1; ++j) { double v = 0.0; for (int l = 0; l < D; ++l) v += b.UltVeh[l][i] * Bn[l][j]; fit.B[i][j] = v; }
  }
  fit.logl = logl_new;
  return logl_new;
}

// chi-square upper tail with D degrees of freedom (gsl_cdf_chisq_Q(x, d), mvlmm.cpp:818)
template <int D>
GB_HD double chisq_Q_int(double x) {
  if (x <= 0.0) return 1.0;
  if (D == 2) return exp(-0.5 * x);
  if (D == 1) return erfc(sqrt(0.5 * x));
  // Q(D/2, x/2) by the finite recurrences over integer / half-integer a
  const double h = 0.5 * x;
  if (D % 2 == 0) { double term = exp(-h), sum = term; for (int k = 1; k < D / 2; ++k) { term *= h / (double)k; sum += term; } return sum; }
  double sum = erfc(sqrt(h)), term = sqrt(h) * exp(-h) / 0.88622692545275801365;      // Gamma(3/2) = sqrt(pi)/2
  sum += term;
  for (int k = 1; k < D / 2; ++k) { term *= h / ((double)k + 0.5); sum += term; }
  return sum;
}

// MphCalcP (mvlmm.cpp:727-831).  The SNP is x row C1-1; the covariates are rows 0..C1-2.  beta: D, Vbeta: D x D.
template <int D, int C1>
GB_HD_BIG double mph_calc_p(const MvData<C1 + D> &dat, const double (&V_g)[D][D], const double (&V_e)[D][D], double (&beta)[D], double (&Vbeta)[D][D]) {
  constexpr int NZ = C1 + D, NP = NZ * (NZ + 1) / 2, C = C1 - 1;
  Basis<D> b;
  eigen_proc<D>(V_g, V_e, b);
  double Dl[D];
  for (int l = 0; l < D; ++l) Dl[l] = b.Dl[l];
  double M1[D][NP], sw[D];
  moments<NZ, D>(dat, [&](double delta, double (&w)[D]) { for (int l = 0; l < D; ++l) w[l] = 1.0 / (delta * Dl[l] + 1.0); }, M1, sw);
  double bl[D], xPy[D], ixPx[D], stat = 0.0;
  for (int l = 0; l < D; ++l) {
    double Q[C > 0 ? C : 1][C > 0 ? C : 1], Qi[C > 0 ? C : 1][C > 0 ? C : 1], ld;
    for (int i = 0; i < C; ++i) for (int j = 0; j < C; ++j) Q[i][j] = M1[l][pair_index(i, j, NZ)];
    if (C > 0) inv_small<(C > 0 ? C : 1)>(Q, Qi, ld);
    double xpy = 0.0, xpx = M1[l][pair_index(C, C, NZ)];
    for (int s = 0; s < D; ++s) xpy += b.UltVehi[l][s] * M1[l][pair_index(C, C1 + s, NZ)];
    for (int i = 0; i < C; ++i) {
      double qx = 0.0;
      for (int j = 0; j < C; ++j) qx += Qi[i][j] * M1[l][pair_index(j, C, NZ)];           // (Qi WHix)_i
      double why = 0.0;
      for (int s = 0; s < D; ++s) why += b.UltVehi[l][s] * M1[l][pair_index(i, C1 + s, NZ)];
      xpx -= M1[l][pair_index(i, C, NZ)] * qx;
      xpy -= qx * why;
    }
    xPy[l] = xpy; ixPx[l] = 1.0 / xpx; bl[l] = xpy / xpx; stat += bl[l] * xpy;
  }
  for (int i = 0; i < D; ++i) {
    double v = 0.0;
    for (int l = 0; l < D; ++l) v += b.UltVeh[l][i] * bl[l];
    beta[i] = v;
    for (int j = 0; j < D; ++j) { double t = 0.0; for (int l = 0; l < D; ++l) t += b.UltVeh[l][i] * ixPx[l] * b.UltVeh[l][j]; Vbeta[i][j] = t; }
  }
  return chisq_Q_int<D>(stat);
}

// GLS B for given (V_g, V_e): MphCalcBeta (mvlmm.cpp:835-937) / the last part of MphInitial (:2882-2935)
template <int D, int C1>
GB_HD_BIG void mph_calc_beta(const MvData<C1 + D> &dat, const double (&V_g)[D][D], const double (&V_e)[D][D], double (&B)[D][C1]) {
  Eval<D, C1> e;
  if (!evaluate<D, C1>(dat, V_g, V_e, true, e)) return;
  for (int i = 0; i < D; ++i) for (int j = 0; j < C1; ++j) { double v = 0.0; for (int l = 0; l < D; ++l) v += e.bs.UltVeh[l][i] * e.Bp[l][j]; B[i][j] = v; }
}

// ---- Newton-Raphson (MphNR, mvlmm.cpp:2608-2760) ----------------------------------------------------------------------------
template <int D>
GB_HD bool is_pd(const double (&V)[D][D]) {
  double ev[D], U[D][D];
  sym_eig<D>(V, ev, U);
  for (int i = 0; i < D; ++i) if (!(ev[i] > 0)) return false;
  return true;
}

constexpr int tri_index(int l, int m, int o, int D) {      // multiset index of (l, m, o), D <= 3
  // sort
  return 0;
}

template <int D, int C1>
struct NrState {
  static constexpr int V = D * (D + 1) / 2;
  double grad[2 * V];
  double Hinv[2 * V][2 * V];
  double logl;
};

// gradient and Hessian of the log (restricted) likelihood at (V_g, V_e): closed block form of CalcDev: grad_t = -1/2 tr(P D_t) + 1/2 y'P D_t P y, Hess_tu = 1/2 tr(P D_t P D_u) - y'P D_t P D_u P y
template <int D, int C1>
GB_HD_BIG bool nr_quantities(bool reml, const MvData<C1 + D> &dat, const double (&V_g)[D][D], const double (&V_e)[D][D], double cst, NrState<D, C1> &st) {
  constexpr int NZ = C1 + D, NP = NZ * (NZ + 1) / 2, DD = D * (D + 1) / 2, V = D * (D + 1) / 2, V2 = 2 * V;
  Eval<D, C1> e;
  if (!evaluate<D, C1>(dat, V_g, V_e, reml, e)) return false;
  st.logl = cst + e.logl;
  const Basis<D> &b = e.bs;
  double Dl[D];
  for (int l = 0; l < D; ++l) Dl[l] = b.Dl[l];
  // second-order moments with delta^p, p = 0, 1, 2 (pairs l <= m)
  double M2[3][DD][NP], s2[3][DD];
  for (int p = 0; p < 3; ++p)
    moments<NZ, DD>(dat, [&](double delta, double (&w)[DD]) {
      double wl[D]; for (int l = 0; l < D; ++l) wl[l] = 1.0 / (delta * Dl[l] + 1.0);
      const double dp = (p == 0) ? 1.0 : (p == 1 ? delta : delta * delta);
      for (int l = 0; l < D; ++l) for (int m = l; m < D; ++m) w[dd_index(l, m, D)] = dp * wl[l] * wl[m]; }, M2[p], s2[p]);
  // third-order moments delta^p w_l w_m w_o for every ordered (l, m, o) (symmetric; computed per multiset and looked up)
  constexpr int D3 = D * D * D;
  double M4[3][D3][NP];
  {
    for (int l = 0; l < D; ++l) for (int m = l; m < D; ++m) for (int o = m; o < D; ++o)
      for (int p = 0; p < 3; ++p) {
        double out[1][NP], sc[1];
        moments<NZ, 1>(dat, [&](double delta, double (&w)[1]) {
          const double dp = (p == 0) ? 1.0 : (p == 1 ? delta : delta * delta);
          w[0] = dp / ((delta * Dl[l] + 1.0) * (delta * Dl[m] + 1.0) * (delta * Dl[o] + 1.0)); }, out, sc);
        const int idx[6][3] = {{l, m, o}, {l, o, m}, {m, l, o}, {m, o, l}, {o, l, m}, {o, m, l}};
        for (int q = 0; q < 6; ++q) { const int t = (idx[q][0] * D + idx[q][1]) * D + idx[q][2]; for (int a = 0; a < NP; ++a) M4[p][t][a] = out[0][a]; }
      }
  }
  // r_l coefficients (B' = Qi xHiy in both REML and ML: the profile over B is the GLS one)
  double co[D][NZ];
  for (int l = 0; l < D; ++l) { for (int j = 0; j < C1; ++j) co[l][j] = -e.Bp[l][j]; for (int s = 0; s < D; ++s) co[l][C1 + s] = b.UltVehi[l][s]; }
  // free elements: t < V -> V_g element (scale delta, power 1), t >= V -> V_e element (power 0); G_t = A E_t A'
  double G[V2][D][D]; int pw[V2];
  {
    int t = 0;
    for (int half = 0; half < 2; ++half)
      for (int i = 0; i < D; ++i) for (int j = i; j < D; ++j) {
        for (int l = 0; l < D; ++l) for (int m = 0; m < D; ++m)
          G[t][l][m] = (i == j) ? b.UltVehi[l][i] * b.UltVehi[m][i] : b.UltVehi[l][i] * b.UltVehi[m][j] + b.UltVehi[l][j] * b.UltVehi[m][i];
        pw[t] = (half == 0) ? 1 : 0;
        ++t;
      }
  }
  // per element: tr(Hi D), tr(Qi A_t), y'PDPy, XHu_t (D x C1), A_t blocks (x-pair matrices per (l,m))
  double XHu[V2][D][C1];
  for (int t = 0; t < V2; ++t) {
    const int p = pw[t];
    double trHiD = 0.0, trQA = 0.0, yPDPy = 0.0;
    for (int l = 0; l < D; ++l) {
      const double swl = (p == 0) ? e.sw[l] : ((Dl[l] != 0.0) ? ((double)dat.n - e.sw[l]) / Dl[l] : 0.0);      // sum_k delta^p w_l
      // for D_l == 0 : sum delta w = sum delta ; handle through the second-order scalars instead
      (void)swl;
    }
    // sum_k delta^p w_l is not a second-order scalar; accumulate it directly (cheap pass)
    double sdw[D];
    for (int l = 0; l < D; ++l) sdw[l] = 0.0;
    for (int k = GB_MV_LANE; k < dat.n; k += GB_MV_NLANE) { const double dlt = dat.delta[k]; for (int l = 0; l < D; ++l) sdw[l] += ((p == 0) ? 1.0 : dlt) / (dlt * Dl[l] + 1.0); }
    for (int l = 0; l < D; ++l) { sdw[l] = gb_mv_allsum(sdw[l]); trHiD += G[t][l][l] * sdw[l]; }
    for (int l = 0; l < D; ++l) for (int m = 0; m < D; ++m) yPDPy += G[t][l][m] * quad<NZ>(M2[p][dd_index(l, m, D)], co[l], co[m]);
    if (reml) for (int l = 0; l < D; ++l) { const int ll = dd_index(l, l, D); double q = 0.0; for (int i = 0; i < C1; ++i) for (int j = 0; j < C1; ++j) q += e.Qi[l][i][j] * M2[p][ll][pair_index(i, j, NZ)]; trQA += G[t][l][l] * q; }
    st.grad[t] = reml ? (-0.5 * (trHiD - trQA) + 0.5 * yPDPy) : (-0.5 * trHiD + 0.5 * yPDPy);
    for (int l = 0; l < D; ++l) for (int j = 0; j < C1; ++j) {
      double v = 0.0;
      for (int m = 0; m < D; ++m) { const double *M = M2[p][dd_index(l, m, D)]; double lin = 0.0; for (int a = 0; a < NZ; ++a) lin += co[m][a] * M[pair_index(j, a, NZ)]; v += G[t][l][m] * lin; }
      XHu[t][l][j] = v;
    }
  }
  double Hess[V2][V2];
  for (int t = 0; t < V2; ++t) for (int r = t; r < V2; ++r) {
    const int p = pw[t] + pw[r];
    double uHu = 0.0, trHH = 0.0;
    for (int l = 0; l < D; ++l) for (int m = 0; m < D; ++m) for (int o = 0; o < D; ++o)
      uHu += G[t][l][m] * G[r][l][o] * quad<NZ>(M4[p][(l * D + m) * D + o], co[m], co[o]);
    double xqx = 0.0;
    for (int l = 0; l < D; ++l) for (int i = 0; i < C1; ++i) for (int j = 0; j < C1; ++j) xqx += XHu[t][l][i] * e.Qi[l][i][j] * XHu[r][l][j];
    const double yPDPDPy = uHu - xqx;
    for (int l = 0; l < D; ++l) for (int m = 0; m < D; ++m) trHH += G[t][l][m] * G[r][m][l] * s2[p][dd_index(l, m, D)];
    double h;
    if (reml) {
      double trQA3 = 0.0, trQAQA = 0.0;
      for (int l = 0; l < D; ++l) for (int m = 0; m < D; ++m) {
        const double gg = G[t][l][m] * G[r][m][l];
        double q = 0.0;
        for (int i = 0; i < C1; ++i) for (int j = 0; j < C1; ++j) q += e.Qi[l][i][j] * M4[p][(l * D + m) * D + l][pair_index(i, j, NZ)];
        trQA3 += gg * q;
        // tr(Qi_l A_t[l,m] Qi_m A_r[m,l]),  A_t[l,m] = G_t[l,m] * X-pair matrix of M2[pw_t][(l,m)]
        const double *Mt = M2[pw[t]][dd_index(l, m, D)], *Mr = M2[pw[r]][dd_index(l, m, D)];
        double tr = 0.0;
        for (int i = 0; i < C1; ++i) for (int j = 0; j < C1; ++j) {
          double left = 0.0, right = 0.0;                    // (Qi_l Xt)_{i j} and (Qi_m Xr)_{j i}
          for (int a = 0; a < C1; ++a) { left += e.Qi[l][i][a] * Mt[pair_index(a, j, NZ)]; right += e.Qi[m][j][a] * Mr[pair_index(a, i, NZ)]; }
          tr += left * right;
        }
        trQAQA += gg * tr;
      }
      h = 0.5 * (trHH - 2.0 * trQA3 + trQAQA) - yPDPDPy;
    } else h = 0.5 * trHH - yPDPDPy;
    Hess[t][r] = h; Hess[r][t] = h;
  }
  double ld;
  return inv_small<V2>(Hess, st.Hinv, ld);
}

template <int D, int C1>
GB_HD_BIG double mph_nr(bool reml, int max_iter, double max_prec, const MvData<C1 + D> &dat, Fit<D, C1> &fit) {
  constexpr int V = D * (D + 1) / 2, V2 = 2 * V;
  const double cst = logl_const<D, C1>(dat, reml);
  double logl_old = 0.0, logl_new = 0.0;
  NrState<D, C1> st, st_new;
  for (int q = 0; q < V2; ++q) { st.grad[q] = 0.0; for (int r = 0; r < V2; ++r) st.Hinv[q][r] = 0.0; }
  for (int t = 0; t < max_iter; ++t) {
    double Vg_save[D][D], Ve_save[D][D];
    for (int i = 0; i < D; ++i) for (int j = 0; j < D; ++j) { Vg_save[i][j] = fit.V_g[i][j]; Ve_save[i][j] = fit.V_e[i][j]; }
    double step_scale = 1.0;
    int step_iter = 0;
    bool flag_pd = false;
    do {
      for (int i = 0; i < D; ++i) for (int j = 0; j < D; ++j) { fit.V_g[i][j] = Vg_save[i][j]; fit.V_e[i][j] = Ve_save[i][j]; }
      if (t != 0) {                                        // UpdateVgVe
        int q = 0;
        for (int i = 0; i < D; ++i) for (int j = i; j < D; ++j, ++q) {
          double sg = 0.0, se = 0.0;
          for (int r = 0; r < V2; ++r) { sg += st.Hinv[q][r] * st.grad[r]; se += st.Hinv[q + V][r] * st.grad[r]; }
          fit.V_g[i][j] = fit.V_g[j][i] = Vg_save[i][j] - step_scale * sg;
          fit.V_e[i][j] = fit.V_e[j][i] = Ve_save[i][j] - step_scale * se;
        }
      }
      flag_pd = is_pd<D>(fit.V_e) && is_pd<D>(fit.V_g);
      if (flag_pd) { if (nr_quantities<D, C1>(reml, dat, fit.V_g, fit.V_e, cst, st_new)) logl_new = st_new.logl; else flag_pd = false; }
      step_scale /= 2.0;
      step_iter++;
    } while ((!flag_pd || logl_new < logl_old || logl_new - logl_old > 10) && step_iter < 10 && t != 0);
    if (t != 0) {
      if (logl_new < logl_old || !flag_pd) {
        for (int i = 0; i < D; ++i) for (int j = 0; j < D; ++j) { fit.V_g[i][j] = Vg_save[i][j]; fit.V_e[i][j] = Ve_save[i][j]; }
        break;
      }
      if (logl_new - logl_old < max_prec) break;
    }
    logl_old = logl_new;
    st = st_new;
  }
  fit.logl = logl_new;
  return logl_new;
}

// per-SNP body of MVLMM::AnalyzeBimbam / AnalyzePlink for -lmm 1/2/3/4 (mvlmm.cpp:3286-3360, crt = 0).  fit enters with the null
// (ML) estimates; out = {beta_1..D, Vbeta upper triangle, p_wald, p_lrt, p_score}.
template <int D, int C1>
GB_HD_BIG void analyze_snp(const MvData<C1 + D> &dat, Fit<D, C1> &fit, int a_mode, int em_iter, double em_prec, int nr_iter, double nr_prec,
                       double p_nr, double logl_mle_H0, double *out) {
  double beta[D], Vb[D][D], p_wald = 0.0, p_lrt = 0.0, p_score = 0.0;
  for (int i = 0; i < D; ++i) { beta[i] = 0.0; for (int j = 0; j < D; ++j) Vb[i][j] = 0.0; }
  if (a_mode == 3 || a_mode == 4) p_score = mph_calc_p<D, C1>(dat, fit.V_g, fit.V_e, beta, Vb);               // at the null estimates
  if (a_mode == 2 || a_mode == 4) {
    double logl_H1 = mph_em<D, C1>(false, em_iter / 10, em_prec * 10.0, dat, fit);
    mph_calc_p<D, C1>(dat, fit.V_g, fit.V_e, beta, Vb);
    p_lrt = chisq_Q_int<D>(2.0 * (logl_H1 - logl_mle_H0));
    if (p_lrt < p_nr) {
      logl_H1 = mph_nr<D, C1>(false, nr_iter / 10, nr_prec * 10.0, dat, fit);
      mph_calc_p<D, C1>(dat, fit.V_g, fit.V_e, beta, Vb);
      p_lrt = chisq_Q_int<D>(2.0 * (logl_H1 - logl_mle_H0));
    }
  }
  if (a_mode == 1 || a_mode == 4) {
    mph_em<D, C1>(true, em_iter / 10, em_prec * 10.0, dat, fit);
    p_wald = mph_calc_p<D, C1>(dat, fit.V_g, fit.V_e, beta, Vb);
    if (p_wald < p_nr) {
      mph_nr<D, C1>(true, nr_iter / 10, nr_prec * 10.0, dat, fit);
      p_wald = mph_calc_p<D, C1>(dat, fit.V_g, fit.V_e, beta, Vb);
    }
  }
  int o = 0;
  for (int i = 0; i < D; ++i) out[o++] = beta[i];
  for (int i = 0; i < D; ++i) for (int j = i; j < D; ++j) out[o++] = Vb[i][j];
  out[o++] = p_wald; out[o++] = p_lrt; out[o++] = p_score;
}

}  // namespace gbmv
