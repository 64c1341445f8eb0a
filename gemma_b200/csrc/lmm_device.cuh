// lmm_device.cuh -- fused per-SNP LMM evaluator (device side).
//
// One WARP owns one SNP.  Every likelihood evaluation is a single coalesced pass over
// the n rotated individuals: lane t reads elements t, t+32, ... of the eigenvalue
// vector, the rotated covariates / phenotype and this SNP's U^T x column, forms
// h = 1/(lambda*delta+1) and accumulates ALL weighted cross products
//   S^k_ab = sum_i h_i^k v_a,i v_b,i     (k = 1..3, a<=b over {w_1..w_c, x, y})
// in registers; a shuffle-xor butterfly leaves bit-identical totals in every lane,
// so the scalar control flow below (grid scan, Brent, Newton, Wald/score/LRT) is
// executed redundantly and divergence-free by all 32 lanes.
//
// What it replaces in the reference (paths relative to the GEMMA tree):
//   CalcUab            src/lmm.cpp:1213-1280   (never materialised: products formed in registers)
//   CalcPab/PPab/PPPab src/lmm.cpp:283-482     (row 0 = the register sums, rows 1.. = in-place sweeps)
//   LogL_* / LogRL_*   src/lmm.cpp:484-1125
//   CalcLambda         src/lmm.cpp:1945-2140   (grid + GSL Brent + GSL Newton, same control flow)
//   CalcRLWald/Score   src/lmm.cpp:1127-1211
//   batch_compute body src/lmm.cpp:1526-1562
// The reference performs ~72 passes over n per SNP for -lmm 4 (each pass itself ~20 GSL
// vector sweeps); sharing sums between REML and ML at the common grid lambdas, re-using
// grid values as Brent end points and folding f(l_min) into the first grid pass brings
// this to ~35 passes with identical arithmetic results (same lambdas, same formulas).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <float.h>
#include "../../include/gemma_b200.h"

namespace gb {

struct LmmConst {
  int n;                 // analysed individuals
  int n_c;               // n rounded up to the pipeline chunk (512); every vector is zero-padded to n_c
  int ldv;               // leading dimension of Wt rows (== n_c)
  const double *delta;   // eigenvalues (n)
  const double *Wt;      // rotated covariates, TRANSPOSED: n_cvt rows of n (coalesced per covariate)
  const double *y;       // rotated phenotype (n)
  // common-lambda tables of the lockstep kernel (lmm_v2.cuh "hoisted" passes); null = not available
  const double *Hrows;   // n_common rows of n_c doubles: h_i = 1/(lambda_j delta_i + 1)
  const double *ctab;    // n_common records: SNP-independent sums at lambda_j (v2c_stride doubles each)
  int n_common;
  // Chebyshev tables of the interpolated root refinement (lmm_v2.cuh): null = off.  Hrows / ctab then hold, after the n_common
  // shared rows, cheb_M node rows per grid interval; cheb = [cos table (4 M) | per interval: coefficients of the SNP-independent sums]
  const double *cheb;
  double cheb_marg;      // half-width added to every interval in log(lambda)
  // Exact x-sums at lambda* = l_mle_null for this batch (int8 projection only; null = off): xex[s * (n_cvt + 1) + q] =
  // sum_i h_i(lambda*) (U^T x_s)_i q_i for q over (w_1..w_c, y), formed in GENOTYPE space as x_s . v_q with v_q = U (h (.) q) -- an
  // FP64 dot product with no digit-plane rounding.  The lockstep kernel adds (exact - projected) to the order-1 x-sums of the score
  // test and of the final f / Wald evaluation: the plane rounding of U^T x then reaches beta only in second order.
  const double *xex;
  // Exact LINEAR x-sums of this batch at every hoisted lambda (null = off): row s holds, for block b (the shared rows 0..n_common-1,
  // then the x-node rows) and power k = 1, 2: sum_i h_i^k (U^T x)_i q_i for q over (w_1..w_c, y) at [(b * 2 + (k - 1)) * (c + 1) + q],
  // and behind the last block the unit-weight sums.  Formed as x . v with v = U (h^k (.) q) by a side GEMM in genotype space (int8
  // digit planes of the few hundred v's): independent of the rounding of the projected U^T x, which then only feeds the sums
  // quadratic in x -- where independent rounding noise averages out -- so that the projection itself can run on 3 planes.
  const double *xsum;
  int xsum_ld, xsum_nblocks;   // row stride in doubles; number of (lambda) blocks
  int xsum_nblocks_skip;       // rows between the shared rows and the x-node rows (the table-node rows, which carry no x-sums)
  const double *xcov;    // G x E: covariate column xcov_idx is this per-SNP vector (U^T x) instead of a row of Wt; null otherwise
  int xcov_idx;
  unsigned long long *cnt;  // optional work counters of the lockstep kernel (gb200_lmm_counters); null = off
  int nc_gen;            // generic-covariate path (NC < 0 instantiations): number of swept covariates
  int gen_stride;        // doubles of shared-memory scratch per warp on that path (3 tables of (nc_gen+3)(nc_gen+2)/2)
};

struct LmmParams {
  int a_mode;
  int n_region;
  double l_min, l_max;
  double l_mle_null, logl_mle_H0;
  int plink_rule;        // 1: AnalyzePlink semantics (src/lmm.cpp:1866-1884): no Wald when the REML search failed,
                         //    p_wald = p_lrt = NaN when the reported logl_H1 is NaN; 0: Analyze (BIMBAM) semantics
};

__host__ __device__ constexpr int abidx(int a, int b, int nv) {
  // 0-based a<=b version of GetabIndex (src/param.cpp:1400-1415)
  return (2 * nv - a + 1) * a / 2 + (b - a);
}

__device__ __forceinline__ double warp_allsum(double v) {
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
  return v;
}

// ---------------------------------------------------------------------------------
// special functions: gsl_cdf_fdist_Q / gsl_cdf_chisq_Q (GSL cdf/fdist.c, cdf/beta_inc.c,
// cdf/gamma.c; call sites src/lmm.cpp:1161,1206,1553)
__device__ inline double stirling_tail(double z) {
  double zi = 1.0 / z, zi2 = zi * zi;
  return zi * (1.0 / 12.0 - zi2 * (1.0 / 360.0 - zi2 * (1.0 / 1260.0 - zi2 * (1.0 / 1680.0 - zi2 * (1.0 / 1188.0)))));
}
__device__ inline double ln_beta_dev(double a, double b) {
  double big = a > b ? a : b, small = a > b ? b : a;
  if (big >= 10.0) {
    double s = big + small;
    double d = (big - 0.5) * (-log1p(small / big)) - small * log(s) + small +
               (stirling_tail(big) - stirling_tail(s));
    return lgamma(small) + d;
  }
  return lgamma(a) + lgamma(b) - lgamma(a + b);
}
__device__ inline double beta_cont_frac_dev(double a, double b, double x, double epsabs) {
  const unsigned max_iter = 512;
  const double cutoff = 2.0 * DBL_MIN;
  unsigned iter = 0;
  double num_term = 1.0;
  double den_term = 1.0 - (a + b) * x / (a + 1.0);
  if (fabs(den_term) < cutoff) den_term = nan("");
  den_term = 1.0 / den_term;
  double cf = den_term;
  while (iter < max_iter) {
    const int k = (int)iter + 1;
    double coeff = k * (b - k) * x / (((a - 1.0) + 2 * k) * (a + 2 * k));
    double delta_frac;
    den_term = 1.0 + coeff * den_term;
    num_term = 1.0 + coeff / num_term;
    if (fabs(den_term) < cutoff) den_term = nan("");
    if (fabs(num_term) < cutoff) num_term = nan("");
    den_term = 1.0 / den_term;
    delta_frac = den_term * num_term;
    cf *= delta_frac;
    coeff = -(a + k) * (a + b + k) * x / ((a + 2 * k) * (a + 2 * k + 1.0));
    den_term = 1.0 + coeff * den_term;
    num_term = 1.0 + coeff / num_term;
    if (fabs(den_term) < cutoff) den_term = nan("");
    if (fabs(num_term) < cutoff) num_term = nan("");
    den_term = 1.0 / den_term;
    delta_frac = den_term * num_term;
    cf *= delta_frac;
    if (fabs(delta_frac - 1.0) < 2.0 * DBL_EPSILON) break;
    if (cf * fabs(delta_frac - 1.0) < epsabs) break;
    ++iter;
  }
  if (iter >= max_iter) return nan("");
  return cf;
}
__device__ inline double beta_inc_AXPY_dev(double A, double Y, double a, double b, double x) {
  if (x == 0.0) return A * 0 + Y;
  if (x == 1.0) return A * 1 + Y;
  // GSL's asymptotic branches [A&S 26.5.17] for a or b > 1e5 (df/2 > 1e5: more than 2e5 individuals); gsl_sf_gamma_inc_P/Q at the
  // only first argument this path produces (nu1 = 1 -> 1/2): P(1/2, z) = erf(sqrt z), Q(1/2, z) = erfc(sqrt z)
  if (a > 1e5 && b < 10 && x > a / (a + b) && b == 0.5) {
    const double N = a + (b - 1.0) / 2.0;
    return A * erfc(sqrt(-N * log(x))) + Y;
  }
  if (b > 1e5 && a < 10 && x < b / (a + b) && a == 0.5) {
    const double N = b + (a - 1.0) / 2.0;
    return A * erf(sqrt(-N * log1p(-x))) + Y;
  }
  double lnb = ln_beta_dev(a, b);
  double ln_pre = -lnb + a * log(x) + b * log1p(-x);
  double prefactor = exp(ln_pre);
  if (x < (a + 1.0) / (a + b + 2.0)) {
    double epsabs = fabs(Y / (A * prefactor / a)) * DBL_EPSILON;
    double cf = beta_cont_frac_dev(a, b, x, epsabs);
    return A * (prefactor * cf / a) + Y;
  } else {
    double epsabs = fabs((A + Y) / (A * prefactor / b)) * DBL_EPSILON;
    double cf = beta_cont_frac_dev(b, a, 1.0 - x, epsabs);
    double term = prefactor * cf / b;
    if (A == -Y) return -A * term;
    return A * (1 - term) + Y;
  }
}
__device__ inline double fdist_Q_dev(double x, double nu1, double nu2) {
  double r = nu2 / nu1;
  if (x < r) {
    double u = x / (r + x);
    return beta_inc_AXPY_dev(-1.0, 1.0, nu1 / 2.0, nu2 / 2.0, u);
  } else {
    double u = r / (r + x);
    return beta_inc_AXPY_dev(1.0, 0.0, nu2 / 2.0, nu1 / 2.0, u);
  }
}
__device__ inline double chisq1_Q_dev(double x) {
  if (x <= 0.0) return 1.0;
  return erfc(sqrt(0.5 * x));
}
// src/mathfunc.cpp:122-131: every d < 0.001 becomes |d| (misplaced parenthesis kept for parity)
__device__ inline double safe_sqrt_dev(double d) {
  double d1 = d;
  if (d < 0.001) d1 = fabs(d);
  if (d1 < 0.0) return nan("");
  return sqrt(d1);
}

// ---------------------------------------------------------------------------------
// One pass over the n individuals for powers KLO..KHI of h (power 0 = unit weights, the
// "Iab" table of LogRL_f, src/lmm.cpp:838-843).
template <int NC, int KLO, int KHI, bool LOGDET>
struct PassOut {
  static constexpr int NV = NC + 2;
  static constexpr int NIDX = (NC + 3) * (NC + 2) / 2;
  static constexpr int NK = KHI - KLO + 1;
  double S[NK][NIDX];
  double tr[NK];        // sum_i h_i^k
  double logdet;        // sum_i log|lambda*delta_i + 1|
};

template <int NC, int KLO, int KHI, bool LOGDET>
__device__ __forceinline__ void lmm_pass(const LmmConst &D, const double *__restrict__ x, double lam,
                                         PassOut<NC, KLO, KHI, LOGDET> &o) {
  constexpr int NV = NC + 2;
  constexpr int NIDX = (NC + 3) * (NC + 2) / 2;
  constexpr int NK = KHI - KLO + 1;
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < NK; ++k) {
    o.tr[k] = 0.0;
#pragma unroll
    for (int j = 0; j < NIDX; ++j) o.S[k][j] = 0.0;
  }
  o.logdet = 0.0;
  const double *__restrict__ dl = D.delta;
  const double *__restrict__ yy = D.y;
  const double *__restrict__ Wt = D.Wt;
  const int n = D.n;
#pragma unroll 2
  for (int i = lane; i < n; i += 32) {
    double v[NV];
#pragma unroll
    for (int a = 0; a < NC; ++a) v[a] = __ldg(Wt + (size_t)a * D.ldv + i);
    v[NC] = __ldg(x + i);
    v[NC + 1] = __ldg(yy + i);
    const double den = fma(lam, __ldg(dl + i), 1.0);
    const double h = 1.0 / den;
    if (LOGDET) o.logdet += log(fabs(den));
    double hk = (KLO == 0) ? 1.0 : h;
#pragma unroll
    for (int k = 0; k < NK; ++k) {
      o.tr[k] += hk;
#pragma unroll
      for (int a = 0; a < NV; ++a) {
        const double t = hk * v[a];
#pragma unroll
        for (int b = a; b < NV; ++b) o.S[k][abidx(a, b, NV)] = fma(t, v[b], o.S[k][abidx(a, b, NV)]);
      }
      hk *= h;
    }
  }
#pragma unroll
  for (int k = 0; k < NK; ++k) {
    o.tr[k] = warp_allsum(o.tr[k]);
#pragma unroll
    for (int j = 0; j < NIDX; ++j) o.S[k][j] = warp_allsum(o.S[k][j]);
  }
  if (LOGDET) o.logdet = warp_allsum(o.logdet);
}

// ---------------------------------------------------------------------------------
// Derived quantities.  The tables P/PP/PPP start as row 0 of CalcPab/CalcPPab/CalcPPPab
// (the weighted sums) and are swept IN PLACE one variable at a time: after sweeping
// variables 0..p-1 the surviving entries equal row p of the reference's tables
// (src/lmm.cpp:342-345, :399-405, :462-472).  Entries that involve the pivot variable
// are read-only during its sweep, so in-place is exact.
template <int NC, int ORD>
struct Derived {
  // after sweeping the NC covariates (row n_cvt of Pab)
  double P_xx, P_xy, P_yy;
  // after sweeping x as well (row n_cvt+1)
  double Px_yy, PPx_yy, PPPx_yy;
  double trace_P_corr;    // sum_{i<=NC} PP_i[ii]/P_i[ii]
  double trace_PP_corr;   // sum_{i<=NC} (PP_i[ii]/P_i[ii])^2 - 2 PPP_i[ii]/P_i[ii]
  double logdet_piv;      // sum_{i<=NC} log P_i[ii]
};

template <int NC, int ORD>
__device__ __forceinline__ void sweep_tables(double (&P)[(NC + 3) * (NC + 2) / 2],
                                             double (&PP)[(NC + 3) * (NC + 2) / 2],
                                             double (&PPP)[(NC + 3) * (NC + 2) / 2],
                                             Derived<NC, ORD> &d) {
  constexpr int NV = NC + 2;
  d.trace_P_corr = 0.0; d.trace_PP_corr = 0.0; d.logdet_piv = 0.0;
#pragma unroll
  for (int p = 0; p <= NC; ++p) {
    if (p == NC) {
      d.P_xx = P[abidx(NC, NC, NV)];
      d.P_xy = P[abidx(NC, NC + 1, NV)];
      d.P_yy = P[abidx(NC + 1, NC + 1, NV)];
    }
    const double ww = P[abidx(p, p, NV)];
    const double ww2 = (ORD >= 2) ? PP[abidx(p, p, NV)] : 0.0;
    const double ww3 = (ORD >= 3) ? PPP[abidx(p, p, NV)] : 0.0;
    d.logdet_piv += log(ww);
    if (ORD >= 2) {
      const double r = ww2 / ww;
      d.trace_P_corr += r;
      if (ORD >= 3) d.trace_PP_corr += r * r - 2.0 * ww3 / ww;
    }
    if (ww != 0) {
      const double iw = 1.0 / ww;
#pragma unroll
      for (int a = p + 1; a < NV; ++a) {
        const double aw = P[abidx(p, a, NV)];
        const double aw2 = (ORD >= 2) ? PP[abidx(p, a, NV)] : 0.0;
        const double aw3 = (ORD >= 3) ? PPP[abidx(p, a, NV)] : 0.0;
#pragma unroll
        for (int b = a; b < NV; ++b) {
          const double bw = P[abidx(p, b, NV)];
          const double bw2 = (ORD >= 2) ? PP[abidx(p, b, NV)] : 0.0;
          const double bw3 = (ORD >= 3) ? PPP[abidx(p, b, NV)] : 0.0;
          const int ab = abidx(a, b, NV);
          if (ORD >= 3) {
            double p3 = PPP[ab] - aw * bw * ww2 * ww2 * (iw * iw * iw);
            p3 -= (aw * bw3 + bw * aw3 + aw2 * bw2) * iw;
            p3 += (aw * bw2 * ww2 + bw * aw2 * ww2 + aw * bw * ww3) * (iw * iw);
            PPP[ab] = p3;
          }
          if (ORD >= 2) {
            double p2 = PP[ab] + aw * bw * ww2 * (iw * iw);
            p2 -= (aw * bw2 + bw * aw2) * iw;
            PP[ab] = p2;
          }
          P[ab] = P[ab] - aw * bw * iw;
        }
      }
    }
  }
  d.Px_yy = P[abidx(NC + 1, NC + 1, NV)];
  d.PPx_yy = (ORD >= 2) ? PP[abidx(NC + 1, NC + 1, NV)] : 0.0;
  d.PPPx_yy = (ORD >= 3) ? PPP[abidx(NC + 1, NC + 1, NV)] : 0.0;
}

// ---------------------------------------------------------------------------------
// Any number of covariates (instantiations with NC < 0; the count is D.nc_gen at run time).  The register-resident
// tables above grow as (c+2)(c+3)/2 per power of h and stop fitting at c = 6; here the weighted sums are accumulated in
// 4 x 4 register blocks of (a, b) pairs -- one pass over the individuals per block, the columns re-read from L1/L2 --
// and the totals live in a per-warp shared-memory table that is swept lane-parallel (one (a, b) entry per lane and
// step).  Every sum is accumulated in the same order and with the same fma sequence as lmm_pass, so for c <= 6 this
// path is bit-identical to the templated one (tests force it to check exactly that).
__device__ __forceinline__ double *gen_tables(const LmmConst &D) {
  extern __shared__ __align__(16) double gb_gen_smem[];
  return gb_gen_smem + (size_t)(threadIdx.x >> 5) * (size_t)D.gen_stride;
}
__device__ __forceinline__ int gen_nidx(int nc) { return (nc + 3) * (nc + 2) / 2; }

template <int KLO, int KHI, bool LOGDET>
__device__ __noinline__ void gen_pass(const LmmConst &D, const double *__restrict__ x, double lam, double *__restrict__ S,
                                      double (&tr)[KHI - KLO + 1], double &logdet) {
  constexpr int NK = KHI - KLO + 1;
  const int nc = D.nc_gen, NV = nc + 2, NIDX = gen_nidx(nc);
  const int lane = threadIdx.x & 31, n = D.n;
  const double *__restrict__ dl = D.delta;
#pragma unroll
  for (int k = 0; k < NK; ++k) tr[k] = 0.0;
  logdet = 0.0;
  for (int a0 = 0; a0 < NV; a0 += 4) {
    const double *ca[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) { const int a = a0 + r; ca[r] = (a < nc) ? ((D.xcov && a == D.xcov_idx) ? D.xcov : D.Wt + (size_t)a * D.ldv) : (a == nc ? x : D.y); }
    for (int b0 = a0; b0 < NV; b0 += 4) {
      const double *cb[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) { const int b = b0 + r; cb[r] = (b < nc) ? ((D.xcov && b == D.xcov_idx) ? D.xcov : D.Wt + (size_t)b * D.ldv) : (b == nc ? x : D.y); }
      const bool first = (a0 == 0 && b0 == 0);
      double acc[NK][4][4];
#pragma unroll
      for (int k = 0; k < NK; ++k)
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[k][r][q] = 0.0;
      for (int i = lane; i < n; i += 32) {
        double va[4], vb[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) { va[r] = __ldg(ca[r] + i); vb[r] = __ldg(cb[r] + i); }
        const double den = fma(lam, __ldg(dl + i), 1.0);
        const double h = 1.0 / den;
        if (LOGDET && first) logdet += log(fabs(den));
        double hk = (KLO == 0) ? 1.0 : h;
#pragma unroll
        for (int k = 0; k < NK; ++k) {
          if (first) tr[k] += hk;
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const double t = hk * va[r];
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[k][r][q] = fma(t, vb[q], acc[k][r][q]);
          }
          hk *= h;
        }
      }
#pragma unroll
      for (int k = 0; k < NK; ++k)
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const double tot = warp_allsum(acc[k][r][q]);
            const int a = a0 + r, b = b0 + q;
            if (a < NV && b < NV && a <= b && lane == 0) S[k * NIDX + abidx(a, b, NV)] = tot;
          }
    }
  }
#pragma unroll
  for (int k = 0; k < NK; ++k) tr[k] = warp_allsum(tr[k]);
  if (LOGDET) logdet = warp_allsum(logdet);
  __syncwarp();
}

// lane-parallel form of sweep_tables on shared-memory tables (same per-entry arithmetic)
template <int ORD>
__device__ __noinline__ void gen_sweep(double *__restrict__ P, double *__restrict__ PP, double *__restrict__ PPP, int nc,
                                       Derived<-1, ORD> &d) {
  const int NV = nc + 2, lane = threadIdx.x & 31;
  d.trace_P_corr = 0.0; d.trace_PP_corr = 0.0; d.logdet_piv = 0.0;
  d.P_xx = d.P_xy = d.P_yy = 0.0;
  for (int p = 0; p <= nc; ++p) {
    if (p == nc) {
      d.P_xx = P[abidx(nc, nc, NV)];
      d.P_xy = P[abidx(nc, nc + 1, NV)];
      d.P_yy = P[abidx(nc + 1, nc + 1, NV)];
    }
    const double ww = P[abidx(p, p, NV)];
    const double ww2 = (ORD >= 2) ? PP[abidx(p, p, NV)] : 0.0;
    const double ww3 = (ORD >= 3) ? PPP[abidx(p, p, NV)] : 0.0;
    d.logdet_piv += log(ww);
    if (ORD >= 2) {
      const double r = ww2 / ww;
      d.trace_P_corr += r;
      if (ORD >= 3) d.trace_PP_corr += r * r - 2.0 * ww3 / ww;
    }
    if (ww != 0) {
      const double iw = 1.0 / ww;
      for (int a = p + 1; a < NV; ++a) {
        const double aw = P[abidx(p, a, NV)];
        const double aw2 = (ORD >= 2) ? PP[abidx(p, a, NV)] : 0.0;
        const double aw3 = (ORD >= 3) ? PPP[abidx(p, a, NV)] : 0.0;
        for (int b = a + lane; b < NV; b += 32) {
          const double bw = P[abidx(p, b, NV)];
          const double bw2 = (ORD >= 2) ? PP[abidx(p, b, NV)] : 0.0;
          const double bw3 = (ORD >= 3) ? PPP[abidx(p, b, NV)] : 0.0;
          const int ab = abidx(a, b, NV);
          if (ORD >= 3) {
            double p3 = PPP[ab] - aw * bw * ww2 * ww2 * (iw * iw * iw);
            p3 -= (aw * bw3 + bw * aw3 + aw2 * bw2) * iw;
            p3 += (aw * bw2 * ww2 + bw * aw2 * ww2 + aw * bw * ww3) * (iw * iw);
            PPP[ab] = p3;
          }
          if (ORD >= 2) {
            double p2 = PP[ab] + aw * bw * ww2 * (iw * iw);
            p2 -= (aw * bw2 + bw * aw2) * iw;
            PP[ab] = p2;
          }
          P[ab] = P[ab] - aw * bw * iw;
        }
      }
    }
    __syncwarp();
  }
  d.Px_yy = P[abidx(nc + 1, nc + 1, NV)];
  d.PPx_yy = (ORD >= 2) ? PP[abidx(nc + 1, nc + 1, NV)] : 0.0;
  d.PPPx_yy = (ORD >= 3) ? PPP[abidx(nc + 1, nc + 1, NV)] : 0.0;
  __syncwarp();
}

struct DevVals { double d1R, d2R, d1L, d2L; };

// dev1 (and dev2 when ORD==3) of the REML and ML log-likelihoods from one pass.
// LogRL_dev1/dev2/dev12 src/lmm.cpp:866-1125 ; LogL_dev1/dev2/dev12 src/lmm.cpp:544-797
template <int NC, int KHI>
__device__ __forceinline__ DevVals eval_devs(const LmmConst &D, const double *x, double lam) {
  constexpr int NIDX = (NC + 3) * (NC + 2) / 2;
  Derived<NC, KHI> d;
  double tr0, tr1 = 0.0;
  int ncv = NC;
  if constexpr (NC < 0) {
    ncv = D.nc_gen;
    double *T = gen_tables(D);
    const int gi = gen_nidx(ncv);
    double tr[KHI], ld;
    gen_pass<1, KHI, false>(D, x, lam, T, tr, ld);
    gen_sweep<KHI>(T, T + gi, T + 2 * gi, ncv, d);
    tr0 = tr[0]; if (KHI >= 2) tr1 = tr[KHI >= 2 ? 1 : 0];
  } else {
    PassOut<NC, 1, KHI, false> o;
    lmm_pass<NC, 1, KHI, false>(D, x, lam, o);
    double dummy[NIDX];
    if constexpr (KHI == 3) sweep_tables<NC, KHI>(o.S[0], o.S[1], o.S[2], d);
    else sweep_tables<NC, KHI>(o.S[0], o.S[1], dummy, d);
    tr0 = o.tr[0]; if (KHI >= 2) tr1 = o.tr[KHI >= 2 ? 1 : 0];
  }
  const double n = (double)D.n;
  const double df = n - (double)ncv - 1.0;
  const double P_yy = d.Px_yy, PP_yy = d.PPx_yy;
  const double yPKPy = (P_yy - PP_yy) / lam;
  const double trace_Hi = tr0;
  const double trace_P = trace_Hi - d.trace_P_corr;
  DevVals r;
  r.d1R = -0.5 * ((df - trace_P) / lam) + 0.5 * df * yPKPy / P_yy;
  r.d1L = -0.5 * ((n - trace_Hi) / lam) + 0.5 * n * yPKPy / P_yy;
  r.d2R = 0.0; r.d2L = 0.0;
  if (KHI == 3) {
    const double trace_HiHi = tr1;
    const double trace_PP = trace_HiHi + d.trace_PP_corr;
    const double PPP_yy = d.PPPx_yy;
    const double yPKPKPy = (P_yy + PPP_yy - 2.0 * PP_yy) / (lam * lam);
    const double quad = (2.0 * yPKPKPy * P_yy - yPKPy * yPKPy) / (P_yy * P_yy);
    r.d2R = 0.5 * ((df + trace_PP - 2.0 * trace_P) / (lam * lam)) - 0.5 * df * quad;
    r.d2L = 0.5 * ((n + trace_HiHi - 2.0 * trace_Hi) / (lam * lam)) - 0.5 * n * quad;
  }
  return r;
}

struct FVals { double fR, fL; };

// LogRL_f src/lmm.cpp:799-864, LogL_f src/lmm.cpp:484-542.  logdetI = sum_{i<=NC} log I_i[ii]
// (unit-weight pivots) is lambda-independent and passed in.
__device__ __forceinline__ FVals f_from(double n, int nc, double logdet_h, double logdet_piv,
                                        double logdetI, double P_yy) {
  const double df = n - (double)nc - 1.0;
  if (P_yy >= 0.0 && P_yy < 1e-8) P_yy = 1e-8;      // P_YY_MIN, src/lmm.cpp:52,527,854
  const double lp = log(P_yy);
  const double l2pi = 1.8378770664093453;            // log(2*pi)
  FVals r;
  r.fR = 0.5 * df * (log(df) - l2pi - 1.0) - 0.5 * logdet_h - 0.5 * (logdet_piv - logdetI) - 0.5 * df * lp;
  r.fL = 0.5 * n * (log(n) - l2pi - 1.0) - 0.5 * logdet_h - 0.5 * n * lp;
  return r;
}

template <int NC>
__device__ __forceinline__ FVals eval_f(const LmmConst &D, const double *x, double lam, double logdetI,
                                        Derived<NC, 1> *dout = nullptr) {
  constexpr int NIDX = (NC + 3) * (NC + 2) / 2;
  Derived<NC, 1> d;
  if constexpr (NC < 0) {
    double *T = gen_tables(D);
    double tr[1], ld;
    gen_pass<1, 1, true>(D, x, lam, T, tr, ld);
    gen_sweep<1>(T, T, T, D.nc_gen, d);
    if (dout) *dout = d;
    return f_from((double)D.n, D.nc_gen, ld, d.logdet_piv, logdetI, d.Px_yy);
  } else {
    PassOut<NC, 1, 1, true> o;
    lmm_pass<NC, 1, 1, true>(D, x, lam, o);
    double dm1[NIDX], dm2[NIDX];
    sweep_tables<NC, 1>(o.S[0], dm1, dm2, d);
    if (dout) *dout = d;
    return f_from((double)D.n, NC, o.logdet, d.logdet_piv, logdetI, d.Px_yy);
  }
}

// CalcRLWald src/lmm.cpp:1127-1167 / CalcRLScore :1170-1211 from a swept order-1 table.
template <int NC>
__device__ __forceinline__ void wald_score_from(const Derived<NC, 1> &d, int n, bool score,
                                                double &beta, double &se, double &pval, int nc_run = NC) {
  const int df = n - nc_run - 1;
  beta = d.P_xy / d.P_xx;
  const double tau = (double)df / d.Px_yy;
  se = safe_sqrt_dev(1.0 / (tau * d.P_xx));
  if (score) pval = fdist_Q_dev((double)n * d.P_xy * d.P_xy / (d.P_yy * d.P_xx), 1.0, (double)df);
  else pval = fdist_Q_dev((d.P_yy - d.Px_yy) * tau, 1.0, (double)df);
}

template <int NC>
__device__ __forceinline__ void eval_wald_score(const LmmConst &D, const double *x, double lam, bool score,
                                                double &beta, double &se, double &pval) {
  constexpr int NIDX = (NC + 3) * (NC + 2) / 2;
  Derived<NC, 1> d;
  if constexpr (NC < 0) {
    double *T = gen_tables(D);
    double tr[1], ld;
    gen_pass<1, 1, false>(D, x, lam, T, tr, ld);
    gen_sweep<1>(T, T, T, D.nc_gen, d);
    wald_score_from<NC>(d, D.n, score, beta, se, pval, D.nc_gen);
  } else {
    PassOut<NC, 1, 1, false> o;
    lmm_pass<NC, 1, 1, false>(D, x, lam, o);
    double dm1[NIDX], dm2[NIDX];
    sweep_tables<NC, 1>(o.S[0], dm1, dm2, d);
    wald_score_from<NC>(d, D.n, score, beta, se, pval);
  }
}

// ---------------------------------------------------------------------------------
// GSL root finders restated (roots/brent.c, roots/newton.c, roots/convergence.c) with the
// control flow of CalcLambda (src/lmm.cpp:2024-2116).  `fnR` selects REML vs ML values.
#define GB_ST_SUCCESS 0
#define GB_ST_CONTINUE (-2)
#define GB_ST_ERR 1

struct RootState {
  double l, l_temp;      // persist across intervals like the reference's locals (:2003)
  double lambda, logf;   // best so far
  bool have;             // at least one interval processed (i == 0 test of :2109)
  bool aborted;          // NaN return of :2087-2094
  bool stopped;          // Brent hit max_iter -> `break` out of the interval loop (:2057-2060)
};

template <int NC>
__device__ __noinline__ void refine_interval(const LmmConst &D, const double *x, bool fnR, double x_lower,
                                             double x_upper, double f_lower, double f_upper,
                                             double l_min, double l_max, double logdetI, RootState &rs) {
  // --- gsl_root_fsolver_set: brent_init (end-point values are the grid values, bit-identical lambdas)
  double a = x_lower, fa = f_lower, b = x_upper, fb = f_upper, c = x_upper, fc = f_upper;
  double d = x_upper - x_lower, e = x_upper - x_lower;
  double root = 0.5 * (x_lower + x_upper), xl = x_lower, xu = x_upper;
  int status = GB_ST_ERR;
  unsigned iter = 0;
  const unsigned max_iter = 100;
  do {
    iter++;
    // --- brent_iterate
    {
      double tol, m;
      bool ac_equal = false;
      if ((fb < 0 && fc < 0) || (fb > 0 && fc > 0)) { ac_equal = true; c = a; fc = fa; d = b - a; e = b - a; }
      if (fabs(fc) < fabs(fb)) { ac_equal = true; a = b; b = c; c = a; fa = fb; fb = fc; fc = fa; }
      tol = 0.5 * DBL_EPSILON * fabs(b);
      m = 0.5 * (c - b);
      if (fb == 0) { root = b; xl = b; xu = b; status = GB_ST_SUCCESS; }
      else if (fabs(m) <= tol) {
        root = b;
        if (b < c) { xl = b; xu = c; } else { xl = c; xu = b; }
        status = GB_ST_SUCCESS;
      } else {
        if (fabs(e) < tol || fabs(fa) <= fabs(fb)) { d = m; e = m; }
        else {
          double p, q, r, s = fb / fa;
          if (ac_equal) { p = 2 * m * s; q = 1 - s; }
          else {
            q = fa / fc; r = fb / fc;
            p = s * (2 * m * q * (q - r) - (b - a) * (r - 1));
            q = (q - 1) * (r - 1) * (s - 1);
          }
          if (p > 0) q = -q; else p = -p;
          const double lim1 = 3 * m * q - fabs(tol * q), lim2 = fabs(e * q);
          if (2 * p < (lim1 < lim2 ? lim1 : lim2)) { e = d; d = p / q; }
          else { d = m; e = m; }
        }
        // note: on a non-finite value GSL returns before storing the state; we keep locals
        const double a_new = b, fa_new = fb;
        double b_new = b;
        if (fabs(d) > tol) b_new += d; else b_new += (m > 0 ? +tol : -tol);
        DevVals dv = eval_devs<NC, 2>(D, x, b_new);
        const double fb_new = fnR ? dv.d1R : dv.d1L;
        if (!isfinite(fb_new)) { status = GB_ST_ERR; }
        else {
          a = a_new; fa = fa_new; b = b_new; fb = fb_new;
          root = b;
          double cc = c;
          if ((fb < 0 && fc < 0) || (fb > 0 && fc > 0)) cc = a;
          if (b < cc) { xl = b; xu = cc; } else { xl = cc; xu = b; }
          status = GB_ST_SUCCESS;
        }
      }
    }
    if (status != GB_ST_SUCCESS) break;
    rs.l = root;
    // --- gsl_root_test_interval(xl, xu, 0, 0.1)
    {
      if (xl > xu) { status = GB_ST_ERR; break; }
      double min_abs = 0.0;
      if ((xl > 0.0 && xu > 0.0) || (xl < 0.0 && xu < 0.0)) min_abs = fmin(fabs(xl), fabs(xu));
      status = (fabs(xu - xl) < 0.1 * min_abs) ? GB_ST_SUCCESS : GB_ST_CONTINUE;
    }
  } while (status == GB_ST_CONTINUE && iter < max_iter);
  if (status == GB_ST_CONTINUE) { rs.stopped = true; return; }

  // --- Newton (gsl_root_fdfsolver_newton) from Brent's root
  double nroot = rs.l, nf, ndf;
  {
    DevVals dv = eval_devs<NC, 3>(D, x, nroot);
    nf = fnR ? dv.d1R : dv.d1L; ndf = fnR ? dv.d2R : dv.d2L;
  }
  unsigned iter2 = 0;
  do {
    iter2++;
    if (ndf == 0.0) { status = GB_ST_ERR; break; }
    nroot = nroot - (nf / ndf);
    {
      DevVals dv = eval_devs<NC, 3>(D, x, nroot);
      nf = fnR ? dv.d1R : dv.d1L; ndf = fnR ? dv.d2R : dv.d2L;
    }
    if (!isfinite(nf) || !isfinite(ndf)) { status = GB_ST_ERR; break; }
    rs.l_temp = rs.l;
    rs.l = nroot;
    // gsl_root_test_delta(l, l_temp, 0, 1e-5)
    status = (fabs(rs.l - rs.l_temp) < 1e-5 * fabs(rs.l) || rs.l == rs.l_temp) ? GB_ST_SUCCESS : GB_ST_CONTINUE;
  } while (status == GB_ST_CONTINUE && iter2 < max_iter && rs.l > l_min && rs.l < l_max);
  if (status != GB_ST_SUCCESS) {
    rs.aborted = true; rs.lambda = nan(""); rs.logf = nan("");
    return;
  }
  double l = rs.l_temp;                    // the PREVIOUS iterate (src/lmm.cpp:2096)
  rs.l = l;
  if (l < l_min) l = l_min;
  if (l > l_max) l = l_max;
  rs.l = l;
  FVals fv = eval_f<NC>(D, x, l, logdetI);
  const double logf_l = fnR ? fv.fR : fv.fL;
  if (!rs.have) { rs.logf = logf_l; rs.lambda = l; rs.have = true; }
  else if (rs.logf < logf_l) { rs.logf = logf_l; rs.lambda = l; }
}

// CalcLambda for REML and/or ML in one interleaved sweep of the lambda grid
// (src/lmm.cpp:1945-2140).  On return R/L hold lambda/logf exactly as two independent
// reference calls would produce them.
template <int NC>
__device__ __forceinline__ void calc_lambda_both(const LmmConst &D, const double *x, double l_min,
                                                 double l_max, int n_region, bool needR, bool needL,
                                                 RootState &R, RootState &L) {
  constexpr int NIDX = (NC + 3) * (NC + 2) / 2;
  const double lambda_interval = log(l_max / l_min) / (double)n_region;
  R.l = R.l_temp = 0.0; R.lambda = nan(""); R.logf = nan(""); R.have = R.aborted = R.stopped = false;
  L = R;
  // first grid point lambda = l_min*exp(0) == l_min: powers 0..2 + logdet in one pass gives
  // dev1 (both), f(l_min) (both) and the unit-weight pivots of LogRL_f's Iab.
  double logdetI, fRmin, fLmin, dR_prev, dL_prev;
  double lam_l = l_min * exp(lambda_interval * 0.0);
  {
    Derived<NC, 2> d;
    double tr1v, logdet_h;
    int ncv = NC;
    if constexpr (NC < 0) {
      ncv = D.nc_gen;
      double *T = gen_tables(D);
      const int gi = gen_nidx(ncv);
      double tr[3];
      gen_pass<0, 2, true>(D, x, lam_l, T, tr, logdet_h);
      Derived<NC, 1> dI;
      gen_sweep<1>(T, T, T, ncv, dI);
      logdetI = dI.logdet_piv;
      gen_sweep<2>(T + gi, T + 2 * gi, T, ncv, d);
      tr1v = tr[1];
    } else {
      PassOut<NC, 0, 2, true> o;
      lmm_pass<NC, 0, 2, true>(D, x, lam_l, o);
      {
        Derived<NC, 1> dI;
        double dm1[NIDX], dm2[NIDX];
        sweep_tables<NC, 1>(o.S[0], dm1, dm2, dI);
        logdetI = dI.logdet_piv;
      }
      double dm3[NIDX];
      sweep_tables<NC, 2>(o.S[1], o.S[2], dm3, d);
      tr1v = o.tr[1]; logdet_h = o.logdet;
    }
    const double n = (double)D.n, df = n - (double)ncv - 1.0;
    const double yPKPy = (d.Px_yy - d.PPx_yy) / lam_l;
    const double trace_P = tr1v - d.trace_P_corr;
    dR_prev = -0.5 * ((df - trace_P) / lam_l) + 0.5 * df * yPKPy / d.Px_yy;
    dL_prev = -0.5 * ((n - tr1v) / lam_l) + 0.5 * n * yPKPy / d.Px_yy;
    FVals fv = f_from(n, ncv, logdet_h, d.logdet_piv, logdetI, d.Px_yy);
    fRmin = fv.fR; fLmin = fv.fL;
  }
  for (int i = 0; i < n_region; ++i) {
    const double lam_h = l_min * exp(lambda_interval * (i + 1.0));
    DevVals dv = eval_devs<NC, 2>(D, x, lam_h);
    if (needR && !R.aborted && !R.stopped && dR_prev * dv.d1R <= 0)
      refine_interval<NC>(D, x, true, lam_l, lam_h, dR_prev, dv.d1R, l_min, l_max, logdetI, R);
    if (needL && !L.aborted && !L.stopped && dL_prev * dv.d1L <= 0)
      refine_interval<NC>(D, x, false, lam_l, lam_h, dL_prev, dv.d1L, l_min, l_max, logdetI, L);
    dR_prev = dv.d1R; dL_prev = dv.d1L; lam_l = lam_h;
  }
  FVals fmax = eval_f<NC>(D, x, l_max, logdetI);
  // end-point comparison (:1985-2000 when no interval, :2121-2136 otherwise)
  if (needR && !R.aborted) {
    if (!R.have && !R.stopped) {
      if (fRmin >= fmax.fR) { R.lambda = l_min; R.logf = fRmin; } else { R.lambda = l_max; R.logf = fmax.fR; }
    } else {
      if (fRmin > R.logf) { R.lambda = l_min; R.logf = fRmin; }
      if (fmax.fR > R.logf) { R.lambda = l_max; R.logf = fmax.fR; }
    }
  }
  if (needL && !L.aborted) {
    if (!L.have && !L.stopped) {
      if (fLmin >= fmax.fL) { L.lambda = l_min; L.logf = fLmin; } else { L.lambda = l_max; L.logf = fmax.fL; }
    } else {
      if (fLmin > L.logf) { L.lambda = l_min; L.logf = fLmin; }
      if (fmax.fL > L.logf) { L.lambda = l_max; L.logf = fmax.fL; }
    }
  }
}

// Whole per-SNP analysis (batch_compute body, src/lmm.cpp:1526-1562).  x = this SNP's U^T x.
template <int NC>
__device__ __forceinline__ void analyze_snp(const LmmConst &D, const LmmParams &prm, const double *x,
                                            gb200_sumstat &out) {
  const int mode = prm.a_mode;
  const bool needR = (mode == 1 || mode == 4);
  const bool needL = (mode == 2 || mode == 4 || mode == 9);
  const bool needS = (mode == 3 || mode == 4 || mode == 9);
  double lambda_mle = 0.0, lambda_remle = 0.0, beta = 0.0, se = 0.0, p_wald = 0.0;
  double p_lrt = 0.0, p_score = 0.0, logl_H1 = 0.0;

  if (needS) eval_wald_score<NC>(D, x, prm.l_mle_null, true, beta, se, p_score);   // "3 is before 1"

  if (needR || needL) {
    RootState R, L;
    calc_lambda_both<NC>(D, x, prm.l_min, prm.l_max, prm.n_region, needR, needL, R, L);
    if (needR) {
      lambda_remle = R.lambda; logl_H1 = R.logf;
      if (!(prm.plink_rule && isnan(logl_H1))) eval_wald_score<NC>(D, x, lambda_remle, false, beta, se, p_wald);
    }
    if (needL) {
      lambda_mle = L.lambda; logl_H1 = L.logf;
      p_lrt = chisq1_Q_dev(2.0 * (logl_H1 - prm.logl_mle_H0));
    }
    if (prm.plink_rule && isnan(logl_H1)) { p_wald = logl_H1; p_lrt = logl_H1; }
  }
  out.beta = beta; out.se = se; out.lambda_remle = lambda_remle; out.lambda_mle = lambda_mle;
  out.p_wald = p_wald; out.p_lrt = p_lrt; out.p_score = p_score; out.logl_H1 = logl_H1;
}

// Null model (src/gemma.cpp:2711-2753; CalcLambda(func,eval,UtW,Uty) src/lmm.cpp:2143-2180,
// CalcPve :2183-2205, the P_yy part of CalcLmmVgVeBeta :2268-2271).  The null model with c
// covariates (calc_null: nc_total = c, df = n-c) is algebraically the alternative model with
// c-1 covariates whose "x" is the last covariate, so the same evaluator is reused with
// NC = c-1 and x = Wt row c-1.  S1 sums at both lambdas are returned for the c x c solve
// of beta on the host.
struct NullOut {
  double l_mle, logl_mle, l_remle, logl_remle;
  double dev2_remle;                 // LogRL_dev2 at l_remle (CalcPve)
  double Pyy_mle, Pyy_remle;         // row n_cvt P_yy at the two lambdas
  double S1_mle[(GB200_MAX_CVT + 2) * (GB200_MAX_CVT + 1) / 2];
  double S1_remle[(GB200_MAX_CVT + 2) * (GB200_MAX_CVT + 1) / 2];
};

template <int NC>
__device__ __forceinline__ void null_model(const LmmConst &D, const double *x, double l_min, double l_max,
                                           int n_region, NullOut &out) {
  constexpr int NIDX = (NC + 3) * (NC + 2) / 2;
  RootState R, L;
  calc_lambda_both<NC>(D, x, l_min, l_max, n_region, true, true, R, L);
  out.l_mle = L.lambda; out.logl_mle = L.logf; out.l_remle = R.lambda; out.logl_remle = R.logf;
  DevVals dv = eval_devs<NC, 3>(D, x, R.lambda);
  out.dev2_remle = dv.d2R;
  for (int which = 0; which < 2; ++which) {
    const double lam = which == 0 ? L.lambda : R.lambda;
    double *dst = which == 0 ? out.S1_mle : out.S1_remle;
    Derived<NC, 1> d;
    if constexpr (NC < 0) {
      double *T = gen_tables(D);
      double tr[1], ld;
      gen_pass<1, 1, false>(D, x, lam, T, tr, ld);
      const int gi = gen_nidx(D.nc_gen);
      for (int j = 0; j < gi; ++j) dst[j] = T[j];
      __syncwarp();
      gen_sweep<1>(T, T, T, D.nc_gen, d);
    } else {
      PassOut<NC, 1, 1, false> o;
      lmm_pass<NC, 1, 1, false>(D, x, lam, o);
#pragma unroll
      for (int j = 0; j < NIDX; ++j) dst[j] = o.S[0][j];
      double dm1[NIDX], dm2[NIDX];
      sweep_tables<NC, 1>(o.S[0], dm1, dm2, d);
    }
    if (which == 0) out.Pyy_mle = d.Px_yy; else out.Pyy_remle = d.Px_yy;
  }
}

}  // namespace gb
