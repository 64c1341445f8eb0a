/*
 * gemma_oracle.c -- CPU restatement of GEMMA's univariate-LMM hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT THE PRODUCT.  Only tests/, the smoke check in
 * __graft_entry__.py and bench.py's cpu_baseline / --impl reference legs may
 * load this file's shared object.  The product path (libgemma_b200.so) never
 * links or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks this restatement
 * against the reference's documented outputs: example/demo.txt:9-12,31-36,41-42
 * (mouse_hs1940 kinship block, first five -lmm 1 rows, pve / se(pve)) and the cell
 * pins of test/dev_tests.rb:42-43,53-54 (BXD -lmm 2 / -lmm 9 with covariates).
 *
 * Every function cites the reference file:line (relative to /root/reference) whose
 * behaviour it restates.  GSL / LAPACK pieces are not vendored in the reference
 * (system libgsl 2.x, OpenBLAS): their published algorithms are restated here
 * (roots/brent.c, roots/newton.c, roots/convergence.c, cdf/fdist.c, cdf/beta_inc.c,
 * cdf/gamma.c) and anchored on the reference's call sites src/lmm.cpp:2024-2078,
 * :1161, :1206, :1553.
 *
 * Layout conventions follow the reference: matrices are row-major with an
 * explicit leading dimension; Uab is n x n_index with the (a,b) product vector in
 * column GetabIndex(a,b) (stride n_index), exactly like src/lmm.cpp:1213-1280.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#include <stdint.h>

#define GO_EXPORT __attribute__((visibility("default")))

/* src/lmm.cpp:52 */
#define P_YY_MIN 1e-8

/* ------------------------------------------------------------------------- */
/* src/param.cpp:1400-1415  GetabIndex: packed upper-triangle index, 1-based a,b */
GO_EXPORT size_t go_getab_index(size_t a, size_t b, size_t n_cvt) {
  size_t cols = n_cvt + 2;
  size_t a1 = a, b1 = b;
  if (b <= a) { a1 = b; b1 = a; }
  return (2 * cols - a1 + 2) * (a1 - 1) / 2 + b1 - a1;
}

/* ------------------------------------------------------------------------- */
/* Special functions used by gsl_cdf_fdist_Q / gsl_cdf_chisq_Q (external GSL,
 * restated: cdf/beta_inc.c beta_cont_frac + beta_inc_AXPY, cdf/fdist.c,
 * cdf/gamma.c).  ln B(a,b) is evaluated with a Stirling-difference for large
 * arguments so that the prefactor keeps ~1e-14 relative accuracy at a ~ 25000. */

static double stirling_tail(double z) {
  /* lgamma(z) - [(z-0.5)ln z - z + 0.5 ln 2pi], asymptotic series, z >= 10 */
  double zi = 1.0 / z, zi2 = zi * zi;
  return zi * (1.0 / 12.0 - zi2 * (1.0 / 360.0 - zi2 * (1.0 / 1260.0 - zi2 * (1.0 / 1680.0 - zi2 * (1.0 / 1188.0)))));
}

static double ln_beta(double a, double b) {
  /* ln B(a,b) = lgamma(a) + lgamma(b) - lgamma(a+b) */
  double big = a > b ? a : b, small = a > b ? b : a;
  if (big >= 10.0) {
    /* lgamma(big) - lgamma(big+small) without cancellation */
    double s = big + small;
    double d = (big - 0.5) * (-log1p(small / big)) - small * log(s) + small
               + (stirling_tail(big) - stirling_tail(s));
    return lgamma(small) + d;
  }
  return lgamma(a) + lgamma(b) - lgamma(a + b);
}

static double beta_cont_frac(double a, double b, double x, double epsabs) {
  /* modified Lentz evaluation, GSL cdf/beta_inc.c */
  const unsigned max_iter = 512;
  const double cutoff = 2.0 * DBL_MIN;
  unsigned iter = 0;
  double num_term = 1.0;
  double den_term = 1.0 - (a + b) * x / (a + 1.0);
  double cf;
  if (fabs(den_term) < cutoff) den_term = NAN;
  den_term = 1.0 / den_term;
  cf = den_term;
  while (iter < max_iter) {
    const int k = (int)iter + 1;
    double coeff = k * (b - k) * x / (((a - 1.0) + 2 * k) * (a + 2 * k));
    double delta_frac;
    den_term = 1.0 + coeff * den_term;
    num_term = 1.0 + coeff / num_term;
    if (fabs(den_term) < cutoff) den_term = NAN;
    if (fabs(num_term) < cutoff) num_term = NAN;
    den_term = 1.0 / den_term;
    delta_frac = den_term * num_term;
    cf *= delta_frac;
    coeff = -(a + k) * (a + b + k) * x / ((a + 2 * k) * (a + 2 * k + 1.0));
    den_term = 1.0 + coeff * den_term;
    num_term = 1.0 + coeff / num_term;
    if (fabs(den_term) < cutoff) den_term = NAN;
    if (fabs(num_term) < cutoff) num_term = NAN;
    den_term = 1.0 / den_term;
    delta_frac = den_term * num_term;
    cf *= delta_frac;
    if (fabs(delta_frac - 1.0) < 2.0 * DBL_EPSILON) break;
    if (cf * fabs(delta_frac - 1.0) < epsabs) break;
    ++iter;
  }
  if (iter >= max_iter) return NAN;
  return cf;
}

static double beta_inc_AXPY(double A, double Y, double a, double b, double x) {
  /* A * I_x(a,b) + Y ; GSL cdf/beta_inc.c, including its two asymptotic branches [Abramowitz-Stegun 26.5.17] for
   * a or b > 1e5 (more than 2e5 individuals: df/2 > 1e5 at the call sites src/lmm.cpp:1161,1206).  They go through
   * gsl_sf_gamma_inc_Q / _P (specfunc/gamma_inc.c, external GSL), restated below for the only first argument the LMM
   * path produces (nu1 = 1 -> 1/2): P(1/2, z) = erf(sqrt z), Q(1/2, z) = erfc(sqrt z).  Other first arguments keep
   * the continued fraction (never reached from src/lmm.cpp). */
  if (x == 0.0) return A * 0 + Y;
  if (x == 1.0) return A * 1 + Y;
  if (a > 1e5 && b < 10 && x > a / (a + b) && b == 0.5) {
    double N = a + (b - 1.0) / 2.0;
    return A * erfc(sqrt(-N * log(x))) + Y;
  }
  if (b > 1e5 && a < 10 && x < b / (a + b) && a == 0.5) {
    double N = b + (a - 1.0) / 2.0;
    return A * erf(sqrt(-N * log1p(-x))) + Y;
  }
  {
    double lnb = ln_beta(a, b);
    double ln_pre = -lnb + a * log(x) + b * log1p(-x);
    double prefactor = exp(ln_pre);
    if (x < (a + 1.0) / (a + b + 2.0)) {
      double epsabs = fabs(Y / (A * prefactor / a)) * DBL_EPSILON;
      double cf = beta_cont_frac(a, b, x, epsabs);
      return A * (prefactor * cf / a) + Y;
    } else {
      double epsabs = fabs((A + Y) / (A * prefactor / b)) * DBL_EPSILON;
      double cf = beta_cont_frac(b, a, 1.0 - x, epsabs);
      double term = prefactor * cf / b;
      if (A == -Y) return -A * term;
      return A * (1 - term) + Y;
    }
  }
}

/* gsl_cdf_fdist_Q(x, nu1, nu2)  (GSL cdf/fdist.c); call sites src/lmm.cpp:1161,1206 */
GO_EXPORT double go_cdf_fdist_Q(double x, double nu1, double nu2) {
  double r = nu2 / nu1;
  if (x < r) {
    double u = x / (r + x);
    return beta_inc_AXPY(-1.0, 1.0, nu1 / 2.0, nu2 / 2.0, u);
  } else {
    double u = r / (r + x);
    return beta_inc_AXPY(1.0, 0.0, nu2 / 2.0, nu1 / 2.0, u);
  }
}

/* gsl_cdf_chisq_Q(x, 1) = gamma_inc_Q(1/2, x/2) = erfc(sqrt(x/2)); 1 for x<=0
 * (GSL cdf/gamma.c); call site src/lmm.cpp:1553 */
GO_EXPORT double go_cdf_chisq1_Q(double x) {
  if (x <= 0.0) return 1.0;     /* NaN falls through to erfc(NaN)=NaN */
  return erfc(sqrt(0.5 * x));
}

/* src/mathfunc.cpp:122-131 safe_sqrt: the misplaced parenthesis makes every
 * d < 0.001 (all negatives included) become |d| */
static double safe_sqrt(double d) {
  double d1 = d;
  if (fabs((double)(d < 0.001))) d1 = fabs(d);
  if (d1 < 0.0) return NAN;
  return sqrt(d1);
}

/* ------------------------------------------------------------------------- */
/* FUNC_PARAM  (src/lmm.h:35-45) */
typedef struct {
  int calc_null;
  size_t ni_test;
  size_t n_cvt;
  const double *eval;   /* n */
  const double *Uab;    /* n x n_index row-major */
  size_t e_mode;        /* always 0 on this path */
  /* scratch, allocated once per go_* call instead of per evaluation */
  double *Hi, *HiHi, *HiHiHi, *vtmp;
  double *Pab, *PPab, *PPPab, *Iab;
  long n_eval;          /* evaluation counter (diagnostic) */
} func_param;

static size_t n_index_of(size_t n_cvt) { return (n_cvt + 3) * (n_cvt + 2) / 2; }

static double strided_dot(const double *w, const double *col, size_t n, size_t stride) {
  /* gsl_blas_ddot(Hi_eval, Uab_col) with Uab_col stride n_index (src/lmm.cpp:317-318) */
  double s = 0.0;
  for (size_t i = 0; i < n; ++i) s += w[i] * col[i * stride];
  return s;
}

/* src/lmm.cpp:283-357 CalcPab */
static void calc_pab(size_t n_cvt, const double *Hi, const double *Uab, size_t n, double *Pab) {
  size_t n_index = n_index_of(n_cvt);
  for (size_t p = 0; p <= n_cvt + 1; ++p) {
    for (size_t a = p + 1; a <= n_cvt + 2; ++a) {
      for (size_t b = a; b <= n_cvt + 2; ++b) {
        size_t index_ab = go_getab_index(a, b, n_cvt);
        double p_ab;
        if (p == 0) {
          p_ab = strided_dot(Hi, Uab + index_ab, n, n_index);
        } else {
          size_t index_aw = go_getab_index(a, p, n_cvt);
          size_t index_bw = go_getab_index(b, p, n_cvt);
          size_t index_ww = go_getab_index(p, p, n_cvt);
          double ps_ab = Pab[(p - 1) * n_index + index_ab];
          double ps_aw = Pab[(p - 1) * n_index + index_aw];
          double ps_bw = Pab[(p - 1) * n_index + index_bw];
          double ps_ww = Pab[(p - 1) * n_index + index_ww];
          if (ps_ww != 0) p_ab = ps_ab - ps_aw * ps_bw / ps_ww;
          else p_ab = ps_ab;
        }
        Pab[p * n_index + index_ab] = p_ab;
      }
    }
  }
}

/* src/lmm.cpp:359-416 CalcPPab */
static void calc_ppab(size_t n_cvt, const double *HiHi, const double *Uab, size_t n,
                      const double *Pab, double *PPab) {
  size_t n_index = n_index_of(n_cvt);
  for (size_t p = 0; p <= n_cvt + 1; ++p) {
    for (size_t a = p + 1; a <= n_cvt + 2; ++a) {
      for (size_t b = a; b <= n_cvt + 2; ++b) {
        size_t index_ab = go_getab_index(a, b, n_cvt);
        double p2_ab;
        if (p == 0) {
          p2_ab = strided_dot(HiHi, Uab + index_ab, n, n_index);
        } else {
          size_t index_aw = go_getab_index(a, p, n_cvt);
          size_t index_bw = go_getab_index(b, p, n_cvt);
          size_t index_ww = go_getab_index(p, p, n_cvt);
          const double *P = Pab + (p - 1) * n_index, *PP = PPab + (p - 1) * n_index;
          double ps2_ab = PP[index_ab];
          double ps_aw = P[index_aw], ps_bw = P[index_bw], ps_ww = P[index_ww];
          double ps2_aw = PP[index_aw], ps2_bw = PP[index_bw], ps2_ww = PP[index_ww];
          if (ps_ww != 0) {
            p2_ab = ps2_ab + ps_aw * ps_bw * ps2_ww / (ps_ww * ps_ww);
            p2_ab -= (ps_aw * ps2_bw + ps_bw * ps2_aw) / ps_ww;
          } else {
            p2_ab = ps2_ab;
          }
        }
        PPab[p * n_index + index_ab] = p2_ab;
      }
    }
  }
}

/* src/lmm.cpp:418-482 CalcPPPab */
static void calc_pppab(size_t n_cvt, const double *HiHiHi, const double *Uab, size_t n,
                       const double *Pab, const double *PPab, double *PPPab) {
  size_t n_index = n_index_of(n_cvt);
  for (size_t p = 0; p <= n_cvt + 1; ++p) {
    for (size_t a = p + 1; a <= n_cvt + 2; ++a) {
      for (size_t b = a; b <= n_cvt + 2; ++b) {
        size_t index_ab = go_getab_index(a, b, n_cvt);
        double p3_ab;
        if (p == 0) {
          p3_ab = strided_dot(HiHiHi, Uab + index_ab, n, n_index);
        } else {
          size_t index_aw = go_getab_index(a, p, n_cvt);
          size_t index_bw = go_getab_index(b, p, n_cvt);
          size_t index_ww = go_getab_index(p, p, n_cvt);
          const double *P = Pab + (p - 1) * n_index, *PP = PPab + (p - 1) * n_index,
                       *PPP = PPPab + (p - 1) * n_index;
          double ps3_ab = PPP[index_ab];
          double ps_aw = P[index_aw], ps_bw = P[index_bw], ps_ww = P[index_ww];
          double ps2_aw = PP[index_aw], ps2_bw = PP[index_bw], ps2_ww = PP[index_ww];
          double ps3_aw = PPP[index_aw], ps3_bw = PPP[index_bw], ps3_ww = PPP[index_ww];
          if (ps_ww != 0) {
            p3_ab = ps3_ab - ps_aw * ps_bw * ps2_ww * ps2_ww / (ps_ww * ps_ww * ps_ww);
            p3_ab -= (ps_aw * ps3_bw + ps_bw * ps3_aw + ps2_aw * ps2_bw) / ps_ww;
            p3_ab += (ps_aw * ps2_bw * ps2_ww + ps_bw * ps2_aw * ps2_ww + ps_aw * ps_bw * ps3_ww) /
                     (ps_ww * ps_ww);
          } else {
            p3_ab = ps3_ab;
          }
        }
        PPPab[p * n_index + index_ab] = p3_ab;
      }
    }
  }
}

/* Hi = 1/(l*eval+1) and its powers; v = l*eval+1 kept in vtmp
 * (the gsl_vector_* prologue of every LogL/LogRL function, e.g. src/lmm.cpp:885-906) */
static void fill_hi(func_param *p, double l, int order) {
  size_t n = p->ni_test;
  for (size_t i = 0; i < n; ++i) {
    double v = p->eval[i] * l;
    v += 1.0;
    p->vtmp[i] = v;
    p->Hi[i] = 1.0 / v;
  }
  if (order >= 2) for (size_t i = 0; i < n; ++i) p->HiHi[i] = p->Hi[i] * p->Hi[i];
  if (order >= 3) for (size_t i = 0; i < n; ++i) p->HiHiHi[i] = p->HiHi[i] * p->Hi[i];
  p->n_eval++;
}

static double vsum(const double *v, size_t n) {
  double s = 0.0;
  for (size_t i = 0; i < n; ++i) s += v[i];
  return s;
}

/* src/lmm.cpp:484-542 LogL_f */
static double LogL_f(double l, func_param *p) {
  size_t n_cvt = p->n_cvt, n = p->ni_test, n_index = n_index_of(n_cvt);
  size_t nc_total = p->calc_null ? n_cvt : n_cvt + 1;
  double logdet_h = 0.0;
  fill_hi(p, l, 1);
  for (size_t i = 0; i < n; ++i) logdet_h += log(fabs(p->vtmp[i]));
  calc_pab(n_cvt, p->Hi, p->Uab, n, p->Pab);
  double c = 0.5 * (double)n * (log((double)n) - log(2 * M_PI) - 1.0);
  size_t index_yy = go_getab_index(n_cvt + 2, n_cvt + 2, n_cvt);
  double P_yy = p->Pab[nc_total * n_index + index_yy];
  if (P_yy >= 0.0 && P_yy < P_YY_MIN) P_yy = P_YY_MIN;
  return c - 0.5 * logdet_h - 0.5 * (double)n * log(P_yy);
}

/* src/lmm.cpp:544-640 LogL_dev1, :642-717 LogL_dev2, :719-797 LogL_dev12 */
static void LogL_dev12(double l, func_param *p, double *dev1, double *dev2, int want2) {
  size_t n_cvt = p->n_cvt, n = p->ni_test, n_index = n_index_of(n_cvt);
  size_t nc_total = p->calc_null ? n_cvt : n_cvt + 1;
  fill_hi(p, l, want2 ? 3 : 2);
  double trace_Hi = vsum(p->Hi, n);
  calc_pab(n_cvt, p->Hi, p->Uab, n, p->Pab);
  calc_ppab(n_cvt, p->HiHi, p->Uab, n, p->Pab, p->PPab);
  size_t index_yy = go_getab_index(n_cvt + 2, n_cvt + 2, n_cvt);
  double P_yy = p->Pab[nc_total * n_index + index_yy];
  double PP_yy = p->PPab[nc_total * n_index + index_yy];
  double yPKPy = (P_yy - PP_yy) / l;
  double trace_HiK = ((double)n - trace_Hi) / l;
  if (dev1) *dev1 = -0.5 * trace_HiK + 0.5 * (double)n * yPKPy / P_yy;
  if (want2) {
    double trace_HiHi = vsum(p->HiHi, n);
    calc_pppab(n_cvt, p->HiHiHi, p->Uab, n, p->Pab, p->PPab, p->PPPab);
    double trace_HiKHiK = ((double)n + trace_HiHi - 2 * trace_Hi) / (l * l);
    double PPP_yy = p->PPPab[nc_total * n_index + index_yy];
    double yPKPKPy = (P_yy + PPP_yy - 2.0 * PP_yy) / (l * l);
    *dev2 = 0.5 * trace_HiKHiK -
            0.5 * (double)n * (2.0 * yPKPKPy * P_yy - yPKPy * yPKPy) / (P_yy * P_yy);
  }
}

/* src/lmm.cpp:799-864 LogRL_f */
static double LogRL_f(double l, func_param *p) {
  size_t n_cvt = p->n_cvt, n = p->ni_test, n_index = n_index_of(n_cvt);
  size_t nc_total;
  double df;
  if (p->calc_null) { nc_total = n_cvt; df = (double)n - (double)n_cvt; }
  else { nc_total = n_cvt + 1; df = (double)n - (double)n_cvt - 1.0; }
  double logdet_h = 0.0, logdet_hiw = 0.0;
  fill_hi(p, l, 1);
  for (size_t i = 0; i < n; ++i) logdet_h += log(fabs(p->vtmp[i]));
  calc_pab(n_cvt, p->Hi, p->Uab, n, p->Pab);
  for (size_t i = 0; i < n; ++i) p->vtmp[i] = 1.0;
  calc_pab(n_cvt, p->vtmp, p->Uab, n, p->Iab);
  for (size_t i = 0; i < nc_total; ++i) {
    size_t index_ww = go_getab_index(i + 1, i + 1, n_cvt);
    logdet_hiw += log(p->Pab[i * n_index + index_ww]);
    logdet_hiw -= log(p->Iab[i * n_index + index_ww]);
  }
  size_t index_yy = go_getab_index(n_cvt + 2, n_cvt + 2, n_cvt);
  double P_yy = p->Pab[nc_total * n_index + index_yy];
  if (P_yy >= 0.0 && P_yy < P_YY_MIN) P_yy = P_YY_MIN;
  double c = 0.5 * df * (log(df) - log(2 * M_PI) - 1.0);
  return c - 0.5 * logdet_h - 0.5 * logdet_hiw - 0.5 * df * log(P_yy);
}

/* src/lmm.cpp:866-943 LogRL_dev1, :945-1033 LogRL_dev2, :1035-1125 LogRL_dev12 */
static void LogRL_dev12(double l, func_param *p, double *dev1, double *dev2, int want2) {
  size_t n_cvt = p->n_cvt, n = p->ni_test, n_index = n_index_of(n_cvt);
  size_t nc_total;
  double df;
  if (p->calc_null) { nc_total = n_cvt; df = (double)n - (double)n_cvt; }
  else { nc_total = n_cvt + 1; df = (double)n - (double)n_cvt - 1.0; }
  fill_hi(p, l, want2 ? 3 : 2);
  double trace_Hi = vsum(p->Hi, n);
  double trace_HiHi = want2 ? vsum(p->HiHi, n) : 0.0;
  calc_pab(n_cvt, p->Hi, p->Uab, n, p->Pab);
  calc_ppab(n_cvt, p->HiHi, p->Uab, n, p->Pab, p->PPab);
  if (want2) calc_pppab(n_cvt, p->HiHiHi, p->Uab, n, p->Pab, p->PPab, p->PPPab);
  double trace_P = trace_Hi, trace_PP = trace_HiHi;
  for (size_t i = 0; i < nc_total; ++i) {
    size_t index_ww = go_getab_index(i + 1, i + 1, n_cvt);
    double ps_ww = p->Pab[i * n_index + index_ww];
    double ps2_ww = p->PPab[i * n_index + index_ww];
    trace_P -= ps2_ww / ps_ww;
    if (want2) {
      double ps3_ww = p->PPPab[i * n_index + index_ww];
      trace_PP += ps2_ww * ps2_ww / (ps_ww * ps_ww) - 2.0 * ps3_ww / ps_ww;
    }
  }
  size_t index_yy = go_getab_index(n_cvt + 2, n_cvt + 2, n_cvt);
  double P_yy = p->Pab[nc_total * n_index + index_yy];
  double PP_yy = p->PPab[nc_total * n_index + index_yy];
  double yPKPy = (P_yy - PP_yy) / l;
  double trace_PK = (df - trace_P) / l;
  if (dev1) *dev1 = -0.5 * trace_PK + 0.5 * df * yPKPy / P_yy;
  if (want2) {
    double trace_PKPK = (df + trace_PP - 2.0 * trace_P) / (l * l);
    double PPP_yy = p->PPPab[nc_total * n_index + index_yy];
    double yPKPKPy = (P_yy + PPP_yy - 2.0 * PP_yy) / (l * l);
    *dev2 = 0.5 * trace_PKPK -
            0.5 * df * (2.0 * yPKPKPy * P_yy - yPKPy * yPKPy) / (P_yy * P_yy);
  }
}

static double eval_dev1(char fn, double l, func_param *p) {
  double d1;
  if (fn == 'R') LogRL_dev12(l, p, &d1, NULL, 0); else LogL_dev12(l, p, &d1, NULL, 0);
  return d1;
}
static void eval_dev12(char fn, double l, func_param *p, double *d1, double *d2) {
  if (fn == 'R') LogRL_dev12(l, p, d1, d2, 1); else LogL_dev12(l, p, d1, d2, 1);
}
static double eval_f(char fn, double l, func_param *p) {
  return fn == 'R' ? LogRL_f(l, p) : LogL_f(l, p);
}

/* ------------------------------------------------------------------------- */
/* GSL root finders, restated (external: roots/brent.c, roots/newton.c,
 * roots/convergence.c).  Status codes: 0 success, -2 continue, >0 error. */
#define ST_SUCCESS 0
#define ST_CONTINUE (-2)
#define ST_EBADFUNC 9
#define ST_EZERODIV 12
#define ST_EINVAL 4

typedef struct { double a, b, c, d, e, fa, fb, fc; double root, x_lower, x_upper; } brent_state;

static int brent_set(brent_state *s, char fn, func_param *p, double x_lower, double x_upper) {
  double f_lower, f_upper;
  s->root = 0.5 * (x_lower + x_upper);
  s->x_lower = x_lower; s->x_upper = x_upper;
  f_lower = eval_dev1(fn, x_lower, p);
  if (!isfinite(f_lower)) return ST_EBADFUNC;
  f_upper = eval_dev1(fn, x_upper, p);
  if (!isfinite(f_upper)) return ST_EBADFUNC;
  s->a = x_lower; s->fa = f_lower;
  s->b = x_upper; s->fb = f_upper;
  s->c = x_upper; s->fc = f_upper;
  s->d = x_upper - x_lower;
  s->e = x_upper - x_lower;
  if ((f_lower < 0.0 && f_upper < 0.0) || (f_lower > 0.0 && f_upper > 0.0)) return ST_EINVAL;
  return ST_SUCCESS;
}

static int brent_iterate(brent_state *s, char fn, func_param *p) {
  double tol, m;
  int ac_equal = 0;
  double a = s->a, b = s->b, c = s->c, fa = s->fa, fb = s->fb, fc = s->fc, d = s->d, e = s->e;
  if ((fb < 0 && fc < 0) || (fb > 0 && fc > 0)) {
    ac_equal = 1; c = a; fc = fa; d = b - a; e = b - a;
  }
  if (fabs(fc) < fabs(fb)) {
    ac_equal = 1; a = b; b = c; c = a; fa = fb; fb = fc; fc = fa;
  }
  tol = 0.5 * DBL_EPSILON * fabs(b);
  m = 0.5 * (c - b);
  if (fb == 0) {
    s->root = b; s->x_lower = b; s->x_upper = b;
    return ST_SUCCESS;
  }
  if (fabs(m) <= tol) {
    s->root = b;
    if (b < c) { s->x_lower = b; s->x_upper = c; } else { s->x_lower = c; s->x_upper = b; }
    return ST_SUCCESS;
  }
  if (fabs(e) < tol || fabs(fa) <= fabs(fb)) {
    d = m; e = m;                       /* bisection */
  } else {
    double pp, q, r, sfrac = fb / fa;   /* inverse interpolation */
    if (ac_equal) { pp = 2 * m * sfrac; q = 1 - sfrac; }
    else {
      q = fa / fc; r = fb / fc;
      pp = sfrac * (2 * m * q * (q - r) - (b - a) * (r - 1));
      q = (q - 1) * (r - 1) * (sfrac - 1);
    }
    if (pp > 0) q = -q; else pp = -pp;
    {
      double lim1 = 3 * m * q - fabs(tol * q), lim2 = fabs(e * q);
      if (2 * pp < (lim1 < lim2 ? lim1 : lim2)) { e = d; d = pp / q; }
      else { d = m; e = m; }
    }
  }
  a = b; fa = fb;
  if (fabs(d) > tol) b += d; else b += (m > 0 ? +tol : -tol);
  fb = eval_dev1(fn, b, p);
  if (!isfinite(fb)) return ST_EBADFUNC;       /* SAFE_FUNC_CALL: state not stored */
  s->a = a; s->b = b; s->c = c; s->d = d; s->e = e; s->fa = fa; s->fb = fb; s->fc = fc;
  s->root = b;
  if ((fb < 0 && fc < 0) || (fb > 0 && fc > 0)) c = a;
  if (b < c) { s->x_lower = b; s->x_upper = c; } else { s->x_lower = c; s->x_upper = b; }
  return ST_SUCCESS;
}

static int root_test_interval(double x_lower, double x_upper, double epsabs, double epsrel) {
  double abs_lower = fabs(x_lower), abs_upper = fabs(x_upper), min_abs, tolerance;
  if (x_lower > x_upper) return ST_EINVAL;
  if ((x_lower > 0.0 && x_upper > 0.0) || (x_lower < 0.0 && x_upper < 0.0))
    min_abs = abs_lower < abs_upper ? abs_lower : abs_upper;
  else min_abs = 0;
  tolerance = epsabs + epsrel * min_abs;
  if (fabs(x_upper - x_lower) < tolerance) return ST_SUCCESS;
  return ST_CONTINUE;
}

static int root_test_delta(double x1, double x0, double epsabs, double epsrel) {
  double tolerance = epsabs + epsrel * fabs(x1);
  if (fabs(x1 - x0) < tolerance || x1 == x0) return ST_SUCCESS;
  return ST_CONTINUE;
}

typedef struct { double f, df, root; } newton_state;

static int newton_set(newton_state *s, char fn, func_param *p, double root) {
  s->root = root;
  eval_dev12(fn, root, p, &s->f, &s->df);
  return ST_SUCCESS;
}
static int newton_iterate(newton_state *s, char fn, func_param *p) {
  double root_new, f_new, df_new;
  if (s->df == 0.0) return ST_EZERODIV;
  root_new = s->root - (s->f / s->df);
  s->root = root_new;
  eval_dev12(fn, root_new, p, &f_new, &df_new);
  s->f = f_new; s->df = df_new;
  if (!isfinite(f_new)) return ST_EBADFUNC;
  if (!isfinite(df_new)) return ST_EBADFUNC;
  return ST_SUCCESS;
}

/* src/lmm.cpp:1945-2140 CalcLambda(func_name, params, l_min, l_max, n_region, lambda, logf) */
static void calc_lambda(char fn, func_param *p, double l_min, double l_max, size_t n_region,
                        double *lambda, double *logf) {
  *logf = NAN; *lambda = NAN;
  /* at most n_region sign-change intervals */
  double *lo = (double *)malloc(sizeof(double) * (n_region + 1));
  double *hi = (double *)malloc(sizeof(double) * (n_region + 1));
  size_t n_iv = 0;
  double lambda_interval = log(l_max / l_min) / (double)n_region;
  double lambda_l, lambda_h, dev1_l, dev1_h, logf_l, logf_h;
  for (size_t i = 0; i < n_region; ++i) {
    lambda_l = l_min * exp(lambda_interval * i);
    lambda_h = l_min * exp(lambda_interval * (i + 1.0));
    dev1_l = eval_dev1(fn, lambda_l, p);
    dev1_h = eval_dev1(fn, lambda_h, p);
    if (dev1_l * dev1_h <= 0) { lo[n_iv] = lambda_l; hi[n_iv] = lambda_h; n_iv++; }
  }
  if (n_iv == 0) {
    logf_l = eval_f(fn, l_min, p);
    logf_h = eval_f(fn, l_max, p);
    if (logf_l >= logf_h) { *lambda = l_min; *logf = logf_l; }
    else { *lambda = l_max; *logf = logf_h; }
    free(lo); free(hi);
    return;
  }
  {
    double l = 0.0, l_temp = 0.0;
    int broke_out = 0;
    for (size_t i = 0; i < n_iv; ++i) {
      brent_state bs; newton_state ns;
      memset(&bs, 0, sizeof(bs)); memset(&ns, 0, sizeof(ns));
      int status = -1;                /* GSL_FAILURE */
      unsigned iter = 0, iter2 = 0;
      const unsigned max_iter = 100;
      lambda_l = lo[i]; lambda_h = hi[i];
      brent_set(&bs, fn, p, lambda_l, lambda_h);   /* return value ignored, handler off (:2033-2034) */
      do {
        iter++;
        status = brent_iterate(&bs, fn, p);
        if (status != ST_SUCCESS && status != ST_CONTINUE) break;
        l = bs.root; lambda_l = bs.x_lower; lambda_h = bs.x_upper;
        status = root_test_interval(lambda_l, lambda_h, 0, 1e-1);
        if (status != ST_SUCCESS && status != ST_CONTINUE) break;
      } while (status == ST_CONTINUE && iter < max_iter);
      if (status == ST_CONTINUE) { broke_out = 1; break; }      /* :2057-2060 */
      newton_set(&ns, fn, p, l);
      do {
        iter2++;
        status = newton_iterate(&ns, fn, p);
        if (status != ST_SUCCESS && status != ST_CONTINUE) break;
        l_temp = l;
        l = ns.root;
        status = root_test_delta(l, l_temp, 0, 1e-5);
        if (status != ST_SUCCESS && status != ST_CONTINUE) break;
      } while (status == ST_CONTINUE && iter2 < max_iter && l > l_min && l < l_max);
      if (status == ST_CONTINUE || status != ST_SUCCESS) {       /* :2087-2094 */
        *logf = NAN; *lambda = NAN;
        free(lo); free(hi);
        return;
      }
      l = l_temp;                                                /* :2096 previous iterate */
      if (l < l_min) l = l_min;
      if (l > l_max) l = l_max;
      logf_l = eval_f(fn, l, p);
      if (i == 0) { *logf = logf_l; *lambda = l; }
      else if (*logf < logf_l) { *logf = logf_l; *lambda = l; }
    }
    (void)broke_out;
    logf_l = eval_f(fn, l_min, p);
    logf_h = eval_f(fn, l_max, p);
    if (logf_l > *logf) { *lambda = l_min; *logf = logf_l; }
    if (logf_h > *logf) { *lambda = l_max; *logf = logf_h; }
  }
  free(lo); free(hi);
}

/* ------------------------------------------------------------------------- */
static void param_alloc(func_param *p, int calc_null, size_t n, size_t n_cvt, const double *eval,
                        const double *Uab) {
  size_t n_index = n_index_of(n_cvt);
  memset(p, 0, sizeof(*p));
  p->calc_null = calc_null; p->ni_test = n; p->n_cvt = n_cvt; p->eval = eval; p->Uab = Uab;
  p->Hi = (double *)malloc(sizeof(double) * n);
  p->HiHi = (double *)malloc(sizeof(double) * n);
  p->HiHiHi = (double *)malloc(sizeof(double) * n);
  p->vtmp = (double *)malloc(sizeof(double) * n);
  p->Pab = (double *)calloc((n_cvt + 2) * n_index, sizeof(double));
  p->PPab = (double *)calloc((n_cvt + 2) * n_index, sizeof(double));
  p->PPPab = (double *)calloc((n_cvt + 2) * n_index, sizeof(double));
  p->Iab = (double *)calloc((n_cvt + 2) * n_index, sizeof(double));
}
static void param_free(func_param *p) {
  free(p->Hi); free(p->HiHi); free(p->HiHiHi); free(p->vtmp);
  free(p->Pab); free(p->PPab); free(p->PPPab); free(p->Iab);
}

/* src/lmm.cpp:1213-1256 CalcUab(UtW,Uty,Uab): SNP-independent columns */
static void calc_uab_base(const double *UtW, size_t ldw, const double *Uty, size_t n, size_t n_cvt,
                          double *Uab) {
  size_t n_index = n_index_of(n_cvt);
  for (size_t a = 1; a <= n_cvt + 2; ++a) {
    if (a == n_cvt + 1) continue;
    for (size_t b = a; b >= 1; --b) {
      if (b == n_cvt + 1) continue;
      size_t index_ab = go_getab_index(a, b, n_cvt);
      for (size_t i = 0; i < n; ++i) {
        double ua = (a == n_cvt + 2) ? Uty[i] : UtW[i * ldw + (a - 1)];
        double ub = (b == n_cvt + 2) ? Uty[i] : UtW[i * ldw + (b - 1)];
        Uab[i * n_index + index_ab] = ub * ua;
      }
    }
  }
}

/* src/lmm.cpp:1258-1280 CalcUab(UtW,Uty,Utx,Uab): the c+2 columns involving x */
static void calc_uab_x(const double *UtW, size_t ldw, const double *Uty, const double *Utx,
                       size_t incx, size_t n, size_t n_cvt, double *Uab) {
  size_t n_index = n_index_of(n_cvt);
  for (size_t b = 1; b <= n_cvt + 2; ++b) {
    size_t index_ab = go_getab_index(n_cvt + 1, b, n_cvt);
    for (size_t i = 0; i < n; ++i) {
      double x = Utx[i * incx];
      double ub = (b == n_cvt + 2) ? Uty[i] : (b == n_cvt + 1) ? x : UtW[i * ldw + (b - 1)];
      Uab[i * n_index + index_ab] = ub * x;
    }
  }
}

/* src/lmm.cpp:1127-1167 CalcRLWald ; :1170-1211 CalcRLScore */
static void calc_rl_wald_score(int score, double l, func_param *p, double *beta, double *se,
                               double *pval) {
  size_t n_cvt = p->n_cvt, n = p->ni_test, n_index = n_index_of(n_cvt);
  int df = (int)n - (int)n_cvt - 1;
  fill_hi(p, l, 1);
  calc_pab(n_cvt, p->Hi, p->Uab, n, p->Pab);
  size_t index_yy = go_getab_index(n_cvt + 2, n_cvt + 2, n_cvt);
  size_t index_xx = go_getab_index(n_cvt + 1, n_cvt + 1, n_cvt);
  size_t index_xy = go_getab_index(n_cvt + 2, n_cvt + 1, n_cvt);
  double P_yy = p->Pab[n_cvt * n_index + index_yy];
  double P_xx = p->Pab[n_cvt * n_index + index_xx];
  double P_xy = p->Pab[n_cvt * n_index + index_xy];
  double Px_yy = p->Pab[(n_cvt + 1) * n_index + index_yy];
  *beta = P_xy / P_xx;
  double tau = (double)df / Px_yy;
  *se = safe_sqrt(1.0 / (tau * P_xx));
  if (score) *pval = go_cdf_fdist_Q((double)n * P_xy * P_xy / (P_yy * P_xx), 1.0, df);
  else *pval = go_cdf_fdist_Q((P_yy - Px_yy) * tau, 1.0, df);
}

/* SUMSTAT (src/param.h:54-66) */
typedef struct {
  double beta, se, lambda_remle, lambda_mle, p_wald, p_lrt, p_score, logl_H1;
} go_sumstat;

/*
 * The per-SNP body of LMM::Analyze's batch_compute closure (src/lmm.cpp:1526-1562):
 * UtX is n x l row-major with leading dimension ldx (SNP i is column i, as in
 * UtXlarge).  Results are appended in order.  a_mode in {1,2,3,4,9}.
 * n_eval_out (optional) receives the number of likelihood-function evaluations.
 */
static int lmm_analyze_impl(size_t n, size_t n_cvt, const double *eval, const double *UtW,
                                 size_t ldw, const double *Uty, const double *UtX, size_t l,
                                 size_t ldx, int a_mode, double l_min, double l_max,
                                 size_t n_region, double l_mle_null, double logl_mle_H0,
                                 go_sumstat *out, long *n_eval_out, int plink_rule) {
  size_t n_index = n_index_of(n_cvt);
  double *Uab = (double *)calloc(n * n_index, sizeof(double));
  func_param p;
  if (!Uab) return 1;
  param_alloc(&p, 0, n, n_cvt, eval, Uab);
  calc_uab_base(UtW, ldw, Uty, n, n_cvt, Uab);                  /* :1508 */
  for (size_t i = 0; i < l; ++i) {
    double lambda_mle = 0.0, lambda_remle = 0.0, beta = 0.0, se = 0.0, p_wald = 0.0;
    double p_lrt = 0.0, p_score = 0.0, logl_H1 = 0.0;
    calc_uab_x(UtW, ldw, Uty, UtX + i, ldx, n, n_cvt, Uab);     /* :1528-1531 */
    if (a_mode == 3 || a_mode == 4 || a_mode == 9)              /* :1541-1543 */
      calc_rl_wald_score(1, l_mle_null, &p, &beta, &se, &p_score);
    if (a_mode == 1 || a_mode == 4) {                           /* :1546-1549 */
      calc_lambda('R', &p, l_min, l_max, n_region, &lambda_remle, &logl_H1);
      if (!(plink_rule && isnan(logl_H1)))                       /* AnalyzePlink: src/lmm.cpp:1869-1870 */
        calc_rl_wald_score(0, lambda_remle, &p, &beta, &se, &p_wald);
    }
    if (a_mode == 2 || a_mode == 4 || a_mode == 9) {            /* :1551-1554 */
      calc_lambda('L', &p, l_min, l_max, n_region, &lambda_mle, &logl_H1);
      p_lrt = go_cdf_chisq1_Q(2.0 * (logl_H1 - logl_mle_H0));
    }
    if (plink_rule && isnan(logl_H1)) { p_wald = logl_H1; p_lrt = logl_H1; }   /* src/lmm.cpp:1882-1884 */
    out[i].beta = beta; out[i].se = se; out[i].lambda_remle = lambda_remle;
    out[i].lambda_mle = lambda_mle; out[i].p_wald = p_wald; out[i].p_lrt = p_lrt;
    out[i].p_score = p_score; out[i].logl_H1 = logl_H1;
  }
  if (n_eval_out) *n_eval_out = p.n_eval;
  param_free(&p);
  free(Uab);
  return 0;
}

GO_EXPORT int go_lmm_analyze_utx(size_t n, size_t n_cvt, const double *eval, const double *UtW,
                                 size_t ldw, const double *Uty, const double *UtX, size_t l,
                                 size_t ldx, int a_mode, double l_min, double l_max,
                                 size_t n_region, double l_mle_null, double logl_mle_H0,
                                 go_sumstat *out, long *n_eval_out) {
  return lmm_analyze_impl(n, n_cvt, eval, UtW, ldw, Uty, UtX, l, ldx, a_mode, l_min, l_max, n_region, l_mle_null,
                          logl_mle_H0, out, n_eval_out, 0);
}

/* Same with the NaN handling of LMM::AnalyzePlink (src/lmm.cpp:1866-1884). */
GO_EXPORT int go_lmm_analyze_utx_plink(size_t n, size_t n_cvt, const double *eval, const double *UtW,
                                       size_t ldw, const double *Uty, const double *UtX, size_t l,
                                       size_t ldx, int a_mode, double l_min, double l_max,
                                       size_t n_region, double l_mle_null, double logl_mle_H0,
                                       go_sumstat *out, long *n_eval_out) {
  return lmm_analyze_impl(n, n_cvt, eval, UtW, ldw, Uty, UtX, l, ldx, a_mode, l_min, l_max, n_region, l_mle_null,
                          logl_mle_H0, out, n_eval_out, 1);
}

/* Null model: src/lmm.cpp:2143-2180 CalcLambda(func_name, eval, UtW, Uty, ...) */
GO_EXPORT int go_calc_lambda_null(char func_name, size_t n, size_t n_cvt, const double *eval,
                                  const double *UtW, size_t ldw, const double *Uty, double l_min,
                                  double l_max, size_t n_region, double *lambda, double *logl_H0) {
  size_t n_index = n_index_of(n_cvt);
  double *Uab = (double *)calloc(n * n_index, sizeof(double));
  func_param p;
  if (!Uab) return 1;
  if (func_name == 'r') func_name = 'R';
  if (func_name == 'l') func_name = 'L';
  param_alloc(&p, 1, n, n_cvt, eval, Uab);
  calc_uab_base(UtW, ldw, Uty, n, n_cvt, Uab);
  calc_lambda(func_name, &p, l_min, l_max, n_region, lambda, logl_H0);
  param_free(&p);
  free(Uab);
  return 0;
}

/* src/lmm.cpp:2183-2205 CalcPve */
GO_EXPORT int go_calc_pve(size_t n, size_t n_cvt, const double *eval, const double *UtW, size_t ldw,
                          const double *Uty, double lambda, double trace_G, double *pve,
                          double *pve_se) {
  size_t n_index = n_index_of(n_cvt);
  double *Uab = (double *)calloc(n * n_index, sizeof(double));
  func_param p;
  double d1, d2;
  if (!Uab) return 1;
  param_alloc(&p, 1, n, n_cvt, eval, Uab);
  calc_uab_base(UtW, ldw, Uty, n, n_cvt, Uab);
  LogRL_dev12(lambda, &p, &d1, &d2, 1);
  double se = safe_sqrt(-1.0 / d2);
  *pve = trace_G * lambda / (trace_G * lambda + 1.0);
  *pve_se = trace_G / ((trace_G * lambda + 1.0) * (trace_G * lambda + 1.0)) * se;
  param_free(&p);
  free(Uab);
  return 0;
}

/* Small dense LU with partial pivoting (gsl_linalg_LU_decomp/solve/invert via
 * src/lapack.cpp:307-352), used only for the c x c null-model system. */
static int lu_solve_invert(size_t c, double *A, const double *rhs, double *x, double *Ainv) {
  size_t *piv = (size_t *)malloc(sizeof(size_t) * c);
  for (size_t i = 0; i < c; ++i) piv[i] = i;
  for (size_t k = 0; k < c; ++k) {
    size_t pr = k; double mx = fabs(A[k * c + k]);
    for (size_t i = k + 1; i < c; ++i) if (fabs(A[i * c + k]) > mx) { mx = fabs(A[i * c + k]); pr = i; }
    if (pr != k) {
      for (size_t j = 0; j < c; ++j) { double t = A[k * c + j]; A[k * c + j] = A[pr * c + j]; A[pr * c + j] = t; }
      size_t t = piv[k]; piv[k] = piv[pr]; piv[pr] = t;
    }
    if (A[k * c + k] != 0.0)
      for (size_t i = k + 1; i < c; ++i) {
        double f = A[i * c + k] / A[k * c + k];
        A[i * c + k] = f;
        for (size_t j = k + 1; j < c; ++j) A[i * c + j] -= f * A[k * c + j];
      }
  }
  double *col = (double *)malloc(sizeof(double) * c);
  for (size_t r = 0; r <= c; ++r) {           /* r<c: columns of the inverse; r==c: rhs */
    for (size_t i = 0; i < c; ++i) col[i] = (r == c) ? rhs[piv[i]] : (piv[i] == r ? 1.0 : 0.0);
    for (size_t i = 0; i < c; ++i) for (size_t j = 0; j < i; ++j) col[i] -= A[i * c + j] * col[j];
    for (size_t ii = c; ii-- > 0;) {
      for (size_t j = ii + 1; j < c; ++j) col[ii] -= A[ii * c + j] * col[j];
      col[ii] /= A[ii * c + ii];
    }
    for (size_t i = 0; i < c; ++i) { if (r == c) x[i] = col[i]; else Ainv[i * c + r] = col[i]; }
  }
  free(col); free(piv);
  return 0;
}

/* src/lmm.cpp:2210-2281 CalcLmmVgVeBeta */
GO_EXPORT int go_calc_vgvebeta(size_t n, size_t n_cvt, const double *eval, const double *UtW,
                               size_t ldw, const double *Uty, double lambda, double *vg, double *ve,
                               double *beta, double *se_beta) {
  size_t n_index = n_index_of(n_cvt);
  double *Uab = (double *)calloc(n * n_index, sizeof(double));
  double *WHiW = (double *)calloc(n_cvt * n_cvt, sizeof(double));
  double *WHiy = (double *)calloc(n_cvt, sizeof(double));
  double *Vbeta = (double *)calloc(n_cvt * n_cvt, sizeof(double));
  func_param p;
  if (!Uab) return 1;
  param_alloc(&p, 1, n, n_cvt, eval, Uab);
  calc_uab_base(UtW, ldw, Uty, n, n_cvt, Uab);
  fill_hi(&p, lambda, 1);
  for (size_t i = 0; i < n; ++i)
    for (size_t a = 0; a < n_cvt; ++a) {
      double hw = UtW[i * ldw + a] * p.Hi[i];
      WHiy[a] += hw * Uty[i];
      for (size_t b = 0; b < n_cvt; ++b) WHiW[a * n_cvt + b] += hw * UtW[i * ldw + b];
    }
  lu_solve_invert(n_cvt, WHiW, WHiy, beta, Vbeta);
  calc_pab(n_cvt, p.Hi, Uab, n, p.Pab);
  size_t index_yy = go_getab_index(n_cvt + 2, n_cvt + 2, n_cvt);
  double P_yy = p.Pab[n_cvt * n_index + index_yy];
  *ve = P_yy / (double)(n - n_cvt);
  *vg = *ve * lambda;
  for (size_t i = 0; i < n_cvt; ++i) se_beta[i] = safe_sqrt(Vbeta[i * n_cvt + i] * (*ve));
  param_free(&p);
  free(Uab); free(WHiW); free(WHiy); free(Vbeta);
  return 0;
}

/* Expose the likelihood functions for direct spot checks.
 * which: 0 = f, 1 = dev1, 2 = dev2; fn 'R' or 'L'.  Utx may be NULL with calc_null=1. */
GO_EXPORT double go_eval_fn(char fn, int which, int calc_null, double l, size_t n, size_t n_cvt,
                            const double *eval, const double *UtW, size_t ldw, const double *Uty,
                            const double *Utx) {
  size_t n_index = n_index_of(n_cvt);
  double *Uab = (double *)calloc(n * n_index, sizeof(double));
  func_param p;
  double r = NAN, d1, d2;
  param_alloc(&p, calc_null, n, n_cvt, eval, Uab);
  calc_uab_base(UtW, ldw, Uty, n, n_cvt, Uab);
  if (Utx) calc_uab_x(UtW, ldw, Uty, Utx, 1, n, n_cvt, Uab);
  if (which == 0) r = eval_f(fn, l, &p);
  else { eval_dev12(fn, l, &p, &d1, &d2); r = which == 1 ? d1 : d2; }
  param_free(&p);
  free(Uab);
  return r;
}

/* ------------------------------------------------------------------------- */
/* Kinship.  src/gemma_io.cpp:1486-1570 (BimbamKin per-SNP transform + batch
 * accumulate) and :1650-1729 (PlinkKin).  G is SNP-major: l SNPs x n individuals
 * (ldg >= n); missing genotypes are NaN.  Xc (n x l row-major, optional) receives
 * the centred/scaled columns exactly as Xlarge holds them; K (n x n, ldk) is
 * accumulated K += Xc Xc^T.  k_mode 1 = centred, 2 = standardised. */
GO_EXPORT int go_kin_transform(const double *G, size_t l, size_t n, size_t ldg, int k_mode,
                               double *Xc, size_t ldx) {
  for (size_t s = 0; s < l; ++s) {
    const double *g = G + s * ldg;
    double mean = 0.0, var = 0.0;
    size_t n_miss = 0;
    for (size_t i = 0; i < n; ++i) {
      if (isnan(g[i])) n_miss++;
      else { mean += g[i]; var += g[i] * g[i]; }
    }
    /* :1511-1516 */
    mean /= (double)(n - n_miss);
    var += mean * mean * (double)n_miss;
    var /= (double)n;
    var -= mean * mean;
    for (size_t i = 0; i < n; ++i) {
      double v = isnan(g[i]) ? mean : g[i];   /* :1518-1522 */
      v -= mean;                               /* :1524 */
      if (k_mode == 2 && var != 0) v /= sqrt(var);   /* :1526-1528 */
      Xc[i * ldx + s] = v;
    }
  }
  return 0;
}

/* Plain K += X X^T (lower+upper), the job fast_eigen_dgemm does at
 * src/gemma_io.cpp:1554; O(n^2 l) reference loop for small cases only. */
GO_EXPORT int go_kin_accumulate(const double *Xc, size_t n, size_t l, size_t ldx, double *K,
                                size_t ldk) {
  for (size_t i = 0; i < n; ++i)
    for (size_t j = 0; j <= i; ++j) {
      double s = 0.0;
      const double *xi = Xc + i * ldx, *xj = Xc + j * ldx;
      for (size_t t = 0; t < l; ++t) s += xi[t] * xj[t];
      K[i * ldk + j] += s;
      if (j != i) K[j * ldk + i] += s;
    }
  return 0;
}

/* src/mathfunc.cpp:147-177 CenterMatrix: G <- G - (Gw w^T + w Gw^T)/n + (w^T G w / n^2) w w^T
 * computed on the upper triangle then mirrored to the lower. */
GO_EXPORT int go_center_matrix(double *G, size_t n, size_t ldg) {
  double *Gw = (double *)malloc(sizeof(double) * n);
  double d = 0.0;
  if (!Gw) return 1;
  for (size_t i = 0; i < n; ++i) {
    double s = 0.0;
    for (size_t j = 0; j < n; ++j) s += G[i * ldg + j];
    Gw[i] = s;
  }
  for (size_t i = 0; i < n; ++i) d += Gw[i];
  {
    double alpha = -1.0 / (double)n, beta = d / ((double)n * (double)n);
    for (size_t i = 0; i < n; ++i)
      for (size_t j = i; j < n; ++j) {
        double v = G[i * ldg + j];
        v += alpha * (Gw[i] + Gw[j]);      /* dsyr2 upper */
        v += beta;                         /* dsyr upper  */
        G[i * ldg + j] = v;
      }
  }
  for (size_t i = 0; i < n; ++i)
    for (size_t j = 0; j < i; ++j) G[i * ldg + j] = G[j * ldg + i];
  free(Gw);
  return 0;
}

/* src/lapack.cpp:260-291: eigenvalues < 1e-10 -> 0; returns mean(eval) = trace_G */
GO_EXPORT double go_zero_small_eval(double *eval, size_t n) {
  double d = 0.0;
  for (size_t i = 0; i < n; ++i) { if (eval[i] < 1e-10) eval[i] = 0.0; d += eval[i]; }
  return d / (double)n;
}

/* PLINK .bed 2-bit decode of one SNP row (src/gemma_io.cpp:1665-1682,
 * src/lmm.cpp:1783-1817): sample j uses bits (2j mod 8, 2j+1 mod 8) of byte j/4,
 * low bit first.  00 -> 2, 01(low=0,high=1) -> 1, 11 -> 0, 10(low=1,high=0) -> missing (NaN). */
GO_EXPORT void go_bed_decode(const unsigned char *row, size_t n, double *g) {
  for (size_t j = 0; j < n; ++j) {
    unsigned b = row[j >> 2] >> (2 * (j & 3));
    unsigned lo = b & 1u, hi = (b >> 1) & 1u;
    if (lo == 0) g[j] = hi == 0 ? 2.0 : 1.0;
    else g[j] = hi == 1 ? 0.0 : NAN;
  }
}

/* Mean imputation of LMM::Analyze (src/lmm.cpp:1590-1618): G is l x n SNP-major
 * with NaN = missing, X (n x l row-major, ldx) receives raw genotypes with the
 * per-SNP mean over non-missing plugged into the holes (no centring). */
GO_EXPORT void go_lmm_impute(const double *G, size_t l, size_t n, size_t ldg, double *X, size_t ldx) {
  for (size_t s = 0; s < l; ++s) {
    const double *g = G + s * ldg;
    double tot = 0.0; size_t n_miss = 0;
    for (size_t i = 0; i < n; ++i) { if (isnan(g[i])) n_miss++; else tot += g[i]; }
    double mean = tot / (double)(n - n_miss);
    for (size_t i = 0; i < n; ++i) X[i * ldx + s] = isnan(g[i]) ? mean : g[i];
  }
}
