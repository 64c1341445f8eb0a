// Host instantiation of gemma_b200/csrc/mvlmm_core.cuh (GB_MV_HOST: one "lane") for tests/test_mvlmm_core.py: the moment
// formulation that the CUDA kernels run is checked against the numpy oracle (oracle/mvlmm_oracle.py) without a GPU.
#define GB_MV_HOST 1
#include "../../gemma_b200/csrc/mvlmm_core.cuh"

using namespace gbmv;

template <int D, int C>
static int null_t(int n, const double *ev, const double *X, const double *Y, const double *Vg0, const double *Ve0, double *out) {
  MvData<C + D> dat; dat.n = n; dat.delta = ev;
  for (int j = 0; j < C; ++j) dat.z[j] = X + (size_t)j * n;
  for (int s = 0; s < D; ++s) dat.z[C + s] = Y + (size_t)s * n;
  Fit<D, C> fit;
  for (int i = 0; i < D; ++i) for (int j = 0; j < D; ++j) { fit.V_g[i][j] = Vg0[i * D + j]; fit.V_e[i][j] = Ve0[i * D + j]; }
  mph_calc_beta<D, C>(dat, fit.V_g, fit.V_e, fit.B);
  mph_em<D, C>(true, 10000, 1e-4, dat, fit);
  const double lr = mph_nr<D, C>(true, 100, 1e-4, dat, fit);
  mph_calc_beta<D, C>(dat, fit.V_g, fit.V_e, fit.B);
  int o = 0;
  for (int i = 0; i < D; ++i) for (int j = 0; j < D; ++j) out[o++] = fit.V_g[i][j];
  for (int i = 0; i < D; ++i) for (int j = 0; j < D; ++j) out[o++] = fit.V_e[i][j];
  out[o++] = lr;
  mph_em<D, C>(false, 10000, 1e-4, dat, fit);
  const double lm = mph_nr<D, C>(false, 100, 1e-4, dat, fit);
  mph_calc_beta<D, C>(dat, fit.V_g, fit.V_e, fit.B);
  for (int i = 0; i < D; ++i) for (int j = 0; j < D; ++j) out[o++] = fit.V_g[i][j];
  for (int i = 0; i < D; ++i) for (int j = 0; j < D; ++j) out[o++] = fit.V_e[i][j];
  out[o++] = lm;
  for (int i = 0; i < D; ++i) for (int j = 0; j < C; ++j) out[o++] = fit.B[i][j];
  return o;
}

template <int D, int C>
static int snp_t(int n, const double *ev, const double *X, const double *x, const double *Y, const double *Vg, const double *Ve, const double *Bn,
                 int a_mode, double logl_mle_H0, double *out) {
  constexpr int C1 = C + 1;
  MvData<C1 + D> dat; dat.n = n; dat.delta = ev;
  for (int j = 0; j < C; ++j) dat.z[j] = X + (size_t)j * n;
  dat.z[C] = x;
  for (int s = 0; s < D; ++s) dat.z[C1 + s] = Y + (size_t)s * n;
  Fit<D, C1> fit;
  for (int i = 0; i < D; ++i) for (int j = 0; j < D; ++j) { fit.V_g[i][j] = Vg[i * D + j]; fit.V_e[i][j] = Ve[i * D + j]; }
  for (int i = 0; i < D; ++i) { for (int j = 0; j < C; ++j) fit.B[i][j] = Bn[i * C + j]; fit.B[i][C] = 0.0; }
  analyze_snp<D, C1>(dat, fit, a_mode, 10000, 1e-4, 100, 1e-4, 0.001, logl_mle_H0, out);     // src/param.cpp:98-99 defaults
  return D + D * (D + 1) / 2 + 3;
}

extern "C" {
int mvh_null(int n, int c, const double *ev, const double *X, const double *Y, const double *Vg0, const double *Ve0, double *out) {
  switch (c) { case 1: return null_t<2, 1>(n, ev, X, Y, Vg0, Ve0, out); case 2: return null_t<2, 2>(n, ev, X, Y, Vg0, Ve0, out); case 3: return null_t<2, 3>(n, ev, X, Y, Vg0, Ve0, out); }
  return -1;
}
int mvh_snp(int n, int c, const double *ev, const double *X, const double *x, const double *Y, const double *Vg, const double *Ve, const double *Bn,
            int a_mode, double logl_mle_H0, double *out) {
  switch (c) { case 1: return snp_t<2, 1>(n, ev, X, x, Y, Vg, Ve, Bn, a_mode, logl_mle_H0, out); case 2: return snp_t<2, 2>(n, ev, X, x, Y, Vg, Ve, Bn, a_mode, logl_mle_H0, out);
               case 3: return snp_t<2, 3>(n, ev, X, x, Y, Vg, Ve, Bn, a_mode, logl_mle_H0, out); }
  return -1;
}
// three phenotypes (the kernels of this round instantiate two; the core itself is generic in D)
int mvh3_null(int n, int c, const double *ev, const double *X, const double *Y, const double *Vg0, const double *Ve0, double *out) {
  switch (c) { case 1: return null_t<3, 1>(n, ev, X, Y, Vg0, Ve0, out); case 2: return null_t<3, 2>(n, ev, X, Y, Vg0, Ve0, out); }
  return -1;
}
int mvh3_snp(int n, int c, const double *ev, const double *X, const double *x, const double *Y, const double *Vg, const double *Ve, const double *Bn,
             int a_mode, double logl_mle_H0, double *out) {
  switch (c) { case 1: return snp_t<3, 1>(n, ev, X, x, Y, Vg, Ve, Bn, a_mode, logl_mle_H0, out); case 2: return snp_t<3, 2>(n, ev, X, x, Y, Vg, Ve, Bn, a_mode, logl_mle_H0, out); }
  return -1;
}
}
