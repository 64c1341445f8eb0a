// mvlmm_kernel.cu -- kernels of the multivariate LMM (two phenotypes): null model and per-SNP tests on the moment-form core
// (mvlmm_core.cuh).  Replaces MVLMM::AnalyzeBimbam / AnalyzePlink (src/mvlmm.cpp:2972-3899).  Own translation unit: the EM /
// Newton-Raphson routines are large and would otherwise dominate the build time of lmm_kernel.cu.
#include "common.cuh"
#include "mvlmm_core.cuh"

namespace gb {

// ---- multivariate LMM, two phenotypes (MVLMM::AnalyzeBimbam / AnalyzePlink, src/mvlmm.cpp:2972-3899; math in mvlmm_core.cuh) ----------
template <int C>
__device__ __forceinline__ void mv_fill(gbmv::MvData<C + 2> &d, const MvConst &K) {
  d.n = K.n; d.delta = K.delta;
  for (int j = 0; j < C; ++j) d.z[j] = K.Wt + (size_t)j * K.ld;
  for (int s2 = 0; s2 < 2; ++s2) d.z[C + s2] = K.Yt + (size_t)s2 * K.ld;
}

// null model: MphInitial's B, EM + NR for REML, then EM + NR for ML (mvlmm.cpp:3047-3133); one warp
template <int C>
__global__ void __launch_bounds__(32) mv_null_kernel(MvConst K, MvNull *out) {
  using namespace gbmv;
  constexpr int D = 2;
  MvData<C + D> dat; mv_fill<C>(dat, K);
  Fit<D, C> fit;
  for (int i = 0; i < D; ++i) for (int j = 0; j < D; ++j) { fit.V_g[i][j] = (i == j) ? K.vg0[i] : 0.0; fit.V_e[i][j] = (i == j) ? K.ve0[i] : 0.0; }
  for (int i = 0; i < D; ++i) for (int j = 0; j < C; ++j) fit.B[i][j] = 0.0;
  mph_calc_beta<D, C>(dat, fit.V_g, fit.V_e, fit.B);
  mph_em<D, C>(true, K.em_iter, K.em_prec, dat, fit);
  const double lr = mph_nr<D, C>(true, K.nr_iter, K.nr_prec, dat, fit);
  mph_calc_beta<D, C>(dat, fit.V_g, fit.V_e, fit.B);
  MvNull r;
  for (int i = 0; i < D; ++i) for (int j = 0; j < D; ++j) { r.Vg_remle[i * D + j] = fit.V_g[i][j]; r.Ve_remle[i * D + j] = fit.V_e[i][j]; }
  for (int i = 0; i < D; ++i) for (int j = 0; j < 4; ++j) r.B_remle[i * 4 + j] = (j < C) ? fit.B[i][j] : 0.0;
  r.logl_remle = lr;
  mph_em<D, C>(false, K.em_iter, K.em_prec, dat, fit);
  const double lm = mph_nr<D, C>(false, K.nr_iter, K.nr_prec, dat, fit);
  mph_calc_beta<D, C>(dat, fit.V_g, fit.V_e, fit.B);
  for (int i = 0; i < D; ++i) for (int j = 0; j < D; ++j) { r.Vg_mle[i * D + j] = fit.V_g[i][j]; r.Ve_mle[i * D + j] = fit.V_e[i][j]; }
  for (int i = 0; i < D; ++i) for (int j = 0; j < 4; ++j) r.B_mle[i * 4 + j] = (j < C) ? fit.B[i][j] : 0.0;
  r.logl_mle = lm;
  if ((threadIdx.x & 31) == 0) *out = r;
}

// per SNP: -lmm 1/2/3/4 body of the reference loop (mvlmm.cpp:3286-3360) from the null estimates
template <int C>
__global__ void __launch_bounds__(128) mv_assoc_kernel(MvConst K, const MvNull *__restrict__ nm, const double *__restrict__ UtXt, size_t ldu, int l,
                                                       double *__restrict__ out, unsigned int *__restrict__ ticket) {
  using namespace gbmv;
  constexpr int D = 2, C1 = C + 1;
  const int lane = threadIdx.x & 31;
  for (;;) {
    unsigned int s = 0;
    if (lane == 0) s = atomicAdd(ticket, 1u);
    s = __shfl_sync(0xffffffffu, s, 0);
    if (s >= (unsigned int)l) break;
    MvData<C1 + D> dat; dat.n = K.n; dat.delta = K.delta;
    for (int j = 0; j < C; ++j) dat.z[j] = K.Wt + (size_t)j * K.ld;
    dat.z[C] = UtXt + (size_t)s * ldu;
    for (int q = 0; q < D; ++q) dat.z[C1 + q] = K.Yt + (size_t)q * K.ld;
    Fit<D, C1> fit;
    for (int i = 0; i < D; ++i) for (int j = 0; j < D; ++j) { fit.V_g[i][j] = nm->Vg_mle[i * D + j]; fit.V_e[i][j] = nm->Ve_mle[i * D + j]; }
    for (int i = 0; i < D; ++i) { for (int j = 0; j < C; ++j) fit.B[i][j] = nm->B_mle[i * 4 + j]; fit.B[i][C] = 0.0; }
    double o8[8];
    analyze_snp<D, C1>(dat, fit, K.a_mode, K.em_iter, K.em_prec, K.nr_iter, K.nr_prec, K.p_nr, nm->logl_mle, o8);
    if (lane == 0) { double *o = out + (size_t)s * 8; for (int q = 0; q < 8; ++q) o[q] = o8[q]; }
  }
}

cudaError_t launch_mv_null(int c, const MvConst &K, MvNull *out, cudaStream_t st) {
  switch (c) {
    case 1: mv_null_kernel<1><<<1, 32, 0, st>>>(K, out); break;
    case 2: mv_null_kernel<2><<<1, 32, 0, st>>>(K, out); break;
    case 3: mv_null_kernel<3><<<1, 32, 0, st>>>(K, out); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}
cudaError_t launch_mv_assoc(int c, const MvConst &K, const MvNull *nm, const double *UtXt, size_t ldu, int l, double *out, unsigned int *ticket,
                            int num_sms, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(ticket, 0, sizeof(unsigned int), st);
  if (e != cudaSuccess) return e;
  long want = ((long)l + 3) / 4, grid = (long)num_sms * 4;
  if (grid > want) grid = want;
  if (grid < 1) grid = 1;
  switch (c) {
    case 1: mv_assoc_kernel<1><<<(unsigned)grid, 128, 0, st>>>(K, nm, UtXt, ldu, l, out, ticket); break;
    case 2: mv_assoc_kernel<2><<<(unsigned)grid, 128, 0, st>>>(K, nm, UtXt, ldu, l, out, ticket); break;
    case 3: mv_assoc_kernel<3><<<(unsigned)grid, 128, 0, st>>>(K, nm, UtXt, ldu, l, out, ticket); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

}  // namespace gb
