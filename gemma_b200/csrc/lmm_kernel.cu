// lmm_kernel.cu -- kernels wrapping the fused per-SNP evaluator (lmm_device.cuh).
//
// lmm_assoc_kernel: warp-per-SNP association tests over a rotated batch UtXt (l x n,
// SNP-major so that each warp streams ONE contiguous U^T x row with fully coalesced
// 256-byte warp loads; eigenvalues / rotated covariates / phenotype are shared by all
// warps and stay L1/L2 resident).  Grid = all SMs x resident CTAs, SNPs handed out by an
// atomic ticket so uneven Brent/Newton iteration counts do not leave SMs idle.
// Replaces the per-SNP loop of LMM::Analyze's batch_compute (src/lmm.cpp:1526-1562).
#include "common.cuh"
#include "lmm_device.cuh"
#include "lmm_v2.cuh"

namespace gb {

template <int NC>
__global__ void __launch_bounds__(128) lmm_assoc_kernel(LmmConst D, LmmParams prm,
                                                        const double *__restrict__ UtXt, size_t ldu,
                                                        int l, gb200_sumstat *__restrict__ out,
                                                        unsigned int *__restrict__ ticket) {
  const int lane = threadIdx.x & 31;
  for (;;) {
    unsigned int s = 0;
    if (lane == 0) s = atomicAdd(ticket, 1u);
    s = __shfl_sync(0xffffffffu, s, 0);
    if (s >= (unsigned int)l) break;
    gb200_sumstat r;
    analyze_snp<NC>(D, prm, UtXt + (size_t)s * ldu, r);
    if (lane == 0) out[s] = r;
  }
}

// v2: one CTA = 8 warps = 8 SNPs in lockstep passes over shared-memory stages (lmm_v2.cuh)
template <int NC>
__global__ void __launch_bounds__(V2_THREADS, (NC <= GB_V2_CTAS2_MAXNC) ? GB_V2_CTAS : 1) lmm_assoc_v2_kernel(LmmConst D, LmmParams prm,
                                                                    const double *__restrict__ UtXt, size_t ldu, int l,
                                                                    gb200_sumstat *__restrict__ out,
                                                                    unsigned int *__restrict__ ticket) {
  extern __shared__ __align__(16) double v2_smem[];
  __shared__ const double *xrows[V2_WARPS];
  __shared__ unsigned int grp_s;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nchunks = D.n_c / V2_CHUNK, pad = D.n_c - D.n;
#if GB_V2_TMA
  v2_pipeline_init(v2_smem, NC);
#endif
  unsigned int pipe_it = 0;            // chunks that went through the stage ring so far (identical in every thread)
  for (;;) {
    if (threadIdx.x == 0) grp_s = atomicAdd(ticket, 1u);
    __syncthreads();
    const unsigned int g = grp_s;
    if ((size_t)g * V2_WARPS >= (size_t)l) break;
    const int s = (int)(g * V2_WARPS) + warp;
    const bool valid = s < l;
    if (lane == 0) xrows[warp] = valid ? UtXt + (size_t)s * ldu : nullptr;
    __syncthreads();
    gb200_sumstat r;
    v2_analyze_group<NC>(D, prm, xrows, v2_smem, nchunks, pad, valid, r, pipe_it,
                         (D.xex && valid) ? D.xex + (size_t)s * (NC + 1) : nullptr,
                         (D.xsum && valid) ? D.xsum + (size_t)s * D.xsum_ld : nullptr);
    if (valid && lane == 0) out[s] = r;
    __syncthreads();
  }
}

template <int NC>
static cudaError_t launch_assoc_v2_nc(const LmmConst &D, const LmmParams &prm, const double *UtXt, size_t ldu, int l,
                                      gb200_sumstat *out, unsigned int *ticket, int num_sms, cudaStream_t st) {
  const size_t smem = v2_smem_bytes(NC);
  cudaError_t e = cudaFuncSetAttribute(lmm_assoc_v2_kernel<NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  long groups = ((long)l + V2_WARPS - 1) / V2_WARPS;
  int per_sm = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lmm_assoc_v2_kernel<NC>, V2_THREADS, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
  long grid = groups < (long)num_sms * per_sm ? groups : (long)num_sms * per_sm;
  if (grid < 1) grid = 1;
  e = cudaMemsetAsync(ticket, 0, sizeof(unsigned int), st);
  if (e != cudaSuccess) return e;
  lmm_assoc_v2_kernel<NC><<<(unsigned)grid, V2_THREADS, smem, st>>>(D, prm, UtXt, ldu, l, out, ticket);
  return cudaGetLastError();
}

// SNP-independent quantities at the lambdas every SNP visits (grid 0..n_region, exactly l_max, l_mle_null): one CTA per
// lambda writes the h row and the record read by the hoisted passes of lmm_v2.cuh.  Once per (setup, params) pair.
// Rows j >= n_region + 3 are the Chebyshev nodes of the interpolated refinement: their lambdas come from `node_lams`.
template <int NC>
__global__ void __launch_bounds__(256) lmm_common_kernel(LmmConst D, LmmParams prm, double *__restrict__ H, double *__restrict__ ctab,
                                                         const double *__restrict__ node_lams) {
  constexpr int CN = v2c_nidx(NC), CS = v2c_stride(NC), NVC = NC + 1, NVAL = 3 * CN + 3;
  __shared__ double part[8][NVAL];
  const int j = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int n_region = prm.n_region;
  const double lambda_interval = log(prm.l_max / prm.l_min) / (double)n_region;
  const double lam = (j <= n_region) ? prm.l_min * exp(lambda_interval * (double)j)
                                     : (j == n_region + 1 ? prm.l_max : (j == n_region + 2 ? prm.l_mle_null : node_lams[j - (n_region + 3)]));
  double acc[NVAL];
#pragma unroll
  for (int q = 0; q < NVAL; ++q) acc[q] = 0.0;
  double *hrow = H + (size_t)j * D.n_c;
  for (int i = tid; i < D.n_c; i += 256) {
    double h = 1.0;                                     // zero-padded tail: delta = 0
    if (i < D.n) {
      const double den = fma(lam, __ldg(D.delta + i), 1.0);
      h = 1.0 / den;
      const double h2 = h * h;
      double v[NVC];
#pragma unroll
      for (int a = 0; a < NC; ++a) v[a] = __ldg(D.Wt + (size_t)a * D.ldv + i);
      v[NC] = __ldg(D.y + i);
#pragma unroll
      for (int a = 0; a < NVC; ++a)
#pragma unroll
        for (int b = a; b < NVC; ++b) {
          const double pr = v[a] * v[b];
          const int q = abidx(a, b, NVC);
          acc[q] += pr; acc[CN + q] = fma(h, pr, acc[CN + q]); acc[2 * CN + q] = fma(h2, pr, acc[2 * CN + q]);
        }
      acc[3 * CN] += h; acc[3 * CN + 1] += h2; acc[3 * CN + 2] += log(fabs(den));
    }
    hrow[i] = h;
  }
#pragma unroll
  for (int q = 0; q < NVAL; ++q) {
    const double t = warp_allsum(acc[q]);
    if (lane == 0) part[wid][q] = t;
  }
  __syncthreads();
  if (tid < NVAL) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += part[w][tid];
    ctab[(size_t)j * CS + tid] = t;
  }
  if (tid == 0) ctab[(size_t)j * CS + 3 * CN + 3] = lam;
}

// Chebyshev coefficients (first-kind nodes, tau_m = cos(pi (m + 1/2) / M)) of the SNP-independent sums over one grid interval:
// f = S^1 pairs (CN), S^2 pairs (CN), sum h, sum h^2, sum log(l d + 1); block g = interval, thread = (f, k).  Block 0 also
// writes the cosine tables cos(pi j / (2 M)), j < 4 M, for M = V2_CM and V2_XM (the per-SNP kernel's own transform uses the second).
template <int NC>
__global__ void __launch_bounds__(512) lmm_cheb_coef_kernel(const double *__restrict__ ctab, int row0, double *__restrict__ cheb) {
  constexpr int CN = v2c_nidx(NC), CS = v2c_stride(NC), NF = 2 * CN + 3, M = V2_CM;
  const int g = blockIdx.x, tid = threadIdx.x;
  if (g == 0 && tid < 4 * M) cheb[tid] = cospi((double)tid / (double)(2 * M));
  if (g == 0 && tid < 4 * V2_XM) cheb[4 * M + tid] = cospi((double)tid / (double)(2 * V2_XM));
  if (tid >= NF * M) return;
  const int f = tid / M, k = tid - f * M;
  const int off = (f < 2 * CN) ? CN + f : 3 * CN + (f - 2 * CN);        // S^1 | S^2 are contiguous behind S^0; then tr1, tr2, logdet
  double acc = 0.0;
  for (int m = 0; m < M; ++m) {
    const double v = ctab[(size_t)(row0 + g * M + m) * CS + off];
    acc = fma(v, cospi((double)(k * (2 * m + 1)) / (double)(2 * M)), acc);
  }
  cheb[V2_CHEB_BASE + ((size_t)g * NF + f) * M + k] = acc * (k == 0 ? 1.0 / M : 2.0 / M);
}

cudaError_t launch_lmm_common(int n_cvt, const LmmConst &D, const LmmParams &prm, double *H, double *ctab, const double *node_lams,
                              int n_nodes, double *cheb, cudaStream_t st) {
  const int J0 = prm.n_region + 3, J = J0 + n_nodes;
  switch (n_cvt) {
    case 1: lmm_common_kernel<1><<<J, 256, 0, st>>>(D, prm, H, ctab, node_lams); break;
    case 2: lmm_common_kernel<2><<<J, 256, 0, st>>>(D, prm, H, ctab, node_lams); break;
    case 3: lmm_common_kernel<3><<<J, 256, 0, st>>>(D, prm, H, ctab, node_lams); break;
    default: return cudaErrorInvalidValue;
  }
  if (n_nodes > 0) {
    switch (n_cvt) {
      case 1: lmm_cheb_coef_kernel<1><<<prm.n_region, 512, 0, st>>>(ctab, J0, cheb); break;
      case 2: lmm_cheb_coef_kernel<2><<<prm.n_region, 512, 0, st>>>(ctab, J0, cheb); break;
      case 3: lmm_cheb_coef_kernel<3><<<prm.n_region, 512, 0, st>>>(ctab, J0, cheb); break;
    }
  }
  return cudaGetLastError();
}
int lmm_cheb_nodes() { return V2_CM; }
int lmm_cheb_xnodes() { return V2_XM; }
size_t lmm_cheb_doubles(int n_cvt, int n_region) { return (size_t)V2_CHEB_BASE + (size_t)n_region * (2 * (size_t)v2c_nidx(n_cvt) + 3) * V2_CM; }
size_t lmm_common_record_doubles(int n_cvt) { return (size_t)v2c_stride(n_cvt); }

bool lmm_v2_supported(int n_cvt, int n_region) { return n_cvt >= 1 && n_cvt <= 3 && n_region <= V2_MAX_REGION; }

cudaError_t launch_lmm_assoc_v2(int n_cvt, const LmmConst &D, const LmmParams &prm, const double *UtXt, size_t ldu,
                                int l, gb200_sumstat *out, unsigned int *ticket, int num_sms, cudaStream_t st) {
  switch (n_cvt) {
    case 1: return launch_assoc_v2_nc<1>(D, prm, UtXt, ldu, l, out, ticket, num_sms, st);
    case 2: return launch_assoc_v2_nc<2>(D, prm, UtXt, ldu, l, out, ticket, num_sms, st);
    case 3: return launch_assoc_v2_nc<3>(D, prm, UtXt, ldu, l, out, ticket, num_sms, st);
    default: return cudaErrorInvalidValue;
  }
}

template <int NC>
__global__ void __launch_bounds__(32) lmm_null_kernel(LmmConst D, double l_min, double l_max,
                                                      int n_region, NullOut *out) {
  NullOut r;
  const int nc = NC < 0 ? D.nc_gen : NC;
  null_model<NC>(D, D.Wt + (size_t)nc * D.ldv, l_min, l_max, n_region, r);
  if ((threadIdx.x & 31) == 0) *out = r;
}

// any number of covariates: NC = -1 instantiations (lmm_device.cuh, "generic" section); shared memory holds the
// per-warp sum tables
static size_t gen_smem_bytes(int nc, int warps) { return (size_t)warps * 3 * (size_t)((nc + 3) * (nc + 2) / 2) * sizeof(double); }

static cudaError_t launch_assoc_generic(int n_cvt, LmmConst D, const LmmParams &prm, const double *UtXt, size_t ldu, int l,
                                        gb200_sumstat *out, unsigned int *ticket, int num_sms, cudaStream_t st) {
  D.nc_gen = n_cvt; D.gen_stride = 3 * ((n_cvt + 3) * (n_cvt + 2) / 2);
  const size_t smem = gen_smem_bytes(n_cvt, 4);
  cudaError_t e = cudaFuncSetAttribute(lmm_assoc_kernel<-1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lmm_assoc_kernel<-1>, 128, smem);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) per_sm = 1;
  long want = ((long)l + 3) / 4, grid = (long)num_sms * per_sm;
  if (grid > want) grid = want;
  if (grid < 1) grid = 1;
  e = cudaMemsetAsync(ticket, 0, sizeof(unsigned int), st);
  if (e != cudaSuccess) return e;
  lmm_assoc_kernel<-1><<<(unsigned)grid, 128, smem, st>>>(D, prm, UtXt, ldu, l, out, ticket);
  return cudaGetLastError();
}

// ---- G x E (LMM::AnalyzePlinkGXE / AnalyzeBimbamGXE, src/lmm.cpp:2283-2608): covariates = [W, env, x], tested variable = x * env.
// D.Wt holds c_base rows of U^T W followed by the U^T env row; the per-SNP covariate U^T x comes in through D.xcov.
__global__ void __launch_bounds__(128) lmm_gxe_kernel(LmmConst D, LmmParams prm, const double *__restrict__ UtX1t,
                                                      const double *__restrict__ UtX2t, size_t ldu, int l, int c_base,
                                                      const unsigned char *__restrict__ flip, gb200_sumstat *__restrict__ out,
                                                      unsigned int *__restrict__ ticket) {
  const int lane = threadIdx.x & 31;
  for (;;) {
    unsigned int s = 0;
    if (lane == 0) s = atomicAdd(ticket, 1u);
    s = __shfl_sync(0xffffffffu, s, 0);
    if (s >= (unsigned int)l) break;
    const double *x1 = UtX1t + (size_t)s * ldu, *x2 = UtX2t + (size_t)s * ldu;
    double logl_H0 = 0.0;                                  // stays 0 in modes 3 / 9 like the reference's local (:2301, never assigned there)
    if (prm.a_mode == 2 || prm.a_mode == 4) {              // param0: calc_null with c+2 covariates == alternative with [W, env] and x = U^T x (:2560-2563)
      LmmConst Dn = D; Dn.nc_gen = c_base + 1; Dn.xcov = nullptr;
      RootState R, L;
      calc_lambda_both<-1>(Dn, x1, prm.l_min, prm.l_max, prm.n_region, false, true, R, L);
      logl_H0 = L.logf;
    }
    LmmConst Da = D; Da.nc_gen = c_base + 2; Da.xcov = x1; Da.xcov_idx = c_base + 1;
    LmmParams p2 = prm; p2.logl_mle_H0 = logl_H0; p2.plink_rule = 0;
    gb200_sumstat r;
    analyze_snp<-1>(Da, p2, x2, r);
    if (flip[s]) r.beta = -r.beta;                          // allele flip when the mean genotype exceeds 1 (:2537-2540, :2587-2589)
    if (lane == 0) out[s] = r;
  }
}

cudaError_t launch_lmm_gxe(int c_base, LmmConst D, const LmmParams &prm, const double *UtX1t, const double *UtX2t, size_t ldu, int l,
                           const unsigned char *flip, gb200_sumstat *out, unsigned int *ticket, int num_sms, cudaStream_t st) {
  const int nc = c_base + 2;
  D.nc_gen = nc; D.gen_stride = 3 * ((nc + 3) * (nc + 2) / 2); D.xcov = nullptr; D.xcov_idx = 0;
  const size_t smem = gen_smem_bytes(nc, 4);
  cudaError_t e = cudaFuncSetAttribute(lmm_gxe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lmm_gxe_kernel, 128, smem);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) per_sm = 1;
  long want = ((long)l + 3) / 4, grid = (long)num_sms * per_sm;
  if (grid > want) grid = want;
  if (grid < 1) grid = 1;
  e = cudaMemsetAsync(ticket, 0, sizeof(unsigned int), st);
  if (e != cudaSuccess) return e;
  lmm_gxe_kernel<<<(unsigned)grid, 128, smem, st>>>(D, prm, UtX1t, UtX2t, ldu, l, c_base, flip, out, ticket);
  return cudaGetLastError();
}

// one CTA per SNP row of the mean-imputed SNP-major batch X1 (l x n): flip = mean > 1 -> x := 2 - x; X2 := x * env
__global__ void __launch_bounds__(256) gxe_prepare_kernel(double *__restrict__ X1, double *__restrict__ X2, const double *__restrict__ env,
                                                          int n, unsigned char *__restrict__ flip) {
  __shared__ double part[8];
  __shared__ int do_flip;
  const int s = blockIdx.x, tid = threadIdx.x;
  double *row = X1 + (size_t)s * n, *row2 = X2 + (size_t)s * n;
  double acc = 0.0;
  for (int i = tid; i < n; i += 256) acc += row[i];
  acc = warp_allsum(acc);
  if ((tid & 31) == 0) part[tid >> 5] = acc;
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += part[w];
    do_flip = (t / (double)n > 1.0) ? 1 : 0;                // mean of the imputed vector == mean over the observed entries
    flip[s] = (unsigned char)do_flip;
  }
  __syncthreads();
  const bool f = do_flip != 0;
  for (int i = tid; i < n; i += 256) {
    double v = row[i];
    if (f) { v = 2.0 - v; row[i] = v; }
    row2[i] = v * env[i];
  }
}
cudaError_t launch_gxe_prepare(double *X1, double *X2, const double *env, size_t l, size_t n, unsigned char *flip, cudaStream_t st) {
  if (l == 0) return cudaSuccess;
  gxe_prepare_kernel<<<(unsigned)l, 256, 0, st>>>(X1, X2, env, (int)n, flip);
  return cudaGetLastError();
}

// ---- -lm: linear model without random effect (LM::AnalyzeBimbam / AnalyzePlink, CalcvPv, LmCalcP: src/lm.cpp:224-288, 382-640).
// One CTA per SNP row of the mean-imputed SNP-major batch X: x'x, x'y and W'x in one pass; then
// xPwx = x'x - (W'x)' (W'W)^-1 (W'x), xPwy = x'y - (W'x)' (W'W)^-1 (W'y), and the three tests of LmCalcP.
__global__ void __launch_bounds__(128) lm_kernel(const double *__restrict__ X, int n, int n_cvt, const double *__restrict__ Wt /* c rows of n */,
                                                 const double *__restrict__ y, const double *__restrict__ WtWi, const double *__restrict__ Wty,
                                                 double yPwy, int test_mode, gb200_sumstat *__restrict__ out) {
  __shared__ double sh[4][GB200_MAX_CVT + 2];
  const double *x = X + (size_t)blockIdx.x * n;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double acc[GB200_MAX_CVT + 2];
#pragma unroll
  for (int a = 0; a < GB200_MAX_CVT + 2; ++a) acc[a] = 0.0;
  for (int i = threadIdx.x; i < n; i += 128) {
    const double xi = x[i];
    acc[GB200_MAX_CVT] = fma(xi, xi, acc[GB200_MAX_CVT]);
    acc[GB200_MAX_CVT + 1] = fma(xi, __ldg(y + i), acc[GB200_MAX_CVT + 1]);
    for (int a = 0; a < n_cvt; ++a) acc[a] = fma(__ldg(Wt + (size_t)a * n + i), xi, acc[a]);
  }
#pragma unroll
  for (int a = 0; a < GB200_MAX_CVT + 2; ++a) {
    const double v = warp_allsum(acc[a]);
    if (lane == 0) sh[warp][a] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double wtx[GB200_MAX_CVT];
    for (int a = 0; a < n_cvt; ++a) wtx[a] = sh[0][a] + sh[1][a] + sh[2][a] + sh[3][a];
    double xPwx = sh[0][GB200_MAX_CVT] + sh[1][GB200_MAX_CVT] + sh[2][GB200_MAX_CVT] + sh[3][GB200_MAX_CVT];
    double xPwy = sh[0][GB200_MAX_CVT + 1] + sh[1][GB200_MAX_CVT + 1] + sh[2][GB200_MAX_CVT + 1] + sh[3][GB200_MAX_CVT + 1];
    double d1 = 0.0, d2 = 0.0;
    for (int a = 0; a < n_cvt; ++a) {
      double t = 0.0;
      for (int b = 0; b < n_cvt; ++b) t += WtWi[a * n_cvt + b] * wtx[b];      // WtWiWtx
      d1 += t * wtx[a]; d2 += t * Wty[a];
    }
    xPwx -= d1; xPwy -= d2;
    const double df = (double)n - (double)n_cvt - 1.0;
    const double yPxy = yPwy - xPwy * xPwy / xPwx;
    const double beta = xPwy / xPwx;
    const double se_wald = sqrt(yPxy / (df * xPwx)), se_score = sqrt(yPwy / ((double)n * xPwx));
    gb200_sumstat r;
    r.beta = beta; r.se = (test_mode == 3) ? se_score : se_wald; r.lambda_remle = 0.0; r.lambda_mle = 0.0;
    r.p_wald = fdist_Q_dev(beta * beta / (se_wald * se_wald), 1.0, df);
    r.p_score = fdist_Q_dev(beta * beta / (se_score * se_score), 1.0, df);
    r.p_lrt = chisq1_Q_dev((double)n * (log(yPwy) - log(yPxy)));
    r.logl_H1 = -0.0;
    out[blockIdx.x] = r;
  }
}
cudaError_t launch_lm(const double *X, size_t l, int n, int n_cvt, const double *Wt, const double *y, const double *WtWi, const double *Wty,
                      double yPwy, int test_mode, gb200_sumstat *out, cudaStream_t st) {
  if (l == 0) return cudaSuccess;
  lm_kernel<<<(unsigned)l, 128, 0, st>>>(X, n, n_cvt, Wt, y, WtWi, Wty, yPwy, test_mode, out);
  return cudaGetLastError();
}

template <int NC>
static cudaError_t launch_assoc_nc(const LmmConst &D, const LmmParams &prm, const double *UtXt, size_t ldu,
                                   int l, gb200_sumstat *out, unsigned int *ticket, int num_sms,
                                   cudaStream_t st) {
  int per_sm = 0;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lmm_assoc_kernel<NC>, 128, 0);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) per_sm = 1;
  long want = ((long)l + 3) / 4;
  long grid = (long)num_sms * per_sm;
  if (grid > want) grid = want;
  if (grid < 1) grid = 1;
  e = cudaMemsetAsync(ticket, 0, sizeof(unsigned int), st);
  if (e != cudaSuccess) return e;
  lmm_assoc_kernel<NC><<<(unsigned)grid, 128, 0, st>>>(D, prm, UtXt, ldu, l, out, ticket);
  return cudaGetLastError();
}

cudaError_t launch_lmm_assoc(int n_cvt, const LmmConst &D, const LmmParams &prm, const double *UtXt,
                             size_t ldu, int l, gb200_sumstat *out, unsigned int *ticket, int num_sms,
                             cudaStream_t st) {
  if (n_cvt > 6 || D.nc_gen > 0) return launch_assoc_generic(n_cvt, D, prm, UtXt, ldu, l, out, ticket, num_sms, st);
  switch (n_cvt) {
    case 1: return launch_assoc_nc<1>(D, prm, UtXt, ldu, l, out, ticket, num_sms, st);
    case 2: return launch_assoc_nc<2>(D, prm, UtXt, ldu, l, out, ticket, num_sms, st);
    case 3: return launch_assoc_nc<3>(D, prm, UtXt, ldu, l, out, ticket, num_sms, st);
    case 4: return launch_assoc_nc<4>(D, prm, UtXt, ldu, l, out, ticket, num_sms, st);
    case 5: return launch_assoc_nc<5>(D, prm, UtXt, ldu, l, out, ticket, num_sms, st);
    case 6: return launch_assoc_nc<6>(D, prm, UtXt, ldu, l, out, ticket, num_sms, st);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t launch_lmm_null(int n_cvt, const LmmConst &D, double l_min, double l_max, int n_region,
                            NullOut *out, cudaStream_t st) {
  // null model with c covariates == alternative model with c-1 covariates and x = last covariate
  if (n_cvt > 6 || D.nc_gen > 0) {
    LmmConst G = D;
    G.nc_gen = n_cvt - 1; G.gen_stride = 3 * ((n_cvt + 2) * (n_cvt + 1) / 2);
    const size_t smem = gen_smem_bytes(n_cvt - 1, 1);
    cudaError_t e = cudaFuncSetAttribute(lmm_null_kernel<-1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    lmm_null_kernel<-1><<<1, 32, smem, st>>>(G, l_min, l_max, n_region, out);
    return cudaGetLastError();
  }
  switch (n_cvt) {
    case 1: lmm_null_kernel<0><<<1, 32, 0, st>>>(D, l_min, l_max, n_region, out); break;
    case 2: lmm_null_kernel<1><<<1, 32, 0, st>>>(D, l_min, l_max, n_region, out); break;
    case 3: lmm_null_kernel<2><<<1, 32, 0, st>>>(D, l_min, l_max, n_region, out); break;
    case 4: lmm_null_kernel<3><<<1, 32, 0, st>>>(D, l_min, l_max, n_region, out); break;
    case 5: lmm_null_kernel<4><<<1, 32, 0, st>>>(D, l_min, l_max, n_region, out); break;
    case 6: lmm_null_kernel<5><<<1, 32, 0, st>>>(D, l_min, l_max, n_region, out); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

}  // namespace gb

// ---- diagnostic: the device restatement of the GSL tails on host arrays (include/gemma_b200.h: gb200_cdf_tails) ----------------
namespace gb {
__global__ void cdf_tails_kernel(int kind, const double *x, double nu1, const double *nu2, double *out, size_t count) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  out[i] = (kind == 0) ? fdist_Q_dev(x[i], nu1, nu2[i]) : chisq1_Q_dev(x[i]);
}
}  // namespace gb

extern "C" int gb200_cdf_tails(gb200_ctx *c, int kind, const double *x, double nu1, const double *nu2, double *out, size_t count) {
  if (!c) return GB200_ERR_ARG;
  if (!x || !out || count == 0 || (kind != 0 && kind != 1) || (kind == 0 && !nu2)) return gb::set_err(c, GB200_ERR_ARG, "gb200_cdf_tails: bad argument");
  gb::DevBuf dx, dn, dout;
  const size_t by = count * sizeof(double);
  GB_CUDA(c, dx.reserve(by)); GB_CUDA(c, dn.reserve(by)); GB_CUDA(c, dout.reserve(by));
  GB_CUDA(c, cudaMemcpyAsync(dx.p, x, by, cudaMemcpyHostToDevice, c->stream));
  if (kind == 0) GB_CUDA(c, cudaMemcpyAsync(dn.p, nu2, by, cudaMemcpyHostToDevice, c->stream));
  gb::cdf_tails_kernel<<<(unsigned)((count + 127) / 128), 128, 0, c->stream>>>(kind, dx.as<double>(), nu1, dn.as<double>(), dout.as<double>(), count);
  GB_CUDA(c, cudaGetLastError());
  GB_CUDA(c, cudaMemcpyAsync(out, dout.p, by, cudaMemcpyDeviceToHost, c->stream));
  GB_CUDA(c, cudaStreamSynchronize(c->stream));
  dx.release(); dn.release(); dout.release();
  return GB200_OK;
}

// ---- exact x-sums at l_mle_null (LmmConst::xex): the SNP-independent vectors v_q = U (h (.) q), q over (w_1..w_c, y) ------------
namespace gb {
__global__ void __launch_bounds__(256) lmm_hq_kernel(LmmConst D, int n_cvt, double lam, double *__restrict__ a) {
  // a[q][i] = q_i / (lam * delta_i + 1), rows of n_c doubles, zero in the padding
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= D.n_c) return;
  const bool in = i < D.n;
  const double h = in ? 1.0 / fma(lam, __ldg(D.delta + i), 1.0) : 0.0;
  for (int q = 0; q < n_cvt; ++q) a[(size_t)q * D.n_c + i] = in ? h * __ldg(D.Wt + (size_t)q * D.ldv + i) : 0.0;
  a[(size_t)n_cvt * D.n_c + i] = in ? h * __ldg(D.y + i) : 0.0;
}
// one warp per individual j: v[q][j] = sum_i U[j][i] a[q][i]
__global__ void __launch_bounds__(256) lmm_vnull_kernel(const double *__restrict__ U, int n, int n_c, int nq, const double *__restrict__ a,
                                                        double *__restrict__ v) {
  const int j = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (j >= n) return;
  const double *u = U + (size_t)j * n;
  double acc[GB200_MAX_CVT + 1];
  for (int q = 0; q < nq; ++q) acc[q] = 0.0;
  for (int i = lane; i < n; i += 32) {
    const double x = __ldg(u + i);
    for (int q = 0; q < nq; ++q) acc[q] = fma(x, __ldg(a + (size_t)q * n_c + i), acc[q]);
  }
  for (int q = 0; q < nq; ++q) {
    const double t = warp_allsum(acc[q]);
    if (lane == 0) v[(size_t)q * n_c + j] = t;
  }
}
}  // namespace gb

// v: (n_cvt + 1) rows of n_c doubles; scratch: same size
cudaError_t gb::launch_lmm_vnull(int n_cvt, const LmmConst &D, double lam, const double *U, double *scratch, double *v, cudaStream_t st) {
  lmm_hq_kernel<<<(D.n_c + 255) / 256, 256, 0, st>>>(D, n_cvt, lam, scratch);
  lmm_vnull_kernel<<<(D.n + 7) / 8, 256, 0, st>>>(U, D.n, D.n_c, n_cvt + 1, scratch, v);
  return cudaGetLastError();
}

// ---- exact linear x-sums (LmmConst::xsum): the columns a_{b,k,q} = h_b^k (.) q of A (n x ncol, row-major), q over (w_1..w_c, y);
// block b: the shared rows 0..J0-1 of Hrows, then the x-node rows (which start at row x0); last block: unit weights.
namespace gb {
__global__ void __launch_bounds__(256) lmm_acols_kernel(LmmConst D, int n_cvt, const double *__restrict__ H, int J0, int x0, int nblocks,
                                                        double *__restrict__ A, int ncol, const int *__restrict__ patch_idx, int npatch, int patch0) {
  const int i = blockIdx.x * 256 + threadIdx.x;            // individual (eigen-coordinate)
  const int b = blockIdx.y;                                // block; b == nblocks: unit weights (+ the one-hot columns of the patched eigenvectors)
  if (i >= D.n) return;
  const int nq = n_cvt + 1;
  if (b == nblocks) {
    // columns [patch0, patch0 + npatch): e_{idx}: V = U A then carries eigenvector idx itself -> exact U^T x entries for the null and
    // the leading eigenvectors (their digit-plane rounding is coherent across individuals: constant / piecewise-constant vectors)
    double *r0 = A + (size_t)i * ncol;
    for (int k = 0; k < npatch; ++k) r0[patch0 + k] = (__ldg(patch_idx + k) == i) ? 1.0 : 0.0;
  }
  double q[GB200_MAX_CVT + 1];
  for (int a = 0; a < n_cvt; ++a) q[a] = __ldg(D.Wt + (size_t)a * D.ldv + i);
  q[n_cvt] = __ldg(D.y + i);
  double *row = A + (size_t)i * ncol;
  if (b == nblocks) {
    for (int a = 0; a < nq; ++a) row[(size_t)nblocks * 2 * nq + a] = q[a];
    return;
  }
  const int r = b < J0 ? b : x0 + (b - J0);
  const double h = __ldg(H + (size_t)r * D.n_c + i);
  for (int a = 0; a < nq; ++a) { row[((size_t)b * 2) * nq + a] = h * q[a]; row[((size_t)b * 2 + 1) * nq + a] = h * h * q[a]; }
}
}  // namespace gb
cudaError_t gb::launch_lmm_acols(int n_cvt, const LmmConst &D, const double *H, int J0, int x0, int nblocks, double *A, int ncol,
                                 const int *patch_idx, int npatch, int patch0, cudaStream_t st) {
  dim3 grid((D.n + 255) / 256, nblocks + 1);
  lmm_acols_kernel<<<grid, 256, 0, st>>>(D, n_cvt, H, J0, x0, nblocks, A, ncol, patch_idx, npatch, patch0);
  return cudaGetLastError();
}
