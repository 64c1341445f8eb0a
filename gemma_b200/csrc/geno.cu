// geno.cu -- genotype ingest kernels (HBM-bound byte/FP64 streaming work; one CTA per SNP
// row, coalesced along individuals, block reductions for the per-SNP statistics).
//
//   kin_transform : per-SNP mean/var over non-missing, mean imputation, centring, optional
//                   1/sqrt(var)                      BimbamKin src/gemma_io.cpp:1511-1538,
//                                                    PlinkKin  src/gemma_io.cpp:1688-1706
//   lmm_impute    : mean imputation only            LMM::Analyze src/lmm.cpp:1590-1618
//   bed_decode    : PLINK 2-bit -> FP64 (NaN = missing) with analysed-individual gather
//                                                    src/lmm.cpp:1783-1817, gemma_io.cpp:1665-1682
//   center_matrix : CenterMatrix                     src/mathfunc.cpp:147-177
#include "common.cuh"

namespace gb {

__device__ __forceinline__ double block_sum(double v, double *sh) {
  // blockDim.x == 256
  v = warp_allsum(v);
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  double t = (lane < 8) ? sh[lane] : 0.0;
  t = warp_allsum(t);
  return t;   // identical in every thread
}

__global__ void __launch_bounds__(256) kin_transform_kernel(double *G, size_t n, size_t ldg, int k_mode) {
  __shared__ double sh[8];
  double *g = G + (size_t)blockIdx.x * ldg;
  double s = 0.0, ss = 0.0, miss = 0.0;
  for (size_t i = threadIdx.x; i < n; i += 256) {
    const double v = g[i];
    if (isnan(v)) miss += 1.0; else { s += v; ss += v * v; }
  }
  s = block_sum(s, sh); ss = block_sum(ss, sh); miss = block_sum(miss, sh);
  // src/gemma_io.cpp:1511-1514
  const double mean = s / ((double)n - miss);
  double var = ss + mean * mean * miss;
  var /= (double)n;
  var -= mean * mean;
  const bool scale = (k_mode == 2 && var != 0);
  const double inv_sd = scale ? 1.0 / sqrt(var) : 1.0;
  for (size_t i = threadIdx.x; i < n; i += 256) {
    double v = g[i];
    if (isnan(v)) v = mean;
    v -= mean;
    if (scale) v *= inv_sd;
    g[i] = v;
  }
}

cudaError_t launch_kin_transform(double *G, size_t l, size_t n, size_t ldg, int k_mode, cudaStream_t st) {
  if (l == 0) return cudaSuccess;
  kin_transform_kernel<<<(unsigned)l, 256, 0, st>>>(G, n, ldg, k_mode);
  return cudaGetLastError();
}

__global__ void __launch_bounds__(256) lmm_impute_kernel(double *G, size_t n, size_t ldg) {
  __shared__ double sh[8];
  double *g = G + (size_t)blockIdx.x * ldg;
  double s = 0.0, miss = 0.0;
  for (size_t i = threadIdx.x; i < n; i += 256) {
    const double v = g[i];
    if (isnan(v)) miss += 1.0; else s += v;
  }
  s = block_sum(s, sh); miss = block_sum(miss, sh);
  if (miss == 0.0) return;
  const double mean = s / ((double)n - miss);
  for (size_t i = threadIdx.x; i < n; i += 256) {
    if (isnan(g[i])) g[i] = mean;
  }
}

cudaError_t launch_lmm_impute(double *G, size_t l, size_t n, size_t ldg, cudaStream_t st) {
  if (l == 0) return cudaSuccess;
  lmm_impute_kernel<<<(unsigned)l, 256, 0, st>>>(G, n, ldg);
  return cudaGetLastError();
}

// idx == nullptr: identity (all ni_total individuals analysed, n_out == ni_total)
__global__ void __launch_bounds__(256) bed_decode_kernel(const unsigned char *__restrict__ bed,
                                                         size_t bytes_per_snp, const int *__restrict__ idx,
                                                         size_t n_out, double *__restrict__ G, size_t ldg) {
  const unsigned char *row = bed + (size_t)blockIdx.x * bytes_per_snp;
  double *g = G + (size_t)blockIdx.x * ldg;
  for (size_t p = threadIdx.x; p < n_out; p += 256) {
    const size_t j = idx ? (size_t)idx[p] : p;
    const unsigned b = (unsigned)row[j >> 2] >> (2 * (j & 3));
    const unsigned lo = b & 1u, hi = (b >> 1) & 1u;
    double v;
    if (lo == 0) v = hi == 0 ? 2.0 : 1.0;
    else v = hi == 1 ? 0.0 : nan("");
    g[p] = v;
  }
}

cudaError_t launch_bed_decode(const unsigned char *bed, size_t l, size_t bytes_per_snp, const int *idx,
                              size_t n_out, double *G, size_t ldg, cudaStream_t st) {
  if (l == 0) return cudaSuccess;
  bed_decode_kernel<<<(unsigned)l, 256, 0, st>>>(bed, bytes_per_snp, idx, n_out, G, ldg);
  return cudaGetLastError();
}

// SNP QC statistics from PLINK rows (ReadFile_bed counting pass + r2 terms, src/gemma_io.cpp:951-1046)
__global__ void __launch_bounds__(128) qc_bed_kernel(const unsigned char *__restrict__ bed, size_t bytes_per_snp,
                                                     const int *__restrict__ idx, int n_test, const double *__restrict__ W,
                                                     const double *__restrict__ WtWi, int n_cvt, gb200_snpqc *__restrict__ out) {
  __shared__ int sh_i[4][4];
  __shared__ double sh_d[4][GB200_MAX_CVT + 1];
  const unsigned char *row = bed + (size_t)blockIdx.x * bytes_per_snp;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int cnt[4] = {0, 0, 0, 0};                  // n_miss, n_0, n_1, n_2
  for (int p = threadIdx.x; p < n_test; p += 128) {
    const size_t j = idx ? (size_t)idx[p] : (size_t)p;
    const unsigned b = ((unsigned)row[j >> 2] >> (2 * (j & 3))) & 3u;
    // bits (hi,lo): 00 -> 2 copies, 10 -> 1, 11 -> 0, 01 -> missing
    cnt[b == 1u ? 0 : b == 3u ? 1 : b == 2u ? 2 : 3]++;
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    int v = cnt[q];
    for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    if (lane == 0) sh_i[warp][q] = v;
  }
  __syncthreads();
  int tot[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) tot[q] = sh_i[0][q] + sh_i[1][q] + sh_i[2][q] + sh_i[3][q];
  const double maf = (double)(tot[2] + 2 * tot[3]) / (2.0 * (double)(n_test - tot[0]));
  double v_x = 0.0, v_w = 0.0;
  if (W) {
    double acc[GB200_MAX_CVT + 1];
#pragma unroll
    for (int a = 0; a <= GB200_MAX_CVT; ++a) acc[a] = 0.0;
    for (int p = threadIdx.x; p < n_test; p += 128) {
      const size_t j = idx ? (size_t)idx[p] : (size_t)p;
      const unsigned b = ((unsigned)row[j >> 2] >> (2 * (j & 3))) & 3u;
      const double x = b == 0u ? 2.0 : b == 2u ? 1.0 : b == 3u ? 0.0 : maf * 2.0;
      acc[GB200_MAX_CVT] += x * x;
      for (int a = 0; a < n_cvt; ++a) acc[a] += W[(size_t)p * n_cvt + a] * x;
    }
    __syncthreads();
#pragma unroll
    for (int a = 0; a <= GB200_MAX_CVT; ++a) {
      const double v = warp_allsum(acc[a]);
      if (lane == 0) sh_d[warp][a] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double wtx[GB200_MAX_CVT];
      for (int a = 0; a < n_cvt; ++a) wtx[a] = sh_d[0][a] + sh_d[1][a] + sh_d[2][a] + sh_d[3][a];
      v_x = sh_d[0][GB200_MAX_CVT] + sh_d[1][GB200_MAX_CVT] + sh_d[2][GB200_MAX_CVT] + sh_d[3][GB200_MAX_CVT];
      for (int a = 0; a < n_cvt; ++a) {
        double t = 0.0;
        for (int b2 = 0; b2 < n_cvt; ++b2) t += WtWi[a * n_cvt + b2] * wtx[b2];
        v_w += wtx[a] * t;
      }
    }
  }
  if (threadIdx.x == 0) {
    gb200_snpqc r;
    r.n_miss = tot[0]; r.n_0 = tot[1]; r.n_1 = tot[2]; r.n_2 = tot[3]; r.maf = maf; r.v_x = v_x; r.v_w = v_w;
    out[blockIdx.x] = r;
  }
}

cudaError_t launch_qc_bed(const unsigned char *bed, size_t l, size_t bytes_per_snp, const int *idx, int n_test,
                          const double *W, const double *WtWi, int n_cvt, gb200_snpqc *out, cudaStream_t st) {
  if (l == 0) return cudaSuccess;
  qc_bed_kernel<<<(unsigned)l, 128, 0, st>>>(bed, bytes_per_snp, idx, n_test, W, WtWi, n_cvt, out);
  return cudaGetLastError();
}

// CenterMatrix: G <- G - (Gw 1^T + 1 Gw^T)/n + (1^T G 1 / n^2) 11^T, evaluated from the
// upper triangle and mirrored (the reference updates the upper triangle with dsyr2/dsyr
// and copies it to the lower one, src/mathfunc.cpp:157-171).
__global__ void __launch_bounds__(256) row_sum_kernel(const double *__restrict__ G, size_t n, size_t ldg,
                                                      double *__restrict__ row_sums) {
  __shared__ double sh[8];
  const double *g = G + (size_t)blockIdx.x * ldg;
  double s = 0.0;
  for (size_t j = threadIdx.x; j < n; j += 256) s += g[j];
  s = block_sum(s, sh);
  if (threadIdx.x == 0) row_sums[blockIdx.x] = s;
}
__global__ void __launch_bounds__(256) total_kernel(const double *__restrict__ row_sums, size_t n,
                                                    double *__restrict__ total) {
  __shared__ double sh[8];
  double s = 0.0;
  for (size_t j = threadIdx.x; j < n; j += 256) s += row_sums[j];
  s = block_sum(s, sh);
  if (threadIdx.x == 0) *total = s;
}
__global__ void center_apply_kernel(double *G, size_t n, size_t ldg, const double *__restrict__ row_sums,
                                    const double *__restrict__ total) {
  // one thread per (i, j >= i) pair of a 32x32 tile pair; tiles below the diagonal idle
  const size_t bi = blockIdx.y, bj = blockIdx.x;
  if (bj < bi) return;
  const double alpha = -1.0 / (double)n, beta = (*total) / ((double)n * (double)n);
  const size_t j = bj * 32 + threadIdx.x;
  for (int r = threadIdx.y; r < 32; r += 8) {
    const size_t i = bi * 32 + r;
    if (i < n && j < n && j >= i) {
      double v = G[i * ldg + j];
      v += alpha * (row_sums[i] + row_sums[j]);
      v += beta;
      G[i * ldg + j] = v;
      if (j != i) G[j * ldg + i] = v;
    }
  }
}

cudaError_t launch_center_matrix(double *G, size_t n, size_t ldg, double *row_sums, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  // row_sums has n+1 doubles: [0..n) sums, [n] grand total
  row_sum_kernel<<<(unsigned)n, 256, 0, st>>>(G, n, ldg, row_sums);
  total_kernel<<<1, 256, 0, st>>>(row_sums, n, row_sums + n);
  dim3 grid((unsigned)((n + 31) / 32), (unsigned)((n + 31) / 32)), block(32, 8);
  center_apply_kernel<<<grid, block, 0, st>>>(G, n, ldg, row_sums, row_sums + n);
  return cudaGetLastError();
}

}  // namespace gb
