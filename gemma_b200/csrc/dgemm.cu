// dgemm.cu -- FP64 GEMM seam of the reference: fast_dgemm / fast_eigen_dgemm
// (src/fastblas.cpp:175-236 -> cblas_dgemm).  Exact-FP64 path used for dosage-valued
// (non-integer) genotypes, for U^T W / U^T y, and as the generic kinship accumulator;
// integer genotypes take the tensor-core path in i8gemm_sm100.cu.
//
// 128x128x16 CTA tile, 8 warps x (8x4) DMMA m8n8k4 tiles on the FP64 tensor pipe, operands staged in two shared-memory
// stages k-major (register prefetch of the next k-step).  Element strides are arbitrary (covers N/T on either operand and gsl_matrix sub-views with
// tda != size2).
#include "common.cuh"

namespace gb {

constexpr int BM = 128, BN = 128, BK = 16;
constexpr int PITCH = 132;      // doubles per k-row of the staged tiles: 132 = 4 (mod 16) keeps the DMMA fragment loads conflict-free

// D(8x8) += A(8x4) * B(4x8) on the FP64 tensor pipe.  Fragment layout (PTX mma.m8n8k4.f64):
// a: row = lane/4, k = lane%4 ; b: k = lane%4, col = lane/4 ; c/d: row = lane/4, cols = 2*(lane%4) + {0,1}
__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// 128x128x16 CTA tile, 8 warps in a 2 (m) x 4 (n) grid, each warp 64x32 = 8x4 DMMA tiles (64 accumulators / thread).
// Operands are staged k-major in shared memory; element strides are arbitrary (N/T, gsl sub-views); the loader picks
// the thread mapping that makes the unit-stride dimension the fastest varying one (coalesced).
template <bool LOWER_ONLY>
__global__ void __launch_bounds__(256) dgemm_kernel(size_t M, size_t N, size_t K, double alpha,
                                                    const double *__restrict__ A, size_t sam, size_t sak,
                                                    const double *__restrict__ B, size_t sbk, size_t sbn,
                                                    double beta, double *__restrict__ C, size_t ldc) {
  // two stages of (A tile | B tile), k-major: the global loads of k-step t + 1 are issued into registers before the 128 DMMAs of
  // k-step t and stored into the other stage afterwards -- one barrier per k-step, global latency behind the tensor work
  // (round 1: synchronous loads, two barriers per k-step, 20 TFLOP/s)
  extern __shared__ __align__(16) double dg_smem[];
  double (*As)[BK][PITCH] = reinterpret_cast<double (*)[BK][PITCH]>(dg_smem);
  double (*Bs)[BK][PITCH] = reinterpret_cast<double (*)[BK][PITCH]>(dg_smem + 2 * BK * PITCH);
  const size_t m0 = (size_t)blockIdx.y * BM, n0 = (size_t)blockIdx.x * BN;
  if (LOWER_ONLY && n0 > m0 + BM - 1) return;     // tile entirely above the diagonal
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = (warp >> 2) * 64, wn = (warp & 3) * 32;     // warp tile origin inside the CTA tile
  const int fr = lane >> 2, fk = lane & 3;
  double acc[8][4][2];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

  const bool a_kfast = (sak == 1);
  const bool b_nfast = (sbn == 1);
  constexpr int RA = (BM * BK) / 256, RB = (BN * BK) / 256;
  double ra[RA], rb[RB];
  auto gload = [&](size_t k0) {
#pragma unroll
    for (int r = 0; r < RA; ++r) {
      int mm, kk;
      if (a_kfast) { kk = tid % BK; mm = tid / BK + r * (256 / BK); }
      else { mm = tid % BM; kk = tid / BM + r * (256 / BM); }
      const size_t gm = m0 + mm, gk = k0 + kk;
      ra[r] = (gm < M && gk < K) ? A[gm * sam + gk * sak] : 0.0;
    }
#pragma unroll
    for (int r = 0; r < RB; ++r) {
      int nn, kk;
      if (b_nfast) { nn = tid % BN; kk = tid / BN + r * (256 / BN); }
      else { kk = tid % BK; nn = tid / BK + r * (256 / BK); }
      const size_t gn = n0 + nn, gk = k0 + kk;
      rb[r] = (gn < N && gk < K) ? B[gk * sbk + gn * sbn] : 0.0;
    }
  };
  auto sstore = [&](int st) {
#pragma unroll
    for (int r = 0; r < RA; ++r) {
      int mm, kk;
      if (a_kfast) { kk = tid % BK; mm = tid / BK + r * (256 / BK); }
      else { mm = tid % BM; kk = tid / BM + r * (256 / BM); }
      As[st][kk][mm] = ra[r];
    }
#pragma unroll
    for (int r = 0; r < RB; ++r) {
      int nn, kk;
      if (b_nfast) { nn = tid % BN; kk = tid / BN + r * (256 / BN); }
      else { kk = tid % BK; nn = tid / BK + r * (256 / BK); }
      Bs[st][kk][nn] = rb[r];
    }
  };
  gload(0);
  sstore(0);
  __syncthreads();
  int cur = 0;
  for (size_t k0 = 0; k0 < K; k0 += BK) {
    const bool more = k0 + BK < K;
    if (more) gload(k0 + BK);
#pragma unroll
    for (int kk = 0; kk < BK; kk += 4) {
      double a[8], b[4];
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = As[cur][kk + fk][wm + i * 8 + fr];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[cur][kk + fk][wn + j * 8 + fr];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
    if (more) sstore(cur ^ 1);      // the other stage was last read one k-step ago, before the previous barrier
    __syncthreads();
    cur ^= 1;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const size_t gm = m0 + wm + i * 8 + fr;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const size_t gn = n0 + wn + j * 8 + fk * 2 + q;
        if (gn >= N) continue;
        if (LOWER_ONLY && gn > gm) continue;
        double *c = C + gm * ldc + gn;
        const double prev = (beta == 0.0) ? 0.0 : beta * (*c);
        *c = fma(alpha, acc[i][j][q], prev);
      }
    }
  }
}

cudaError_t launch_dgemm(size_t M, size_t N, size_t K, double alpha, const double *A, size_t sam, size_t sak,
                         const double *B, size_t sbk, size_t sbn, double beta, double *C, size_t ldc,
                         bool lower_only, cudaStream_t st) {
  if (M == 0 || N == 0) return cudaSuccess;
  dim3 grid((unsigned)((N + BN - 1) / BN), (unsigned)((M + BM - 1) / BM));
  constexpr size_t smem = 4 * (size_t)BK * PITCH * sizeof(double);         // 2 stages x (A | B) = 66 KB
  cudaError_t e;
  if (lower_only) {
    e = cudaFuncSetAttribute(dgemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    dgemm_kernel<true><<<grid, 256, smem, st>>>(M, N, K, alpha, A, sam, sak, B, sbk, sbn, beta, C, ldc);
  } else {
    e = cudaFuncSetAttribute(dgemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    dgemm_kernel<false><<<grid, 256, smem, st>>>(M, N, K, alpha, A, sam, sak, B, sbk, sbn, beta, C, ldc);
  }
  return cudaGetLastError();
}

// mirror the lower triangle into the upper one (PlinkKin symmetrises explicitly,
// src/gemma_io.cpp:1724-1729; the dgemm of BimbamKin produces both halves)
__global__ void symmetrize_kernel(double *C, size_t n, size_t ldc) {
  __shared__ double tile[32][33];
  const size_t bi = blockIdx.y, bj = blockIdx.x;
  if (bj > bi) return;
  const size_t i = bi * 32 + threadIdx.y, j = bj * 32 + threadIdx.x;
  // read lower tile (bi,bj), write transposed into (bj,bi)
  for (int r = 0; r < 32; r += 8) {
    const size_t ii = i + r;
    if (ii < n && j < n) tile[threadIdx.y + r][threadIdx.x] = C[ii * ldc + j];
  }
  __syncthreads();
  const size_t oi = bj * 32 + threadIdx.y, oj = bi * 32 + threadIdx.x;
  for (int r = 0; r < 32; r += 8) {
    const size_t ii = oi + r;
    if (ii < n && oj < n && oj > ii) C[ii * ldc + oj] = tile[threadIdx.x][threadIdx.y + r];
  }
}

cudaError_t launch_symmetrize_from_lower(double *C, size_t n, size_t ldc, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  dim3 grid((unsigned)((n + 31) / 32), (unsigned)((n + 31) / 32)), block(32, 8);
  symmetrize_kernel<<<grid, block, 0, st>>>(C, n, ldc);
  return cudaGetLastError();
}

__global__ void scale_kernel(double *C, size_t count, double alpha) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < count; i += stride) C[i] *= alpha;
}
cudaError_t launch_scale(double *C, size_t count, double alpha, cudaStream_t st) {
  if (count == 0) return cudaSuccess;
  scale_kernel<<<148 * 8, 256, 0, st>>>(C, count, alpha);
  return cudaGetLastError();
}

__global__ void transpose_kernel(const double *__restrict__ in, size_t rows, size_t cols, size_t ldi,
                                 double *__restrict__ out, size_t ldo) {
  __shared__ double tile[32][33];
  const size_t r0 = (size_t)blockIdx.y * 32, c0 = (size_t)blockIdx.x * 32;
  for (int r = threadIdx.y; r < 32; r += 8) {
    const size_t rr = r0 + r, cc = c0 + threadIdx.x;
    if (rr < rows && cc < cols) tile[r][threadIdx.x] = in[rr * ldi + cc];
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += 8) {
    const size_t orow = c0 + r, ocol = r0 + threadIdx.x;   // out is cols x rows
    if (orow < cols && ocol < rows) out[orow * ldo + ocol] = tile[threadIdx.x][r];
  }
}
cudaError_t launch_transpose(const double *in, size_t rows, size_t cols, size_t ldi, double *out, size_t ldo,
                             cudaStream_t st) {
  if (rows == 0 || cols == 0) return cudaSuccess;
  dim3 grid((unsigned)((cols + 31) / 32), (unsigned)((rows + 31) / 32)), block(32, 8);
  transpose_kernel<<<grid, block, 0, st>>>(in, rows, cols, ldi, out, ldo);
  return cudaGetLastError();
}

}  // namespace gb
