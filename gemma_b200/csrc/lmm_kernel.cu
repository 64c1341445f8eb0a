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

template <int NC>
__global__ void __launch_bounds__(32) lmm_null_kernel(LmmConst D, double l_min, double l_max,
                                                      int n_region, NullOut *out) {
  NullOut r;
  null_model<NC>(D, D.Wt + (size_t)NC * D.ldv, l_min, l_max, n_region, r);
  if ((threadIdx.x & 31) == 0) *out = r;
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
