// eigh.cu -- once-per-run symmetric eigendecomposition of the (centred) kinship matrix on
// the device.  Replaces EigenDecomp_Zeroed -> lapack_eigen_symmv -> dsyevr_
// (src/lapack.cpp:149-291).  The O(n^3) factorisation itself is cuSOLVER's syevd (64-bit
// API so the workspace may exceed 2^31 elements at n = 50 000); centring, the transpose
// that puts eigenvectors in the COLUMNS of row-major U (lapack.cpp:228) and the <1e-10
// zeroing (lapack.cpp:268-269) are ours.
#include "common.cuh"
#include <cusolverDn.h>

namespace gb {

__global__ void zero_small_kernel(double *eval, size_t n, double *stats) {
  // stats[0] = sum (after zeroing), stats[1] = #zero, stats[2] = #(< -1e-10 before zeroing: always 0 after)
  __shared__ double sh[3][8];
  double s = 0.0, nz = 0.0, nneg = 0.0;
  for (size_t i = threadIdx.x; i < n; i += blockDim.x) {
    double v = eval[i];
    if (v < 1e-10) { v = 0.0; eval[i] = 0.0; }
    if (v == 0.0) nz += 1.0;
    s += v;
  }
  s = warp_allsum(s); nz = warp_allsum(nz); nneg = warp_allsum(nneg);
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { sh[0][w] = s; sh[1][w] = nz; sh[2][w] = nneg; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, b = 0, c = 0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) { a += sh[0][k]; b += sh[1][k]; c += sh[2][k]; }
    stats[0] = a; stats[1] = b; stats[2] = c;
  }
}

}  // namespace gb

using namespace gb;

extern "C" int gb200_eigh(gb200_ctx *ctx, double *G, size_t n, size_t ldg, int center, double *U,
                          size_t ldu, double *eval, double *trace_G, int *n_zero, int *n_negative) {
  if (!ctx) return GB200_ERR_ARG;
  if (!G || !U || !eval || n == 0 || ldg < n || ldu < n) return set_err(ctx, GB200_ERR_ARG, "gb200_eigh: bad argument");
  cudaStream_t st = ctx->stream;
  DevBuf dA, dV, dW, dWork, dInfo, dRow;
  int rc = GB200_OK;
  cusolverDnHandle_t h = nullptr;
  cusolverDnParams_t params = nullptr;
  void *hWork = nullptr;
  auto cleanup = [&]() {
    dA.release(); dV.release(); dW.release(); dWork.release(); dInfo.release(); dRow.release();
    if (params) cusolverDnDestroyParams(params);
    if (h) cusolverDnDestroy(h);
    if (hWork) free(hWork);
  };
#define EIGH_CUDA(call)                                                                              \
  do {                                                                                               \
    cudaError_t _e = (call);                                                                         \
    if (_e != cudaSuccess) {                                                                         \
      cleanup();                                                                                     \
      return set_err(ctx, GB200_ERR_CUDA, std::string(#call " failed: ") + cudaGetErrorString(_e)); \
    }                                                                                                \
  } while (0)
#define EIGH_SOLVER(call)                                                                  \
  do {                                                                                     \
    cusolverStatus_t _s = (call);                                                          \
    if (_s != CUSOLVER_STATUS_SUCCESS) {                                                   \
      cleanup();                                                                           \
      return set_err(ctx, GB200_ERR_NUMERIC, std::string(#call " failed, cusolver status ") + \
                                                 std::to_string((int)_s));                 \
    }                                                                                      \
  } while (0)

  EIGH_CUDA(dA.reserve(n * n * sizeof(double)));
  EIGH_CUDA(dW.reserve((n + 4) * sizeof(double)));
  EIGH_CUDA(dInfo.reserve(sizeof(int)));
  EIGH_CUDA(cudaMemcpy2DAsync(dA.p, n * sizeof(double), G, ldg * sizeof(double), n * sizeof(double), n,
                              cudaMemcpyHostToDevice, st));
  {
    ProfScope ps(ctx, "eigh");
    if (center) {
      EIGH_CUDA(dRow.reserve((n + 1) * sizeof(double)));
      EIGH_CUDA(launch_center_matrix(dA.as<double>(), n, n, dRow.as<double>(), st));
    }
    EIGH_SOLVER(cusolverDnCreate(&h));
    EIGH_SOLVER(cusolverDnSetStream(h, st));
    EIGH_SOLVER(cusolverDnCreateParams(&params));
    size_t wdev = 0, whost = 0;
    // Row-major symmetric == column-major symmetric.  The reference hands its row-major G to
    // Fortran with UPLO='L' (src/lapack.cpp:205), i.e. the row-major UPPER triangle; for
    // cuSOLVER's column-major view that is CUBLAS_FILL_MODE_LOWER as well.
    EIGH_SOLVER(cusolverDnXsyevd_bufferSize(h, params, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER,
                                            (int64_t)n, CUDA_R_64F, dA.p, (int64_t)n, CUDA_R_64F, dW.p,
                                            CUDA_R_64F, &wdev, &whost));
    EIGH_CUDA(dWork.reserve(wdev ? wdev : 8));
    if (whost) { hWork = malloc(whost); if (!hWork) { cleanup(); return set_err(ctx, GB200_ERR_NOMEM, "host workspace"); } }
    EIGH_SOLVER(cusolverDnXsyevd(h, params, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, (int64_t)n,
                                 CUDA_R_64F, dA.p, (int64_t)n, CUDA_R_64F, dW.p, CUDA_R_64F, dWork.p, wdev,
                                 hWork, whost, dInfo.as<int>()));
    int info = 0;
    EIGH_CUDA(cudaMemcpyAsync(&info, dInfo.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    EIGH_CUDA(cudaStreamSynchronize(st));
    if (info != 0) {
      cleanup();
      return set_err(ctx, GB200_ERR_NUMERIC, "gb200_eigh: syevd did not converge, info=" + std::to_string(info));
    }
    dWork.release();
    // cuSOLVER leaves eigenvector j in column j of the column-major array == row j of the
    // row-major view; transpose so that U[i][j] = v_j[i] (eigenvectors in columns).
    EIGH_CUDA(dV.reserve(n * n * sizeof(double)));
    EIGH_CUDA(launch_transpose(dA.as<double>(), n, n, n, dV.as<double>(), n, st));
    zero_small_kernel<<<1, 256, 0, st>>>(dW.as<double>(), n, dW.as<double>() + n);
    EIGH_CUDA(cudaGetLastError());
  }
  double stats[3];
  EIGH_CUDA(cudaMemcpy2DAsync(U, ldu * sizeof(double), dV.p, n * sizeof(double), n * sizeof(double), n,
                              cudaMemcpyDeviceToHost, st));
  EIGH_CUDA(cudaMemcpyAsync(eval, dW.p, n * sizeof(double), cudaMemcpyDeviceToHost, st));
  EIGH_CUDA(cudaMemcpyAsync(stats, dW.as<double>() + n, 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
  EIGH_CUDA(cudaStreamSynchronize(st));
  if (trace_G) *trace_G = stats[0] / (double)n;
  if (n_zero) *n_zero = (int)stats[1];
  if (n_negative) *n_negative = (int)stats[2];
  cleanup();
  return rc;
}
