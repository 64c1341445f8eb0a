// eigh.cu -- once-per-run symmetric eigendecomposition of the (centred) kinship matrix on
// the device.  Replaces EigenDecomp_Zeroed -> lapack_eigen_symmv -> dsyevr_
// (src/lapack.cpp:149-291).  The O(n^3) factorisation itself is cuSOLVER's syevd (cusolverDnXsyevd
// up to n = 32768, where that entry point stops; cusolverMgSyevd on this one device beyond); centring, the transpose
// that puts eigenvectors in the COLUMNS of row-major U (lapack.cpp:228) and the <1e-10
// zeroing (lapack.cpp:268-269) are ours.
#include "common.cuh"
#include <cusolverDn.h>
#include <cusolverMg.h>
#include <dlfcn.h>

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


// ---- cusolverMgSyevd on ONE device -----------------------------------------------------------------------------------------
// cusolverDnXsyevd refuses n > 32768 (CUSOLVER_STATUS_INVALID_VALUE from its bufferSize; probed on CUDA 12.9: n = 40 000, 46 340,
// 50 000, 65 536 all fail, as do the legacy Dsyevd / Dormtr entry points), and BASELINE config 4 has n = 50 000.  The multi-GPU
// solver of the same library has no such limit and runs on a 1 x 1 device grid: measured 110 s at n = 50 000 on one B200 (64 GB
// workspace), eigenvalues equal to Xsyevd's to 2e-15 where both run (scripts/eig_probe.cu).  Loaded with dlopen so that the rest
// of the library does not depend on libcusolverMg being installed.
struct MgApi {
  void *lib = nullptr;
  cusolverStatus_t (*Create)(cusolverMgHandle_t *) = nullptr;
  cusolverStatus_t (*Destroy)(cusolverMgHandle_t) = nullptr;
  cusolverStatus_t (*DeviceSelect)(cusolverMgHandle_t, int, int[]) = nullptr;
  cusolverStatus_t (*CreateDeviceGrid)(cudaLibMgGrid_t *, int32_t, int32_t, const int32_t[], cusolverMgGridMapping_t) = nullptr;
  cusolverStatus_t (*DestroyGrid)(cudaLibMgGrid_t) = nullptr;
  cusolverStatus_t (*CreateMatrixDesc)(cudaLibMgMatrixDesc_t *, int64_t, int64_t, int64_t, int64_t, cudaDataType, const cudaLibMgGrid_t) = nullptr;
  cusolverStatus_t (*DestroyMatrixDesc)(cudaLibMgMatrixDesc_t) = nullptr;
  cusolverStatus_t (*Syevd_bufferSize)(cusolverMgHandle_t, cusolverEigMode_t, cublasFillMode_t, int, void *[], int, int, cudaLibMgMatrixDesc_t,
                                       void *, cudaDataType, cudaDataType, int64_t *) = nullptr;
  cusolverStatus_t (*Syevd)(cusolverMgHandle_t, cusolverEigMode_t, cublasFillMode_t, int, void *[], int, int, cudaLibMgMatrixDesc_t, void *,
                            cudaDataType, cudaDataType, void *[], int64_t, int *) = nullptr;
  bool ok = false;
  std::string why;
};
static MgApi &mg_api() {
  static MgApi a;
  if (a.lib || a.ok) return a;
  // First choice: the libcusolverMg that ships NEXT TO the libcusolver this process has loaded (a Python process with torch has
  // the pip-packaged cuBLAS / cuSOLVER of torch's CUDA minor version in memory; the toolkit's libcusolverMg then fails to bind
  // against that older libcublas -- "undefined symbol: cublasSetEnvironmentMode").  Then the default search path.
  std::vector<std::string> names;
  Dl_info di;
  if (dladdr((void *)&cusolverDnCreate, &di) && di.dli_fname) {
    std::string path(di.dli_fname);
    const size_t slash = path.rfind('/');
    if (slash != std::string::npos) names.push_back(path.substr(0, slash + 1) + "libcusolverMg.so.11");
  }
  names.push_back("libcusolverMg.so.11"); names.push_back("libcusolverMg.so"); names.push_back("/usr/local/cuda/lib64/libcusolverMg.so.11");
  for (const std::string &name : names) {
    a.lib = dlopen(name.c_str(), RTLD_NOW | RTLD_LOCAL);
    if (a.lib) break;
    const char *e = dlerror();
    a.why += name + ": " + (e ? e : "?") + "; ";
  }
  if (!a.lib) return a;
#define MG_SYM(field, sym) *(void **)(&a.field) = dlsym(a.lib, sym)
  MG_SYM(Create, "cusolverMgCreate"); MG_SYM(Destroy, "cusolverMgDestroy"); MG_SYM(DeviceSelect, "cusolverMgDeviceSelect");
  MG_SYM(CreateDeviceGrid, "cusolverMgCreateDeviceGrid"); MG_SYM(DestroyGrid, "cusolverMgDestroyGrid");
  MG_SYM(CreateMatrixDesc, "cusolverMgCreateMatrixDesc"); MG_SYM(DestroyMatrixDesc, "cusolverMgDestroyMatrixDesc");
  MG_SYM(Syevd_bufferSize, "cusolverMgSyevd_bufferSize"); MG_SYM(Syevd, "cusolverMgSyevd");
#undef MG_SYM
  a.ok = a.Create && a.Destroy && a.DeviceSelect && a.CreateDeviceGrid && a.DestroyGrid && a.CreateMatrixDesc && a.DestroyMatrixDesc &&
         a.Syevd_bufferSize && a.Syevd;
  return a;
}

}  // namespace gb

using namespace gb;

// dA (n x n, ld n, symmetric; destroyed) -> eigenvectors in the ROWS of the row-major view of dA (== columns of the column-major
// array, like Xsyevd leaves them), eigenvalues ascending in dW.  Blocks the host; uses cuSOLVER's own streams.
static int eigh_mg(gb200_ctx *ctx, double *dA, size_t n, double *dW) {
  MgApi &mg = mg_api();
  if (!mg.ok) return set_err(ctx, GB200_ERR_UNSUPPORTED, "gb200_eigh: n > 32768 needs libcusolverMg (cusolverDnXsyevd stops at 32768), which could not be loaded: " + mg.why);
  const int tile = 256;
  const size_t ncols = (n + tile - 1) / tile * tile;            // the last column tile is stored in full
  cusolverMgHandle_t h = nullptr; cudaLibMgGrid_t grid = nullptr; cudaLibMgMatrixDesc_t desc = nullptr;
  DevBuf dP, dWork;
  std::vector<double> W(n);
  auto fail = [&](int code, const std::string &msg) {
    dP.release(); dWork.release();
    if (desc) mg.DestroyMatrixDesc(desc);
    if (grid) mg.DestroyGrid(grid);
    if (h) mg.Destroy(h);
    return set_err(ctx, code, msg);
  };
#define MG_CALL(call)                                                                                                    \
  do { cusolverStatus_t _s = (call); if (_s != CUSOLVER_STATUS_SUCCESS) return fail(GB200_ERR_NUMERIC, std::string(#call " failed, cusolver status ") + std::to_string((int)_s)); } while (0)
#define MG_CUDA(call)                                                                                                    \
  do { cudaError_t _e = (call); if (_e != cudaSuccess) return fail(GB200_ERR_CUDA, std::string(#call " failed: ") + cudaGetErrorString(_e)); } while (0)
  int dev[1] = {ctx->device};
  int32_t dev32[1] = {ctx->device};
  MG_CALL(mg.Create(&h));
  MG_CALL(mg.DeviceSelect(h, 1, dev));
  MG_CALL(mg.CreateDeviceGrid(&grid, 1, 1, dev32, CUDALIBMG_GRID_MAPPING_COL_MAJOR));
  MG_CALL(mg.CreateMatrixDesc(&desc, (int64_t)n, (int64_t)n, (int64_t)n, (int64_t)tile, CUDA_R_64F, grid));
  double *A = dA;
  if (ncols != n) {
    MG_CUDA(dP.reserve(n * ncols * sizeof(double)));
    MG_CUDA(cudaMemcpyAsync(dP.p, dA, n * n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    MG_CUDA(cudaMemsetAsync(dP.as<double>() + n * n, 0, n * (ncols - n) * sizeof(double), ctx->stream));
    A = dP.as<double>();
  }
  MG_CUDA(cudaStreamSynchronize(ctx->stream));
  void *arrA[1] = {A};
  int64_t lwork = 0;
  MG_CALL(mg.Syevd_bufferSize(h, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, (int)n, arrA, 1, 1, desc, W.data(), CUDA_R_64F, CUDA_R_64F, &lwork));
  ctx->eigh_workspace_bytes = (size_t)lwork * sizeof(double) + (A != dA ? n * ncols * sizeof(double) : 0);
  MG_CUDA(dWork.reserve((size_t)(lwork > 0 ? lwork : 1) * sizeof(double)));
  void *arrW[1] = {dWork.p};
  int info = -1;
  MG_CALL(mg.Syevd(h, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, (int)n, arrA, 1, 1, desc, W.data(), CUDA_R_64F, CUDA_R_64F, arrW, lwork, &info));
  MG_CUDA(cudaDeviceSynchronize());
  if (info != 0) return fail(GB200_ERR_NUMERIC, "gb200_eigh: cusolverMgSyevd did not converge, info=" + std::to_string(info));
  dWork.release();
  if (A != dA) MG_CUDA(cudaMemcpyAsync(dA, A, n * n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  MG_CUDA(cudaMemcpyAsync(dW, W.data(), n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  MG_CUDA(cudaStreamSynchronize(ctx->stream));
  dP.release();
  mg.DestroyMatrixDesc(desc); mg.DestroyGrid(grid); mg.Destroy(h);
#undef MG_CALL
#undef MG_CUDA
  return GB200_OK;
}

// Device-resident core: dA (n x n, ld n; destroyed) -> dV (n x n, eigenvectors in COLUMNS of the row-major matrix), dW (n + 4 doubles:
// eigenvalues ascending with the < 1e-10 zeroing applied, then 3 statistics).  The only large temporary is cuSOLVER's workspace.
static int eigh_device(gb200_ctx *ctx, double *dA, size_t n, int center, double *dV, double *dW, double *trace_G, int *n_zero,
                       int *n_negative) {
  cudaStream_t st = ctx->stream;
  DevBuf dWork, dInfo, dRow;
  cusolverDnHandle_t h = nullptr;
  cusolverDnParams_t params = nullptr;
  void *hWork = nullptr;
  auto cleanup = [&]() {
    dWork.release(); dInfo.release(); dRow.release();
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
  EIGH_CUDA(dInfo.reserve(sizeof(int)));
  {
    ProfScope ps(ctx, "eigh");
    if (center) {
      EIGH_CUDA(dRow.reserve((n + 1) * sizeof(double)));
      EIGH_CUDA(launch_center_matrix(dA, n, n, dRow.as<double>(), st));
    }
    const bool use_mg = ctx->eigh_path == 2 || (ctx->eigh_path == 0 && n > 32768);
    if (use_mg) {
      const int rc = eigh_mg(ctx, dA, n, dW);
      if (rc) { cleanup(); return rc; }
    } else {
    EIGH_SOLVER(cusolverDnCreate(&h));
    EIGH_SOLVER(cusolverDnSetStream(h, st));
    EIGH_SOLVER(cusolverDnCreateParams(&params));
    size_t wdev = 0, whost = 0;
    // Row-major symmetric == column-major symmetric.  The reference hands its row-major G to
    // Fortran with UPLO='L' (src/lapack.cpp:205), i.e. the row-major UPPER triangle; for
    // cuSOLVER's column-major view that is CUBLAS_FILL_MODE_LOWER as well.
    EIGH_SOLVER(cusolverDnXsyevd_bufferSize(h, params, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER,
                                            (int64_t)n, CUDA_R_64F, dA, (int64_t)n, CUDA_R_64F, dW,
                                            CUDA_R_64F, &wdev, &whost));
    ctx->eigh_workspace_bytes = wdev;
    EIGH_CUDA(dWork.reserve(wdev ? wdev : 8));
    if (whost) { hWork = malloc(whost); if (!hWork) { cleanup(); return set_err(ctx, GB200_ERR_NOMEM, "host workspace"); } }
    EIGH_SOLVER(cusolverDnXsyevd(h, params, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, (int64_t)n,
                                 CUDA_R_64F, dA, (int64_t)n, CUDA_R_64F, dW, CUDA_R_64F, dWork.p, wdev,
                                 hWork, whost, dInfo.as<int>()));
    int info = 0;
    EIGH_CUDA(cudaMemcpyAsync(&info, dInfo.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    EIGH_CUDA(cudaStreamSynchronize(st));
    if (info != 0) {
      cleanup();
      return set_err(ctx, GB200_ERR_NUMERIC, "gb200_eigh: syevd did not converge, info=" + std::to_string(info));
    }
    dWork.release();
    }
    // cuSOLVER leaves eigenvector j in column j of the column-major array == row j of the
    // row-major view; transpose so that U[i][j] = v_j[i] (eigenvectors in columns).
    EIGH_CUDA(launch_transpose(dA, n, n, n, dV, n, st));
    zero_small_kernel<<<1, 256, 0, st>>>(dW, n, dW + n);
    EIGH_CUDA(cudaGetLastError());
  }
  double stats[3];
  EIGH_CUDA(cudaMemcpyAsync(stats, dW + n, 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
  EIGH_CUDA(cudaStreamSynchronize(st));
  if (trace_G) *trace_G = stats[0] / (double)n;
  if (n_zero) *n_zero = (int)stats[1];
  if (n_negative) *n_negative = (int)stats[2];
  cleanup();
  return GB200_OK;
#undef EIGH_CUDA
#undef EIGH_SOLVER
}

extern "C" int gb200_eigh(gb200_ctx *ctx, double *G, size_t n, size_t ldg, int center, double *U,
                          size_t ldu, double *eval, double *trace_G, int *n_zero, int *n_negative) {
  if (!ctx) return GB200_ERR_ARG;
  if (!G || !U || !eval || n == 0 || ldg < n || ldu < n) return set_err(ctx, GB200_ERR_ARG, "gb200_eigh: bad argument");
  cudaStream_t st = ctx->stream;
  DevBuf dA, dV, dW;
  auto fail = [&](cudaError_t e, const char *what) {
    dA.release(); dV.release(); dW.release();
    return set_err(ctx, GB200_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
  };
  cudaError_t e;
  if ((e = dA.reserve(n * n * sizeof(double))) != cudaSuccess) return fail(e, "gb200_eigh: alloc A");
  if ((e = dV.reserve(n * n * sizeof(double))) != cudaSuccess) return fail(e, "gb200_eigh: alloc V");
  if ((e = dW.reserve((n + 4) * sizeof(double))) != cudaSuccess) return fail(e, "gb200_eigh: alloc W");
  if ((e = cudaMemcpy2DAsync(dA.p, n * sizeof(double), G, ldg * sizeof(double), n * sizeof(double), n, cudaMemcpyHostToDevice, st)) != cudaSuccess)
    return fail(e, "gb200_eigh: copy in");
  int rc = eigh_device(ctx, dA.as<double>(), n, center, dV.as<double>(), dW.as<double>(), trace_G, n_zero, n_negative);
  if (rc) { dA.release(); dV.release(); dW.release(); return rc; }
  dA.release();
  if ((e = cudaMemcpy2DAsync(U, ldu * sizeof(double), dV.p, n * sizeof(double), n * sizeof(double), n, cudaMemcpyDeviceToHost, st)) != cudaSuccess)
    return fail(e, "gb200_eigh: copy U");
  if ((e = cudaMemcpyAsync(eval, dW.p, n * sizeof(double), cudaMemcpyDeviceToHost, st)) != cudaSuccess) return fail(e, "gb200_eigh: copy eval");
  if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return fail(e, "gb200_eigh: sync");
  dA.release(); dV.release(); dW.release();
  return GB200_OK;
}

// Device-resident variant: G_dev (n x n, ld n) is destroyed, U_dev (n x n, ld n) and eval_dev (n) are caller-owned device buffers.
extern "C" int gb200_eigh_dev(gb200_ctx *ctx, double *G_dev, size_t n, int center, double *U_dev, double *eval_dev, double *trace_G,
                              int *n_zero, int *n_negative) {
  if (!ctx) return GB200_ERR_ARG;
  if (!G_dev || !U_dev || !eval_dev || n == 0 || G_dev == U_dev) return set_err(ctx, GB200_ERR_ARG, "gb200_eigh_dev: bad argument");
  DevBuf dW;
  if (dW.reserve((n + 4) * sizeof(double)) != cudaSuccess) return set_err(ctx, GB200_ERR_NOMEM, "gb200_eigh_dev: alloc");
  int rc = eigh_device(ctx, G_dev, n, center, U_dev, dW.as<double>(), trace_G, n_zero, n_negative);
  if (rc == GB200_OK) {
    cudaError_t e = cudaMemcpyAsync(eval_dev, dW.p, n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) rc = set_err(ctx, GB200_ERR_CUDA, std::string("gb200_eigh_dev: ") + cudaGetErrorString(e));
  }
  dW.release();
  return rc;
}
