// eig_probe.cu -- which cuSOLVER symmetric eigensolver entry points accept the benchmark's n = 50 000, and how long they take.
// Build here:  nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o scripts/eig_probe scripts/eig_probe.cu -lcusolver -lcusolverMg -lcublas
// Run on the GPU box:  scripts/eig_probe [n_big]            (prints one line per experiment)
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <chrono>
#include <cmath>
#include <cuda_runtime.h>
#include <cusolverDn.h>
#include <cusolverMg.h>

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

__global__ void fill_sym(double *A, size_t n, unsigned long long seed) {
  // symmetric pseudo-random matrix with entries ~ U(-1,1)/sqrt(n) plus a diagonal ramp (distinct, well spread eigenvalues)
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * n) return;
  size_t i = idx / n, j = idx % n;
  size_t a = i < j ? i : j, b = i < j ? j : i;
  unsigned long long z = seed + a * 0x9E3779B97F4A7C15ull + b * 0xC2B2AE3D27D4EB4Full;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;
  double u = (double)(z >> 11) * (1.0 / 9007199254740992.0) * 2.0 - 1.0;
  double v = u / sqrt((double)n);
  if (i == j) v += 1.0 + (double)i / (double)n;
  A[idx] = v;
}

static void buffer_sizes(cusolverDnHandle_t h, cusolverDnParams_t p, int64_t n) {
  double *dA = nullptr, *dW = nullptr;
  cudaMalloc(&dA, 1024); cudaMalloc(&dW, 1024);
  size_t wd = 0, wh = 0;
  cusolverStatus_t s = cusolverDnXsyevd_bufferSize(h, p, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, CUDA_R_64F, dA, n, CUDA_R_64F, dW, CUDA_R_64F, &wd, &wh);
  printf("n=%lld Xsyevd_bufferSize(VECTOR): status %d dev %.3f GB host %zu\n", (long long)n, (int)s, wd / 1e9, wh);
  wd = wh = 0;
  s = cusolverDnXsyevd_bufferSize(h, p, CUSOLVER_EIG_MODE_NOVECTOR, CUBLAS_FILL_MODE_LOWER, n, CUDA_R_64F, dA, n, CUDA_R_64F, dW, CUDA_R_64F, &wd, &wh);
  printf("n=%lld Xsyevd_bufferSize(NOVECTOR): status %d dev %.3f GB host %zu\n", (long long)n, (int)s, wd / 1e9, wh);
  int64_t hm = 0; wd = wh = 0;
  s = cusolverDnXsyevdx_bufferSize(h, p, CUSOLVER_EIG_MODE_VECTOR, CUSOLVER_EIG_RANGE_ALL, CUBLAS_FILL_MODE_LOWER, n, CUDA_R_64F, dA, n, nullptr, nullptr, 0, 0, &hm,
                                   CUDA_R_64F, dW, CUDA_R_64F, &wd, &wh);
  printf("n=%lld Xsyevdx_bufferSize(VECTOR, ALL): status %d dev %.3f GB host %zu\n", (long long)n, (int)s, wd / 1e9, wh);
  int lw = 0;
  s = cusolverDnDsyevd_bufferSize(h, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, (int)n, dA, (int)n, dW, &lw);
  printf("n=%lld Dsyevd_bufferSize(VECTOR): status %d lwork %d\n", (long long)n, (int)s, lw);
  lw = 0;
  s = cusolverDnDsytrd_bufferSize(h, CUBLAS_FILL_MODE_LOWER, (int)n, dA, (int)n, dW, dW, dW, &lw);
  printf("n=%lld Dsytrd_bufferSize: status %d lwork %d\n", (long long)n, (int)s, lw);
  lw = 0;
  s = cusolverDnDormtr_bufferSize(h, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, (int)n, (int)n, dA, (int)n, dW, dA, (int)n, &lw);
  printf("n=%lld Dormtr_bufferSize: status %d lwork %d\n", (long long)n, (int)s, lw);
  cudaFree(dA); cudaFree(dW);
  fflush(stdout);
}

static double run_dn(cusolverDnHandle_t h, cusolverDnParams_t p, int64_t n, std::vector<double> *evals) {
  double *dA = nullptr, *dW = nullptr; void *dWork = nullptr, *hWork = nullptr; int *dInfo = nullptr;
  if (cudaMalloc(&dA, n * n * sizeof(double)) != cudaSuccess) { printf("n=%lld Xsyevd: cudaMalloc A failed\n", (long long)n); return -1; }
  cudaMalloc(&dW, n * sizeof(double)); cudaMalloc(&dInfo, sizeof(int));
  fill_sym<<<(unsigned)((n * n + 255) / 256), 256>>>(dA, n, 7);
  size_t wd = 0, wh = 0;
  cusolverStatus_t s = cusolverDnXsyevd_bufferSize(h, p, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, CUDA_R_64F, dA, n, CUDA_R_64F, dW, CUDA_R_64F, &wd, &wh);
  if (s != CUSOLVER_STATUS_SUCCESS) { printf("n=%lld Xsyevd: bufferSize status %d\n", (long long)n, (int)s); cudaFree(dA); cudaFree(dW); cudaFree(dInfo); return -1; }
  cudaMalloc(&dWork, wd ? wd : 8); if (wh) hWork = malloc(wh);
  cudaDeviceSynchronize();
  double t0 = now();
  s = cusolverDnXsyevd(h, p, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, CUDA_R_64F, dA, n, CUDA_R_64F, dW, CUDA_R_64F, dWork, wd, hWork, wh, dInfo);
  cudaError_t e = cudaDeviceSynchronize();
  double dt = now() - t0;
  int info = -1; cudaMemcpy(&info, dInfo, sizeof(int), cudaMemcpyDeviceToHost);
  printf("n=%lld Xsyevd(VECTOR): status %d cuda %d info %d  %.2f s  workspace %.2f GB\n", (long long)n, (int)s, (int)e, info, dt, wd / 1e9);
  if (evals) { evals->resize(n); cudaMemcpy(evals->data(), dW, n * sizeof(double), cudaMemcpyDeviceToHost); }
  cudaFree(dA); cudaFree(dW); cudaFree(dWork); cudaFree(dInfo); free(hWork);
  fflush(stdout);
  return dt;
}

static double run_mg(int n, int tile, std::vector<double> *evals) {
  cusolverMgHandle_t h = nullptr;
  int dev[1] = {0};
  cusolverStatus_t s = cusolverMgCreate(&h);
  if (s != CUSOLVER_STATUS_SUCCESS) { printf("cusolverMgCreate status %d\n", (int)s); return -1; }
  s = cusolverMgDeviceSelect(h, 1, dev);
  cudaLibMgGrid_t grid = nullptr; cudaLibMgMatrixDesc_t desc = nullptr;
  int32_t dev32[1] = {0};
  s = cusolverMgCreateDeviceGrid(&grid, 1, 1, dev32, CUDALIBMG_GRID_MAPPING_COL_MAJOR);
  if (s != CUSOLVER_STATUS_SUCCESS) { printf("MgCreateDeviceGrid status %d\n", (int)s); return -1; }
  s = cusolverMgCreateMatrixDesc(&desc, n, n, n, tile, CUDA_R_64F, grid);
  if (s != CUSOLVER_STATUS_SUCCESS) { printf("MgCreateMatrixDesc status %d\n", (int)s); return -1; }
  // one device: all column tiles are local and contiguous (local leading dimension n); columns padded to a multiple of the tile
  const size_t ncols = ((size_t)n + tile - 1) / tile * tile;
  double *dA = nullptr;
  if (cudaMalloc(&dA, (size_t)n * ncols * sizeof(double)) != cudaSuccess) { printf("n=%d Mg: cudaMalloc A failed\n", n); return -1; }
  cudaMemset(dA, 0, (size_t)n * ncols * sizeof(double));
  fill_sym<<<(unsigned)(((size_t)n * n + 255) / 256), 256>>>(dA, n, 7);
  void *arrA[1] = {dA};
  std::vector<double> W(n);
  int64_t lwork = 0;
  s = cusolverMgSyevd_bufferSize(h, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, arrA, 1, 1, desc, W.data(), CUDA_R_64F, CUDA_R_64F, &lwork);
  printf("n=%d MgSyevd_bufferSize: status %d lwork %lld elements (%.2f GB)\n", n, (int)s, (long long)lwork, lwork * 8.0 / 1e9);
  fflush(stdout);
  if (s != CUSOLVER_STATUS_SUCCESS) { cudaFree(dA); return -1; }
  void *dWork = nullptr;
  if (cudaMalloc(&dWork, (size_t)lwork * sizeof(double)) != cudaSuccess) { printf("n=%d Mg: cudaMalloc work failed\n", n); cudaFree(dA); return -1; }
  void *arrW[1] = {dWork};
  int info = -1;
  cudaDeviceSynchronize();
  double t0 = now();
  s = cusolverMgSyevd(h, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, arrA, 1, 1, desc, W.data(), CUDA_R_64F, CUDA_R_64F, arrW, lwork, &info);
  cudaError_t e = cudaDeviceSynchronize();
  double dt = now() - t0;
  printf("n=%d MgSyevd(VECTOR, 1 GPU, tile %d): status %d cuda %d info %d  %.2f s\n", n, tile, (int)s, (int)e, info, dt);
  if (evals) *evals = W;
  // residual of a few eigenpairs: ||A v - w v|| with A regenerated on the host would be O(n^2); check orthonormality of 2 columns instead
  if (s == CUSOLVER_STATUS_SUCCESS) {
    std::vector<double> v0(n), v1(n);
    cudaMemcpy(v0.data(), dA, n * sizeof(double), cudaMemcpyDeviceToHost);
    cudaMemcpy(v1.data(), dA + (size_t)n * (n / 2), n * sizeof(double), cudaMemcpyDeviceToHost);
    double a = 0, b = 0, c = 0;
    for (int i = 0; i < n; ++i) { a += v0[i] * v0[i]; b += v1[i] * v1[i]; c += v0[i] * v1[i]; }
    printf("   column norms^2 %.15f %.15f  dot %.3e   W[0] %.12g W[n-1] %.12g\n", a, b, c, W[0], W[n - 1]);
  }
  cudaFree(dA); cudaFree(dWork);
  cusolverMgDestroyMatrixDesc(desc); cusolverMgDestroyGrid(grid); cusolverMgDestroy(h);
  fflush(stdout);
  return dt;
}

int main(int argc, char **argv) {
  int n_big = argc > 1 ? atoi(argv[1]) : 50000;
  cusolverDnHandle_t h; cusolverDnParams_t p;
  cusolverDnCreate(&h); cusolverDnCreateParams(&p);
  for (int64_t n : {16384LL, 32768LL, 40000LL, 46340LL, 46341LL, 50000LL, 65536LL}) buffer_sizes(h, p, n);
  std::vector<double> e_dn, e_mg;
  run_dn(h, p, 4096, &e_dn);
  run_mg(4096, 256, &e_mg);
  if (e_dn.size() == e_mg.size() && !e_dn.empty()) {
    double m = 0; for (size_t i = 0; i < e_dn.size(); ++i) m = fmax(m, fabs(e_dn[i] - e_mg[i]));
    printf("n=4096 max |eval_Dn - eval_Mg| = %.3e\n", m);
  }
  run_dn(h, p, 16384, nullptr);
  run_mg(16384, 256, nullptr);
  run_dn(h, p, 32768, nullptr);
  if (n_big > 0) {
    run_dn(h, p, n_big, nullptr);
    run_mg(n_big, 256, nullptr);
  }
  return 0;
}
