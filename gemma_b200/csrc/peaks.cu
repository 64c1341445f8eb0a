// peaks.cu -- measured denominators for the rooflines bench.py prints.  MEASURED_PEAKS.json (driver-written) holds the HBM copy
// bandwidth and the bf16 tensor rate; the per-SNP kernel is bound by the plain (non-tensor) FP64 pipe, for which no
// driver-measured figure exists, so the library measures it: independent DFMA chains, all SMs, full occupancy.
#include "common.cuh"

namespace gb {

__global__ void __launch_bounds__(256) fp64_fma_probe_kernel(double *__restrict__ out, int iters, double a, double b) {
  double x[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) x[k] = (double)(threadIdx.x + k) * 1e-3;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int k = 0; k < 8; ++k) x[k] = fma(x[k], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += x[k];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace gb

using namespace gb;

// seconds: length of the measurement (the power cap needs a while to settle); *tflops counts an FMA as 2 flops
extern "C" int gb200_measure_fp64_fma(gb200_ctx *c, double seconds, double *tflops, double *ms_out) {
  if (!c || !tflops) return GB200_ERR_ARG;
  const int ctas = c->num_sms * 8, threads = 256;
  DevBuf buf;
  GB_CUDA(c, buf.reserve((size_t)ctas * threads * sizeof(double)));
  cudaEvent_t e0, e1;
  GB_CUDA(c, cudaEventCreate(&e0)); GB_CUDA(c, cudaEventCreate(&e1));
  int iters = 2000;
  double ms = 0.0, flops = 0.0;
  // calibrate on one launch, then run a single launch of about `seconds`
  for (int round = 0; round < 2; ++round) {
    cudaEventRecord(e0, c->stream);
    fp64_fma_probe_kernel<<<ctas, threads, 0, c->stream>>>(buf.as<double>(), iters, 0.999999, 1e-9);
    cudaEventRecord(e1, c->stream);
    cudaError_t e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) { buf.release(); return set_err(c, GB200_ERR_CUDA, cudaGetErrorString(e)); }
    float f = 0.f; cudaEventElapsedTime(&f, e0, e1); ms = f;
    flops = 2.0 * 64.0 * (double)iters * (double)ctas * threads;
    if (round == 0) {
      double want = seconds > 0.0 ? seconds * 1e3 : 200.0;
      double scale = want / (ms > 1e-3 ? ms : 1e-3);
      if (scale > 2000.0) scale = 2000.0;
      long it2 = (long)((double)iters * scale);
      if (it2 < 1000) it2 = 1000;
      if (it2 > 200000000L) it2 = 200000000L;
      iters = (int)it2;
    }
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  buf.release();
  *tflops = flops / (ms * 1e-3) / 1e12;
  if (ms_out) *ms_out = ms;
  c->kernel_launches += 2;
  return GB200_OK;
}
