// lmm_v2.cuh -- second-generation fused per-SNP LMM kernel.
//
// Same arithmetic and control flow as lmm_device.cuh (one warp owns one SNP; grid scan, GSL Brent,
// GSL Newton, Wald / score / LRT -- src/lmm.cpp:1526-1562, :1945-2140), re-organised around memory:
//
//  * a CTA holds 8 warps = 8 SNPs that walk the n rotated individuals in LOCKSTEP passes.  The
//    SNP-independent vectors (eigenvalues, rotated covariates, rotated phenotype) are staged once per
//    CTA and chunk in shared memory by a 3-stage TMA bulk-copy pipeline (one elected thread issues the 2 KB row copies,
//    completion counted in bytes on an mbarrier per stage; GB_V2_TMA=0 selects the older per-thread cp.async
//    variant) and shared by the 8 SNPs; every
//    warp's own U^T x row streams through its private slice of the same stages.  At n = 50 000 the
//    v1 kernel re-read 1.6 MB per SNP and pass through L2 with ~2 KB in flight per warp
//    (latency-bound, 17% of the FP64 pipe); here a pass moves 0.55 MB per SNP with two 22 KB stages in
//    flight per CTA and two CTAs per SM.
//  * several lambdas per pass ("slots"): the 11 grid lambdas are evaluated 2 at a time, f(l_max) and the
//    score-test lambda share one pass, and the REML and ML root refinements (independent chains) advance
//    side by side, so a SNP needs ~18 passes over its row instead of ~35.
//  * 1/(lambda*delta+1) by MUFU.RCP64H seed + two Newton steps (delta >= 0 after the <1e-10 zeroing of
//    lapack.cpp:268, so the denominator is >= 1: no special cases).
//
// Results are identical to v1 up to summation order (both are checked against the oracle).
#pragma once
#include "lmm_device.cuh"

namespace gb {

#ifndef GB_V2_WARPS
#define GB_V2_WARPS 8
#endif
#ifndef GB_V2_CTAS2_MAXNC
#define GB_V2_CTAS2_MAXNC 1     // covariate counts up to this run 2 CTAs per SM (<= 128 registers); larger tables get the full register file
#endif
#ifndef GB_V2_CTAS
#define GB_V2_CTAS 2
#endif
constexpr int V2_WARPS = GB_V2_WARPS;
constexpr int V2_THREADS = V2_WARPS * 32;
constexpr int V2_CHUNK = 256;          // individuals per pipeline stage
constexpr int V2_STAGES = 3;
#ifndef GB_V2_AHEAD
#define GB_V2_AHEAD 1
#endif
// chunks queued ahead of the one being consumed.  With V2_STAGES - 1 the producing lane has to wait until EVERY warp has released the
// chunk just before its own (source-level ncu: 27 % of the kernel's stall samples sat in those waits -- the ring degenerated into a
// per-chunk barrier); with one chunk less the other warps may trail by two chunks before a producer blocks, and one 2.9 us chunk of
// lead is still several memory latencies.
constexpr int V2_AHEAD = GB_V2_AHEAD;
constexpr int V2_MAX_REGION = 64;
constexpr int V2_NSG = 2;              // grid lambdas per pass (register budget: 2 CTAs per SM need <= 128 registers)
constexpr int V2_NSC = 5;              // common-lambda slots per hoisted pass (h rows staged next to the data rows)

constexpr int V2_CM = 20;              // Chebyshev nodes per grid interval for the SNP-independent sums (tabulated once per run: free)
constexpr int V2_XM = 15;              // Chebyshev nodes per grid interval for the x-sums of a SNP (a multiple of V2_NSC).  The x-sums enter the
                                       // derivatives with weight O(z^2 / n) (sweep of x) or through one O(1) pivot ratio against a trace of
                                       // order n, but the iterates must still be reproduced to ~1e-11: the reference reports the PREVIOUS
                                       // Newton iterate once |l_k - l_{k-1}| < 1e-5 l_k (src/lmm.cpp:2096), so an evaluation error eps flips
                                       // the iteration count -- and moves the reported lambda by ~1e-5 -- for a fraction eps / 1e-5 of the
                                       // SNPs (10 nodes, 1e-7: one SNP in a hundred, measured; 15 nodes, 3e-11 x weight: < 1e-6).
                                       // f(lambda_hat) and the Wald tables are NOT taken from these interpolants but from one exact pass.
constexpr int V2_CHEB_BASE = 4 * (V2_CM + V2_XM);   // LmmConst::cheb: cos(pi j / (2 V2_CM)), j < 4 V2_CM | cos(pi j / (2 V2_XM)), j < 4 V2_XM | coefficients

__host__ __device__ constexpr size_t v2_stage_doubles(int nc) { return (size_t)(nc + 2 + V2_WARPS + V2_NSC) * V2_CHUNK; }
// per warp: Chebyshev coefficients of its SNP's x-sums over the current interval (2 powers x (nc + 2) sums x V2_CM) + one pass of node values
__host__ __device__ constexpr size_t v2_cheb_warp_doubles(int nc) { return (size_t)(2 * (nc + 2)) * (V2_XM + V2_NSC); }
// SNP-independent sums at one common lambda: S^k_ab over (w_1..w_c, y) for k = 0,1,2, then sum h, sum h^2, sum log(l d+1), lambda
__host__ __device__ constexpr int v2c_nidx(int nc) { return (nc + 2) * (nc + 1) / 2; }
__host__ __device__ constexpr int v2c_stride(int nc) { return 3 * v2c_nidx(nc) + 4; }
__host__ __device__ constexpr size_t v2_smem_bytes(int nc) {
  return v2_stage_doubles(nc) * V2_STAGES * sizeof(double) + 64 + v2_cheb_warp_doubles(nc) * V2_WARPS * sizeof(double);
}

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---- TMA (bulk async copy) producer: one elected thread queues the 2 KB rows of a stage, completion is counted in bytes
// on the stage's mbarrier; the other 255 threads issue nothing (the per-thread cp.async variant spent ~15% of the kernel's
// instructions on copy addressing).  Barriers + the pass counter live in the 64 spare bytes behind the stages.
#ifndef GB_V2_TMA
#define GB_V2_TMA 1
#endif
__device__ __forceinline__ uint32_t v2_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void v2_mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(v2_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void v2_mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(v2_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void v2_mbar_wait(uint64_t *bar, uint32_t parity) {
  const uint32_t addr = v2_smem_u32(bar);
  uint32_t done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(addr), "r"(parity), "r"(1000u) : "memory");      // suspend-time hint (ns): the ring waits were 14 % of all issued instructions as bare spins
  } while (!done);
}
__device__ __forceinline__ void v2_bulk_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(v2_smem_u32(dst)), "l"(src), "r"(bytes), "r"(v2_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint64_t *v2_bars(double *smem, int nc) { return reinterpret_cast<uint64_t *>(smem + v2_stage_doubles(nc) * V2_STAGES); }
__device__ __forceinline__ void v2_mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(v2_smem_u32(bar)) : "memory");
}
// The stages form ONE ring for the whole kernel: chunk number k (counted across passes, identical in every thread) lives in stage
// k % V2_STAGES, its arrival is phase k / V2_STAGES of full[stage], and it may be overwritten once all V2_WARPS warps have arrived on
// empty[stage] for it.  No CTA-wide barrier inside or between passes: chunk k + 2 is queued (by lane 0 of warp (k + 2) % V2_WARPS) as soon as chunk k - 1 has been
// released by every warp, each warp consumes at its own pace (at most V2_STAGES - 1 chunks apart), and the first chunks of the next
// pass are in flight while slower warps still finish the previous one.
__device__ __forceinline__ void v2_acquire_stage(uint64_t *empty, unsigned int k) {
  if (k >= (unsigned int)V2_STAGES) v2_mbar_wait(&empty[k % V2_STAGES], ((k / V2_STAGES) - 1u) & 1u);
}

// once per kernel, before the first pass
__device__ __forceinline__ void v2_pipeline_init(double *smem, int nc) {
  if (threadIdx.x == 0) {
    uint64_t *bars = v2_bars(smem, nc);
    for (int s = 0; s < V2_STAGES; ++s) { v2_mbar_init(&bars[s], 1); v2_mbar_init(&bars[V2_STAGES + s], V2_WARPS); }   // full[] (TMA bytes), empty[] (one arrival per warp)
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
}

// reciprocal of den >= 1 (finite): hardware seed + two Newton-Raphson steps, ~1 ulp
__device__ __forceinline__ double rcp_ge1(double den) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(den));
  double e = fma(-den, r, 1.0);
  r = fma(r, e, r);
  e = fma(-den, r, 1.0);
  r = fma(r, e, r);
  return r;
}

template <int NC>
struct V2Ctx {
  const LmmConst *D;
  const double *xrow;        // this warp's U^T x row (nullptr for an idle warp)
  double *smem;              // stage 0 base
  int nchunks;
  int pad;                   // n_c - n (zero-padded tail elements: each adds 1 to every sum of h^k)
};

// all 256 threads: queue the cp.async copies of chunk `c` into stage `st`
template <int NC>
__device__ __forceinline__ void v2_issue(const LmmConst &D, const double *const *xrows, double *stage, int c) {
  constexpr int HALF = V2_CHUNK / 2;
  const size_t off0 = (size_t)c * V2_CHUNK;
  for (int op = threadIdx.x; op < (NC + 2) * HALF; op += V2_THREADS) {
    const int arr = op / HALF, o = (op - arr * HALF) * 2;
    const double *src = (arr == 0) ? D.delta : (arr == NC + 1) ? D.y : D.Wt + (size_t)(arr - 1) * D.ldv;
    cp_async16(stage + arr * V2_CHUNK + o, src + off0 + o);
  }
  double *xs = stage + (NC + 2) * V2_CHUNK;
  for (int op = threadIdx.x; op < V2_WARPS * HALF; op += V2_THREADS) {
    const int w = op / HALF, o = (op - w * HALF) * 2;
    const double *src = xrows[w];
    if (src) cp_async16(xs + w * V2_CHUNK + o, src + off0 + o);
  }
}

// thread 0 only: queue chunk `c` into `stage` with bulk copies that complete on `bar`
template <int NC>
__device__ __forceinline__ void v2_issue_tma(const LmmConst &D, const double *const *xrows, double *stage, int c, uint64_t *bar) {
  constexpr uint32_t ROW = V2_CHUNK * sizeof(double);
  const size_t off0 = (size_t)c * V2_CHUNK;
  uint32_t rows = NC + 2;
#pragma unroll
  for (int w = 0; w < V2_WARPS; ++w) rows += (xrows[w] != nullptr) ? 1u : 0u;
  v2_mbar_expect_tx(bar, rows * ROW);
  v2_bulk_load(stage, D.delta + off0, ROW, bar);
#pragma unroll
  for (int a = 0; a < NC; ++a) v2_bulk_load(stage + (a + 1) * V2_CHUNK, D.Wt + (size_t)a * D.ldv + off0, ROW, bar);
  v2_bulk_load(stage + (NC + 1) * V2_CHUNK, D.y + off0, ROW, bar);
  double *xs = stage + (NC + 2) * V2_CHUNK;
#pragma unroll
  for (int w = 0; w < V2_WARPS; ++w)
    if (xrows[w]) v2_bulk_load(xs + w * V2_CHUNK, xrows[w] + off0, ROW, bar);
}

template <int NC, int NS, int KLO, int KHI>
struct V2Acc {
  static constexpr int NIDX = (NC + 3) * (NC + 2) / 2;
  static constexpr int NK = KHI - KLO + 1;
  double S[NS][NK][NIDX];
  double tr[NS][NK];
  double ld[NS];
};

// arithmetic of one staged chunk for this warp's SNP: 8 individuals per lane, all slots and powers
template <int NC, int NS, int KLO, int KHI, bool LD>
__device__ __forceinline__ void v2_chunk(const double *__restrict__ st, int warp, int lane, const double (&lam)[NS],
                                         V2Acc<NC, NS, KLO, KHI> &acc) {
  constexpr int NV = NC + 2;
  constexpr int NIDX = (NC + 3) * (NC + 2) / 2;
  constexpr int NK = KHI - KLO + 1;
  const double *sl = st + lane;
  const double *xl = st + (NC + 2) * V2_CHUNK + warp * V2_CHUNK + lane;
#pragma unroll
  for (int u = 0; u < V2_CHUNK / 32; ++u) {
    const int j = u * 32;
    double v[NV];
    const double dl = sl[j];
#pragma unroll
    for (int a = 0; a < NC; ++a) v[a] = sl[(a + 1) * V2_CHUNK + j];
    v[NC] = xl[j];
    v[NC + 1] = sl[(NC + 1) * V2_CHUNK + j];
    // products shared by all slots and powers
    double pr[NIDX];
#pragma unroll
    for (int a = 0; a < NV; ++a)
#pragma unroll
      for (int b = a; b < NV; ++b) pr[abidx(a, b, NV)] = v[a] * v[b];
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      const double den = fma(lam[s], dl, 1.0);
      const double h = rcp_ge1(den);
      if (LD) acc.ld[s] += log(den);
      double hk = (KLO == 0) ? 1.0 : h;
#pragma unroll
      for (int k = 0; k < NK; ++k) {
        acc.tr[s][k] += hk;
#pragma unroll
        for (int q = 0; q < NIDX; ++q) acc.S[s][k][q] = fma(hk, pr[q], acc.S[s][k][q]);
        hk *= h;
      }
    }
  }
}

// One lockstep pass.  Every thread of the CTA must call it (pipeline + barriers); `active` selects
// whether this warp does arithmetic.  want_ld[s] adds sum log|lambda*delta+1| for slot s.
template <int NC, int NS, int KLO, int KHI, bool LD>
__device__ __forceinline__ void v2_pass(const LmmConst &D, const double *const *xrows, double *smem, int nchunks,
                                        int pad, bool active, const double (&lam)[NS], V2Acc<NC, NS, KLO, KHI> &acc, unsigned int &pipe_it) {
  constexpr int NIDX = (NC + 3) * (NC + 2) / 2;
  constexpr int NK = KHI - KLO + 1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t stage_d = v2_stage_doubles(NC);
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    acc.ld[s] = 0.0;
#pragma unroll
    for (int k = 0; k < NK; ++k) {
      acc.tr[s][k] = 0.0;
#pragma unroll
      for (int j = 0; j < NIDX; ++j) acc.S[s][k][j] = 0.0;
    }
  }
#if GB_V2_TMA
  uint64_t *full = v2_bars(smem, NC), *empty = full + V2_STAGES;
  const unsigned int it0 = pipe_it;
  // the producer duty rotates over the warps (chunk k is queued by lane 0 of warp k % V2_WARPS): a fixed producer thread made its
  // warp ~40 % slower than the other seven and the whole ring ran at that warp's pace
#pragma unroll
  for (int c = 0; c < V2_AHEAD; ++c)
    if (c < nchunks) {
      const unsigned int k = it0 + (unsigned int)c;
      if (lane == 0 && warp == (int)(k % V2_WARPS)) {
        v2_acquire_stage(empty, k);
        v2_issue_tma<NC>(D, xrows, smem + (size_t)(k % V2_STAGES) * stage_d, c, &full[k % V2_STAGES]);
      }
    }
  for (int c = 0; c < nchunks; ++c) {
    {
      const int cn = c + V2_AHEAD;
      const unsigned int kn = it0 + (unsigned int)cn;
      if (cn < nchunks && lane == 0 && warp == (int)(kn % V2_WARPS)) {
        v2_acquire_stage(empty, kn);
        v2_issue_tma<NC>(D, xrows, smem + (size_t)(kn % V2_STAGES) * stage_d, cn, &full[kn % V2_STAGES]);
      }
    }
    const unsigned int k = it0 + (unsigned int)c;
    const int sidx = (int)(k % V2_STAGES);
    v2_mbar_wait(&full[sidx], (k / V2_STAGES) & 1u);        // chunk c has landed
    if (active) v2_chunk<NC, NS, KLO, KHI, LD>(smem + (size_t)sidx * stage_d, warp, lane, lam, acc);
    __syncwarp();
    if (lane == 0) v2_mbar_arrive(&empty[sidx]);            // this warp is done with the stage
  }
  pipe_it = it0 + (unsigned int)nchunks;
#else
  // prologue: stages 0 .. STAGES-2
#pragma unroll
  for (int c = 0; c < V2_STAGES - 1; ++c) {
    if (c < nchunks) v2_issue<NC>(D, xrows, smem + (size_t)c * stage_d, c);
    cp_async_commit();
  }
  for (int c = 0; c < nchunks; ++c) {
    cp_async_wait<V2_STAGES - 2>();                       // this thread's copies of chunk c have landed
    __syncthreads();                                      // everyone's have; everyone is done with chunk c-1
    {
      const int cn = c + V2_STAGES - 1;                   // refill the stage chunk c-1 just vacated
      if (cn < nchunks) v2_issue<NC>(D, xrows, smem + (size_t)(cn % V2_STAGES) * stage_d, cn);
      cp_async_commit();
    }
    if (active) v2_chunk<NC, NS, KLO, KHI, LD>(smem + (size_t)(c % V2_STAGES) * stage_d, warp, lane, lam, acc);
  }
  cp_async_wait<0>();
  __syncthreads();                                        // stages are free for the next pass
#endif
  if (active) {
#pragma unroll
    for (int s = 0; s < NS; ++s) {
#pragma unroll
      for (int k = 0; k < NK; ++k) {
        acc.tr[s][k] = warp_allsum(acc.tr[s][k]) - (double)pad;     // padded tail: delta=0 -> h=1
#pragma unroll
        for (int j = 0; j < NIDX; ++j) acc.S[s][k][j] = warp_allsum(acc.S[s][k][j]);
      }
      acc.ld[s] = warp_allsum(acc.ld[s]);
    }
  }
}

// ---- hoisted passes ---------------------------------------------------------------------------------------------
// The lambdas of the grid scan, the two end points and the score test are the same for every SNP of a run, so everything
// that does not involve x is computed ONCE (lmm_common_kernel): h rows, S^k_ab over (w, y), sum h^k, sum log(l d + 1).
// A hoisted pass then only accumulates the c+2 sums per power that involve this SNP's x, for V2_NSC lambdas at a time,
// with h read from the staged rows instead of being recomputed (no reciprocal, no log): 7 FP64 operations per
// (individual, lambda) for c = 1 instead of 21-54, and 3 passes instead of 7 for the default 10-interval grid.
template <int NC>
struct V2CAcc {
  double X[V2_NSC][2][NC + 2];     // sum h^k v_q x  for q over (w_1..w_c, x, y), k = 1, 2
  double I[NC + 2];                // unit-weight sums (the Iab table of LogRL_f)
  double dl[NC + 1];               // exact minus projected order-1 sums at the l_mle_null row (when the exact linear sums are on)
};

template <int NC>
__device__ __forceinline__ void v2_issue_common(const LmmConst &D, const double *const *xrows, double *stage, int c,
                                                uint64_t *bar, const int (&jrow)[V2_NSC]) {
  constexpr uint32_t ROW = V2_CHUNK * sizeof(double);
  const size_t off0 = (size_t)c * V2_CHUNK;
  uint32_t rows = NC + 1 + V2_NSC;
#pragma unroll
  for (int w = 0; w < V2_WARPS; ++w) rows += (xrows[w] != nullptr) ? 1u : 0u;
  v2_mbar_expect_tx(bar, rows * ROW);
#pragma unroll
  for (int a = 0; a < NC; ++a) v2_bulk_load(stage + (a + 1) * V2_CHUNK, D.Wt + (size_t)a * D.ldv + off0, ROW, bar);
  v2_bulk_load(stage + (NC + 1) * V2_CHUNK, D.y + off0, ROW, bar);
  double *xs = stage + (NC + 2) * V2_CHUNK;
#pragma unroll
  for (int w = 0; w < V2_WARPS; ++w)
    if (xrows[w]) v2_bulk_load(xs + w * V2_CHUNK, xrows[w] + off0, ROW, bar);
  double *hs = stage + (NC + 2 + V2_WARPS) * V2_CHUNK;
#pragma unroll
  for (int s = 0; s < V2_NSC; ++s) v2_bulk_load(hs + s * V2_CHUNK, D.Hrows + (size_t)jrow[s] * D.n_c + off0, ROW, bar);
}

// XONLY: the sums linear in x come from the side GEMM (LmmConst::xsum), the pass only accumulates x'x per lambda and power
// (3 FP64 operations per individual and slot instead of 7; no covariate / phenotype operands)
template <int NC, bool WITH_I, bool XONLY = false>
__device__ __forceinline__ void v2_pass_common(const LmmConst &D, const double *const *xrows, double *smem, int nchunks,
                                               bool active, const int (&jrow)[V2_NSC], V2CAcc<NC> &acc, unsigned int &pipe_it,
                                               const double *xs = nullptr, int jscore = -1) {
  constexpr int NQ = NC + 2;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t stage_d = v2_stage_doubles(NC);
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    acc.I[q] = 0.0;
#pragma unroll
    for (int s = 0; s < V2_NSC; ++s) { acc.X[s][0][q] = 0.0; acc.X[s][1][q] = 0.0; }
  }
  uint64_t *full = v2_bars(smem, NC), *empty = full + V2_STAGES;
  const unsigned int it0 = pipe_it;
#pragma unroll
  for (int c = 0; c < V2_AHEAD; ++c)
    if (c < nchunks) {
      const unsigned int k = it0 + (unsigned int)c;
      if (lane == 0 && warp == (int)(k % V2_WARPS)) {         // rotating producer, see v2_pass
        v2_acquire_stage(empty, k);
        v2_issue_common<NC>(D, xrows, smem + (size_t)(k % V2_STAGES) * stage_d, c, &full[k % V2_STAGES], jrow);
      }
    }
  for (int c = 0; c < nchunks; ++c) {
    {
      const int cn = c + V2_AHEAD;
      const unsigned int kn = it0 + (unsigned int)cn;
      if (cn < nchunks && lane == 0 && warp == (int)(kn % V2_WARPS)) {
        v2_acquire_stage(empty, kn);
        v2_issue_common<NC>(D, xrows, smem + (size_t)(kn % V2_STAGES) * stage_d, cn, &full[kn % V2_STAGES], jrow);
      }
    }
    const unsigned int k = it0 + (unsigned int)c;
    const int sidx = (int)(k % V2_STAGES);
    v2_mbar_wait(&full[sidx], (k / V2_STAGES) & 1u);
    if (active) {
      const double *st = smem + (size_t)sidx * stage_d;
      const double *sl = st + lane;
      const double *xl = st + (NC + 2) * V2_CHUNK + warp * V2_CHUNK + lane;
      const double *hl = st + (NC + 2 + V2_WARPS) * V2_CHUNK + lane;
      // register double buffer: the 8 + NC shared-memory values of individual u + 1 are requested before the 38 FP64 operations of
      // individual u (the h loads sat directly in front of their first use: `short scoreboard` was 17 % of the stall samples)
      double cx, cw[NC], cy = 0.0, ch[V2_NSC];
      cx = xl[0];
      if (!XONLY) {
        cy = sl[(NC + 1) * V2_CHUNK];
#pragma unroll
        for (int a = 0; a < NC; ++a) cw[a] = sl[(a + 1) * V2_CHUNK];
      }
#pragma unroll
      for (int s = 0; s < V2_NSC; ++s) ch[s] = hl[s * V2_CHUNK];
#pragma unroll
      for (int u = 0; u < V2_CHUNK / 32; ++u) {
        double nx = 0.0, nw[NC], ny = 0.0, nh[V2_NSC];
        if (u + 1 < V2_CHUNK / 32) {
          const int j = (u + 1) * 32;
          nx = xl[j];
          if (!XONLY) {
            ny = sl[(NC + 1) * V2_CHUNK + j];
#pragma unroll
            for (int a = 0; a < NC; ++a) nw[a] = sl[(a + 1) * V2_CHUNK + j];
          }
#pragma unroll
          for (int s = 0; s < V2_NSC; ++s) nh[s] = hl[s * V2_CHUNK + j];
        }
        const double x = cx;
        if (XONLY) {
          const double xx = x * x;
          if (WITH_I) acc.I[NC] += xx;
#pragma unroll
          for (int s = 0; s < V2_NSC; ++s) {
            const double h = ch[s];
            const double hx = h * xx;
            acc.X[s][0][NC] += hx;
            acc.X[s][1][NC] = fma(h, hx, acc.X[s][1][NC]);
          }
        } else {
          double px[NQ];
#pragma unroll
          for (int a = 0; a < NC; ++a) px[a] = cw[a] * x;
          px[NC] = x * x;
          px[NC + 1] = x * cy;
          if (WITH_I) {
#pragma unroll
            for (int q = 0; q < NQ; ++q) acc.I[q] += px[q];
          }
#pragma unroll
          for (int s = 0; s < V2_NSC; ++s) {
            const double h = ch[s];
            const double h2 = h * h;
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
              acc.X[s][0][q] = fma(h, px[q], acc.X[s][0][q]);
              acc.X[s][1][q] = fma(h2, px[q], acc.X[s][1][q]);
            }
          }
        }
        if (u + 1 < V2_CHUNK / 32) {
          cx = nx;
          if (!XONLY) {
            cy = ny;
#pragma unroll
            for (int a = 0; a < NC; ++a) cw[a] = nw[a];
          }
#pragma unroll
          for (int s = 0; s < V2_NSC; ++s) ch[s] = nh[s];
        }
      }
    }
    __syncwarp();
    if (lane == 0) v2_mbar_arrive(&empty[sidx]);
  }
  pipe_it = it0 + (unsigned int)nchunks;
  if (active) {
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      if (XONLY && q != NC) continue;                       // only x'x was accumulated; the linear sums are filled in below
      if (WITH_I) acc.I[q] = warp_allsum(acc.I[q]);
#pragma unroll
      for (int s = 0; s < V2_NSC; ++s) { acc.X[s][0][q] = warp_allsum(acc.X[s][0][q]); acc.X[s][1][q] = warp_allsum(acc.X[s][1][q]); }
    }
    if (xs) {
      // the sums LINEAR in x come from the side GEMM in genotype space (LmmConst::xsum); the pass keeps its own value only for x'x
#pragma unroll
      for (int s = 0; s < V2_NSC; ++s) {
        const int b = jrow[s] < D.n_common ? jrow[s] : jrow[s] - (D.xsum_nblocks_skip);
        const double *e = xs + (size_t)b * 2 * (NC + 1);
#pragma unroll
        for (int k = 0; k < 2; ++k) {
#pragma unroll
          for (int a = 0; a < NC; ++a) acc.X[s][k][a] = __ldg(e + k * (NC + 1) + a);
          acc.X[s][k][NC + 1] = __ldg(e + k * (NC + 1) + NC);
        }
      }
      if (WITH_I) {
        const double *e = xs + (size_t)D.xsum_nblocks * 2 * (NC + 1);
#pragma unroll
        for (int a = 0; a < NC; ++a) acc.I[a] = __ldg(e + a);
        acc.I[NC + 1] = __ldg(e + NC);
      }
    }
  }
}

// full (w, x, y) table of one power from the common record (pairs without x) and this SNP's x sums
template <int NC>
__device__ __forceinline__ void v2_assemble(const double *__restrict__ ct_k, const double (&Xq)[NC + 2],
                                            double (&S)[(NC + 3) * (NC + 2) / 2]) {
  constexpr int NV = NC + 2;
#pragma unroll
  for (int a = 0; a < NV; ++a)
#pragma unroll
    for (int b = a; b < NV; ++b) {
      double v;
      if (a == NC) v = (b == NC) ? Xq[NC] : Xq[NC + 1];                 // xx, xy
      else if (b == NC) v = Xq[a];                                       // w_a x
      else v = __ldg(ct_k + abidx(a > NC ? NC : a, b > NC ? NC : b, NC + 1));   // pair over (w, y): y sits at index NC there
      S[abidx(a, b, NV)] = v;
    }
}

// what one slot's sums turn into
struct V2Eval {
  double d1R, d1L, d2R, d2L, fR, fL;
  double P_xx, P_xy, P_yy, Px_yy;       // Wald / score inputs (order-1 tables)
};

// S1/S2/S3 = order 1..3 tables of one slot (S2/S3 may be dummies), tr1/tr2 = sum h, sum h^2
template <int NC, int ORD>
__device__ __forceinline__ void v2_derive(double (&S1)[(NC + 3) * (NC + 2) / 2], double (&S2)[(NC + 3) * (NC + 2) / 2],
                                          double (&S3)[(NC + 3) * (NC + 2) / 2], double tr1, double tr2, double lam,
                                          double n, bool want_f, double logdet_h, double logdetI, V2Eval &o) {
  Derived<NC, ORD> d;
  sweep_tables<NC, ORD>(S1, S2, S3, d);
  const double df = n - (double)NC - 1.0;
  o.P_xx = d.P_xx; o.P_xy = d.P_xy; o.P_yy = d.P_yy; o.Px_yy = d.Px_yy;
  o.d1R = o.d1L = o.d2R = o.d2L = 0.0; o.fR = o.fL = 0.0;
  if (ORD >= 2) {
    const double P_yy = d.Px_yy, PP_yy = d.PPx_yy;
    const double yPKPy = (P_yy - PP_yy) / lam;
    const double trace_P = tr1 - d.trace_P_corr;
    o.d1R = -0.5 * ((df - trace_P) / lam) + 0.5 * df * yPKPy / P_yy;
    o.d1L = -0.5 * ((n - tr1) / lam) + 0.5 * n * yPKPy / P_yy;
    if (ORD >= 3) {
      const double trace_PP = tr2 + d.trace_PP_corr;
      const double yPKPKPy = (P_yy + d.PPPx_yy - 2.0 * PP_yy) / (lam * lam);
      const double quad = (2.0 * yPKPKPy * P_yy - yPKPy * yPKPy) / (P_yy * P_yy);
      o.d2R = 0.5 * ((df + trace_PP - 2.0 * trace_P) / (lam * lam)) - 0.5 * df * quad;
      o.d2L = 0.5 * ((n + tr2 - 2.0 * tr1) / (lam * lam)) - 0.5 * n * quad;
    }
  }
  if (want_f) {
    FVals fv = f_from(n, NC, logdet_h, d.logdet_piv, logdetI, d.Px_yy);
    o.fR = fv.fR; o.fL = fv.fL;
  }
}

// ---- resumable CalcLambda state machine for one likelihood (REML or ML) ---------------------------
enum { V2_SCAN = 0, V2_BRENT_PREP, V2_BRENT_EVAL, V2_NEWTON_INIT, V2_NEWTON_INIT_EVAL, V2_NEWTON_PREP, V2_NEWTON_EVAL,
       V2_F_EVAL, V2_DONE };

struct V2Fn {
  int stage, next_g, status;
  unsigned iter, iter2;
  double a, b, c, d, e, fa, fb, fc, root, xl, xu;        // Brent state (gsl roots/brent.c)
  double a_new, fa_new, b_new;
  double nroot, nf, ndf;                                  // Newton state (gsl roots/newton.c)
  RootState rs;
  // f-evaluation cache for the Wald test
  double cache_lam, cP_xx, cP_xy, cP_yy, cPx_yy;
  // order-1 sums linear in x (w_1 x .. w_c x, x y) at the last f-evaluation (cX) and at the estimate kept so far (bX), taken from the
  // interpolants of the EXACT node values when LmmConst::xsum is on: the final exact pass uses them instead of its own projected sums
  double cX[4], bX[4];
};

struct V2Req { bool need; double lam; int K; bool logdet; bool ext; double ld_ext; };   // ext: f wanted with sum log(l d + 1) supplied (ld_ext, from the run's table) instead of accumulated

__device__ __forceinline__ void v2fn_init(V2Fn &F) {
  F.stage = V2_SCAN; F.next_g = 0; F.status = GB_ST_ERR; F.iter = F.iter2 = 0;
  F.rs.l = F.rs.l_temp = 0.0; F.rs.lambda = nan(""); F.rs.logf = nan("");
  F.rs.have = F.rs.aborted = F.rs.stopped = false;
  F.cache_lam = nan("");
#pragma unroll
  for (int a = 0; a < 4; ++a) { F.cX[a] = nan(""); F.bX[a] = nan(""); }
}

// Advance until an evaluation is required (returns req.need) or the interval list is exhausted.
// `ev_*` carry the result of the evaluation requested by the previous call.
// `single`: the caller walks the grid intervals itself (v2fn_due / v2fn_begin) and gets control back when the chain returns to
// the scan state, i.e. when the current interval is finished.
__device__ __noinline__ void v2fn_advance(V2Fn &F, const double *glam, const double *gd1, int n_region, double l_min,
                                          double l_max, double ev_d1, double ev_d2, double ev_f, V2Req &req, bool single = false) {
  req.need = false; req.ext = false; req.ld_ext = 0.0;
  const unsigned max_iter = 100;
  for (;;) {
    switch (F.stage) {
      case V2_SCAN: {
        if (F.rs.aborted || F.rs.stopped) { F.stage = V2_DONE; return; }
        if (single) return;
        while (F.next_g < n_region && !(gd1[F.next_g] * gd1[F.next_g + 1] <= 0)) F.next_g++;
        if (F.next_g >= n_region) { F.stage = V2_DONE; return; }
        const int g = F.next_g++;
        const double x_lower = glam[g], x_upper = glam[g + 1], f_lower = gd1[g], f_upper = gd1[g + 1];
        F.a = x_lower; F.fa = f_lower; F.b = x_upper; F.fb = f_upper; F.c = x_upper; F.fc = f_upper;
        F.d = x_upper - x_lower; F.e = x_upper - x_lower;
        F.root = 0.5 * (x_lower + x_upper); F.xl = x_lower; F.xu = x_upper;
        F.status = GB_ST_ERR; F.iter = 0;
        F.stage = V2_BRENT_PREP;
        break;
      }
      case V2_BRENT_PREP: {
        F.iter++;
        double a = F.a, b = F.b, c = F.c, d = F.d, e = F.e, fa = F.fa, fb = F.fb, fc = F.fc;
        bool ac_equal = false;
        if ((fb < 0 && fc < 0) || (fb > 0 && fc > 0)) { ac_equal = true; c = a; fc = fa; d = b - a; e = b - a; }
        if (fabs(fc) < fabs(fb)) { ac_equal = true; a = b; b = c; c = a; fa = fb; fb = fc; fc = fa; }
        const double tol = 0.5 * DBL_EPSILON * fabs(b);
        const double m = 0.5 * (c - b);
        bool immediate = false;
        if (fb == 0) { F.root = b; F.xl = b; F.xu = b; F.status = GB_ST_SUCCESS; immediate = true; }
        else if (fabs(m) <= tol) {
          F.root = b;
          if (b < c) { F.xl = b; F.xu = c; } else { F.xl = c; F.xu = b; }
          F.status = GB_ST_SUCCESS; immediate = true;
        }
        F.a = a; F.b = b; F.c = c; F.d = d; F.e = e; F.fa = fa; F.fb = fb; F.fc = fc;
        if (immediate) { F.stage = V2_BRENT_EVAL; ev_d1 = nan(""); F.b_new = nan(""); goto brent_post; }
        if (fabs(e) < tol || fabs(fa) <= fabs(fb)) { d = m; e = m; }
        else {
          double p, q, r, s = fb / fa;
          if (ac_equal) { p = 2 * m * s; q = 1 - s; }
          else {
            q = fa / fc; r = fb / fc;
            p = s * (2 * m * q * (q - r) - (b - a) * (r - 1));
            q = (q - 1) * (r - 1) * (s - 1);
          }
          if (p > 0) q = -q; else p = -p;
          const double lim1 = 3 * m * q - fabs(tol * q), lim2 = fabs(e * q);
          if (2 * p < (lim1 < lim2 ? lim1 : lim2)) { e = d; d = p / q; }
          else { d = m; e = m; }
        }
        F.d = d; F.e = e;
        F.a_new = b; F.fa_new = fb;
        F.b_new = b + ((fabs(d) > tol) ? d : (m > 0 ? +tol : -tol));
        F.stage = V2_BRENT_EVAL;
        req.need = true; req.lam = F.b_new; req.K = 2; req.logdet = false;
        return;
      }
      case V2_BRENT_EVAL: {
        {
          const double fb_new = ev_d1;
          if (!isfinite(fb_new)) F.status = GB_ST_ERR;     // GSL returns before storing the state
          else {
            F.a = F.a_new; F.fa = F.fa_new; F.b = F.b_new; F.fb = fb_new;
            F.root = F.b;
            double cc = F.c;
            if ((F.fb < 0 && F.fc < 0) || (F.fb > 0 && F.fc > 0)) cc = F.a;
            if (F.b < cc) { F.xl = F.b; F.xu = cc; } else { F.xl = cc; F.xu = F.b; }
            F.status = GB_ST_SUCCESS;
          }
        }
      brent_post:
        if (F.status != GB_ST_SUCCESS) { F.stage = V2_NEWTON_INIT; break; }       // `break` out of the do-loop (:2040)
        F.rs.l = F.root;
        if (F.xl > F.xu) { F.status = GB_ST_ERR; F.stage = V2_NEWTON_INIT; break; }
        {
          double min_abs = 0.0;
          if ((F.xl > 0.0 && F.xu > 0.0) || (F.xl < 0.0 && F.xu < 0.0)) min_abs = fmin(fabs(F.xl), fabs(F.xu));
          F.status = (fabs(F.xu - F.xl) < 0.1 * min_abs) ? GB_ST_SUCCESS : GB_ST_CONTINUE;
        }
        if (F.status == GB_ST_CONTINUE && F.iter < max_iter) { F.stage = V2_BRENT_PREP; break; }
        if (F.status == GB_ST_CONTINUE) { F.rs.stopped = true; F.stage = V2_DONE; return; }   // :2057-2060
        F.stage = V2_NEWTON_INIT;
        break;
      }
      case V2_NEWTON_INIT:
        F.stage = V2_NEWTON_INIT_EVAL;
        req.need = true; req.lam = F.rs.l; req.K = 3; req.logdet = false;
        return;
      case V2_NEWTON_INIT_EVAL:
        F.nroot = F.rs.l; F.nf = ev_d1; F.ndf = ev_d2; F.iter2 = 0;
        F.stage = V2_NEWTON_PREP;
        break;
      case V2_NEWTON_PREP:
        F.iter2++;
        if (F.ndf == 0.0) { F.status = GB_ST_ERR; goto newton_end; }
        F.nroot = F.nroot - (F.nf / F.ndf);
        F.stage = V2_NEWTON_EVAL;
        req.need = true; req.lam = F.nroot; req.K = 3; req.logdet = false;
        return;
      case V2_NEWTON_EVAL: {
        F.nf = ev_d1; F.ndf = ev_d2;
        if (!isfinite(F.nf) || !isfinite(F.ndf)) { F.status = GB_ST_ERR; goto newton_end; }
        F.rs.l_temp = F.rs.l;
        F.rs.l = F.nroot;
        F.status = (fabs(F.rs.l - F.rs.l_temp) < 1e-5 * fabs(F.rs.l) || F.rs.l == F.rs.l_temp) ? GB_ST_SUCCESS : GB_ST_CONTINUE;
        if (F.status == GB_ST_CONTINUE && F.iter2 < max_iter && F.rs.l > l_min && F.rs.l < l_max) { F.stage = V2_NEWTON_PREP; break; }
      newton_end:
        if (F.status != GB_ST_SUCCESS) {
          F.rs.aborted = true; F.rs.lambda = nan(""); F.rs.logf = nan("");
          F.stage = V2_DONE; return;
        }
        double l = F.rs.l_temp;                     // the PREVIOUS iterate (src/lmm.cpp:2096)
        if (l < l_min) l = l_min;
        if (l > l_max) l = l_max;
        F.rs.l = l;
        F.stage = V2_F_EVAL;
        req.need = true; req.lam = l; req.K = 1; req.logdet = true;
        return;
      }
      case V2_F_EVAL: {
        const double logf_l = ev_f;
        if (!F.rs.have) { F.rs.logf = logf_l; F.rs.lambda = F.rs.l; F.rs.have = true; }
        else if (F.rs.logf < logf_l) { F.rs.logf = logf_l; F.rs.lambda = F.rs.l; }
        F.stage = V2_SCAN;
        break;
      }
      default:
        return;
    }
  }
}

__device__ __forceinline__ void v2_finalize(RootState &R, double f_min, double f_max, double l_min, double l_max) {
  if (R.aborted) return;
  if (!R.have && !R.stopped) {
    if (f_min >= f_max) { R.lambda = l_min; R.logf = f_min; } else { R.lambda = l_max; R.logf = f_max; }
  } else {
    if (f_min > R.logf) { R.lambda = l_min; R.logf = f_min; }
    if (f_max > R.logf) { R.lambda = l_max; R.logf = f_max; }
  }
}


// ---- interpolated refinement ---------------------------------------------------------------------------------------
// After the grid scan every remaining evaluation of CalcLambda (Brent, Newton, the final f) asks for the sums S^k_ab at a
// lambda INSIDE one grid interval.  As functions of t = log(lambda) they are analytic in the strip |Im t| < pi (the poles sit
// at t = -log(delta_i) +- i pi), so on an interval of width log(10) + 2 x 0.15 their Chebyshev interpolant through V2_CM = 20
// nodes is exact to ~1e-14 relative (measured; the geometric rate is 5^-M).  The node lambdas are the same for every SNP, hence
// "common" rows like the grid lambdas: the x-dependent sums at the nodes cost 4 hoisted passes, the SNP-independent ones are
// tabulated once per run (lmm_cheb_coef_kernel).  The root searches then run on scalars -- same Brent / Newton control flow,
// same iterates to ~1e-12 -- instead of ~10 more passes of 62 FP64 operations per individual.  Third powers come from
// S^3 = S^2 + 1/2 dS^2/dt (dh/dt = h^2 - h).  A lambda outside the padded interval (a diverging Newton step) is served by an
// exact pass, as before.
__device__ __forceinline__ bool v2fn_due(const V2Fn &F, const double *gd1, int g) {
  return F.stage == V2_SCAN && !F.rs.aborted && !F.rs.stopped && (gd1[g] * gd1[g + 1] <= 0);
}
__device__ __forceinline__ void v2fn_begin(V2Fn &F, const double *glam, const double *gd1, int g) {
  const double x_lower = glam[g], x_upper = glam[g + 1], f_lower = gd1[g], f_upper = gd1[g + 1];
  F.next_g = g + 1;
  F.a = x_lower; F.fa = f_lower; F.b = x_upper; F.fb = f_upper; F.c = x_upper; F.fc = f_upper;
  F.d = x_upper - x_lower; F.e = x_upper - x_lower;
  F.root = 0.5 * (x_lower + x_upper); F.xl = x_lower; F.xu = x_upper;
  F.status = GB_ST_ERR; F.iter = 0;
  F.stage = V2_BRENT_PREP;
}

// Clenshaw: value / tau-derivative of sum_k c_k T_k(tau)
template <bool GLOBAL, int M>
__device__ __forceinline__ double v2_cheb_val(const double *c, double tau) {
  double b1 = 0.0, b2 = 0.0;
  const double t2 = tau + tau;
#pragma unroll
  for (int k = M - 1; k >= 1; --k) {
    const double ck = GLOBAL ? __ldg(c + k) : c[k];
    const double b0 = fma(t2, b1, ck) - b2;
    b2 = b1; b1 = b0;
  }
  return fma(tau, b1, (GLOBAL ? __ldg(c) : c[0])) - b2;
}
template <bool GLOBAL, int M>
__device__ __forceinline__ double v2_cheb_der(const double *c, double tau) {
  // p' = sum_{j=0}^{M-2} (j + 1) c_{j+1} U_j(tau)
  double b1 = 0.0, b2 = 0.0;
  const double t2 = tau + tau;
#pragma unroll
  for (int j = M - 2; j >= 0; --j) {
    const double dj = (double)(j + 1) * (GLOBAL ? __ldg(c + j + 1) : c[j + 1]);
    const double b0 = fma(t2, b1, dj) - b2;
    b2 = b1; b1 = b0;
  }
  return b1;
}

template <int NC>
__device__ __forceinline__ void v2_assemble_arr(const double (&C)[(NC + 2) * (NC + 1) / 2], const double (&Xq)[NC + 2],
                                                double (&S)[(NC + 3) * (NC + 2) / 2]) {
  constexpr int NV = NC + 2;
#pragma unroll
  for (int a = 0; a < NV; ++a)
#pragma unroll
    for (int b = a; b < NV; ++b) {
      double v;
      if (a == NC) v = (b == NC) ? Xq[NC] : Xq[NC + 1];
      else if (b == NC) v = Xq[a];
      else v = C[abidx(a > NC ? NC : a, b > NC ? NC : b, NC + 1)];
      S[abidx(a, b, NV)] = v;
    }
}

// all sums at lambda from the interpolants of interval g; coef = this warp's x-sum coefficients (shared memory)
template <int NC, int ORD>
__device__ __noinline__ void v2_interp_eval(const LmmConst &D, const double *coef, int g, double tau, double dtau_dt, double lam,
                                            double n, bool want_f, double logdetI, const double *dlt, V2Eval &ev, double *xlin = nullptr) {
  constexpr int NQ = NC + 2, NIDX = (NC + 3) * (NC + 2) / 2, CN = v2c_nidx(NC), M = V2_CM, XM = V2_XM;
  const double *gc = D.cheb + V2_CHEB_BASE + (size_t)g * (2 * CN + 3) * M;
  double X1[NQ], X2[NQ], X3[NQ], C1[CN], C2[CN], C3[CN];
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    X1[q] = v2_cheb_val<false, XM>(coef + q * XM, tau);
    X2[q] = v2_cheb_val<false, XM>(coef + (NQ + q) * XM, tau);
    X3[q] = (ORD >= 3) ? fma(0.5 * dtau_dt, v2_cheb_der<false, XM>(coef + (NQ + q) * XM, tau), X2[q]) : 0.0;
  }
  if (xlin) {
#pragma unroll
    for (int a = 0; a < NC; ++a) xlin[a] = X1[a];
    xlin[NC] = X1[NC + 1];
  }
  if (dlt && want_f) {           // final f / Wald tables: exact-minus-projected x-sums (LmmConst::xex), order 1 only
#pragma unroll
    for (int a = 0; a < NC; ++a) X1[a] += dlt[a];
    X1[NC + 1] += dlt[NC];
  }
#pragma unroll
  for (int q = 0; q < CN; ++q) {
    C1[q] = v2_cheb_val<true, M>(gc + q * M, tau);
    C2[q] = v2_cheb_val<true, M>(gc + (CN + q) * M, tau);
    C3[q] = (ORD >= 3) ? fma(0.5 * dtau_dt, v2_cheb_der<true, M>(gc + (CN + q) * M, tau), C2[q]) : 0.0;
  }
  const double tr1 = v2_cheb_val<true, M>(gc + 2 * CN * M, tau), tr2 = v2_cheb_val<true, M>(gc + (2 * CN + 1) * M, tau);
  const double ld = want_f ? v2_cheb_val<true, M>(gc + (2 * CN + 2) * M, tau) : 0.0;
  double S1[NIDX], S2[NIDX], S3[NIDX];
  v2_assemble_arr<NC>(C1, X1, S1);
  v2_assemble_arr<NC>(C2, X2, S2);
  if (ORD >= 3) v2_assemble_arr<NC>(C3, X3, S3);
  v2_derive<NC, ORD>(S1, S2, S3, tr1, tr2, lam, n, want_f, ld, logdetI, ev);
}

// Runs one chain inside interval g until the interval is finished (returns false) or it asks for a lambda the interpolant does
// not cover (returns true with `rq` filled; the caller delivers the values through `evv` and calls again).
template <int NC>
__device__ __forceinline__ bool v2_drive(const LmmConst &D, V2Fn &F, bool fnR, const double *coef, int g, double lo, double hi,
                                         const double *glam, const double *gd1, int n_region, double l_min, double l_max,
                                         double n, double logdetI, const double *dlt, double (&evv)[3], V2Req &rq) {
  const double dtau_dt = 2.0 / (hi - lo);
  for (;;) {
    const double prev_lambda = F.rs.lambda; const bool prev_have = F.rs.have;
    v2fn_advance(F, glam, gd1, n_region, l_min, l_max, evv[0], evv[1], evv[2], rq, true);
    if (F.rs.have && (!prev_have || F.rs.lambda != prev_lambda)) {      // the candidate of the last f-evaluation is the estimate kept so far
#pragma unroll
      for (int a = 0; a < 4; ++a) F.bX[a] = F.cX[a];
    }
    if (!rq.need) return false;
    const double tau = (2.0 * log(rq.lam) - (lo + hi)) / (hi - lo);
    if (!(tau >= -1.0 && tau <= 1.0)) return true;          // also catches NaN
    V2Eval ev;
    if (rq.K >= 3) v2_interp_eval<NC, 3>(D, coef, g, tau, dtau_dt, rq.lam, n, rq.logdet, logdetI, dlt, ev, rq.logdet ? F.cX : nullptr);
    else v2_interp_eval<NC, 2>(D, coef, g, tau, dtau_dt, rq.lam, n, rq.logdet, logdetI, dlt, ev, rq.logdet ? F.cX : nullptr);
    evv[0] = fnR ? ev.d1R : ev.d1L; evv[1] = fnR ? ev.d2R : ev.d2L; evv[2] = fnR ? ev.fR : ev.fL;
    if (fnR && rq.logdet) { F.cache_lam = rq.lam; F.cP_xx = ev.P_xx; F.cP_xy = ev.P_xy; F.cP_yy = ev.P_yy; F.cPx_yy = ev.Px_yy; }
    rq.need = false;
  }
}

// One exact lockstep pass for up to two lambdas per warp (slot 0 = REML chain, slot 1 = ML chain).  Every thread of the CTA calls it.
// `dlt` (or null): exact-minus-projected order-1 x-sums (LmmConst::xex), added for requests flagged `ext` (the final evaluations).
template <int NC>
__device__ __noinline__ void v2_exact_pair(const LmmConst &D, const double *const *xrows, double *smem, int nchunks, int pad,
                                              bool active, const V2Req (&rq)[2], double n, double logdetI, V2Eval (&ev)[2],
                                              unsigned int &tally2, unsigned int &tally3, unsigned int &tallyld, unsigned int &pipe_it,
                                              const double *dlt = nullptr, const double *xl0 = nullptr, const double *xl1 = nullptr) {
  constexpr int NIDX = (NC + 3) * (NC + 2) / 2, NV = NC + 2;
  const int k0 = rq[0].need ? rq[0].K : 0, k1 = rq[1].need ? rq[1].K : 0;
  const int Kloc = active ? (k0 > k1 ? k0 : k1) : 0;
  // per warp, like the passes themselves: all shapes walk the same stage ring (same rows, same chunk count)
  const double lam[2] = {rq[0].need ? rq[0].lam : 1.0, rq[1].need ? rq[1].lam : 1.0};
  const bool wf[2] = {rq[0].need && (rq[0].logdet || rq[0].ext), rq[1].need && (rq[1].logdet || rq[1].ext)};     // f wanted
  const bool any_ld = (rq[0].need && rq[0].logdet && !rq[0].ext) || (rq[1].need && rq[1].logdet && !rq[1].ext);   // log accumulated in the pass
  if (active) { if (Kloc >= 3) tally3++; else tally2++; if (any_ld) tallyld++; }
  double dummy[NIDX];
  auto fix = [&](double (&S1)[NIDX], int s2) {            // exact x-sums for the final evaluations
    const double *xl = s2 == 0 ? xl0 : xl1;
    if (xl && rq[s2].ext) {                                // interpolated from the exact node values (LmmConst::xsum)
#pragma unroll
      for (int a = 0; a < NC; ++a) S1[abidx(a, NC, NV)] = xl[a];
      S1[abidx(NC, NC + 1, NV)] = xl[NC];
    } else if (dlt && rq[s2].ext) {
#pragma unroll
      for (int a = 0; a < NC; ++a) S1[abidx(a, NC, NV)] += dlt[a];
      S1[abidx(NC, NC + 1, NV)] += dlt[NC];
    }
  };
  if (Kloc >= 3) {
    V2Acc<NC, 2, 1, 3> acc;
    if (any_ld) v2_pass<NC, 2, 1, 3, true>(D, xrows, smem, nchunks, pad, active, lam, acc, pipe_it);
    else v2_pass<NC, 2, 1, 3, false>(D, xrows, smem, nchunks, pad, active, lam, acc, pipe_it);
    if (active) {
#pragma unroll
      for (int s2 = 0; s2 < 2; ++s2) {
        fix(acc.S[s2][0], s2);
        v2_derive<NC, 3>(acc.S[s2][0], acc.S[s2][1], acc.S[s2][2], acc.tr[s2][0], acc.tr[s2][1], lam[s2], n, wf[s2],
                         rq[s2].ext ? rq[s2].ld_ext : acc.ld[s2], logdetI, ev[s2]);
      }
    }
  } else if (Kloc == 2 || any_ld) {
    V2Acc<NC, 2, 1, 2> acc;
    if (any_ld) v2_pass<NC, 2, 1, 2, true>(D, xrows, smem, nchunks, pad, active, lam, acc, pipe_it);
    else v2_pass<NC, 2, 1, 2, false>(D, xrows, smem, nchunks, pad, active, lam, acc, pipe_it);
    if (active) {
#pragma unroll
      for (int s2 = 0; s2 < 2; ++s2) {
        fix(acc.S[s2][0], s2);
        v2_derive<NC, 2>(acc.S[s2][0], acc.S[s2][1], dummy, acc.tr[s2][0], acc.tr[s2][1], lam[s2], n, wf[s2],
                         rq[s2].ext ? rq[s2].ld_ext : acc.ld[s2], logdetI, ev[s2]);
      }
    }
  } else {                                                 // order-1 tables only (final f / Wald evaluations; idle warps of such a pass)
    V2Acc<NC, 2, 1, 1> acc;
    v2_pass<NC, 2, 1, 1, false>(D, xrows, smem, nchunks, pad, active, lam, acc, pipe_it);
    if (active) {
#pragma unroll
      for (int s2 = 0; s2 < 2; ++s2) {
        fix(acc.S[s2][0], s2);
        v2_derive<NC, 1>(acc.S[s2][0], dummy, dummy, acc.tr[s2][0], 0.0, lam[s2], n, wf[s2], rq[s2].ext ? rq[s2].ld_ext : 0.0, logdetI, ev[s2]);
      }
    }
  }
}

// sum_i log(lambda delta_i + 1) from the run's Chebyshev table of the grid interval that contains lambda
template <int NC>
__device__ __forceinline__ double v2_table_logdet(const LmmConst &D, double lam, double l_min, double interval, int n_region) {
  constexpr int CN = v2c_nidx(NC), M = V2_CM;
  const double t = log(lam), t0 = log(l_min);
  int g = (int)floor((t - t0) / interval);
  g = g < 0 ? 0 : (g >= n_region ? n_region - 1 : g);
  const double lo = t0 + interval * (double)g - D.cheb_marg, hi = t0 + interval * (double)(g + 1) + D.cheb_marg;
  const double tau = (2.0 * t - (lo + hi)) / (hi - lo);
  return v2_cheb_val<true, M>(D.cheb + V2_CHEB_BASE + ((size_t)g * (2 * CN + 3) + (2 * CN + 2)) * M, tau);
}

// One CTA = 8 SNPs.  xrows[w] (shared memory) = U^T x row of warp w or nullptr.
template <int NC>
__device__ __forceinline__ void v2_analyze_group(const LmmConst &D, const LmmParams &prm, const double *const *xrows,
                                                 double *smem, int nchunks, int pad, bool valid, gb200_sumstat &out, unsigned int &pipe_it,
                                                 const double *xe = nullptr, const double *xs = nullptr) {
  constexpr int NIDX = (NC + 3) * (NC + 2) / 2;
  const int mode = prm.a_mode;
  const bool needR = (mode == 1 || mode == 4), needL = (mode == 2 || mode == 4 || mode == 9);
  const bool needS = (mode == 3 || mode == 4 || mode == 9);
  const double l_min = prm.l_min, l_max = prm.l_max, n = (double)D.n;
  const int n_region = prm.n_region;
  const double lambda_interval = log(l_max / l_min) / (double)n_region;
  double beta = 0.0, se = 0.0, p_wald = 0.0, p_lrt = 0.0, p_score = 0.0, logl_H1 = 0.0, lambda_remle = 0.0, lambda_mle = 0.0;
  double glam[V2_MAX_REGION + 1], gd1R[V2_MAX_REGION + 1], gd1L[V2_MAX_REGION + 1];
  double logdetI = 0.0, fRmin = 0.0, fLmin = 0.0, fRmax = 0.0, fLmax = 0.0;
  const bool need_search = needR || needL;
  double dummy[NIDX];
  unsigned int tally2 = 0, tally3 = 0, tallyld = 0, tallyc = 0;   // executed passes of this warp by kind (work counters)
  // order-1 quantities at exactly l_min / l_max from the hoisted grid passes: the Wald test when the REML estimate is an end point
  bool have_bound = false;
  double wminP[4] = {0, 0, 0, 0}, wmaxP[4] = {0, 0, 0, 0};
  // exact-minus-projected order-1 x-sums at l_mle_null (LmmConst::xex), over (w_1..w_c, y)
  bool have_dlt = false;
  double dlt[NC + 1];
#pragma unroll
  for (int a = 0; a <= NC; ++a) dlt[a] = 0.0;

#if GB_V2_TMA
  const bool hoist = (D.ctab != nullptr);
#else
  const bool hoist = false;
#endif
  if (hoist) {
#if GB_V2_TMA
    // ---- hoisted passes: every lambda shared by all SNPs (grid 0..n_region, exactly l_max, l_mle_null), V2_NSC at a time
    constexpr int CS = v2c_stride(NC), CN = v2c_nidx(NC);
    const int n_grid = need_search ? n_region + 2 : 0;
    const bool comp = (xe != nullptr) && (xs == nullptr) && prm.l_mle_null > 0.0;     // the slot at l_mle_null also serves the exact-sum correction (x_exact = 1)
    const int nslots = n_grid + ((needS || comp) ? 1 : 0);
    const int j_score = n_region + 2;
    double S1[NIDX], S2[NIDX];
    for (int s0 = 0; s0 < nslots; s0 += V2_NSC) {
      int jrow[V2_NSC];
#pragma unroll
      for (int s = 0; s < V2_NSC; ++s) {
        const int slot = s0 + s;
        jrow[s] = (slot < nslots) ? (slot < n_grid ? slot : j_score) : 0;
      }
      V2CAcc<NC> acc;
      const bool with_I = (s0 == 0) && need_search;
      // with the exact linear sums on (CTA-uniform: D.xsum) the hoisted passes accumulate x'x only
      const bool xonly = (D.xsum != nullptr);
      if (with_I) { if (xonly) v2_pass_common<NC, true, true>(D, xrows, smem, nchunks, valid, jrow, acc, pipe_it, xs, j_score);
                    else v2_pass_common<NC, true, false>(D, xrows, smem, nchunks, valid, jrow, acc, pipe_it, xs, j_score); }
      else { if (xonly) v2_pass_common<NC, false, true>(D, xrows, smem, nchunks, valid, jrow, acc, pipe_it, xs, j_score);
             else v2_pass_common<NC, false, false>(D, xrows, smem, nchunks, valid, jrow, acc, pipe_it, xs, j_score); }
      if (valid) {
        tallyc += V2_NSC;
        if (with_I) {
          v2_assemble<NC>(D.ctab, acc.I, S1);
          Derived<NC, 1> dI;
          sweep_tables<NC, 1>(S1, dummy, dummy, dI);
          logdetI = dI.logdet_piv;
        }
#pragma unroll
        for (int s = 0; s < V2_NSC; ++s) {
          const int slot = s0 + s;
          if (slot < nslots) {
            const double *ct = D.ctab + (size_t)jrow[s] * CS;
            const double tr1 = __ldg(ct + 3 * CN), tr2 = __ldg(ct + 3 * CN + 1), ldj = __ldg(ct + 3 * CN + 2), lamj = __ldg(ct + 3 * CN + 3);
            v2_assemble<NC>(ct + CN, acc.X[s][0], S1);
            if (slot < n_grid && slot <= n_region) {               // grid point: dev1 of both likelihoods (+ f at l_min)
              v2_assemble<NC>(ct + 2 * CN, acc.X[s][1], S2);
              V2Eval ev;
              v2_derive<NC, 2>(S1, S2, dummy, tr1, tr2, lamj, n, slot == 0, ldj, logdetI, ev);
              glam[slot] = lamj; gd1R[slot] = ev.d1R; gd1L[slot] = ev.d1L;
              if (slot == 0) { fRmin = ev.fR; fLmin = ev.fL; wminP[0] = ev.P_xx; wminP[1] = ev.P_xy; wminP[2] = ev.P_yy; wminP[3] = ev.Px_yy; }
            } else if (slot < n_grid) {                            // exactly l_max: f only
              V2Eval ev;
              v2_derive<NC, 1>(S1, dummy, dummy, tr1, 0.0, lamj, n, true, ldj, logdetI, ev);
              fRmax = ev.fR; fLmax = ev.fL;
              wmaxP[0] = ev.P_xx; wmaxP[1] = ev.P_xy; wmaxP[2] = ev.P_yy; wmaxP[3] = ev.Px_yy; have_bound = true;
            } else {                                               // the slot at l_mle_null
              if (comp && xs) {
                // the pass has already put the exact sums in place; the final exact pass takes its linear sums from the interpolants
              } else if (comp) {
#pragma unroll
                for (int a = 0; a < NC; ++a) { const double e = __ldg(xe + a); dlt[a] = e - acc.X[s][0][a]; S1[abidx(a, NC, NC + 2)] = e; }
                const double ey = __ldg(xe + NC);
                dlt[NC] = ey - acc.X[s][0][NC + 1]; S1[abidx(NC, NC + 1, NC + 2)] = ey;
                have_dlt = true;
              }
              if (needS) {                                         // score test ("3 is before 1")
                Derived<NC, 1> d;
                sweep_tables<NC, 1>(S1, dummy, dummy, d);
                wald_score_from<NC>(d, D.n, true, beta, se, p_score);
              }
            }
          }
        }
      }
    }
#endif
  } else {
  if (need_search) {
    // ---- pass A: lambda_0 = l_min (powers 0..2 + logdet): unit-weight pivots, dev1, f(l_min)
    {
      V2Acc<NC, 1, 0, 2> acc;
      const double lam[1] = {l_min * exp(lambda_interval * 0.0)};
      v2_pass<NC, 1, 0, 2, true>(D, xrows, smem, nchunks, pad, valid, lam, acc, pipe_it);
      if (valid) {
        Derived<NC, 1> dI;
        sweep_tables<NC, 1>(acc.S[0][0], dummy, dummy, dI);
        logdetI = dI.logdet_piv;
        V2Eval ev;
        v2_derive<NC, 2>(acc.S[0][1], acc.S[0][2], dummy, acc.tr[0][1], acc.tr[0][2], lam[0], n, true, acc.ld[0], logdetI, ev);
        glam[0] = lam[0]; gd1R[0] = ev.d1R; gd1L[0] = ev.d1L; fRmin = ev.fR; fLmin = ev.fL;
      }
    }
    // ---- grid passes: V2_NSG lambdas at a time
    for (int g0 = 1; g0 <= n_region; g0 += V2_NSG) {
      V2Acc<NC, V2_NSG, 1, 2> acc;
      double lam[V2_NSG];
#pragma unroll
      for (int s = 0; s < V2_NSG; ++s) lam[s] = (g0 + s <= n_region) ? l_min * exp(lambda_interval * (double)(g0 + s)) : 1.0;
      v2_pass<NC, V2_NSG, 1, 2, false>(D, xrows, smem, nchunks, pad, valid, lam, acc, pipe_it);
      if (valid) {
#pragma unroll
        for (int s = 0; s < V2_NSG; ++s) {
          if (g0 + s <= n_region) {
            V2Eval ev;
            v2_derive<NC, 2>(acc.S[s][0], acc.S[s][1], dummy, acc.tr[s][0], acc.tr[s][1], lam[s], n, false, 0.0, 0.0, ev);
            glam[g0 + s] = lam[s]; gd1R[g0 + s] = ev.d1R; gd1L[g0 + s] = ev.d1L;
          }
        }
      }
    }
  }
  // ---- pass C: f(l_max) and the score test at l_mle_null share one pass
  if (need_search || needS) {
    V2Acc<NC, 2, 1, 1> acc;
    const double lam[2] = {l_max, needS ? prm.l_mle_null : 1.0};
    v2_pass<NC, 2, 1, 1, true>(D, xrows, smem, nchunks, pad, valid, lam, acc, pipe_it);
    if (valid) {
      V2Eval ev;
      v2_derive<NC, 1>(acc.S[0][0], dummy, dummy, acc.tr[0][0], 0.0, lam[0], n, true, acc.ld[0], logdetI, ev);
      fRmax = ev.fR; fLmax = ev.fL;
      if (needS) {
        Derived<NC, 1> d;
        sweep_tables<NC, 1>(acc.S[1][0], dummy, dummy, d);
        wald_score_from<NC>(d, D.n, true, beta, se, p_score);
      }
    }
  }
  }
  // ---- refinement: REML and ML chains side by side, then (if needed) the Wald pass
  V2Fn FR, FL;
  v2fn_init(FR); v2fn_init(FL);
  if (!needR || !valid) FR.stage = V2_DONE;
  if (!needL || !valid) FL.stage = V2_DONE;
  V2Req rq[2]; rq[0].need = rq[1].need = false; rq[0].ext = rq[1].ext = false; rq[0].ld_ext = rq[1].ld_ext = 0.0;
  if (need_search) {
    double evR[3] = {0, 0, 0}, evL[3] = {0, 0, 0};     // d1, d2, f delivered to each chain
    bool finalized = false, wald_pending = false, wald_done = !needR;
#if GB_V2_TMA
    if (hoist && D.cheb != nullptr) {
      // ---- interpolated refinement: the grid intervals in the reference's order; per interval 4 node passes, then scalars only
      constexpr int NQ = NC + 2, M = V2_XM;                // x-sum nodes; the SNP-independent tables have V2_CM nodes per interval
      double *coef = smem + v2_stage_doubles(NC) * V2_STAGES + 8 + (size_t)(threadIdx.x >> 5) * v2_cheb_warp_doubles(NC);
      double *stg = coef + 2 * NQ * M;
      const int lane = threadIdx.x & 31;
      const double t0 = log(l_min);
      for (int g = 0; g < n_region; ++g) {
        const bool dueR = valid && v2fn_due(FR, gd1R, g), dueL = valid && v2fn_due(FL, gd1L, g);
        const bool due = dueR || dueL;
        if (!__syncthreads_or(due ? 1 : 0)) continue;
        if (due)
          for (int o = lane; o < 2 * NQ * M; o += 32) coef[o] = 0.0;
        for (int p0 = 0; p0 < M; p0 += V2_NSC) {
          int jrow[V2_NSC];
#pragma unroll
          for (int s2 = 0; s2 < V2_NSC; ++s2) jrow[s2] = D.n_common + n_region * V2_CM + g * M + p0 + s2;   // x-node rows follow the table-node rows
          V2CAcc<NC> acc;
          if (D.xsum != nullptr) v2_pass_common<NC, false, true>(D, xrows, smem, nchunks, due, jrow, acc, pipe_it, xs, -1);
          else v2_pass_common<NC, false, false>(D, xrows, smem, nchunks, due, jrow, acc, pipe_it, xs, -1);
          if (due) {
            tallyc += V2_NSC;
            __syncwarp();
            if (lane == 0) {
#pragma unroll
              for (int s2 = 0; s2 < V2_NSC; ++s2)
#pragma unroll
                for (int q = 0; q < NQ; ++q) { stg[(s2 * 2 + 0) * NQ + q] = acc.X[s2][0][q]; stg[(s2 * 2 + 1) * NQ + q] = acc.X[s2][1][q]; }
            }
            __syncwarp();
            // discrete cosine transform of the node values, 5 nodes at a time: c_k += f_m cos(pi k (m + 1/2) / M)
            for (int o = lane; o < 2 * NQ * M; o += 32) {
              const int kq = o / M, kk = o - kq * M;
              double a = coef[o];
#pragma unroll
              for (int s2 = 0; s2 < V2_NSC; ++s2)
                a = fma(stg[s2 * 2 * NQ + kq], __ldg(D.cheb + 4 * V2_CM + (kk * (2 * (p0 + s2) + 1)) % (4 * M)), a);   // cos(pi j / (2 V2_XM))
              coef[o] = a;
            }
            __syncwarp();
          }
        }
        if (due) {
          for (int o = lane; o < 2 * NQ * M; o += 32) coef[o] *= ((o % M) == 0) ? (1.0 / M) : (2.0 / M);
          __syncwarp();
          if (dueR) v2fn_begin(FR, glam, gd1R, g);
          if (dueL) v2fn_begin(FL, glam, gd1L, g);
        }
        const double lo = t0 + lambda_interval * (double)g - D.cheb_marg, hi = t0 + lambda_interval * (double)(g + 1) + D.cheb_marg;
        bool runR = dueR, runL = dueL;
        for (;;) {
          bool blkR = false, blkL = false;
          if (runR) { blkR = v2_drive<NC>(D, FR, true, coef, g, lo, hi, glam, gd1R, n_region, l_min, l_max, n, logdetI, have_dlt ? dlt : nullptr, evR, rq[0]); runR = blkR; }
          if (runL) { blkL = v2_drive<NC>(D, FL, false, coef, g, lo, hi, glam, gd1L, n_region, l_min, l_max, n, logdetI, have_dlt ? dlt : nullptr, evL, rq[1]); runL = blkL; }
          if (!blkR) rq[0].need = false;
          if (!blkL) rq[1].need = false;
          const bool blocked = blkR || blkL;
          if (!__syncthreads_or(blocked ? 1 : 0)) break;
          V2Eval ev[2];
          v2_exact_pair<NC>(D, xrows, smem, nchunks, pad, blocked, rq, n, logdetI, ev, tally2, tally3, tallyld, pipe_it);
          if (blkR && rq[0].logdet) FR.cX[0] = nan("");       // this f-evaluation did not come from the exact interpolants
          if (blkL && rq[1].logdet) FL.cX[0] = nan("");
          if (blkR) {
            evR[0] = ev[0].d1R; evR[1] = ev[0].d2R; evR[2] = ev[0].fR;
            if (rq[0].logdet) { FR.cache_lam = rq[0].lam; FR.cP_xx = ev[0].P_xx; FR.cP_xy = ev[0].P_xy; FR.cP_yy = ev[0].P_yy; FR.cPx_yy = ev[0].Px_yy; }
          }
          if (blkL) { evL[0] = ev[1].d1L; evL[1] = ev[1].d2L; evL[2] = ev[1].fL; }
        }
      }
      if (FR.stage == V2_SCAN) FR.stage = V2_DONE;           // interval list exhausted
      if (FL.stage == V2_SCAN) FL.stage = V2_DONE;
      rq[0].need = rq[1].need = false;
      // ---- final evaluations: f and the order-1 tables at the estimates the searches ended on, by ONE exact pass (the x-sum
      // interpolants only steered the searches).  An end-point estimate already has exact values from the grid passes.
      {
        bool exR = false, exL = false;
        if (valid && needR && FR.rs.have && !FR.rs.aborted) { RootState t = FR.rs; v2_finalize(t, fRmin, fRmax, l_min, l_max); exR = (t.lambda == FR.rs.lambda) && isfinite(t.lambda); }
        if (valid && needL && FL.rs.have && !FL.rs.aborted) { RootState t = FL.rs; v2_finalize(t, fLmin, fLmax, l_min, l_max); exL = (t.lambda == FL.rs.lambda) && isfinite(t.lambda); }
        if (__syncthreads_or((exR || exL) ? 1 : 0)) {
          rq[0].need = exR; rq[0].lam = exR ? FR.rs.lambda : 1.0; rq[0].K = 1; rq[0].logdet = false; rq[0].ext = exR;
          rq[0].ld_ext = exR ? v2_table_logdet<NC>(D, FR.rs.lambda, l_min, lambda_interval, n_region) : 0.0;
          rq[1].need = exL; rq[1].lam = exL ? FL.rs.lambda : 1.0; rq[1].K = 1; rq[1].logdet = false; rq[1].ext = exL;
          rq[1].ld_ext = exL ? v2_table_logdet<NC>(D, FL.rs.lambda, l_min, lambda_interval, n_region) : 0.0;
          V2Eval ev[2];
          v2_exact_pair<NC>(D, xrows, smem, nchunks, pad, exR || exL, rq, n, logdetI, ev, tally2, tally3, tallyld, pipe_it, have_dlt ? dlt : nullptr,
                            (xs && exR && isfinite(FR.bX[0])) ? FR.bX : nullptr, (xs && exL && isfinite(FL.bX[0])) ? FL.bX : nullptr);
          if (exR) { FR.rs.logf = ev[0].fR; FR.cache_lam = FR.rs.lambda; FR.cP_xx = ev[0].P_xx; FR.cP_xy = ev[0].P_xy; FR.cP_yy = ev[0].P_yy; FR.cPx_yy = ev[0].Px_yy; }
          if (exL) FL.rs.logf = ev[1].fL;
          rq[0].need = rq[1].need = false; rq[0].ext = rq[1].ext = false;
        }
      }
    }
#endif
    for (;;) {
      if (FR.stage != V2_DONE) v2fn_advance(FR, glam, gd1R, n_region, l_min, l_max, evR[0], evR[1], evR[2], rq[0]); else rq[0].need = false;
      if (FL.stage != V2_DONE) v2fn_advance(FL, glam, gd1L, n_region, l_min, l_max, evL[0], evL[1], evL[2], rq[1]); else rq[1].need = false;
      if (valid && !finalized && FR.stage == V2_DONE && FL.stage == V2_DONE) {
        if (needR) v2_finalize(FR.rs, fRmin, fRmax, l_min, l_max);
        if (needL) v2_finalize(FL.rs, fLmin, fLmax, l_min, l_max);
        finalized = true;
        if (needR) {
          lambda_remle = FR.rs.lambda; logl_H1 = FR.rs.logf;
          if (prm.plink_rule && isnan(logl_H1)) {          // AnalyzePlink: `if (!isnan(logl_H1)) CalcRLWald(...)` (lmm.cpp:1869-1870)
            wald_done = true;
          } else if (lambda_remle == FR.cache_lam) {            // f(lambda_hat) pass already produced the order-1 table
            Derived<NC, 1> d; d.P_xx = FR.cP_xx; d.P_xy = FR.cP_xy; d.P_yy = FR.cP_yy; d.Px_yy = FR.cPx_yy;
            wald_score_from<NC>(d, D.n, false, beta, se, p_wald);
            wald_done = true;
          } else if (have_bound && (lambda_remle == l_min || lambda_remle == l_max)) {   // end point: tables of the hoisted grid passes
            const double *w = (lambda_remle == l_min) ? wminP : wmaxP;
            Derived<NC, 1> d; d.P_xx = w[0]; d.P_xy = w[1]; d.P_yy = w[2]; d.Px_yy = w[3];
            wald_score_from<NC>(d, D.n, false, beta, se, p_wald);
            wald_done = true;
          } else wald_pending = true;
        }
      }
      const bool want_wald = valid && finalized && wald_pending && !wald_done;
      if (want_wald) { rq[0].need = true; rq[0].lam = lambda_remle; rq[0].K = 1; rq[0].logdet = false; rq[0].ext = false; }
      const bool active = valid && (rq[0].need || rq[1].need);
      if (!__syncthreads_or(active ? 1 : 0)) break;
      const double lam[2] = {rq[0].need ? rq[0].lam : 1.0, rq[1].need ? rq[1].lam : 1.0};
      V2Eval ev[2];
      v2_exact_pair<NC>(D, xrows, smem, nchunks, pad, active, rq, n, logdetI, ev, tally2, tally3, tallyld, pipe_it);
      if (active) {
        if (want_wald) {
          Derived<NC, 1> d; d.P_xx = ev[0].P_xx; d.P_xy = ev[0].P_xy; d.P_yy = ev[0].P_yy; d.Px_yy = ev[0].Px_yy;
          wald_score_from<NC>(d, D.n, false, beta, se, p_wald);
          wald_done = true; rq[0].need = false;
        } else if (rq[0].need) {
          evR[0] = ev[0].d1R; evR[1] = ev[0].d2R; evR[2] = ev[0].fR;
          if (rq[0].logdet) { FR.cache_lam = lam[0]; FR.cP_xx = ev[0].P_xx; FR.cP_xy = ev[0].P_xy; FR.cP_yy = ev[0].P_yy; FR.cPx_yy = ev[0].Px_yy; }
        }
        if (rq[1].need) { evL[0] = ev[1].d1L; evL[1] = ev[1].d2L; evL[2] = ev[1].fL; }
      }
    }
    if (valid && needL) {
      lambda_mle = FL.rs.lambda; logl_H1 = FL.rs.logf;
      p_lrt = chisq1_Q_dev(2.0 * (logl_H1 - prm.logl_mle_H0));
    }
    if (valid && prm.plink_rule && isnan(logl_H1)) { p_wald = logl_H1; p_lrt = logl_H1; }   // lmm.cpp:1882-1884
  }
  out.beta = beta; out.se = se; out.lambda_remle = lambda_remle; out.lambda_mle = lambda_mle;
  out.p_wald = p_wald; out.p_lrt = p_lrt; out.p_score = p_score; out.logl_H1 = logl_H1;
  if (D.cnt && valid && (threadIdx.x & 31) == 0) {
    atomicAdd(D.cnt + 0, (unsigned long long)tallyc); atomicAdd(D.cnt + 1, (unsigned long long)tally2);
    atomicAdd(D.cnt + 2, (unsigned long long)tally3); atomicAdd(D.cnt + 3, (unsigned long long)tallyld);
    atomicAdd(D.cnt + 5, 1ull);
  }
}

}  // namespace gb
