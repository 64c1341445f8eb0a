// common.cuh -- context object and helpers shared by the translation units of libgemma_b200.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>
#include <map>
#include <utility>
#include "../../include/gemma_b200.h"
#include "lmm_device.cuh"

namespace gb {

struct ProfEntry {
  double ms = 0.0;
  long launches = 0;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending;
};

// grow-only device buffer
struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
  bool owned = true;
  cudaError_t reserve(size_t bytes) {
    if (owned && bytes <= cap) return cudaSuccess;
    release();
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e == cudaSuccess) cap = bytes;
    return e;
  }
  void adopt(void *ptr, size_t bytes) { release(); p = ptr; cap = bytes; owned = false; }   // borrowed, never freed
  void release() { if (p && owned) cudaFree(p); p = nullptr; cap = 0; owned = true; }
  template <typename T> T *as() const { return reinterpret_cast<T *>(p); }
};

// int8 tensor-core projection state (i8gemm_sm100.cu)
struct I8State {
  bool ready = false;
  int n_slices = 0;
  int auto_T = 0;                // plane count chosen from this U's column maxima (0 = not evaluated yet)
  double colmax_max = 0.0;       // largest |entry| of U
  size_t n = 0, n_pad = 0;       // individuals, padded to the K tile
  DevBuf slices;                 // n_slices x [n_pad(i: eigvec) x n_pad(j: individual)] int8, j contiguous
  DevBuf scale;                  // per eigenvector: scale s_i, column maximum, 1 / s_i (3 n doubles)
  DevBuf geno;                   // l_pad x n_pad int8 genotype tile source
  DevBuf miss_mean;              // per-SNP mean + hole count (+ holes per 256-SNP tile)
  // exact linear x-sums at every hoisted lambda (LmmConst::xsum): V = U A (n x xs_ncol) in individual space, its digit planes, the result of the last batch
  DevBuf xs_V, xs_planes, xs_scale, xs_out;
  void *tmap_v = nullptr;
  int xs_ncol = 0, xs_T = 0, xs_NE = 0, xs_groups = 0;
  DevBuf xs_patch_idx; int xs_npatch = 0, xs_patch0 = 0;   // eigenvector indices whose U^T x entries are overwritten by exact values (columns xs_patch0.. of the side GEMM)
  size_t xs_ld = 0;
  bool xs_ready = false, xs_valid = false;
  bool no_xsum_consumer = false;   // set by a projection that does not produce the exact sums (dosage rows): the plane count then stays >= 4
  const double *xs_for = nullptr; size_t xs_l = 0;
  DevBuf xex;                    // exact order-1 x-sums of the batch last projected by i8_project_bed (l x (n_cvt + 1)); valid flag below
  bool xex_valid = false;
  const double *xex_for = nullptr; size_t xex_l = 0;   // the U^T X buffer / row count those sums belong to
  DevBuf wave_ctr;               // wave synchronisation counter of the CTA-pair projection kernel
  DevBuf holeq;                  // l_pad x n_pad int8 hole-indicator rows (second GEMM pass of the mean imputation)
  void *tmap_a = nullptr, *tmap_b = nullptr, *tmap_q = nullptr;   // CUtensorMap storage (host)
  // kinship (K = Z Z^T on the int8 tensor pipe)
  DevBuf kin_zt;                 // individual-major int8 genotypes: n rows x kin_cap SNP columns
  DevBuf kin_stats;              // per staged SNP: int sum, int nmiss, double mean
  DevBuf kin_a;                  // a[i] = sum_s mean_s z_s[i]  (n doubles) + beta + flag
  DevBuf kin_tiles;              // lower-triangle tile list (int2)
  DevBuf kin_qbits;              // missing-genotype bit rows: n x (kin_cap / 64) 64-bit words
  DevBuf kin_y;                  // Y (n x n) + b (n): sparse missing-genotype corrections, allocated on first use
  bool kin_y_used = false;
  size_t kin_cap = 0, kin_fill = 0, kin_n = 0;
  int kin_num_tiles = 0, kin_num_tiles_pair = 0;
  bool tmap_b_half = false;
  bool kin_used = false;
  void *tmap_ka = nullptr, *tmap_kb = nullptr;
};

// multivariate LMM (two phenotypes): run constants and null-model results on the device
struct MvConst {
  int n, ld;                      // individuals, leading dimension of the rows below
  const double *delta, *Wt, *Yt;  // eigenvalues; c rows of U^T W; 2 rows of U^T Y
  double vg0[2], ve0[2];          // univariate REML estimates (MphInitial diagonals)
  int em_iter, nr_iter; double em_prec, nr_prec, p_nr;
  int a_mode;                     // 1 Wald, 2 LRT, 3 score, 4 all
};
struct MvNull {
  double Vg_remle[4], Ve_remle[4], B_remle[8], logl_remle;
  double Vg_mle[4], Ve_mle[4], B_mle[8], logl_mle;
};

}  // namespace gb

struct gb200_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int num_sms = 0;
  std::string err;
  // profiling
  bool prof = false;
  std::map<std::string, gb::ProfEntry> profs;
  std::vector<cudaEvent_t> event_pool;
  // -gk
  size_t kin_n = 0;
  int kin_mode = 0;
  size_t kin_ns = 0;
  bool kin_active = false;
  gb::DevBuf dK;
  // -lmm
  size_t n = 0, n_cvt = 0;
  bool lmm_ready = false, prm_ready = false;
  gb::DevBuf dU, dEval, dWt, dY;
  gb::LmmParams prm{};
  gb::DevBuf dNull;
  gb::DevBuf dWtx, dEnv, dX2, dFlip;   // G x E: expanded covariate rows (W, env, -), env, x*env batch, allele-flip flags
  bool gxe_ready = false;
  gb::DevBuf dLmW, dLmY, dLmSmall;     // -lm: W rows (c x n), y, [WtWi (c x c) | Wty (c)]
  size_t lm_n = 0, lm_c = 0; double lm_yPwy = 0.0; bool lm_ready = false;
  gb::DevBuf dMvY, dMvNull, dMvOut;      // multivariate LMM: U^T Y rows (2 x n_c), MvNull, per-SNP output rows
  gb::MvConst mvK; bool mv_ready = false, mv_null_ready = false;
  gb::DevBuf dHrows, dCtab;     // common-lambda h rows / records of the lockstep kernel (lmm_v2.cuh hoisted passes)
  gb::DevBuf dVnull;            // v_q = U (h(l_mle_null) (.) q), q over (w, y): exact x-sums of int8-projected batches (LmmConst::xex)
  bool vnull_ready = false;
  int xs_nblocks = 0, xs_skip = 0;   // LmmConst::xsum layout of the current common tables
  gb::DevBuf dCheb, dNodeLam;   // Chebyshev tables / node lambdas of the interpolated refinement
  bool common_ready = false;
  // scratch
  gb::DevBuf dX, dUtXt, dOut, dBed, dMask, dIdx, dTicket, dTmp;
  std::vector<int> idx_host;          // analysed-individual index cache for bed batches
  std::vector<unsigned char> mask_host;
  // sub-batch software pipeline: projection of sub-batch i+1 (tensor pipe) overlaps the per-SNP tests of
  // sub-batch i (FP64 pipe) on a second stream; both kernels are sized to co-reside on an SM
  cudaStream_t side = nullptr;
  cudaEvent_t evG[2] = {nullptr, nullptr}, evL[2] = {nullptr, nullptr};
  gb::DevBuf dUtXt2;
  // double-buffered host -> device streamer of gb200_lmm_batch_bed
  cudaStream_t copy = nullptr;
  cudaEvent_t evCopy[2] = {nullptr, nullptr}, evUsed[2] = {nullptr, nullptr}, evStart = nullptr;
  gb::DevBuf dBed2;
  long stage_mask = 3;        // measurement knob of the bed entry points: bit 0 = projection, bit 1 = per-SNP tests (kernel-isolated power / ncu runs)
  long batch_chunk = 0;       // SNPs per internal sub-batch of the bed entry points (0 = auto, see lmm_chunk_snps)
  long overlap = 0;           // 1 = pipelined sub-batches.  Measured SLOWER on B200 (136 vs 113 ms per 8192 SNPs at n = 50 000:
                              // the co-resident kernels contend and the power cap bites harder), so off by default
  bool count_work = false;    // lockstep kernel tallies its executed passes (gb200_lmm_counters)
  long kernel_launches = 0;   // kernels of this library launched so far (bench "gpu_launches")
  // options
  long utx_path = 0;     // 0 auto, 1 fp64 tiled, 2 int8 tensor core
  long n_slices = 0;     // 0 = default
  long cta_pair = 1;     // projection kernel as CTA pairs (tcgen05 cta_group::2): -25% time at n = 50 000
  long gemm_groups = 1;  // 2: pair kernel with two eigenvector groups per tile (shared genotype tile) and the hole pass on the tensor pipe; 1: one group, FP64 hole fix-up
  long x_exact = 2;        // int8-projected PLINK batches: 2 = exact LINEAR x-sums at every hoisted lambda from a side GEMM in genotype space
                           // (LmmConst::xsum; lets the projection run on 3 digit planes), 1 = exact sums at l_mle_null only (LmmConst::xex), 0 = off
  long hole_gemm = 1;      // CTA-pair projection: batches with many missing genotypes add mean * U^T q by a second GEMM pass over the hole-indicator rows (decided on the device); 0 = always the gather kernel
  long gemm_wave_sync = 1; // CTA-pair projection: producers start every tile wave together (keeps the K-panels shared through L2)
  long gemm_l2hint = 0;  // CTA-pair projection: L2 eviction hints on the TMA loads (A/B measurement)
  long gemm_stages = 0;  // TMA pipeline stages of the CTA-pair projection kernel (0 = as many 32 KB stages as fit, at most 6)
  long gemm_panel = 0;   // raster panel width of the projection kernels in eigenvector groups (0 = default: 9 for the CTA-pair kernel; 2-group units, 6, for gemm_groups = 2)
  long kin_cta_pair = 0; // kinship kernel as CTA pairs (no gain measured on the short kinship launches)
  long kin_path = 0;     // 0 auto (int8 tensor cores for centred K; sparse FP64 terms for missing genotypes), 1 = FP64 only
  double kin_miss_max = 0.2;   // chunks with a larger fraction of missing genotypes take the dense FP64 path
  long lmm_kernel = 0;   // 0 auto (v2 when supported), 1 = v1 warp-per-SNP, 2 = v2 lockstep CTA, 3 = any-covariate-count kernel
  long lmm_interp = 1;   // lockstep kernel: Brent / Newton evaluations served by Chebyshev interpolants over each grid interval (needs lmm_hoist)
  long lmm_hoist = 1;    // lockstep kernel: SNP-independent sums at the shared lambdas computed once per run
  size_t n_c = 0;        // n rounded up to 512 (vector / UtX row padding)
  long eigh_path = 0;    // 0 auto (cusolverDnXsyevd up to n = 32768, cusolverMgSyevd on this one device beyond), 1 = Xsyevd, 2 = MgSyevd
  size_t eigh_workspace_bytes = 0;   // device workspace of the last eigendecomposition (reported by bench.py)
  gb::I8State i8;
};

namespace gb {

inline int set_err(gb200_ctx *c, int code, const std::string &msg) {
  if (c) c->err = msg;
  return code;
}

#define GB_CUDA(ctx, call)                                                                       \
  do {                                                                                           \
    cudaError_t _e = (call);                                                                     \
    if (_e != cudaSuccess) {                                                                     \
      char _b[512];                                                                              \
      snprintf(_b, sizeof(_b), "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, \
               __LINE__);                                                                        \
      return gb::set_err(ctx, GB200_ERR_CUDA, _b);                                               \
    }                                                                                            \
  } while (0)

// RAII-less profiling scope: records start/stop events on the context stream when enabled.
struct ProfScope {
  gb200_ctx *c;
  ProfEntry *e = nullptr;
  cudaEvent_t a = nullptr, b = nullptr;
  long launches = 1;          // kernels launched inside this scope
  static cudaEvent_t get_event(gb200_ctx *c) {
    if (!c->event_pool.empty()) { cudaEvent_t ev = c->event_pool.back(); c->event_pool.pop_back(); return ev; }
    cudaEvent_t ev; cudaEventCreate(&ev); return ev;
  }
  cudaStream_t st;
  ProfScope(gb200_ctx *ctx, const char *name, long n_launch = 1, cudaStream_t stream = nullptr)
      : c(ctx), launches(n_launch), st(stream ? stream : ctx->stream) {
    if (!c->prof) return;
    e = &c->profs[name];
    a = get_event(c); b = get_event(c);
    cudaEventRecord(a, st);
  }
  ~ProfScope() {
    c->kernel_launches += launches;
    if (!e) return;
    cudaEventRecord(b, st);
    e->pending.emplace_back(a, b);
    e->launches++;
  }
};

// ---- kernels / launchers implemented in the other translation units ----
cudaError_t launch_lmm_assoc(int n_cvt, const LmmConst &D, const LmmParams &prm, const double *UtXt,
                             size_t ldu, int l, gb200_sumstat *out, unsigned int *ticket, int num_sms,
                             cudaStream_t st);
bool lmm_v2_supported(int n_cvt, int n_region);
cudaError_t launch_lmm_gxe(int c_base, LmmConst D, const LmmParams &prm, const double *UtX1t, const double *UtX2t, size_t ldu, int l,
                           const unsigned char *flip, gb200_sumstat *out, unsigned int *ticket, int num_sms, cudaStream_t st);
cudaError_t launch_gxe_prepare(double *X1, double *X2, const double *env, size_t l, size_t n, unsigned char *flip, cudaStream_t st);
cudaError_t launch_mv_null(int c, const MvConst &K, MvNull *out, cudaStream_t st);
cudaError_t launch_mv_assoc(int c, const MvConst &K, const MvNull *nm, const double *UtXt, size_t ldu, int l, double *out, unsigned int *ticket,
                            int num_sms, cudaStream_t st);
cudaError_t launch_lm(const double *X, size_t l, int n, int n_cvt, const double *Wt, const double *y, const double *WtWi, const double *Wty,
                      double yPwy, int test_mode, gb200_sumstat *out, cudaStream_t st);
cudaError_t launch_lmm_common(int n_cvt, const LmmConst &D, const LmmParams &prm, double *H, double *ctab, const double *node_lams,
                              int n_nodes, double *cheb, cudaStream_t st);
cudaError_t launch_lmm_vnull(int n_cvt, const LmmConst &D, double lam, const double *U, double *scratch, double *v, cudaStream_t st);
cudaError_t launch_lmm_acols(int n_cvt, const LmmConst &D, const double *H, int J0, int x0, int nblocks, double *A, int ncol,
                             const int *patch_idx, int npatch, int patch0, cudaStream_t st);
int lmm_cheb_nodes();
int lmm_cheb_xnodes();
size_t lmm_cheb_doubles(int n_cvt, int n_region);
size_t lmm_common_record_doubles(int n_cvt);
cudaError_t launch_lmm_assoc_v2(int n_cvt, const LmmConst &D, const LmmParams &prm, const double *UtXt, size_t ldu,
                                int l, gb200_sumstat *out, unsigned int *ticket, int num_sms, cudaStream_t st);
cudaError_t launch_lmm_null(int n_cvt, const LmmConst &D, double l_min, double l_max, int n_region,
                            NullOut *out, cudaStream_t st);

// C(M x N, row-major ldc) = alpha * A * B + beta * C with arbitrary element strides:
// A(m,k) = A[m*sam + k*sak], B(k,n) = B[k*sbk + n*sbn].
cudaError_t launch_dgemm(size_t M, size_t N, size_t K, double alpha, const double *A, size_t sam, size_t sak,
                         const double *B, size_t sbk, size_t sbn, double beta, double *C, size_t ldc,
                         bool lower_only, cudaStream_t st);
cudaError_t launch_symmetrize_from_lower(double *C, size_t n, size_t ldc, cudaStream_t st);
cudaError_t launch_scale(double *C, size_t count, double alpha, cudaStream_t st);
cudaError_t launch_transpose(const double *in, size_t rows, size_t cols, size_t ldi, double *out, size_t ldo,
                             cudaStream_t st);

// genotype kernels (geno.cu)
cudaError_t launch_kin_transform(double *G, size_t l, size_t n, size_t ldg, int k_mode, cudaStream_t st);
cudaError_t launch_lmm_impute(double *G, size_t l, size_t n, size_t ldg, cudaStream_t st);
cudaError_t launch_bed_decode(const unsigned char *bed, size_t l, size_t bytes_per_snp, const int *idx,
                              size_t n_out, double *G, size_t ldg, cudaStream_t st);
cudaError_t launch_qc_bed(const unsigned char *bed, size_t l, size_t bytes_per_snp, const int *idx, int n_test,
                          const double *W, const double *WtWi, int n_cvt, gb200_snpqc *out, cudaStream_t st);
cudaError_t launch_center_matrix(double *G, size_t n, size_t ldg, double *row_sums, cudaStream_t st);

}  // namespace gb
