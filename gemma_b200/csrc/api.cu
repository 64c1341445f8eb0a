// api.cu -- the C ABI of libgemma_b200.so (include/gemma_b200.h).  Host-side orchestration
// only: every FLOP of the hot path runs in the kernels of lmm_kernel.cu / dgemm.cu /
// i8gemm_sm100.cu / geno.cu / eigh.cu.  There is NO CPU fallback: without a CUDA device
// gb200_create fails.
#include "common.cuh"
#include <string.h>
#include <math.h>

using namespace gb;

namespace gb {
// i8gemm_sm100.cu
int i8_prepare(gb200_ctx *ctx);                       // slice U into int8 planes (after lmm_setup)
int i8_project_bed(gb200_ctx *ctx, const unsigned char *bed_dev, const int *idx_dev, size_t ni_total,
                   size_t l, size_t bytes_per_snp, double *UtXt_dev);   // UtXt l x n (ld n)
bool i8_available(gb200_ctx *ctx);
int i8_xsum_prepare(gb200_ctx *ctx, int ncol);            // digit planes of the exact-sum vectors (c->i8.xs_V)
int i8_project_geno(gb200_ctx *ctx, const double *G_dev, size_t l, size_t ldg, double *UtXt_dev, bool *taken);   // dosage rows as exact digit rows
int i8_default_planes(size_t n);
int i8_effective_planes(gb200_ctx *c, int *T_out);
bool kin_i8_eligible(gb200_ctx *ctx);
int kin_i8_begin(gb200_ctx *ctx);
int kin_i8_add_chunk(gb200_ctx *ctx, const unsigned char *bed_dev, size_t l, size_t bytes_per_snp, bool *taken);
int kin_i8_finish(gb200_ctx *ctx, double inv_ns);
}

// grow-only buffer whose NEW allocations are zero-filled (row tails of padded buffers stay zero)
static cudaError_t reserve_zeroed(gb::DevBuf &b, size_t bytes, cudaStream_t st) {
  if (b.owned && b.p && bytes <= b.cap) return cudaSuccess;
  cudaError_t e = b.reserve(bytes);
  if (e != cudaSuccess) return e;
  return cudaMemsetAsync(b.p, 0, bytes, st);
}
static inline size_t round_up(size_t v, size_t m) { return (v + m - 1) / m * m; }

static bool trans_flag(const char *t, bool *ok) {
  *ok = t && (t[0] == 'N' || t[0] == 'n' || t[0] == 'T' || t[0] == 't');
  return t && (t[0] == 'T' || t[0] == 't');
}

// SNPs per internal sub-batch of the bed entry points: the FP64 U^T X staging buffer (chunk x n_c doubles) stays near 4 GB
// (8192 SNPs at n = 50 000: 32 x 1042 tile pairs = 450.6 waves of the 74 CTA pairs, 99.9% wave efficiency)
static size_t lmm_chunk_snps(const gb200_ctx *c) {
  if (c->batch_chunk > 0) return (size_t)c->batch_chunk;
  size_t ch = ((size_t)4 << 30) / (c->n_c * sizeof(double));
  ch = ch / 2048 * 2048;
  if (ch < 2048) ch = 2048;
  if (ch > 65536) ch = 65536;
  return ch;
}

extern "C" {
static int lmm_ensure_common_public(gb200_ctx *c);

int gb200_abi_version(void) { return GB200_ABI_VERSION; }

int gb200_create(gb200_ctx **out, int device, void *stream) {
  if (!out) return GB200_ERR_ARG;
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) return GB200_ERR_CUDA;   // fail loudly: no CPU fallback
  if (device < 0) { if (cudaGetDevice(&device) != cudaSuccess) return GB200_ERR_CUDA; }
  if (device >= count) return GB200_ERR_ARG;
  if (cudaSetDevice(device) != cudaSuccess) return GB200_ERR_CUDA;
  gb200_ctx *c = new gb200_ctx();
  c->device = device;
  if (stream) { c->stream = (cudaStream_t)stream; c->own_stream = false; }
  else {
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return GB200_ERR_CUDA; }
    c->own_stream = true;
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete c; return GB200_ERR_CUDA; }
  c->num_sms = prop.multiProcessorCount;
  if (c->dTicket.reserve(128) != cudaSuccess || cudaMemset(c->dTicket.p, 0, 128) != cudaSuccess) { delete c; return GB200_ERR_CUDA; }
  *out = c;
  return GB200_OK;
}

void gb200_destroy(gb200_ctx *c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  for (auto &kv : c->profs)
    for (auto &pr : kv.second.pending) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
  for (auto ev : c->event_pool) cudaEventDestroy(ev);
  gb::DevBuf *bufs[] = {&c->dK, &c->dU, &c->dEval, &c->dWt, &c->dY, &c->dNull, &c->dX, &c->dUtXt, &c->dOut,
                        &c->dBed, &c->dMask, &c->dIdx, &c->dTicket, &c->dTmp, &c->i8.slices, &c->i8.scale, &c->i8.wave_ctr,
                        &c->i8.geno, &c->i8.miss_mean, &c->i8.kin_zt, &c->i8.kin_stats, &c->i8.kin_a, &c->i8.kin_tiles,
                        &c->i8.kin_qbits, &c->i8.kin_y, &c->dWtx, &c->dEnv, &c->dX2, &c->dFlip, &c->dLmW, &c->dLmY,
                        &c->dLmSmall, &c->dMvY, &c->dMvNull, &c->dMvOut, &c->dHrows, &c->dCtab, &c->dCheb, &c->dNodeLam, &c->dVnull, &c->i8.xex, &c->i8.xs_V, &c->i8.xs_planes, &c->i8.xs_scale, &c->i8.xs_out, &c->i8.xs_patch_idx};
  for (auto b : bufs) b->release();
  if (c->i8.tmap_a) free(c->i8.tmap_a);
  if (c->i8.tmap_b) free(c->i8.tmap_b);
  if (c->i8.tmap_q) free(c->i8.tmap_q);
  c->i8.holeq.release();
  if (c->i8.tmap_ka) free(c->i8.tmap_ka);
  if (c->i8.tmap_kb) free(c->i8.tmap_kb);
  if (c->side) { cudaStreamSynchronize(c->side); cudaStreamDestroy(c->side); }
  if (c->copy) { cudaStreamSynchronize(c->copy); cudaStreamDestroy(c->copy); }
  for (int k = 0; k < 2; ++k) { if (c->evCopy[k]) cudaEventDestroy(c->evCopy[k]); if (c->evUsed[k]) cudaEventDestroy(c->evUsed[k]); }
  if (c->evStart) cudaEventDestroy(c->evStart);
  c->dBed2.release();
  for (int k = 0; k < 2; ++k) { if (c->evG[k]) cudaEventDestroy(c->evG[k]); if (c->evL[k]) cudaEventDestroy(c->evL[k]); }
  c->dUtXt2.release();
  if (c->own_stream) cudaStreamDestroy(c->stream);
  delete c;
}

const char *gb200_last_error(const gb200_ctx *c) { return c ? c->err.c_str() : "null context"; }
void *gb200_stream(const gb200_ctx *c) { return c ? (void *)c->stream : nullptr; }

int gb200_synchronize(gb200_ctx *c) {
  if (!c) return GB200_ERR_ARG;
  GB_CUDA(c, cudaStreamSynchronize(c->stream));
  return GB200_OK;
}

int gb200_profile_enable(gb200_ctx *c, int on) { if (!c) return GB200_ERR_ARG; c->prof = on != 0; return GB200_OK; }

static void prof_drain(gb200_ctx *c, ProfEntry &e) {
  for (auto &pr : e.pending) {
    float ms = 0.f;
    cudaEventSynchronize(pr.second);
    if (cudaEventElapsedTime(&ms, pr.first, pr.second) == cudaSuccess) e.ms += ms;
    c->event_pool.push_back(pr.first);
    c->event_pool.push_back(pr.second);
  }
  e.pending.clear();
}

int gb200_profile_reset(gb200_ctx *c) {
  if (!c) return GB200_ERR_ARG;
  for (auto &kv : c->profs) { prof_drain(c, kv.second); kv.second.ms = 0.0; kv.second.launches = 0; }
  return GB200_OK;
}

int gb200_profile_get(gb200_ctx *c, const char *name, double *ms, long *launches) {
  if (!c || !name) return GB200_ERR_ARG;
  if (!strcmp(name, "__launches")) { if (ms) *ms = 0.0; if (launches) *launches = c->kernel_launches; return GB200_OK; }
  auto it = c->profs.find(name);
  if (it == c->profs.end()) { if (ms) *ms = 0.0; if (launches) *launches = 0; return GB200_OK; }
  prof_drain(c, it->second);
  if (ms) *ms = it->second.ms;
  if (launches) *launches = it->second.launches;
  return GB200_OK;
}

int gb200_lmm_counters(gb200_ctx *c, unsigned long long counts[6], int reset) {
  if (!c) return GB200_ERR_ARG;
  GB_CUDA(c, cudaStreamSynchronize(c->stream));
  if (counts) GB_CUDA(c, cudaMemcpy(counts, c->dTicket.as<unsigned long long>() + 4, 6 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  if (reset) GB_CUDA(c, cudaMemset(c->dTicket.as<unsigned long long>() + 4, 0, 8 * sizeof(unsigned long long)));
  c->count_work = true;                  // counting starts with the first query
  return GB200_OK;
}

int gb200_set_option(gb200_ctx *c, const char *name, long value) {
  if (!c || !name) return GB200_ERR_ARG;
  if (!strcmp(name, "utx_path")) {
    if (value < 0 || value > 2) return set_err(c, GB200_ERR_ARG, "utx_path must be 0,1,2");
    c->utx_path = value; return GB200_OK;
  }
  if (!strcmp(name, "overlap")) {
    if (value < 0 || value > 1) return set_err(c, GB200_ERR_ARG, "overlap must be 0 or 1");
    c->overlap = value; return GB200_OK;
  }
  if (!strcmp(name, "cta_pair")) {
    if (value < 0 || value > 1) return set_err(c, GB200_ERR_ARG, "cta_pair must be 0 or 1");
    c->cta_pair = value; return GB200_OK;
  }
  if (!strcmp(name, "kin_cta_pair")) {
    if (value < 0 || value > 1) return set_err(c, GB200_ERR_ARG, "kin_cta_pair must be 0 or 1");
    c->kin_cta_pair = value; return GB200_OK;
  }
  if (!strcmp(name, "kin_miss_max_permille")) {
    if (value < 0 || value > 1000) return set_err(c, GB200_ERR_ARG, "kin_miss_max_permille must be in 0..1000");
    c->kin_miss_max = (double)value / 1000.0; return GB200_OK;
  }
  if (!strcmp(name, "kin_path")) {
    if (value < 0 || value > 1) return set_err(c, GB200_ERR_ARG, "kin_path must be 0 or 1");
    c->kin_path = value; return GB200_OK;
  }
  if (!strcmp(name, "lmm_kernel")) {
    if (value < 0 || value > 3) return set_err(c, GB200_ERR_ARG, "lmm_kernel must be 0,1,2,3");
    c->lmm_kernel = value; return GB200_OK;
  }
  if (!strcmp(name, "gemm_wave_sync")) {
    if (value < 0 || value > 1) return set_err(c, GB200_ERR_ARG, "gemm_wave_sync must be 0 or 1");
    c->gemm_wave_sync = value; return GB200_OK;
  }
  if (!strcmp(name, "x_exact")) {
    if (value < 0 || value > 2) return set_err(c, GB200_ERR_ARG, "x_exact must be 0, 1 or 2");
    if (value != c->x_exact) c->common_ready = false;
    c->x_exact = value; return GB200_OK;
  }
  if (!strcmp(name, "hole_gemm")) {
    if (value < 0 || value > 1) return set_err(c, GB200_ERR_ARG, "hole_gemm must be 0 or 1");
    c->hole_gemm = value; return GB200_OK;
  }
  if (!strcmp(name, "gemm_l2hint")) {
    if (value < 0 || value > 1) return set_err(c, GB200_ERR_ARG, "gemm_l2hint must be 0 or 1");
    c->gemm_l2hint = value; return GB200_OK;
  }
  if (!strcmp(name, "gemm_stages")) {
    if (value < 0 || value > 8) return set_err(c, GB200_ERR_ARG, "gemm_stages must be 0..8");
    c->gemm_stages = value; return GB200_OK;
  }
  if (!strcmp(name, "lmm_interp")) {
    if (value < 0 || value > 1) return set_err(c, GB200_ERR_ARG, "lmm_interp must be 0 or 1");
    if (value != c->lmm_interp) c->common_ready = false;
    c->lmm_interp = value; return GB200_OK;
  }
  if (!strcmp(name, "eigh_path")) {
    if (value < 0 || value > 2) return set_err(c, GB200_ERR_ARG, "eigh_path must be 0, 1 or 2");
    c->eigh_path = value; return GB200_OK;
  }
  if (!strcmp(name, "lmm_hoist")) {
    if (value < 0 || value > 1) return set_err(c, GB200_ERR_ARG, "lmm_hoist must be 0 or 1");
    c->lmm_hoist = value; return GB200_OK;
  }
  if (!strcmp(name, "gemm_groups")) {
    if (value < 1 || value > 2) return set_err(c, GB200_ERR_ARG, "gemm_groups must be 1 or 2");
    c->gemm_groups = value; return GB200_OK;
  }
  if (!strcmp(name, "gemm_panel")) {
    if (value < 0 || value > 64) return set_err(c, GB200_ERR_ARG, "gemm_panel must be 0..64");
    c->gemm_panel = value; return GB200_OK;
  }
  if (!strcmp(name, "stage_mask")) {
    if (value < 1 || value > 3) return set_err(c, GB200_ERR_ARG, "stage_mask must be 1 (projection only), 2 (tests only, on the last projection) or 3");
    c->stage_mask = value; return GB200_OK;
  }
  if (!strcmp(name, "batch_chunk")) {
    if (value < 0 || value > (1 << 20) || (value % 256) != 0) return set_err(c, GB200_ERR_ARG, "batch_chunk must be 0 (auto) or a multiple of 256");
    c->batch_chunk = value; return GB200_OK;
  }
  if (!strcmp(name, "n_slices")) {
    if (value < 0 || value > 8) return set_err(c, GB200_ERR_ARG, "n_slices must be 0..8");
    if (value != c->n_slices) c->i8.ready = false;
    c->n_slices = value; return GB200_OK;
  }
  return set_err(c, GB200_ERR_ARG, std::string("unknown option ") + name);
}

int gb200_get_option(gb200_ctx *c, const char *name, long *value) {
  if (!c) return GB200_ERR_ARG;
  if (!name || !value) return set_err(c, GB200_ERR_ARG, "gb200_get_option: null argument");
  if (!strcmp(name, "n_slices")) {      // effective: the forced count, else the count chosen from this U's column maxima (i8gemm_sm100.cu)
    if (c->lmm_ready && c->prm_ready) { const int rc0 = lmm_ensure_common_public(c); if (rc0) return rc0; }   // whether the exact linear sums are on decides it
    int T = 0;
    const int rc = i8_effective_planes(c, &T);
    if (rc) return rc;
    *value = T; return GB200_OK;
  }
  if (!strcmp(name, "utx_path")) { *value = c->utx_path; return GB200_OK; }
  if (!strcmp(name, "overlap")) { *value = c->overlap; return GB200_OK; }
  if (!strcmp(name, "cta_pair")) { *value = c->cta_pair; return GB200_OK; }
  if (!strcmp(name, "kin_cta_pair")) { *value = c->kin_cta_pair; return GB200_OK; }
  if (!strcmp(name, "kin_path")) { *value = c->kin_path; return GB200_OK; }
  if (!strcmp(name, "kin_miss_max_permille")) { *value = (long)(c->kin_miss_max * 1000.0 + 0.5); return GB200_OK; }
  if (!strcmp(name, "lmm_kernel")) { *value = c->lmm_kernel; return GB200_OK; }
  if (!strcmp(name, "lmm_hoist")) { *value = c->lmm_hoist; return GB200_OK; }
  if (!strcmp(name, "lmm_interp")) { *value = c->lmm_interp; return GB200_OK; }
  if (!strcmp(name, "batch_chunk")) { *value = c->n_c ? (long)lmm_chunk_snps(c) : c->batch_chunk; return GB200_OK; }
  if (!strcmp(name, "stage_mask")) { *value = c->stage_mask; return GB200_OK; }
  if (!strcmp(name, "gemm_groups")) { *value = c->gemm_groups; return GB200_OK; }
  if (!strcmp(name, "gemm_panel")) { *value = c->gemm_panel; return GB200_OK; }
  if (!strcmp(name, "eigh_workspace_bytes")) { *value = (long)c->eigh_workspace_bytes; return GB200_OK; }
  return set_err(c, GB200_ERR_ARG, std::string("unknown option ") + name);
}

// ---------------------------------------------------------------------------------------
int gb200_dgemm(gb200_ctx *c, const char *TransA, const char *TransB, double alpha, const double *A,
                size_t a_rows, size_t a_cols, size_t lda, const double *B, size_t b_rows, size_t b_cols,
                size_t ldb, double beta, double *C, size_t c_rows, size_t c_cols, size_t ldc) {
  if (!c) return GB200_ERR_ARG;
  bool oka, okb;
  const bool ta = trans_flag(TransA, &oka), tb = trans_flag(TransB, &okb);
  if (!oka || !okb || !A || !B || !C) return set_err(c, GB200_ERR_ARG, "gb200_dgemm: bad argument");
  const size_t M = ta ? a_cols : a_rows, K = ta ? a_rows : a_cols;
  const size_t Kb = tb ? b_cols : b_rows, N = tb ? b_rows : b_cols;
  // fastblas.cpp:193-195 enforce(M>0, N>0, K>0); :207 "Range error in dgemm"
  if (M == 0 || N == 0 || K == 0) return set_err(c, GB200_ERR_ARG, "gb200_dgemm: empty dimension");
  if (K != Kb || c_rows != M || c_cols != N || lda < a_cols || ldb < b_cols || ldc < c_cols)
    return set_err(c, GB200_ERR_ARG, "Range error in dgemm");
  cudaStream_t st = c->stream;
  DevBuf dA, dB, dC;
  auto fail = [&](cudaError_t e, const char *what) {
    dA.release(); dB.release(); dC.release();
    return set_err(c, GB200_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
  };
  cudaError_t e;
  if ((e = dA.reserve(a_rows * a_cols * sizeof(double))) != cudaSuccess) return fail(e, "alloc A");
  if ((e = dB.reserve(b_rows * b_cols * sizeof(double))) != cudaSuccess) return fail(e, "alloc B");
  if ((e = dC.reserve(M * N * sizeof(double))) != cudaSuccess) return fail(e, "alloc C");
  if ((e = cudaMemcpy2DAsync(dA.p, a_cols * 8, A, lda * 8, a_cols * 8, a_rows, cudaMemcpyHostToDevice, st)) != cudaSuccess) return fail(e, "copy A");
  if ((e = cudaMemcpy2DAsync(dB.p, b_cols * 8, B, ldb * 8, b_cols * 8, b_rows, cudaMemcpyHostToDevice, st)) != cudaSuccess) return fail(e, "copy B");
  if (beta != 0.0)
    if ((e = cudaMemcpy2DAsync(dC.p, N * 8, C, ldc * 8, N * 8, M, cudaMemcpyHostToDevice, st)) != cudaSuccess) return fail(e, "copy C");
  {
    ProfScope ps(c, "dgemm");
    const size_t sam = ta ? 1 : a_cols, sak = ta ? a_cols : 1;
    const size_t sbk = tb ? 1 : b_cols, sbn = tb ? b_cols : 1;
    if ((e = launch_dgemm(M, N, K, alpha, dA.as<double>(), sam, sak, dB.as<double>(), sbk, sbn, beta,
                          dC.as<double>(), N, false, st)) != cudaSuccess) return fail(e, "dgemm kernel");
  }
  if ((e = cudaMemcpy2DAsync(C, ldc * 8, dC.p, N * 8, N * 8, M, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return fail(e, "copy out");
  if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return fail(e, "sync");
  dA.release(); dB.release(); dC.release();
  return GB200_OK;
}

// ---------------------------------------------------------------------------------------
// -gk
int gb200_kin_begin(gb200_ctx *c, size_t n, int k_mode) {
  if (!c) return GB200_ERR_ARG;
  if (n == 0 || (k_mode != 1 && k_mode != 2)) return set_err(c, GB200_ERR_ARG, "gb200_kin_begin: bad argument");
  GB_CUDA(c, c->dK.reserve(n * n * sizeof(double)));
  GB_CUDA(c, cudaMemsetAsync(c->dK.p, 0, n * n * sizeof(double), c->stream));   // gsl_matrix_set_zero, param.cpp:1301
  c->kin_n = n; c->kin_mode = k_mode; c->kin_ns = 0; c->kin_active = true;
  return kin_i8_begin(c);
}

// K(lower) += Xs^T Xs for an SNP-major centred batch Xs (l x n, ld n) already on the device
static int kin_accumulate_dev(gb200_ctx *c, const double *Xs, size_t l) {
  const size_t n = c->kin_n;
  ProfScope ps(c, "kin");
  GB_CUDA(c, launch_dgemm(n, n, l, 1.0, Xs, 1, n, Xs, n, 1, 1.0, c->dK.as<double>(), n, true, c->stream));
  c->kin_ns += l;
  return GB200_OK;
}

int gb200_kin_add(gb200_ctx *c, const double *Xb, size_t n, size_t l, size_t ldx) {
  if (!c) return GB200_ERR_ARG;
  if (!c->kin_active) return set_err(c, GB200_ERR_STATE, "gb200_kin_add before gb200_kin_begin");
  if (!Xb || n != c->kin_n || ldx < l) return set_err(c, GB200_ERR_ARG, "gb200_kin_add: bad argument");
  if (l == 0) return GB200_OK;
  GB_CUDA(c, c->dX.reserve(n * l * sizeof(double)));
  GB_CUDA(c, cudaMemcpy2DAsync(c->dX.p, l * 8, Xb, ldx * 8, l * 8, n, cudaMemcpyHostToDevice, c->stream));
  {
    ProfScope ps(c, "kin");
    // A(i,s) = X[i*l+s], B(s,j) = X[j*l+s]
    GB_CUDA(c, launch_dgemm(n, n, l, 1.0, c->dX.as<double>(), l, 1, c->dX.as<double>(), 1, l, 1.0,
                            c->dK.as<double>(), n, true, c->stream));
  }
  c->kin_ns += l;
  GB_CUDA(c, cudaStreamSynchronize(c->stream));
  return GB200_OK;
}

int gb200_kin_add_geno(gb200_ctx *c, const double *G, size_t l, size_t n, size_t ldg) {
  if (!c) return GB200_ERR_ARG;
  if (!c->kin_active) return set_err(c, GB200_ERR_STATE, "gb200_kin_add_geno before gb200_kin_begin");
  if (!G || n != c->kin_n || ldg < n) return set_err(c, GB200_ERR_ARG, "gb200_kin_add_geno: bad argument");
  if (l == 0) return GB200_OK;
  GB_CUDA(c, c->dX.reserve(n * l * sizeof(double)));
  GB_CUDA(c, cudaMemcpy2DAsync(c->dX.p, n * 8, G, ldg * 8, n * 8, l, cudaMemcpyHostToDevice, c->stream));
  {
    ProfScope ps(c, "decode");
    GB_CUDA(c, launch_kin_transform(c->dX.as<double>(), l, n, n, c->kin_mode, c->stream));
  }
  int rc = kin_accumulate_dev(c, c->dX.as<double>(), l);
  if (rc) return rc;
  GB_CUDA(c, cudaStreamSynchronize(c->stream));
  return GB200_OK;
}

int gb200_kin_add_bed_dev(gb200_ctx *c, const unsigned char *bed_dev, size_t l, size_t bytes_per_snp) {
  if (!c) return GB200_ERR_ARG;
  if (!c->kin_active) return set_err(c, GB200_ERR_STATE, "gb200_kin_add_bed before gb200_kin_begin");
  const size_t n = c->kin_n;
  if (!bed_dev || bytes_per_snp != (n + 3) / 4) return set_err(c, GB200_ERR_ARG, "gb200_kin_add_bed: bytes_per_snp != ceil(n/4)");
  // bounded staging: decode + accumulate in chunks of at most ~1 GiB of FP64 genotypes
  size_t chunk = (size_t(1) << 30) / (n * sizeof(double));
  if (chunk < 256) chunk = 256;
  const bool try_i8 = kin_i8_eligible(c);
  for (size_t s0 = 0; s0 < l; s0 += chunk) {
    const size_t lc = (l - s0 < chunk) ? (l - s0) : chunk;
    if (try_i8) {            // exact Z Z^T on the int8 tensor pipe when the chunk has no missing genotype
      bool taken = false;
      int rc8 = kin_i8_add_chunk(c, bed_dev + s0 * bytes_per_snp, lc, bytes_per_snp, &taken);
      if (rc8) return rc8;
      if (taken) continue;
    }
    GB_CUDA(c, c->dX.reserve(n * lc * sizeof(double)));
    {
      ProfScope ps(c, "decode", 2);
      GB_CUDA(c, launch_bed_decode(bed_dev + s0 * bytes_per_snp, lc, bytes_per_snp, nullptr, n,
                                   c->dX.as<double>(), n, c->stream));
      GB_CUDA(c, launch_kin_transform(c->dX.as<double>(), lc, n, n, c->kin_mode, c->stream));
    }
    int rc = kin_accumulate_dev(c, c->dX.as<double>(), lc);
    if (rc) return rc;
  }
  return GB200_OK;
}

int gb200_kin_add_bed(gb200_ctx *c, const unsigned char *bed, size_t l, size_t bytes_per_snp) {
  if (!c) return GB200_ERR_ARG;
  if (!c->kin_active) return set_err(c, GB200_ERR_STATE, "gb200_kin_add_bed before gb200_kin_begin");
  if (!bed) return set_err(c, GB200_ERR_ARG, "gb200_kin_add_bed: null input");
  if (l == 0) return GB200_OK;
  GB_CUDA(c, c->dBed.reserve(l * bytes_per_snp));
  GB_CUDA(c, cudaMemcpyAsync(c->dBed.p, bed, l * bytes_per_snp, cudaMemcpyHostToDevice, c->stream));
  int rc = gb200_kin_add_bed_dev(c, c->dBed.as<unsigned char>(), l, bytes_per_snp);
  if (rc) return rc;
  GB_CUDA(c, cudaStreamSynchronize(c->stream));
  return GB200_OK;
}

int gb200_kin_finish_dev(gb200_ctx *c, double **K_dev, size_t *ns_used) {
  if (!c) return GB200_ERR_ARG;
  if (!c->kin_active) return set_err(c, GB200_ERR_STATE, "gb200_kin_finish before gb200_kin_begin");
  const size_t n = c->kin_n;
  if (c->i8.kin_used) {  // rank-one centring terms of the int8 batches + the 1/ns_test scaling
    int rc8 = kin_i8_finish(c, c->kin_ns > 0 ? 1.0 / (double)c->kin_ns : 1.0);
    if (rc8) return rc8;
  } else if (c->kin_ns > 0)   // gsl_matrix_scale(matrix_kin, 1.0/ns_test), gemma_io.cpp:1570
    GB_CUDA(c, launch_scale(c->dK.as<double>(), n * n, 1.0 / (double)c->kin_ns, c->stream));
  GB_CUDA(c, launch_symmetrize_from_lower(c->dK.as<double>(), n, n, c->stream));
  if (K_dev) *K_dev = c->dK.as<double>();
  if (ns_used) *ns_used = c->kin_ns;
  c->kin_active = false;
  return GB200_OK;
}

int gb200_kin_finish(gb200_ctx *c, double *K, size_t ldk, size_t *ns_used) {
  if (!c) return GB200_ERR_ARG;
  if (!K || ldk < c->kin_n) return set_err(c, GB200_ERR_ARG, "gb200_kin_finish: bad argument");
  const size_t n = c->kin_n;
  int rc = gb200_kin_finish_dev(c, nullptr, ns_used);
  if (rc) return rc;
  GB_CUDA(c, cudaMemcpy2DAsync(K, ldk * 8, c->dK.p, n * 8, n * 8, n, cudaMemcpyDeviceToHost, c->stream));
  GB_CUDA(c, cudaStreamSynchronize(c->stream));
  return GB200_OK;
}

// ---------------------------------------------------------------------------------------
// -lmm
static int lmm_upload_common(gb200_ctx *c, size_t n, size_t n_cvt, const double *U, size_t ldu,
                             const double *eval) {
  if (n == 0 || n_cvt == 0 || !U || !eval || ldu < n) return set_err(c, GB200_ERR_ARG, "gb200_lmm_setup: bad argument");
  if (n_cvt > GB200_MAX_CVT)
    return set_err(c, GB200_ERR_UNSUPPORTED, "gb200_lmm_setup: n_cvt exceeds GB200_MAX_CVT");
  if (n_cvt + 1 >= n) return set_err(c, GB200_ERR_ARG, "gb200_lmm_setup: need n > n_cvt + 1");
  c->lmm_ready = false; c->i8.ready = false; c->i8.auto_T = 0; c->i8.xs_ready = false; c->i8.no_xsum_consumer = false; c->common_ready = false; c->gxe_ready = false;
  c->mask_host.clear();                           // the cached gather index belongs to the previous n
  const size_t n_c = round_up(n, 512);            // vectors and U^T x rows are zero-padded to the pipeline chunk
  c->dUtXt.release(); c->dUtXt2.release();        // row pitch changes with n: force fresh zeroed buffers
  GB_CUDA(c, c->dU.reserve(n * n * sizeof(double)));
  GB_CUDA(c, c->dEval.reserve(n_c * sizeof(double)));
  GB_CUDA(c, c->dWt.reserve(n_cvt * n_c * sizeof(double)));
  GB_CUDA(c, c->dY.reserve(n_c * sizeof(double)));
  GB_CUDA(c, c->dNull.reserve(sizeof(NullOut)));
  GB_CUDA(c, cudaMemsetAsync(c->dEval.p, 0, n_c * 8, c->stream));
  GB_CUDA(c, cudaMemsetAsync(c->dWt.p, 0, n_cvt * n_c * 8, c->stream));
  GB_CUDA(c, cudaMemsetAsync(c->dY.p, 0, n_c * 8, c->stream));
  GB_CUDA(c, cudaMemcpy2DAsync(c->dU.p, n * 8, U, ldu * 8, n * 8, n, cudaMemcpyHostToDevice, c->stream));
  GB_CUDA(c, cudaMemcpyAsync(c->dEval.p, eval, n * 8, cudaMemcpyHostToDevice, c->stream));
  c->n = n; c->n_cvt = n_cvt; c->n_c = n_c;
  return GB200_OK;
}

int gb200_lmm_setup(gb200_ctx *c, size_t n, size_t n_cvt, const double *U, size_t ldu, const double *eval,
                    const double *W, size_t ldw, const double *y, double *UtW_out, double *Uty_out) {
  if (!c) return GB200_ERR_ARG;
  if (!W || !y || ldw < n_cvt) return set_err(c, GB200_ERR_ARG, "gb200_lmm_setup: bad argument");
  int rc = lmm_upload_common(c, n, n_cvt, U, ldu, eval);
  if (rc) return rc;
  // CalcUtX (mathfunc.cpp:497-510): UtW = U^T W, Uty = U^T y.  Stored transposed (n_cvt x n).
  GB_CUDA(c, c->dTmp.reserve((n_cvt + 1) * n * sizeof(double)));
  double *dW = c->dTmp.as<double>(), *dy = dW + n_cvt * n;
  GB_CUDA(c, cudaMemcpy2DAsync(dW, n_cvt * 8, W, ldw * 8, n_cvt * 8, n, cudaMemcpyHostToDevice, c->stream));
  GB_CUDA(c, cudaMemcpyAsync(dy, y, n * 8, cudaMemcpyHostToDevice, c->stream));
  // Wt[a][i] = sum_j W[j][a] U[j][i]   (M = n_cvt, N = n, K = n)
  const size_t n_c = c->n_c;
  GB_CUDA(c, launch_dgemm(n_cvt, n, n, 1.0, dW, 1, n_cvt, c->dU.as<double>(), n, 1, 0.0, c->dWt.as<double>(), n_c,
                          false, c->stream));
  GB_CUDA(c, launch_dgemm(1, n, n, 1.0, dy, 1, 1, c->dU.as<double>(), n, 1, 0.0, c->dY.as<double>(), n_c, false,
                          c->stream));
  if (UtW_out) {
    // return as n x n_cvt row-major
    std::vector<double> t(n_cvt * n_c);
    GB_CUDA(c, cudaMemcpyAsync(t.data(), c->dWt.p, n_cvt * n_c * 8, cudaMemcpyDeviceToHost, c->stream));
    GB_CUDA(c, cudaStreamSynchronize(c->stream));
    for (size_t a = 0; a < n_cvt; ++a)
      for (size_t i = 0; i < n; ++i) UtW_out[i * n_cvt + a] = t[a * n_c + i];
  }
  if (Uty_out) GB_CUDA(c, cudaMemcpyAsync(Uty_out, c->dY.p, n * 8, cudaMemcpyDeviceToHost, c->stream));
  GB_CUDA(c, cudaStreamSynchronize(c->stream));
  c->lmm_ready = true;
  return GB200_OK;
}

int gb200_lmm_setup_rotated(gb200_ctx *c, size_t n, size_t n_cvt, const double *U, size_t ldu,
                            const double *eval, const double *UtW, size_t ldw, const double *Uty) {
  if (!c) return GB200_ERR_ARG;
  if (!UtW || !Uty || ldw < n_cvt) return set_err(c, GB200_ERR_ARG, "gb200_lmm_setup_rotated: bad argument");
  int rc = lmm_upload_common(c, n, n_cvt, U, ldu, eval);
  if (rc) return rc;
  const size_t n_c = c->n_c;
  std::vector<double> t(n_cvt * n_c, 0.0);
  for (size_t a = 0; a < n_cvt; ++a)
    for (size_t i = 0; i < n; ++i) t[a * n_c + i] = UtW[i * ldw + a];
  GB_CUDA(c, cudaMemcpyAsync(c->dWt.p, t.data(), n_cvt * n_c * 8, cudaMemcpyHostToDevice, c->stream));
  GB_CUDA(c, cudaMemcpyAsync(c->dY.p, Uty, n * 8, cudaMemcpyHostToDevice, c->stream));
  GB_CUDA(c, cudaStreamSynchronize(c->stream));
  c->lmm_ready = true;
  return GB200_OK;
}

int gb200_lmm_setup_rotated_dev(gb200_ctx *c, size_t n, size_t n_cvt, const double *U_dev, const double *eval_dev,
                                const double *UtWt_dev, const double *Uty_dev) {
  if (!c) return GB200_ERR_ARG;
  if (n == 0 || n_cvt == 0 || !U_dev || !eval_dev || !UtWt_dev || !Uty_dev)
    return set_err(c, GB200_ERR_ARG, "gb200_lmm_setup_rotated_dev: bad argument");
  if (n_cvt > GB200_MAX_CVT) return set_err(c, GB200_ERR_UNSUPPORTED, "n_cvt exceeds GB200_MAX_CVT");
  if (n_cvt + 1 >= n) return set_err(c, GB200_ERR_ARG, "need n > n_cvt + 1");
  c->lmm_ready = false; c->i8.ready = false; c->i8.auto_T = 0; c->i8.xs_ready = false; c->i8.no_xsum_consumer = false; c->common_ready = false; c->gxe_ready = false;
  c->mask_host.clear();
  const size_t n_c = round_up(n, 512);
  c->dUtXt.release(); c->dUtXt2.release();
  c->dU.adopt(const_cast<double *>(U_dev), n * n * 8);          // borrowed: caller keeps it alive
  GB_CUDA(c, c->dEval.reserve(n_c * 8));
  GB_CUDA(c, c->dWt.reserve(n_cvt * n_c * 8));
  GB_CUDA(c, c->dY.reserve(n_c * 8));
  GB_CUDA(c, c->dNull.reserve(sizeof(NullOut)));
  GB_CUDA(c, cudaMemsetAsync(c->dEval.p, 0, n_c * 8, c->stream));
  GB_CUDA(c, cudaMemsetAsync(c->dWt.p, 0, n_cvt * n_c * 8, c->stream));
  GB_CUDA(c, cudaMemsetAsync(c->dY.p, 0, n_c * 8, c->stream));
  GB_CUDA(c, cudaMemcpyAsync(c->dEval.p, eval_dev, n * 8, cudaMemcpyDeviceToDevice, c->stream));
  GB_CUDA(c, cudaMemcpy2DAsync(c->dWt.p, n_c * 8, UtWt_dev, n * 8, n * 8, n_cvt, cudaMemcpyDeviceToDevice, c->stream));
  GB_CUDA(c, cudaMemcpyAsync(c->dY.p, Uty_dev, n * 8, cudaMemcpyDeviceToDevice, c->stream));
  GB_CUDA(c, cudaStreamSynchronize(c->stream));
  c->n = n; c->n_cvt = n_cvt; c->n_c = n_c; c->lmm_ready = true;
  return GB200_OK;
}

static LmmConst make_const(gb200_ctx *c) {
  LmmConst D;
  D.n = (int)c->n; D.n_c = (int)c->n_c; D.ldv = (int)c->n_c;
  D.delta = c->dEval.as<double>(); D.Wt = c->dWt.as<double>(); D.y = c->dY.as<double>();
  D.Hrows = nullptr; D.ctab = nullptr; D.n_common = 0; D.cheb = nullptr; D.cheb_marg = 0.0; D.xex = nullptr; D.xsum = nullptr; D.xsum_ld = 0; D.xsum_nblocks = 0; D.xsum_nblocks_skip = 0; D.xcov = nullptr; D.xcov_idx = 0;
  D.cnt = c->count_work ? c->dTicket.as<unsigned long long>() + 4 : nullptr;     // 8 words behind the ticket words
  D.nc_gen = (c->lmm_kernel == 3) ? 1 : 0;      // > 0 forces the any-covariate-count kernel (the launchers fill the real values)
  D.gen_stride = 0;
  return D;
}

// c x c Gaussian elimination with partial pivoting: beta = A^-1 b, diag of A^-1
// (LUDecomp/LUSolve/LUInvert at src/lmm.cpp:2246-2250)
static void solve_small(size_t cN, std::vector<double> A, const std::vector<double> &b, double *x, double *diag_inv) {
  std::vector<double> M(cN * (cN + 1 + cN));
  const size_t w = 2 * cN + 1;
  for (size_t i = 0; i < cN; ++i) {
    for (size_t j = 0; j < cN; ++j) M[i * w + j] = A[i * cN + j];
    M[i * w + cN] = b[i];
    for (size_t j = 0; j < cN; ++j) M[i * w + cN + 1 + j] = (i == j) ? 1.0 : 0.0;
  }
  for (size_t k = 0; k < cN; ++k) {
    size_t pr = k;
    for (size_t i = k + 1; i < cN; ++i) if (fabs(M[i * w + k]) > fabs(M[pr * w + k])) pr = i;
    if (pr != k) for (size_t j = 0; j < w; ++j) std::swap(M[k * w + j], M[pr * w + j]);
    const double piv = M[k * w + k];
    for (size_t j = 0; j < w; ++j) M[k * w + j] /= piv;
    for (size_t i = 0; i < cN; ++i) {
      if (i == k) continue;
      const double f = M[i * w + k];
      if (f != 0.0) for (size_t j = 0; j < w; ++j) M[i * w + j] -= f * M[k * w + j];
    }
  }
  for (size_t i = 0; i < cN; ++i) { x[i] = M[i * w + cN]; diag_inv[i] = M[i * w + cN + 1 + i]; }
}

static double safe_sqrt_host(double d) {   // mathfunc.cpp:122-131
  double d1 = d;
  if (d < 0.001) d1 = fabs(d);
  if (d1 < 0.0) return nan("");
  return sqrt(d1);
}

int gb200_lmm_null(gb200_ctx *c, double l_min, double l_max, size_t n_region, double trace_G,
                   gb200_nullmodel *out, double *beta_mle, double *se_beta_mle, double *beta_remle,
                   double *se_beta_remle) {
  if (!c) return GB200_ERR_ARG;
  if (!c->lmm_ready) return set_err(c, GB200_ERR_STATE, "gb200_lmm_null before gb200_lmm_setup");
  if (!out || !(l_max > l_min) || n_region == 0) return set_err(c, GB200_ERR_ARG, "gb200_lmm_null: bad argument");
  LmmConst D = make_const(c);
  {
    ProfScope ps(c, "lmm");
    GB_CUDA(c, launch_lmm_null((int)c->n_cvt, D, l_min, l_max, (int)n_region, c->dNull.as<NullOut>(), c->stream));
  }
  NullOut r;
  GB_CUDA(c, cudaMemcpyAsync(&r, c->dNull.p, sizeof(NullOut), cudaMemcpyDeviceToHost, c->stream));
  GB_CUDA(c, cudaStreamSynchronize(c->stream));
  const size_t n = c->n, cN = c->n_cvt;
  out->l_mle_null = r.l_mle; out->logl_mle_H0 = r.logl_mle;
  out->l_remle_null = r.l_remle; out->logl_remle_H0 = r.logl_remle;
  // CalcPve, src/lmm.cpp:2195-2199
  const double se = safe_sqrt_host(-1.0 / r.dev2_remle);
  out->pve_null = trace_G * r.l_remle / (trace_G * r.l_remle + 1.0);
  out->pve_se_null = trace_G / ((trace_G * r.l_remle + 1.0) * (trace_G * r.l_remle + 1.0)) * se;
  // CalcLmmVgVeBeta, src/lmm.cpp:2242-2271.  The kernel ran with NC = c-1 and x = last covariate,
  // so its variable order is w_1..w_{c-1}, x(=w_c), y: NV = c+1 variables.
  const int NV = (int)cN + 1;
  for (int which = 0; which < 2; ++which) {
    const double *S1 = which == 0 ? r.S1_mle : r.S1_remle;
    const double lam = which == 0 ? r.l_mle : r.l_remle;
    const double Pyy = which == 0 ? r.Pyy_mle : r.Pyy_remle;
    std::vector<double> A(cN * cN), b(cN);
    for (size_t a = 0; a < cN; ++a) {
      for (size_t bb = 0; bb < cN; ++bb) {
        const int lo = (int)(a < bb ? a : bb), hi = (int)(a < bb ? bb : a);
        A[a * cN + bb] = S1[abidx(lo, hi, NV)];
      }
      b[a] = S1[abidx((int)a, NV - 1, NV)];
    }
    std::vector<double> beta(cN), dinv(cN);
    solve_small(cN, A, b, beta.data(), dinv.data());
    const double ve = Pyy / (double)(n - cN), vg = ve * lam;
    if (which == 0) { out->ve_mle = ve; out->vg_mle = vg; } else { out->ve_remle = ve; out->vg_remle = vg; }
    double *bo = which == 0 ? beta_mle : beta_remle, *so = which == 0 ? se_beta_mle : se_beta_remle;
    for (size_t i = 0; i < cN; ++i) {
      if (bo) bo[i] = beta[i];
      if (so) so[i] = safe_sqrt_host(dinv[i] * ve);
    }
  }
  return GB200_OK;
}

int gb200_lmm_params(gb200_ctx *c, int a_mode, double l_min, double l_max, size_t n_region,
                     double l_mle_null, double logl_mle_H0) {
  if (!c) return GB200_ERR_ARG;
  if (!(a_mode == 1 || a_mode == 2 || a_mode == 3 || a_mode == 4 || a_mode == 9))
    return set_err(c, GB200_ERR_ARG, "gb200_lmm_params: a_mode must be 1,2,3,4 or 9");
  if (!(l_max > l_min) || !(l_min > 0) || n_region == 0 || n_region > 100000)
    return set_err(c, GB200_ERR_ARG, "gb200_lmm_params: need 0 < l_min < l_max, n_region >= 1");
  c->prm.a_mode = a_mode; c->prm.l_min = l_min; c->prm.l_max = l_max; c->prm.n_region = (int)n_region;
  c->prm.l_mle_null = l_mle_null; c->prm.logl_mle_H0 = logl_mle_H0;
  c->common_ready = false;
  c->prm_ready = true;
  return GB200_OK;
}

static int lmm_check_ready(gb200_ctx *c, const char *who) {
  if (!c->lmm_ready) return set_err(c, GB200_ERR_STATE, std::string(who) + " before gb200_lmm_setup");
  if (!c->prm_ready) return set_err(c, GB200_ERR_STATE, std::string(who) + " before gb200_lmm_params");
  return GB200_OK;
}

// Run-constant tables of the lockstep per-SNP kernel, built once per (setup, params) pair: SNP-independent sums and h rows at the
// lambdas every SNP visits; with lmm_interp also at the Chebyshev nodes of every grid interval (the kernel then serves its Brent /
// Newton evaluations from interpolants); with x_exact the vectors v_q = U (h(l_mle_null) (.) q) behind LmmConst::xex.
struct CommonInfo { size_t J0 = 0, n_nodes = 0; double marg = 0.15; bool on = false; };
static int lmm_ensure_common(gb200_ctx *c, CommonInfo &ci) {
  ci.on = lmm_v2_supported((int)c->n_cvt, c->prm.n_region) && c->lmm_kernel != 1 && c->lmm_kernel != 3 && c->lmm_hoist;
  if (!ci.on) return GB200_OK;
  LmmConst D = make_const(c);
  const size_t J0 = (size_t)c->prm.n_region + 3, rec = lmm_common_record_doubles((int)c->n_cvt);
  const int M = lmm_cheb_nodes(), XM = lmm_cheb_xnodes();      // per interval: M nodes for the SNP-independent tables, then XM for the x-sums
  size_t n_nodes = c->lmm_interp ? (size_t)c->prm.n_region * (size_t)(M + XM) : 0;
  if ((J0 + n_nodes) * c->n_c * sizeof(double) > ((size_t)2 << 30)) n_nodes = 0;        // node rows stay below 2 GB
  // 20 nodes resolve an interval of one decade (+ margins) to ~1e-14; wider intervals (a coarse -region grid) keep the exact passes
  if (log(c->prm.l_max / c->prm.l_min) / (double)c->prm.n_region > 2.4) n_nodes = 0;
  ci.J0 = J0; ci.n_nodes = n_nodes;
  if (c->common_ready) return GB200_OK;
  const size_t J = J0 + n_nodes;
  GB_CUDA(c, c->dHrows.reserve(J * c->n_c * sizeof(double)));
  GB_CUDA(c, c->dCtab.reserve(J * rec * sizeof(double)));
  if (n_nodes) {
    std::vector<double> lams(n_nodes);
    const double interval = log(c->prm.l_max / c->prm.l_min) / (double)c->prm.n_region;
    for (int g = 0; g < c->prm.n_region; ++g) {
      const double lo = log(c->prm.l_min) + interval * (double)g - ci.marg, hi = log(c->prm.l_min) + interval * (double)(g + 1) + ci.marg;
      for (int m = 0; m < M; ++m) lams[(size_t)g * M + m] = exp(0.5 * (lo + hi) + 0.5 * (hi - lo) * cos(M_PI * ((double)m + 0.5) / (double)M));
      for (int m = 0; m < XM; ++m)
        lams[(size_t)c->prm.n_region * M + (size_t)g * XM + m] = exp(0.5 * (lo + hi) + 0.5 * (hi - lo) * cos(M_PI * ((double)m + 0.5) / (double)XM));
    }
    GB_CUDA(c, c->dNodeLam.reserve(n_nodes * sizeof(double)));
    GB_CUDA(c, cudaMemcpyAsync(c->dNodeLam.p, lams.data(), n_nodes * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    GB_CUDA(c, cudaStreamSynchronize(c->stream));      // `lams` goes out of scope
    GB_CUDA(c, c->dCheb.reserve(lmm_cheb_doubles((int)c->n_cvt, c->prm.n_region) * sizeof(double)));
  }
  GB_CUDA(c, launch_lmm_common((int)c->n_cvt, D, c->prm, c->dHrows.as<double>(), c->dCtab.as<double>(), c->dNodeLam.as<double>(),
                               (int)n_nodes, c->dCheb.as<double>(), c->stream));
  c->vnull_ready = false;
  if (c->x_exact && c->prm.l_mle_null > 0.0 && c->dU.p && !c->overlap) {
    gb::DevBuf scratch;
    const size_t by = (c->n_cvt + 1) * c->n_c * sizeof(double);
    GB_CUDA(c, scratch.reserve(by));
    GB_CUDA(c, c->dVnull.reserve(by));
    GB_CUDA(c, cudaMemsetAsync(c->dVnull.p, 0, by, c->stream));
    GB_CUDA(c, launch_lmm_vnull((int)c->n_cvt, D, c->prm.l_mle_null, c->dU.as<double>(), scratch.as<double>(), c->dVnull.as<double>(), c->stream));
    GB_CUDA(c, cudaStreamSynchronize(c->stream));
    scratch.release();
    c->vnull_ready = true;
  }
  // exact LINEAR x-sums at every hoisted lambda (LmmConst::xsum): V = U A, A's columns h_b^k (.) q over the shared rows and the x-node rows
  c->i8.xs_ready = false; c->i8.xs_valid = false;
  if (c->x_exact == 2 && n_nodes && c->dU.p && !c->overlap && !c->mv_ready && c->n >= 1024 && c->utx_path != 1) {
    const int M = lmm_cheb_nodes(), XM = lmm_cheb_xnodes();
    const int nq = (int)c->n_cvt + 1, nblocks = (int)J0 + c->prm.n_region * XM, x0 = (int)J0 + c->prm.n_region * M;
    // eigenvectors with a coherent plane rounding: the null ones (constant vector of a centred K) and the 32 leading ones (population
    // structure: near piecewise-constant); their U^T x entries come exactly from the side GEMM as well
    std::vector<double> ev(c->n);
    GB_CUDA(c, cudaMemcpyAsync(ev.data(), c->dEval.p, c->n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    GB_CUDA(c, cudaStreamSynchronize(c->stream));
    std::vector<int> pidx;
    for (size_t i = 0; i < c->n && ev[i] == 0.0; ++i) pidx.push_back((int)i);
    const bool too_many_null = pidx.size() > 64;
    for (size_t k = 0; k < 32 && k < c->n; ++k) { const int i = (int)(c->n - 1 - k); if (ev[i] != 0.0) pidx.push_back(i); }
    const int npatch = (int)pidx.size(), patch0 = nblocks * 2 * nq + nq;
    const int ncol = patch0 + npatch;
    if (!too_many_null) {
    GB_CUDA(c, c->i8.xs_patch_idx.reserve(npatch * sizeof(int)));
    GB_CUDA(c, cudaMemcpyAsync(c->i8.xs_patch_idx.p, pidx.data(), npatch * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    GB_CUDA(c, cudaStreamSynchronize(c->stream));
    c->i8.xs_npatch = npatch; c->i8.xs_patch0 = patch0;
    gb::DevBuf dA;
    GB_CUDA(c, dA.reserve(c->n * (size_t)ncol * sizeof(double)));
    GB_CUDA(c, c->i8.xs_V.reserve(c->n * (size_t)ncol * sizeof(double)));
    GB_CUDA(c, launch_lmm_acols((int)c->n_cvt, D, c->dHrows.as<double>(), (int)J0, x0, nblocks, dA.as<double>(), ncol,
                                c->i8.xs_patch_idx.as<int>(), npatch, patch0, c->stream));
    GB_CUDA(c, launch_dgemm(c->n, (size_t)ncol, c->n, 1.0, c->dU.as<double>(), c->n, 1, dA.as<double>(), (size_t)ncol, 1, 0.0,
                            c->i8.xs_V.as<double>(), (size_t)ncol, false, c->stream));
    GB_CUDA(c, cudaStreamSynchronize(c->stream));
    dA.release();
    const int rc = i8_xsum_prepare(c, ncol);
    if (rc) return rc;
    c->xs_nblocks = nblocks; c->xs_skip = c->prm.n_region * M;
    }
  }
  c->i8.auto_T = 0; c->i8.ready = false;             // the plane count depends on whether the exact linear sums are on
  GB_CUDA(c, cudaStreamSynchronize(c->stream));       // the tests may run on a side stream
  c->common_ready = true;
  return GB200_OK;
}

static int lmm_ensure_common_public(gb200_ctx *c) { CommonInfo ci; return lmm_ensure_common(c, ci); }

// association kernel on a device-resident rotated batch
static int lmm_assoc_dev(gb200_ctx *c, const double *UtXt, size_t l, size_t ldu, gb200_sumstat *out_dev,
                         bool plink_rule = false, cudaStream_t st = nullptr, unsigned int *ticket = nullptr) {
  LmmConst D = make_const(c);
  c->prm.plink_rule = plink_rule ? 1 : 0;
  if (!st) st = c->stream;
  if (!ticket) ticket = c->dTicket.as<unsigned int>();
  const bool v2_ok = lmm_v2_supported((int)c->n_cvt, c->prm.n_region) && ldu == c->n_c;
  if (c->lmm_kernel == 2 && !v2_ok) return set_err(c, GB200_ERR_UNSUPPORTED, "lmm_kernel=2 (lockstep CTA kernel) needs n_cvt <= 3 and n_region <= 64");
  if (v2_ok) {
    CommonInfo ci;
    const int rc = lmm_ensure_common(c, ci);
    if (rc) return rc;
    if (ci.on) {
      D.Hrows = c->dHrows.as<double>(); D.ctab = c->dCtab.as<double>(); D.n_common = (int)ci.J0;
      if (ci.n_nodes) { D.cheb = c->dCheb.as<double>(); D.cheb_marg = ci.marg; }
      // exact x-sums of the batch the int8 projection has just written into this very buffer
      if (c->i8.xex_valid && c->i8.xex_for == UtXt && c->i8.xex_l == l) D.xex = c->i8.xex.as<double>();
      if (c->i8.xs_valid && c->i8.xs_for == UtXt && c->i8.xs_l == l && ci.n_nodes) {
        D.xsum = c->i8.xs_out.as<double>(); D.xsum_ld = (int)c->i8.xs_ld; D.xsum_nblocks = c->xs_nblocks; D.xsum_nblocks_skip = c->xs_skip;
      }
    }
  }
  c->i8.xex_valid = false; c->i8.xs_valid = false;
  ProfScope ps(c, "lmm", 1, st);
  if (v2_ok && c->lmm_kernel != 1 && c->lmm_kernel != 3)
    GB_CUDA(c, launch_lmm_assoc_v2((int)c->n_cvt, D, c->prm, UtXt, ldu, (int)l, out_dev, ticket, c->num_sms, st));
  else
    GB_CUDA(c, launch_lmm_assoc((int)c->n_cvt, D, c->prm, UtXt, ldu, (int)l, out_dev, ticket, c->num_sms, st));
  return GB200_OK;
}

// UtXt (l x n) = Xs (l x n, SNP-major, ld n) * U   -- the fast_dgemm("T","N",U,X) of lmm.cpp:1521
static int project_fp64_snpmajor(gb200_ctx *c, const double *Xs, size_t l, double *UtXt) {
  const size_t n = c->n;
  ProfScope ps(c, "utx");
  GB_CUDA(c, launch_dgemm(l, n, n, 1.0, Xs, n, 1, c->dU.as<double>(), n, 1, 0.0, UtXt, c->n_c, false, c->stream));
  return GB200_OK;
}

int gb200_lmm_assoc_utx(gb200_ctx *c, const double *UtXt, size_t l, size_t ldu, gb200_sumstat *out) {
  if (!c) return GB200_ERR_ARG;
  int rc = lmm_check_ready(c, "gb200_lmm_assoc_utx");
  if (rc) return rc;
  if (l == 0) return GB200_OK;
  if (!UtXt || !out || ldu < c->n) return set_err(c, GB200_ERR_ARG, "gb200_lmm_assoc_utx: bad argument");
  const size_t n = c->n;
  GB_CUDA(c, reserve_zeroed(c->dUtXt, l * c->n_c * 8, c->stream));
  GB_CUDA(c, c->dOut.reserve(l * sizeof(gb200_sumstat)));
  GB_CUDA(c, cudaMemcpy2DAsync(c->dUtXt.p, c->n_c * 8, UtXt, ldu * 8, n * 8, l, cudaMemcpyHostToDevice, c->stream));
  rc = lmm_assoc_dev(c, c->dUtXt.as<double>(), l, c->n_c, c->dOut.as<gb200_sumstat>());
  if (rc) return rc;
  GB_CUDA(c, cudaMemcpyAsync(out, c->dOut.p, l * sizeof(gb200_sumstat), cudaMemcpyDeviceToHost, c->stream));
  GB_CUDA(c, cudaStreamSynchronize(c->stream));
  return GB200_OK;
}

int gb200_lmm_project(gb200_ctx *c, const double *Xb, size_t l, size_t ldx, double *UtXt) {
  if (!c) return GB200_ERR_ARG;
  if (!c->lmm_ready) return set_err(c, GB200_ERR_STATE, "gb200_lmm_project before gb200_lmm_setup");
  if (l == 0) return GB200_OK;
  if (!Xb || !UtXt || ldx < l) return set_err(c, GB200_ERR_ARG, "gb200_lmm_project: bad argument");
  const size_t n = c->n;
  GB_CUDA(c, c->dX.reserve(n * l * 8));
  GB_CUDA(c, reserve_zeroed(c->dUtXt, l * c->n_c * 8, c->stream));
  GB_CUDA(c, cudaMemcpy2DAsync(c->dX.p, l * 8, Xb, ldx * 8, l * 8, n, cudaMemcpyHostToDevice, c->stream));
  {
    ProfScope ps(c, "utx");
    // A(s,j) = X[j*l + s]
    GB_CUDA(c, launch_dgemm(l, n, n, 1.0, c->dX.as<double>(), 1, l, c->dU.as<double>(), n, 1, 0.0,
                            c->dUtXt.as<double>(), c->n_c, false, c->stream));
  }
  GB_CUDA(c, cudaMemcpy2DAsync(UtXt, n * 8, c->dUtXt.p, c->n_c * 8, n * 8, l, cudaMemcpyDeviceToHost, c->stream));
  GB_CUDA(c, cudaStreamSynchronize(c->stream));
  return GB200_OK;
}

int gb200_lmm_batch(gb200_ctx *c, const double *Xb, size_t l, size_t ldx, gb200_sumstat *out) {
  if (!c) return GB200_ERR_ARG;
  int rc = lmm_check_ready(c, "gb200_lmm_batch");
  if (rc) return rc;
  if (l == 0) return GB200_OK;     // the reference aborts here (fastblas.cpp:193); a no-op is the sane drop-in
  if (!Xb || !out || ldx < l) return set_err(c, GB200_ERR_ARG, "gb200_lmm_batch: bad argument");
  const size_t n = c->n;
  GB_CUDA(c, c->dX.reserve(n * l * 8));
  GB_CUDA(c, reserve_zeroed(c->dUtXt, l * c->n_c * 8, c->stream));
  GB_CUDA(c, c->dOut.reserve(l * sizeof(gb200_sumstat)));
  GB_CUDA(c, cudaMemcpy2DAsync(c->dX.p, l * 8, Xb, ldx * 8, l * 8, n, cudaMemcpyHostToDevice, c->stream));
  {
    ProfScope ps(c, "utx");
    GB_CUDA(c, launch_dgemm(l, n, n, 1.0, c->dX.as<double>(), 1, l, c->dU.as<double>(), n, 1, 0.0,
                            c->dUtXt.as<double>(), c->n_c, false, c->stream));
  }
  rc = lmm_assoc_dev(c, c->dUtXt.as<double>(), l, c->n_c, c->dOut.as<gb200_sumstat>());
  if (rc) return rc;
  GB_CUDA(c, cudaMemcpyAsync(out, c->dOut.p, l * sizeof(gb200_sumstat), cudaMemcpyDeviceToHost, c->stream));
  GB_CUDA(c, cudaStreamSynchronize(c->stream));
  return GB200_OK;
}

int gb200_lmm_batch_geno(gb200_ctx *c, const double *G, size_t l, size_t ldg, gb200_sumstat *out) {
  if (!c) return GB200_ERR_ARG;
  int rc = lmm_check_ready(c, "gb200_lmm_batch_geno");
  if (rc) return rc;
  if (l == 0) return GB200_OK;
  const size_t n = c->n;
  if (!G || !out || ldg < n) return set_err(c, GB200_ERR_ARG, "gb200_lmm_batch_geno: bad argument");
  GB_CUDA(c, c->dX.reserve(n * l * 8));
  GB_CUDA(c, reserve_zeroed(c->dUtXt, l * c->n_c * 8, c->stream));
  GB_CUDA(c, c->dOut.reserve(l * sizeof(gb200_sumstat)));
  GB_CUDA(c, cudaMemcpy2DAsync(c->dX.p, n * 8, G, ldg * 8, n * 8, l, cudaMemcpyHostToDevice, c->stream));
  // dosage rows printed with a few decimals are exact integer digit rows: tensor-core projection (i8gemm_sm100.cu); otherwise FP64
  bool taken = false;
  if (c->utx_path != 1 && i8_available(c) && (c->utx_path == 2 || n >= 1024)) {
    rc = i8_project_geno(c, c->dX.as<double>(), l, n, c->dUtXt.as<double>(), &taken);
    if (rc) return rc;
  }
  if (!taken) {
    {
      ProfScope ps(c, "decode");
      GB_CUDA(c, launch_lmm_impute(c->dX.as<double>(), l, n, n, c->stream));
    }
    rc = project_fp64_snpmajor(c, c->dX.as<double>(), l, c->dUtXt.as<double>());
    if (rc) return rc;
  }
  rc = lmm_assoc_dev(c, c->dUtXt.as<double>(), l, c->n_c, c->dOut.as<gb200_sumstat>());
  if (rc) return rc;
  GB_CUDA(c, cudaMemcpyAsync(out, c->dOut.p, l * sizeof(gb200_sumstat), cudaMemcpyDeviceToHost, c->stream));
  GB_CUDA(c, cudaStreamSynchronize(c->stream));
  return GB200_OK;
}

// rotate a device-resident bed batch into UtXt (l x n): int8 tensor-core path for n >= 1024
// (or when forced), FP64 decode + dgemm otherwise.  idx_dev maps analysed position -> ni_total index.
static int project_bed_dev(gb200_ctx *c, const unsigned char *bed_dev, const int *idx_dev, size_t ni_total,
                           size_t l, size_t bytes_per_snp, double *dst = nullptr) {
  const size_t n = c->n;
  if (c->prm_ready) { CommonInfo ci; const int rc0 = lmm_ensure_common(c, ci); if (rc0) return rc0; }   // v_q of LmmConst::xex before the first batch
  if (!dst) { GB_CUDA(c, reserve_zeroed(c->dUtXt, l * c->n_c * 8, c->stream)); dst = c->dUtXt.as<double>(); }
  bool use_i8 = false;
  if (c->utx_path == 2) {
    if (!i8_available(c)) return set_err(c, GB200_ERR_UNSUPPORTED, "int8 tensor-core path not available (no cuTensorMapEncodeTiled)");
    use_i8 = true;
  } else if (c->utx_path == 0) {
    use_i8 = i8_available(c) && n >= 1024;
  }
  if (use_i8) return i8_project_bed(c, bed_dev, idx_dev, ni_total, l, bytes_per_snp, dst);
  GB_CUDA(c, c->dX.reserve(n * l * 8));
  {
    ProfScope ps(c, "decode", 2);
    GB_CUDA(c, launch_bed_decode(bed_dev, l, bytes_per_snp, idx_dev, n, c->dX.as<double>(), n, c->stream));
    GB_CUDA(c, launch_lmm_impute(c->dX.as<double>(), l, n, n, c->stream));
  }
  return project_fp64_snpmajor(c, c->dX.as<double>(), l, dst);
}

static int lmm_bed_core(gb200_ctx *c, const unsigned char *bed_dev, const int *idx_dev, size_t ni_total,
                        size_t l, size_t bytes_per_snp, gb200_sumstat *out_dev) {
  // Software pipeline over sub-batches: the projection of sub-batch i+1 (tensor pipe, main stream) runs while the
  // per-SNP tests of sub-batch i (FP64 pipe, side stream) are in flight; U^T X is double buffered.
  const size_t SUB = 2048;
  if (c->overlap && l >= 2 * SUB) {
    if (!c->side) {
      GB_CUDA(c, cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking));
      for (int k = 0; k < 2; ++k) {
        GB_CUDA(c, cudaEventCreateWithFlags(&c->evG[k], cudaEventDisableTiming));
        GB_CUDA(c, cudaEventCreateWithFlags(&c->evL[k], cudaEventDisableTiming));
      }
    }
    GB_CUDA(c, reserve_zeroed(c->dUtXt, SUB * c->n_c * 8, c->stream));
    GB_CUDA(c, reserve_zeroed(c->dUtXt2, SUB * c->n_c * 8, c->stream));
    double *buf[2] = {c->dUtXt.as<double>(), c->dUtXt2.as<double>()};
    unsigned int *ticket = c->dTicket.as<unsigned int>() + 4;           // separate ticket word for the side stream
    size_t i = 0;
    for (size_t s0 = 0; s0 < l; s0 += SUB, ++i) {
      const size_t lc = (l - s0 < SUB) ? (l - s0) : SUB;
      const int b = (int)(i & 1);
      if (i >= 2) GB_CUDA(c, cudaStreamWaitEvent(c->stream, c->evL[b], 0));     // tests of sub-batch i-2 released the buffer
      int rc = project_bed_dev(c, bed_dev + s0 * bytes_per_snp, idx_dev, ni_total, lc, bytes_per_snp, buf[b]);
      if (rc) return rc;
      GB_CUDA(c, cudaEventRecord(c->evG[b], c->stream));
      GB_CUDA(c, cudaStreamWaitEvent(c->side, c->evG[b], 0));
      rc = lmm_assoc_dev(c, buf[b], lc, c->n_c, out_dev + s0, /*plink_rule=*/true, c->side, ticket);
      if (rc) return rc;
      GB_CUDA(c, cudaEventRecord(c->evL[b], c->side));
    }
    // the caller's stream observes completion of everything
    GB_CUDA(c, cudaStreamWaitEvent(c->stream, c->evL[0], 0));
    GB_CUDA(c, cudaStreamWaitEvent(c->stream, c->evL[1], 0));
    return GB200_OK;
  }
  if (c->stage_mask & 1) {
    int rc = project_bed_dev(c, bed_dev, idx_dev, ni_total, l, bytes_per_snp);
    if (rc) return rc;
  }
  if (!(c->stage_mask & 2)) return GB200_OK;       // measurement runs only: projection without the tests (stage_mask = 1)
  return lmm_assoc_dev(c, c->dUtXt.as<double>(), l, c->n_c, out_dev, /*plink_rule=*/true);   // AnalyzePlink semantics
}

// any number of device-resident rows: sub-batches of lmm_chunk_snps() through lmm_bed_core, serially on the context stream
static int lmm_bed_chunked_dev(gb200_ctx *c, const unsigned char *bed_dev, const int *idx_dev, size_t ni_total, size_t l,
                               size_t bytes_per_snp, gb200_sumstat *out_dev) {
  const size_t chunk = lmm_chunk_snps(c);
  for (size_t s0 = 0; s0 < l; s0 += chunk) {
    const size_t lc = (l - s0 < chunk) ? (l - s0) : chunk;
    int rc = lmm_bed_core(c, bed_dev + s0 * bytes_per_snp, idx_dev, ni_total, lc, bytes_per_snp, out_dev + s0);
    if (rc) return rc;
  }
  return GB200_OK;
}

static int upload_idx_from_mask(gb200_ctx *c, const unsigned char *idv_mask, size_t ni_total, const int **idx_dev) {
  *idx_dev = nullptr;
  if (!idv_mask) {
    if (ni_total != c->n) return set_err(c, GB200_ERR_ARG, "idv_mask == NULL requires ni_total == n");
    return GB200_OK;
  }
  if (c->mask_host.size() != ni_total || memcmp(c->mask_host.data(), idv_mask, ni_total) != 0) {
    c->mask_host.assign(idv_mask, idv_mask + ni_total);
    c->idx_host.clear();
    for (size_t i = 0; i < ni_total; ++i) if (idv_mask[i]) c->idx_host.push_back((int)i);
    if (c->idx_host.size() != c->n) {
      c->mask_host.clear();
      return set_err(c, GB200_ERR_ARG, "idv_mask selects a number of individuals different from n");
    }
    GB_CUDA(c, c->dIdx.reserve(c->n * sizeof(int)));
    GB_CUDA(c, cudaMemcpyAsync(c->dIdx.p, c->idx_host.data(), c->n * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  }
  *idx_dev = c->dIdx.as<int>();
  return GB200_OK;
}

// ---- G x E (src/gemma.cpp:2580-2582, 2809-2828; LMM::AnalyzePlinkGXE / AnalyzeBimbamGXE, src/lmm.cpp:2283-2608) ----------------
int gb200_lmm_gxe_setup(gb200_ctx *c, const double *env) {
  if (!c) return GB200_ERR_ARG;
  if (!c->lmm_ready) return set_err(c, GB200_ERR_STATE, "gb200_lmm_gxe_setup before gb200_lmm_setup");
  if (!env) return set_err(c, GB200_ERR_ARG, "gb200_lmm_gxe_setup: null env");
  if (c->n_cvt + 2 > GB200_MAX_CVT) return set_err(c, GB200_ERR_UNSUPPORTED, "G x E needs n_cvt + 2 <= GB200_MAX_CVT");
  const size_t n = c->n, n_c = c->n_c, cN = c->n_cvt;
  GB_CUDA(c, c->dEnv.reserve(n * 8));
  GB_CUDA(c, c->dWtx.reserve((cN + 2) * n_c * 8));
  GB_CUDA(c, cudaMemsetAsync(c->dWtx.p, 0, (cN + 2) * n_c * 8, c->stream));
  GB_CUDA(c, cudaMemcpyAsync(c->dEnv.p, env, n * 8, cudaMemcpyHostToDevice, c->stream));
  GB_CUDA(c, cudaMemcpyAsync(c->dWtx.p, c->dWt.p, cN * n_c * 8, cudaMemcpyDeviceToDevice, c->stream));
  // U^T env (gsl_blas_dgemv(CblasTrans, 1.0, U, env, ...), src/lmm.cpp:2319): row cN of the expanded covariate rows
  GB_CUDA(c, launch_dgemm(n, 1, n, 1.0, c->dU.as<double>(), 1, n, c->dEnv.as<double>(), 1, 1, 0.0,
                          c->dWtx.as<double>() + cN * n_c, 1, false, c->stream));
  GB_CUDA(c, cudaStreamSynchronize(c->stream));
  c->gxe_ready = true;
  return GB200_OK;
}

// dX holds l mean-imputed SNP rows (n doubles each) on the device
static int lmm_gxe_core(gb200_ctx *c, size_t l, gb200_sumstat *out_dev) {
  const size_t n = c->n;
  GB_CUDA(c, c->dX2.reserve(n * l * 8));
  GB_CUDA(c, c->dFlip.reserve(l));
  GB_CUDA(c, reserve_zeroed(c->dUtXt, l * c->n_c * 8, c->stream));
  GB_CUDA(c, reserve_zeroed(c->dUtXt2, l * c->n_c * 8, c->stream));
  {
    ProfScope ps(c, "decode");
    GB_CUDA(c, launch_gxe_prepare(c->dX.as<double>(), c->dX2.as<double>(), c->dEnv.as<double>(), l, n, c->dFlip.as<unsigned char>(), c->stream));
  }
  int rc = project_fp64_snpmajor(c, c->dX.as<double>(), l, c->dUtXt.as<double>());
  if (rc) return rc;
  rc = project_fp64_snpmajor(c, c->dX2.as<double>(), l, c->dUtXt2.as<double>());
  if (rc) return rc;
  LmmConst D = make_const(c);
  D.Wt = c->dWtx.as<double>();
  ProfScope ps(c, "lmm");
  GB_CUDA(c, launch_lmm_gxe((int)c->n_cvt, D, c->prm, c->dUtXt.as<double>(), c->dUtXt2.as<double>(), c->n_c, (int)l,
                            c->dFlip.as<unsigned char>(), out_dev, c->dTicket.as<unsigned int>(), c->num_sms, c->stream));
  return GB200_OK;
}

static int gxe_check(gb200_ctx *c, const char *who) {
  int rc = lmm_check_ready(c, who);
  if (rc) return rc;
  if (!c->gxe_ready) return set_err(c, GB200_ERR_STATE, std::string(who) + " before gb200_lmm_gxe_setup");
  return GB200_OK;
}

int gb200_lmm_gxe_batch_geno(gb200_ctx *c, const double *G, size_t l, size_t ldg, gb200_sumstat *out) {
  if (!c) return GB200_ERR_ARG;
  int rc = gxe_check(c, "gb200_lmm_gxe_batch_geno");
  if (rc) return rc;
  if (l == 0) return GB200_OK;
  const size_t n = c->n;
  if (!G || !out || ldg < n) return set_err(c, GB200_ERR_ARG, "gb200_lmm_gxe_batch_geno: bad argument");
  GB_CUDA(c, c->dX.reserve(n * l * 8));
  GB_CUDA(c, c->dOut.reserve(l * sizeof(gb200_sumstat)));
  GB_CUDA(c, cudaMemcpy2DAsync(c->dX.p, n * 8, G, ldg * 8, n * 8, l, cudaMemcpyHostToDevice, c->stream));
  {
    ProfScope ps(c, "decode");
    GB_CUDA(c, launch_lmm_impute(c->dX.as<double>(), l, n, n, c->stream));
  }
  rc = lmm_gxe_core(c, l, c->dOut.as<gb200_sumstat>());
  if (rc) return rc;
  GB_CUDA(c, cudaMemcpyAsync(out, c->dOut.p, l * sizeof(gb200_sumstat), cudaMemcpyDeviceToHost, c->stream));
  GB_CUDA(c, cudaStreamSynchronize(c->stream));
  return GB200_OK;
}

int gb200_lmm_gxe_batch_bed(gb200_ctx *c, const unsigned char *bed, const unsigned char *idv_mask, size_t ni_total, size_t l,
                            size_t bytes_per_snp, gb200_sumstat *out) {
  if (!c) return GB200_ERR_ARG;
  int rc = gxe_check(c, "gb200_lmm_gxe_batch_bed");
  if (rc) return rc;
  if (l == 0) return GB200_OK;
  if (!bed || !out || bytes_per_snp != (ni_total + 3) / 4) return set_err(c, GB200_ERR_ARG, "gb200_lmm_gxe_batch_bed: bad argument");
  const int *idx_dev = nullptr;
  rc = upload_idx_from_mask(c, idv_mask, ni_total, &idx_dev);
  if (rc) return rc;
  const size_t n = c->n;
  GB_CUDA(c, c->dBed.reserve(l * bytes_per_snp));
  GB_CUDA(c, c->dX.reserve(n * l * 8));
  GB_CUDA(c, c->dOut.reserve(l * sizeof(gb200_sumstat)));
  GB_CUDA(c, cudaMemcpyAsync(c->dBed.p, bed, l * bytes_per_snp, cudaMemcpyHostToDevice, c->stream));
  {
    ProfScope ps(c, "decode", 2);
    GB_CUDA(c, launch_bed_decode(c->dBed.as<unsigned char>(), l, bytes_per_snp, idx_dev, n, c->dX.as<double>(), n, c->stream));
    GB_CUDA(c, launch_lmm_impute(c->dX.as<double>(), l, n, n, c->stream));
  }
  rc = lmm_gxe_core(c, l, c->dOut.as<gb200_sumstat>());
  if (rc) return rc;
  GB_CUDA(c, cudaMemcpyAsync(out, c->dOut.p, l * sizeof(gb200_sumstat), cudaMemcpyDeviceToHost, c->stream));
  GB_CUDA(c, cudaStreamSynchronize(c->stream));
  return GB200_OK;
}

// ---- -lm (linear model, no kinship): LM::AnalyzeBimbam / AnalyzePlink + CalcvPv + LmCalcP, src/lm.cpp:224-288, 382-640 ----------
int gb200_lm_setup(gb200_ctx *c, size_t n, size_t n_cvt, const double *W, size_t ldw, const double *y) {
  if (!c) return GB200_ERR_ARG;
  if (n == 0 || n_cvt == 0 || !W || !y || ldw < n_cvt) return set_err(c, GB200_ERR_ARG, "gb200_lm_setup: bad argument");
  if (n_cvt > GB200_MAX_CVT) return set_err(c, GB200_ERR_UNSUPPORTED, "gb200_lm_setup: n_cvt exceeds GB200_MAX_CVT");
  if (n <= n_cvt + 1) return set_err(c, GB200_ERR_ARG, "need n > n_cvt + 1");
  c->lm_ready = false;
  // W'W, W'y, y'y are run constants over c <= 32 columns: formed and inverted on the host like the reference does
  // (gsl_blas_dgemm + LUDecomp/LUInvert, src/lm.cpp:404-413); the per-SNP sums run on the device
  std::vector<double> Wt(n_cvt * n), WtW(n_cvt * n_cvt, 0.0), Wty(n_cvt, 0.0), small(n_cvt * n_cvt + n_cvt);
  double yy = 0.0;
  for (size_t i = 0; i < n; ++i) {
    yy += y[i] * y[i];
    for (size_t a = 0; a < n_cvt; ++a) {
      const double wa = W[i * ldw + a];
      Wt[a * n + i] = wa; Wty[a] += wa * y[i];
      for (size_t b = 0; b < n_cvt; ++b) WtW[a * n_cvt + b] += wa * W[i * ldw + b];
    }
  }
  // inverse by Gauss-Jordan with partial pivoting
  std::vector<double> M(n_cvt * 2 * n_cvt);
  const size_t w2 = 2 * n_cvt;
  for (size_t i = 0; i < n_cvt; ++i) for (size_t j = 0; j < n_cvt; ++j) { M[i * w2 + j] = WtW[i * n_cvt + j]; M[i * w2 + n_cvt + j] = (i == j) ? 1.0 : 0.0; }
  for (size_t k = 0; k < n_cvt; ++k) {
    size_t pr = k;
    for (size_t i = k + 1; i < n_cvt; ++i) if (fabs(M[i * w2 + k]) > fabs(M[pr * w2 + k])) pr = i;
    if (M[pr * w2 + k] == 0.0) return set_err(c, GB200_ERR_ARG, "gb200_lm_setup: W'W is singular");
    if (pr != k) for (size_t j = 0; j < w2; ++j) std::swap(M[k * w2 + j], M[pr * w2 + j]);
    const double piv = M[k * w2 + k];
    for (size_t j = 0; j < w2; ++j) M[k * w2 + j] /= piv;
    for (size_t i = 0; i < n_cvt; ++i) if (i != k) { const double f = M[i * w2 + k]; if (f != 0.0) for (size_t j = 0; j < w2; ++j) M[i * w2 + j] -= f * M[k * w2 + j]; }
  }
  double quad = 0.0;
  for (size_t a = 0; a < n_cvt; ++a) {
    double t = 0.0;
    for (size_t b = 0; b < n_cvt; ++b) { small[a * n_cvt + b] = M[a * w2 + n_cvt + b]; t += M[a * w2 + n_cvt + b] * Wty[b]; }
    quad += t * Wty[a];
    small[n_cvt * n_cvt + a] = Wty[a];
  }
  c->lm_yPwy = yy - quad;                                  // CalcvPv(WtWi, Wty, y, yPwy), src/lm.cpp:247-264
  GB_CUDA(c, c->dLmW.reserve(n_cvt * n * 8));
  GB_CUDA(c, c->dLmY.reserve(n * 8));
  GB_CUDA(c, c->dLmSmall.reserve(small.size() * 8));
  GB_CUDA(c, cudaMemcpyAsync(c->dLmW.p, Wt.data(), n_cvt * n * 8, cudaMemcpyHostToDevice, c->stream));
  GB_CUDA(c, cudaMemcpyAsync(c->dLmY.p, y, n * 8, cudaMemcpyHostToDevice, c->stream));
  GB_CUDA(c, cudaMemcpyAsync(c->dLmSmall.p, small.data(), small.size() * 8, cudaMemcpyHostToDevice, c->stream));
  GB_CUDA(c, cudaStreamSynchronize(c->stream));
  c->lm_n = n; c->lm_c = n_cvt; c->lm_ready = true;
  return GB200_OK;
}

static int lm_core(gb200_ctx *c, size_t l, int a_mode, gb200_sumstat *out) {
  const int test_mode = a_mode > 50 ? a_mode - 50 : a_mode;
  if (test_mode < 1 || test_mode > 4) return set_err(c, GB200_ERR_ARG, "-lm mode must be 1..4 (or 51..54)");
  GB_CUDA(c, c->dOut.reserve(l * sizeof(gb200_sumstat)));
  {
    ProfScope ps(c, "lmm");
    GB_CUDA(c, launch_lm(c->dX.as<double>(), l, (int)c->lm_n, (int)c->lm_c, c->dLmW.as<double>(), c->dLmY.as<double>(), c->dLmSmall.as<double>(),
                         c->dLmSmall.as<double>() + c->lm_c * c->lm_c, c->lm_yPwy, test_mode, c->dOut.as<gb200_sumstat>(), c->stream));
  }
  GB_CUDA(c, cudaMemcpyAsync(out, c->dOut.p, l * sizeof(gb200_sumstat), cudaMemcpyDeviceToHost, c->stream));
  GB_CUDA(c, cudaStreamSynchronize(c->stream));
  return GB200_OK;
}

int gb200_lm_batch_geno(gb200_ctx *c, const double *G, size_t l, size_t ldg, int a_mode, gb200_sumstat *out) {
  if (!c) return GB200_ERR_ARG;
  if (!c->lm_ready) return set_err(c, GB200_ERR_STATE, "gb200_lm_batch_geno before gb200_lm_setup");
  if (l == 0) return GB200_OK;
  const size_t n = c->lm_n;
  if (!G || !out || ldg < n) return set_err(c, GB200_ERR_ARG, "gb200_lm_batch_geno: bad argument");
  GB_CUDA(c, c->dX.reserve(n * l * 8));
  GB_CUDA(c, cudaMemcpy2DAsync(c->dX.p, n * 8, G, ldg * 8, n * 8, l, cudaMemcpyHostToDevice, c->stream));
  {
    ProfScope ps(c, "decode");
    GB_CUDA(c, launch_lmm_impute(c->dX.as<double>(), l, n, n, c->stream));
  }
  return lm_core(c, l, a_mode, out);
}

int gb200_lm_batch_bed(gb200_ctx *c, const unsigned char *bed, const unsigned char *idv_mask, size_t ni_total, size_t l, size_t bytes_per_snp,
                       int a_mode, gb200_sumstat *out) {
  if (!c) return GB200_ERR_ARG;
  if (!c->lm_ready) return set_err(c, GB200_ERR_STATE, "gb200_lm_batch_bed before gb200_lm_setup");
  if (l == 0) return GB200_OK;
  if (!bed || !out || bytes_per_snp != (ni_total + 3) / 4) return set_err(c, GB200_ERR_ARG, "gb200_lm_batch_bed: bad argument");
  const size_t n = c->lm_n;
  // analysed-individual index list (same convention as gb200_lmm_batch_bed)
  std::vector<int> idx;
  if (idv_mask) { for (size_t i = 0; i < ni_total; ++i) if (idv_mask[i]) idx.push_back((int)i); }
  else { if (ni_total != n) return set_err(c, GB200_ERR_ARG, "idv_mask == NULL requires ni_total == n"); }
  if (idv_mask && idx.size() != n) return set_err(c, GB200_ERR_ARG, "idv_mask selects a number of individuals different from n");
  const int *idx_dev = nullptr;
  if (idv_mask) {
    GB_CUDA(c, c->dIdx.reserve(n * sizeof(int)));
    GB_CUDA(c, cudaMemcpyAsync(c->dIdx.p, idx.data(), n * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    GB_CUDA(c, cudaStreamSynchronize(c->stream));
    c->mask_host.clear();                                   // the cached mask of the -lmm entry points no longer matches dIdx
    idx_dev = c->dIdx.as<int>();
  }
  GB_CUDA(c, c->dBed.reserve(l * bytes_per_snp));
  GB_CUDA(c, c->dX.reserve(n * l * 8));
  GB_CUDA(c, cudaMemcpyAsync(c->dBed.p, bed, l * bytes_per_snp, cudaMemcpyHostToDevice, c->stream));
  {
    ProfScope ps(c, "decode", 2);
    GB_CUDA(c, launch_bed_decode(c->dBed.as<unsigned char>(), l, bytes_per_snp, idx_dev, n, c->dX.as<double>(), n, c->stream));
    GB_CUDA(c, launch_lmm_impute(c->dX.as<double>(), l, n, n, c->stream));
  }
  return lm_core(c, l, a_mode, out);
}

// ---- multivariate LMM, two phenotypes (MVLMM::AnalyzeBimbam / AnalyzePlink, src/mvlmm.cpp:2972-3899; -lmm 1) -------------------
int gb200_mvlmm_setup(gb200_ctx *c, size_t n, size_t n_cvt, size_t n_ph, const double *U, size_t ldu, const double *eval, const double *W,
                      size_t ldw, const double *Y, size_t ldy) {
  if (!c) return GB200_ERR_ARG;
  if (!Y || !W || ldy < n_ph) return set_err(c, GB200_ERR_ARG, "gb200_mvlmm_setup: bad argument");
  if (n_ph != 2) return set_err(c, GB200_ERR_UNSUPPORTED, "gb200_mvlmm_setup: two phenotypes are supported in this round");
  if (n_cvt < 1 || n_cvt > 3) return set_err(c, GB200_ERR_UNSUPPORTED, "gb200_mvlmm_setup: 1..3 covariates (incl. intercept) are supported");
  c->mv_ready = false; c->mv_null_ready = false;
  std::vector<double> y0(n);
  for (size_t i = 0; i < n; ++i) y0[i] = Y[i * ldy];
  int rc = gb200_lmm_setup(c, n, n_cvt, U, ldu, eval, W, ldw, y0.data(), nullptr, nullptr);      // U, eval, U^T W and U^T y_1 on the device
  if (rc) return rc;
  const size_t n_c = c->n_c;
  GB_CUDA(c, c->dMvY.reserve(2 * n_c * 8));
  GB_CUDA(c, cudaMemsetAsync(c->dMvY.p, 0, 2 * n_c * 8, c->stream));
  GB_CUDA(c, c->dTmp.reserve(2 * n * 8));
  std::vector<double> Yt(2 * n);
  for (size_t i = 0; i < n; ++i) { Yt[i] = Y[i * ldy]; Yt[n + i] = Y[i * ldy + 1]; }
  GB_CUDA(c, cudaMemcpyAsync(c->dTmp.p, Yt.data(), 2 * n * 8, cudaMemcpyHostToDevice, c->stream));
  // U^T Y (CalcUtX, src/gemma.cpp:2700): row s of dMvY = y_s' U
  GB_CUDA(c, launch_dgemm(2, n, n, 1.0, c->dTmp.as<double>(), n, 1, c->dU.as<double>(), n, 1, 0.0, c->dMvY.as<double>(), n_c, false, c->stream));
  // MphInitial diagonals (src/mvlmm.cpp:2786-2796): univariate REML lambda + CalcLmmVgVeBeta per trait, on the univariate device path
  gb::MvConst &K = c->mvK;
  for (int s = 0; s < 2; ++s) {
    GB_CUDA(c, cudaMemcpyAsync(c->dY.p, c->dMvY.as<double>() + (size_t)s * n_c, n_c * 8, cudaMemcpyDeviceToDevice, c->stream));
    gb200_nullmodel nm;
    std::vector<double> b1(n_cvt), b2(n_cvt), b3(n_cvt), b4(n_cvt);
    // CalcLambda('R', ..., cPar.l_min, cPar.l_max, cPar.n_region, ...) of MphInitial: the limits of gb200_lmm_params when the caller set them
    const double lmin = c->prm_ready ? c->prm.l_min : 1e-5, lmax = c->prm_ready ? c->prm.l_max : 1e5;
    const size_t nreg = c->prm_ready ? (size_t)c->prm.n_region : 10;
    rc = gb200_lmm_null(c, lmin, lmax, nreg, 1.0, &nm, b1.data(), b2.data(), b3.data(), b4.data());
    if (rc) return rc;
    K.vg0[s] = nm.vg_remle; K.ve0[s] = nm.ve_remle;
  }
  // leave the univariate state as gb200_lmm_setup built it (U^T y of the FIRST phenotype): a later gb200_lmm_null / gb200_lmm_batch*
  // on this context must not silently analyse phenotype 2
  GB_CUDA(c, cudaMemcpyAsync(c->dY.p, c->dMvY.as<double>(), n_c * 8, cudaMemcpyDeviceToDevice, c->stream));
  GB_CUDA(c, cudaStreamSynchronize(c->stream));
  c->common_ready = false;
  K.n = (int)n; K.ld = (int)n_c; K.delta = c->dEval.as<double>(); K.Wt = c->dWt.as<double>(); K.Yt = c->dMvY.as<double>();
  K.em_iter = 10000; K.nr_iter = 100; K.em_prec = 1e-4; K.nr_prec = 1e-4; K.p_nr = 0.001;           // src/param.cpp:98-99
  c->mv_ready = true;
  return GB200_OK;
}

int gb200_mvlmm_null(gb200_ctx *c, double *Vg_remle, double *Ve_remle, double *B_remle, double *logl_remle, double *Vg_mle, double *Ve_mle,
                     double *B_mle, double *logl_mle) {
  if (!c) return GB200_ERR_ARG;
  if (!c->mv_ready) return set_err(c, GB200_ERR_STATE, "gb200_mvlmm_null before gb200_mvlmm_setup");
  GB_CUDA(c, c->dMvNull.reserve(sizeof(gb::MvNull)));
  {
    ProfScope ps(c, "lmm");
    GB_CUDA(c, launch_mv_null((int)c->n_cvt, c->mvK, c->dMvNull.as<gb::MvNull>(), c->stream));
  }
  gb::MvNull r;
  GB_CUDA(c, cudaMemcpyAsync(&r, c->dMvNull.p, sizeof(r), cudaMemcpyDeviceToHost, c->stream));
  GB_CUDA(c, cudaStreamSynchronize(c->stream));
  const size_t cN = c->n_cvt;
  for (int i = 0; i < 4; ++i) { if (Vg_remle) Vg_remle[i] = r.Vg_remle[i]; if (Ve_remle) Ve_remle[i] = r.Ve_remle[i]; if (Vg_mle) Vg_mle[i] = r.Vg_mle[i]; if (Ve_mle) Ve_mle[i] = r.Ve_mle[i]; }
  for (int i = 0; i < 2; ++i) for (size_t j = 0; j < cN; ++j) { if (B_remle) B_remle[i * cN + j] = r.B_remle[i * 4 + j]; if (B_mle) B_mle[i * cN + j] = r.B_mle[i * 4 + j]; }
  if (logl_remle) *logl_remle = r.logl_remle;
  if (logl_mle) *logl_mle = r.logl_mle;
  c->mv_null_ready = true;
  return GB200_OK;
}

static int mv_assoc_core(gb200_ctx *c, size_t l, int a_mode, double *out) {
  if (a_mode < 1 || a_mode > 4) return set_err(c, GB200_ERR_ARG, "multivariate -lmm mode must be 1..4");
  c->mvK.a_mode = a_mode;
  GB_CUDA(c, c->dMvOut.reserve(l * 8 * 8));
  {
    ProfScope ps(c, "lmm");
    GB_CUDA(c, launch_mv_assoc((int)c->n_cvt, c->mvK, c->dMvNull.as<gb::MvNull>(), c->dUtXt.as<double>(), c->n_c, (int)l, c->dMvOut.as<double>(),
                               c->dTicket.as<unsigned int>(), c->num_sms, c->stream));
  }
  GB_CUDA(c, cudaMemcpyAsync(out, c->dMvOut.p, l * 8 * 8, cudaMemcpyDeviceToHost, c->stream));
  GB_CUDA(c, cudaStreamSynchronize(c->stream));
  return GB200_OK;
}

int gb200_mvlmm_batch_geno(gb200_ctx *c, const double *G, size_t l, size_t ldg, int a_mode, double *out) {
  if (!c) return GB200_ERR_ARG;
  if (!c->mv_ready || !c->mv_null_ready) return set_err(c, GB200_ERR_STATE, "gb200_mvlmm_batch_geno before gb200_mvlmm_setup / gb200_mvlmm_null");
  if (l == 0) return GB200_OK;
  const size_t n = c->n;
  if (!G || !out || ldg < n) return set_err(c, GB200_ERR_ARG, "gb200_mvlmm_batch_geno: bad argument");
  GB_CUDA(c, c->dX.reserve(n * l * 8));
  GB_CUDA(c, reserve_zeroed(c->dUtXt, l * c->n_c * 8, c->stream));
  GB_CUDA(c, cudaMemcpy2DAsync(c->dX.p, n * 8, G, ldg * 8, n * 8, l, cudaMemcpyHostToDevice, c->stream));
  {
    ProfScope ps(c, "decode");
    GB_CUDA(c, launch_lmm_impute(c->dX.as<double>(), l, n, n, c->stream));
  }
  int rc = project_fp64_snpmajor(c, c->dX.as<double>(), l, c->dUtXt.as<double>());
  if (rc) return rc;
  return mv_assoc_core(c, l, a_mode, out);
}

int gb200_mvlmm_batch_bed(gb200_ctx *c, const unsigned char *bed, const unsigned char *idv_mask, size_t ni_total, size_t l, size_t bytes_per_snp,
                          int a_mode, double *out) {
  if (!c) return GB200_ERR_ARG;
  if (!c->mv_ready || !c->mv_null_ready) return set_err(c, GB200_ERR_STATE, "gb200_mvlmm_batch_bed before gb200_mvlmm_setup / gb200_mvlmm_null");
  if (l == 0) return GB200_OK;
  if (!bed || !out || bytes_per_snp != (ni_total + 3) / 4) return set_err(c, GB200_ERR_ARG, "gb200_mvlmm_batch_bed: bad argument");
  const int *idx_dev = nullptr;
  int rc = upload_idx_from_mask(c, idv_mask, ni_total, &idx_dev);
  if (rc) return rc;
  GB_CUDA(c, c->dBed.reserve(l * bytes_per_snp));
  GB_CUDA(c, cudaMemcpyAsync(c->dBed.p, bed, l * bytes_per_snp, cudaMemcpyHostToDevice, c->stream));
  rc = project_bed_dev(c, c->dBed.as<unsigned char>(), idx_dev, ni_total, l, bytes_per_snp);      // int8 tensor-core projection for n >= 1024
  if (rc) return rc;
  return mv_assoc_core(c, l, a_mode, out);
}

int gb200_lmm_batch_bed(gb200_ctx *c, const unsigned char *bed, const unsigned char *idv_mask, size_t ni_total,
                        size_t l, size_t bytes_per_snp, gb200_sumstat *out) {
  if (!c) return GB200_ERR_ARG;
  int rc = lmm_check_ready(c, "gb200_lmm_batch_bed");
  if (rc) return rc;
  if (l == 0) return GB200_OK;
  if (!bed || !out || bytes_per_snp != (ni_total + 3) / 4) return set_err(c, GB200_ERR_ARG, "gb200_lmm_batch_bed: bad argument");
  const int *idx_dev = nullptr;
  rc = upload_idx_from_mask(c, idv_mask, ni_total, &idx_dev);
  if (rc) return rc;
  // Double-buffered streamer: the rows of sub-batch i+1 travel host -> device on a copy stream while the
  // kernels of sub-batch i run on the context stream; with pinned host rows the copies are fully asynchronous, with pageable
  // rows the call blocks in the driver's staging copy while the already queued kernels execute.  One small D2H of the
  // SUMSTAT rows (64 B per SNP) ends the call.
  const size_t chunk = lmm_chunk_snps(c);
  const size_t cl = l < chunk ? l : chunk;
  GB_CUDA(c, c->dBed.reserve(cl * bytes_per_snp));
  if (l > chunk) GB_CUDA(c, c->dBed2.reserve(cl * bytes_per_snp));
  GB_CUDA(c, c->dOut.reserve(l * sizeof(gb200_sumstat)));
  if (!c->copy) {
    GB_CUDA(c, cudaStreamCreateWithFlags(&c->copy, cudaStreamNonBlocking));
    for (int k = 0; k < 2; ++k) {
      GB_CUDA(c, cudaEventCreateWithFlags(&c->evCopy[k], cudaEventDisableTiming));
      GB_CUDA(c, cudaEventCreateWithFlags(&c->evUsed[k], cudaEventDisableTiming));
    }
    GB_CUDA(c, cudaEventCreateWithFlags(&c->evStart, cudaEventDisableTiming));
  }
  unsigned char *buf[2] = {c->dBed.as<unsigned char>(), l > chunk ? c->dBed2.as<unsigned char>() : c->dBed.as<unsigned char>()};
  GB_CUDA(c, cudaEventRecord(c->evStart, c->stream));            // everything queued earlier on the context stream (mask upload, ...)
  GB_CUDA(c, cudaStreamWaitEvent(c->copy, c->evStart, 0));
  size_t i = 0;
  for (size_t s0 = 0; s0 < l; s0 += chunk, ++i) {
    const size_t lc = (l - s0 < chunk) ? (l - s0) : chunk;
    const int b = (int)(i & 1);
    if (i >= 2) GB_CUDA(c, cudaStreamWaitEvent(c->copy, c->evUsed[b], 0));          // the kernels of sub-batch i-2 are done with this buffer
    GB_CUDA(c, cudaMemcpyAsync(buf[b], bed + s0 * bytes_per_snp, lc * bytes_per_snp, cudaMemcpyHostToDevice, c->copy));
    GB_CUDA(c, cudaEventRecord(c->evCopy[b], c->copy));
    GB_CUDA(c, cudaStreamWaitEvent(c->stream, c->evCopy[b], 0));
    rc = lmm_bed_core(c, buf[b], idx_dev, ni_total, lc, bytes_per_snp, c->dOut.as<gb200_sumstat>() + s0);
    if (rc) return rc;
    GB_CUDA(c, cudaEventRecord(c->evUsed[b], c->stream));
  }
  GB_CUDA(c, cudaMemcpyAsync(out, c->dOut.p, l * sizeof(gb200_sumstat), cudaMemcpyDeviceToHost, c->stream));
  GB_CUDA(c, cudaStreamSynchronize(c->stream));
  return GB200_OK;
}

int gb200_lmm_batch_bed_dev(gb200_ctx *c, const unsigned char *bed_dev, const unsigned char *idv_mask_dev,
                            size_t ni_total, size_t l, size_t bytes_per_snp, gb200_sumstat *out_dev) {
  if (!c) return GB200_ERR_ARG;
  int rc = lmm_check_ready(c, "gb200_lmm_batch_bed_dev");
  if (rc) return rc;
  if (l == 0) return GB200_OK;
  if (!bed_dev || !out_dev || bytes_per_snp != (ni_total + 3) / 4) return set_err(c, GB200_ERR_ARG, "gb200_lmm_batch_bed_dev: bad argument");
  const int *idx_dev = nullptr;
  if (idv_mask_dev) {
    // the mask lives on the device: bring it back once (ni_total bytes) to build the gather index
    std::vector<unsigned char> m(ni_total);
    GB_CUDA(c, cudaMemcpyAsync(m.data(), idv_mask_dev, ni_total, cudaMemcpyDeviceToHost, c->stream));
    GB_CUDA(c, cudaStreamSynchronize(c->stream));
    rc = upload_idx_from_mask(c, m.data(), ni_total, &idx_dev);
    if (rc) return rc;
  } else if (ni_total != c->n) {
    return set_err(c, GB200_ERR_ARG, "idv_mask == NULL requires ni_total == n");
  }
  return lmm_bed_chunked_dev(c, bed_dev, idx_dev, ni_total, l, bytes_per_snp, out_dev);
}

int gb200_lmm_project_bed(gb200_ctx *c, const unsigned char *bed, const unsigned char *idv_mask, size_t ni_total,
                          size_t l, size_t bytes_per_snp, double *UtXt) {
  if (!c) return GB200_ERR_ARG;
  if (!c->lmm_ready) return set_err(c, GB200_ERR_STATE, "gb200_lmm_project_bed before gb200_lmm_setup");
  if (l == 0) return GB200_OK;
  if (!bed || !UtXt || bytes_per_snp != (ni_total + 3) / 4) return set_err(c, GB200_ERR_ARG, "gb200_lmm_project_bed: bad argument");
  const int *idx_dev = nullptr;
  int rc = upload_idx_from_mask(c, idv_mask, ni_total, &idx_dev);
  if (rc) return rc;
  GB_CUDA(c, c->dBed.reserve(l * bytes_per_snp));
  GB_CUDA(c, cudaMemcpyAsync(c->dBed.p, bed, l * bytes_per_snp, cudaMemcpyHostToDevice, c->stream));
  rc = project_bed_dev(c, c->dBed.as<unsigned char>(), idx_dev, ni_total, l, bytes_per_snp);
  if (rc) return rc;
  GB_CUDA(c, cudaMemcpy2DAsync(UtXt, c->n * 8, c->dUtXt.p, c->n_c * 8, c->n * 8, l, cudaMemcpyDeviceToHost, c->stream));
  GB_CUDA(c, cudaStreamSynchronize(c->stream));
  return GB200_OK;
}

int gb200_qc_bed(gb200_ctx *c, const unsigned char *bed, const unsigned char *idv_mask, size_t ni_total, size_t l,
                 size_t bytes_per_snp, const double *W, const double *WtWi, size_t n_cvt, gb200_snpqc *out) {
  if (!c) return GB200_ERR_ARG;
  if (l == 0) return GB200_OK;
  if (!bed || !out || bytes_per_snp != (ni_total + 3) / 4 || (W && (!WtWi || n_cvt == 0 || n_cvt > GB200_MAX_CVT)))
    return set_err(c, GB200_ERR_ARG, "gb200_qc_bed: bad argument");
  // analysed-individual gather index (independent of any lmm_setup state)
  std::vector<int> idx;
  size_t n_test = ni_total;
  if (idv_mask) { for (size_t i = 0; i < ni_total; ++i) if (idv_mask[i]) idx.push_back((int)i); n_test = idx.size(); }
  if (n_test == 0) return set_err(c, GB200_ERR_ARG, "gb200_qc_bed: no analysed individuals");
  DevBuf dIdx, dW, dOutQ;
  auto fail = [&](cudaError_t e, const char *what) {
    dIdx.release(); dW.release(); dOutQ.release();
    return set_err(c, GB200_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
  };
  cudaError_t e;
  if (idv_mask) {
    if ((e = dIdx.reserve(n_test * sizeof(int))) != cudaSuccess) return fail(e, "alloc idx");
    if ((e = cudaMemcpyAsync(dIdx.p, idx.data(), n_test * sizeof(int), cudaMemcpyHostToDevice, c->stream)) != cudaSuccess) return fail(e, "copy idx");
  }
  const double *dWp = nullptr, *dWi = nullptr;
  if (W) {
    if ((e = dW.reserve((n_test * n_cvt + n_cvt * n_cvt) * sizeof(double))) != cudaSuccess) return fail(e, "alloc W");
    if ((e = cudaMemcpyAsync(dW.p, W, n_test * n_cvt * 8, cudaMemcpyHostToDevice, c->stream)) != cudaSuccess) return fail(e, "copy W");
    if ((e = cudaMemcpyAsync(dW.as<double>() + n_test * n_cvt, WtWi, n_cvt * n_cvt * 8, cudaMemcpyHostToDevice, c->stream)) != cudaSuccess) return fail(e, "copy WtWi");
    dWp = dW.as<double>(); dWi = dWp + n_test * n_cvt;
  }
  const size_t chunk = 1 << 16;
  if ((e = dOutQ.reserve(chunk * sizeof(gb200_snpqc))) != cudaSuccess) return fail(e, "alloc out");
  for (size_t s0 = 0; s0 < l; s0 += chunk) {
    const size_t lc = (l - s0 < chunk) ? (l - s0) : chunk;
    if ((e = c->dBed.reserve(lc * bytes_per_snp)) != cudaSuccess) return fail(e, "alloc bed");
    if ((e = cudaMemcpyAsync(c->dBed.p, bed + s0 * bytes_per_snp, lc * bytes_per_snp, cudaMemcpyHostToDevice, c->stream)) != cudaSuccess) return fail(e, "copy bed");
    {
      ProfScope ps(c, "decode");
      if ((e = launch_qc_bed(c->dBed.as<unsigned char>(), lc, bytes_per_snp, idv_mask ? dIdx.as<int>() : nullptr, (int)n_test, dWp, dWi,
                             (int)n_cvt, dOutQ.as<gb200_snpqc>(), c->stream)) != cudaSuccess) return fail(e, "qc kernel");
    }
    if ((e = cudaMemcpyAsync(out + s0, dOutQ.p, lc * sizeof(gb200_snpqc), cudaMemcpyDeviceToHost, c->stream)) != cudaSuccess) return fail(e, "copy out");
    if ((e = cudaStreamSynchronize(c->stream)) != cudaSuccess) return fail(e, "sync");
  }
  dIdx.release(); dW.release(); dOutQ.release();
  return GB200_OK;
}

}  // extern "C"
