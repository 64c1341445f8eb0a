/*
 * gemma_b200.h -- C ABI of libgemma_b200.so, the B200-native replacement for the
 * compute seams of GEMMA's -gk / -eigen / -lmm path.
 *
 * The reference (genetics-statistics/GEMMA) has no plugin/FFI layer: it is one C++
 * binary whose hot path sits behind four internal seams taking gsl_matrix* /
 * gsl_vector*.  Each entry point below names the seam (reference file:line,
 * relative to the GEMMA source tree) it replaces.  INTEGRATION.md shows the
 * binding a GEMMA maintainer would add at each seam.
 *
 * Conventions
 *   - plain C, no exceptions across the boundary; every call returns 0 on success
 *     or a non-zero gb200_status, and gb200_last_error(ctx) describes the failure;
 *   - matrices are row-major FP64 with an explicit leading dimension (gsl_matrix's
 *     `tda`), caller-owned HOST buffers unless the name ends in _dev;
 *   - one context per GPU and per host thread (the reference is single-threaded
 *     and non-reentrant; a context is thread-compatible, not thread-safe);
 *   - all work of a context is issued on ONE CUDA stream (gb200_stream) so that
 *     callers can bracket it with their own events.
 */
#ifndef GEMMA_B200_H
#define GEMMA_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GB200_ABI_VERSION 1

#if defined(__GNUC__)
#define GB200_API __attribute__((visibility("default")))
#else
#define GB200_API
#endif

typedef enum {
  GB200_OK = 0,
  GB200_ERR_ARG = 1,        /* bad argument / shape mismatch ("Range error in dgemm", fastblas.cpp:207) */
  GB200_ERR_CUDA = 2,       /* CUDA runtime / driver error                                            */
  GB200_ERR_STATE = 3,      /* call out of order (e.g. lmm_batch before lmm_setup)                    */
  GB200_ERR_UNSUPPORTED = 4,/* e.g. n_cvt larger than the compiled maximum                            */
  GB200_ERR_NUMERIC = 5,    /* eigensolver did not converge                                           */
  GB200_ERR_NOMEM = 6
} gb200_status;

typedef struct gb200_ctx gb200_ctx;

/* SUMSTAT, src/param.h:54-66: what LMM::Analyze pushes per SNP (src/lmm.cpp:1559-1561). */
typedef struct {
  double beta, se, lambda_remle, lambda_mle, p_wald, p_lrt, p_score, logl_H1;
} gb200_sumstat;

/* Null-model results: what BatchRun computes once per run with CalcLambda('L'/'R'),
 * CalcLmmVgVeBeta and CalcPve (src/gemma.cpp:2711-2753, src/lmm.cpp:2143-2281). */
typedef struct {
  double l_mle_null, logl_mle_H0;
  double l_remle_null, logl_remle_H0;
  double pve_null, pve_se_null;
  double vg_mle, ve_mle, vg_remle, ve_remle;
} gb200_nullmodel;

#define GB200_MAX_CVT 32     /* covariates incl. intercept (1..3 lockstep pipeline kernel, 4..6 register kernel, 7..32 shared-memory-table kernel) */

/* ---- context ---------------------------------------------------------------- */
GB200_API int gb200_abi_version(void);
/* device < 0: use the current CUDA device.  stream == NULL: the context creates its own
 * non-blocking stream; otherwise `stream` is a cudaStream_t owned by the caller. */
GB200_API int gb200_create(gb200_ctx **out, int device, void *stream);
GB200_API void gb200_destroy(gb200_ctx *ctx);
GB200_API const char *gb200_last_error(const gb200_ctx *ctx);
GB200_API void *gb200_stream(const gb200_ctx *ctx);           /* cudaStream_t */
GB200_API int gb200_synchronize(gb200_ctx *ctx);

/* Per-kernel device timing (CUDA events on the context stream).  Names: "utx", "lmm",
 * "kin", "decode", "eigh", "dgemm".  Returns accumulated milliseconds and launch count
 * since the last gb200_profile_reset.  Used by bench.py for the roofline figures. */
GB200_API int gb200_profile_enable(gb200_ctx *ctx, int on);
GB200_API int gb200_profile_reset(gb200_ctx *ctx);
GB200_API int gb200_profile_get(gb200_ctx *ctx, const char *name, double *ms, long *launches);

/* Measured plain (non-tensor) FP64 FMA rate of this device, TFLOP/s with an FMA counted as 2 flops: independent DFMA chains on every SM
 * for about `seconds` -- the roofline denominator of the per-SNP kernel, which is FP64-issue bound (MEASURED_PEAKS.json has no FP64 figure). */
GB200_API int gb200_measure_fp64_fma(gb200_ctx *ctx, double seconds, double *tflops, double *ms);
/* Work counters of the lockstep per-SNP kernel since the last reset, summed over SNPs:
 * {0: lambda slots of hoisted passes (grid, end points, score test, Chebyshev nodes), 1: exact two-lambda passes with powers 1..2,
 *  2: exact two-lambda passes with powers 1..3, 3: exact passes that also accumulate the log-determinant, 4: unused, 5: SNPs}.
 * counts may be NULL to reset only.  Used by bench.py to state the kernel's executed FP64 flops. */
GB200_API int gb200_lmm_counters(gb200_ctx *ctx, unsigned long long counts[6], int reset);

/* The device restatement of the two GSL tails the path calls: out[i] = gsl_cdf_fdist_Q(x[i], nu1, nu2[i]) (call sites
 * src/lmm.cpp:1161,1206; GSL cdf/fdist.c + cdf/beta_inc.c incl. the asymptotic branches for nu/2 > 1e5) when kind == 0,
 * gsl_cdf_chisq_Q(x[i], 1) (src/lmm.cpp:1553) when kind == 1.  Host arrays of `count` doubles; nu2 may be NULL for kind 1.
 * Diagnostic entry point used by the parity tests. */
GB200_API int gb200_cdf_tails(gb200_ctx *ctx, int kind, const double *x, double nu1, const double *nu2, double *out, size_t count);

/* ---- dense GEMM seam -------------------------------------------------------- */
/* fast_dgemm / fast_eigen_dgemm (src/fastblas.h:36-41, src/fastblas.cpp:175-236):
 * C = alpha*op(A)*op(B) + beta*C on row-major matrices with leading dimensions.
 * A is a_rows x a_cols as stored, op selected by TransA/TransB ("N"/"T", first char).
 * Shape mismatch or an empty dimension returns GB200_ERR_ARG (the reference enforces
 * M,N,K > 0 at fastblas.cpp:193-195 and aborts on mismatch at :207). */
GB200_API int gb200_dgemm(gb200_ctx *ctx, const char *TransA, const char *TransB, double alpha,
                const double *A, size_t a_rows, size_t a_cols, size_t lda,
                const double *B, size_t b_rows, size_t b_cols, size_t ldb,
                double beta, double *C, size_t c_rows, size_t c_cols, size_t ldc);

/* ---- -gk : kinship ---------------------------------------------------------- */
/* PARAM::CalcKin -> BimbamKin / PlinkKin (src/param.cpp:1300-1321,
 * src/gemma_io.cpp:1418-1597, :1599-1738).  begin zeroes a device-resident n x n
 * accumulator; each add folds one batch of SNPs (K += Xb Xb^T, the fast_eigen_dgemm
 * of gemma_io.cpp:1554/:1711); finish scales by 1/ns_used (:1570/:1722) and copies K
 * (full symmetric matrix) to the host.  k_mode 1 = centred (-gk 1), 2 = standardised. */
GB200_API int gb200_kin_begin(gb200_ctx *ctx, size_t n, int k_mode);
/* Xb: n x l row-major, columns ALREADY centred/scaled by the caller (Xlarge as the
 * reference holds it at gemma_io.cpp:1546-1554). */
GB200_API int gb200_kin_add(gb200_ctx *ctx, const double *Xb, size_t n, size_t l, size_t ldx);
/* G: l x n SNP-major raw genotypes/dosages, NaN = missing.  The per-SNP transform of
 * gemma_io.cpp:1511-1538 (mean over non-missing, variance, mean imputation, centring,
 * optional 1/sqrt(var)) runs on the device. */
GB200_API int gb200_kin_add_geno(gb200_ctx *ctx, const double *G, size_t l, size_t n, size_t ldg);
/* bed: l rows of PLINK SNP-major 2-bit genotypes, bytes_per_snp = ceil(n/4), decoded as
 * gemma_io.cpp:1665-1682 (00->2, 10->1 [low bit 0, high bit 1], 11->0, 01->missing). */
GB200_API int gb200_kin_add_bed(gb200_ctx *ctx, const unsigned char *bed, size_t l, size_t bytes_per_snp);
GB200_API int gb200_kin_add_bed_dev(gb200_ctx *ctx, const unsigned char *bed_dev, size_t l, size_t bytes_per_snp);
GB200_API int gb200_kin_finish(gb200_ctx *ctx, double *K, size_t ldk, size_t *ns_used);
/* Same, but leaves K on the device and returns the device pointer (n x n, ld n). */
GB200_API int gb200_kin_finish_dev(gb200_ctx *ctx, double **K_dev, size_t *ns_used);

/* ---- eigendecomposition ----------------------------------------------------- */
/* CenterMatrix (src/mathfunc.cpp:147-177, applied when center != 0) followed by
 * EigenDecomp_Zeroed (src/lapack.cpp:260-291 -> dsyevr): all eigenpairs of the
 * symmetric G; eigenvalues ascending, those < 1e-10 set to 0; U row-major with
 * eigenvectors in COLUMNS (lapack.cpp:228); *trace_G = mean(eval).  G is destroyed
 * (as in the reference).  n_zero / n_negative (optional) feed the reference's warnings
 * (lapack.cpp:278-289). */
GB200_API int gb200_eigh(gb200_ctx *ctx, double *G, size_t n, size_t ldg, int center,
               double *U, size_t ldu, double *eval, double *trace_G,
               int *n_zero, int *n_negative);

/* Same with DEVICE buffers (e.g. the K gb200_kin_finish_dev left on the device): G_dev (n x n, ld n) is destroyed; U_dev (n x n, ld n)
 * and eval_dev (n) are caller-owned device buffers distinct from G_dev.  At n = 50 000: 20 GB each for G and U plus cuSOLVER's
 * workspace (option "eigh_workspace_bytes" reports it after the call). */
GB200_API int gb200_eigh_dev(gb200_ctx *ctx, double *G_dev, size_t n, int center, double *U_dev, double *eval_dev,
                   double *trace_G, int *n_zero, int *n_negative);

/* ---- SNP QC statistics for PLINK input ------------------------------------- */
/* The per-SNP counting pass of ReadFile_bed (src/gemma_io.cpp:951-1005) and the covariate-correlation terms of its
 * -r2 filter (:1031-1046) on the device: for each of the l SNP rows, over the analysed individuals (idv_mask),
 * n_miss / genotype class counts / maf = sum/(2(n - n_miss)); when W != NULL also v_x = x'x and
 * v_w = (W'x)' WtWi (W'x) with missing genotypes imputed as 2*maf.  The caller applies the thresholds in the
 * reference's order (miss, maf, polymorphism, hwe, r2).  W: n_test x n_cvt row-major; WtWi: n_cvt x n_cvt. */
typedef struct { int n_miss, n_0, n_1, n_2; double maf, v_x, v_w; } gb200_snpqc;
GB200_API int gb200_qc_bed(gb200_ctx *ctx, const unsigned char *bed, const unsigned char *idv_mask, size_t ni_total,
                 size_t l, size_t bytes_per_snp, const double *W, const double *WtWi, size_t n_cvt,
                 gb200_snpqc *out);

/* ---- -lmm : per-run setup, null model, per-batch association ---------------- */
/* Uploads the run-constant state of LMM::Analyze (src/lmm.cpp:1474-1511): U (n x n,
 * eigenvectors in columns), eval, W (n x n_cvt) and y; computes UtW = U^T W and
 * Uty = U^T y on the device (CalcUtX, src/mathfunc.cpp:497-510, called at
 * src/gemma.cpp:2699-2700).  UtW_out (n x n_cvt, ld n_cvt) / Uty_out may be NULL. */
GB200_API int gb200_lmm_setup(gb200_ctx *ctx, size_t n, size_t n_cvt,
                    const double *U, size_t ldu, const double *eval,
                    const double *W, size_t ldw, const double *y,
                    double *UtW_out, double *Uty_out);
/* Variant for callers that already hold UtW / Uty (e.g. -d/-u input without W,y). */
GB200_API int gb200_lmm_setup_rotated(gb200_ctx *ctx, size_t n, size_t n_cvt,
                            const double *U, size_t ldu, const double *eval,
                            const double *UtW, size_t ldw, const double *Uty);
/* Same with DEVICE pointers: U_dev (n x n row-major, ld n) is BORROWED (not copied; the caller
 * keeps it alive while the context uses it), UtWt_dev is U^T W stored TRANSPOSED (n_cvt x n). */
GB200_API int gb200_lmm_setup_rotated_dev(gb200_ctx *ctx, size_t n, size_t n_cvt, const double *U_dev,
                                const double *eval_dev, const double *UtWt_dev, const double *Uty_dev);
/* Null model on the device with the same fused evaluator (src/gemma.cpp:2711-2753):
 * also returns beta / se(beta) of the covariates (n_cvt values each) for MLE and REMLE. */
GB200_API int gb200_lmm_null(gb200_ctx *ctx, double l_min, double l_max, size_t n_region, double trace_G,
                   gb200_nullmodel *out, double *beta_mle, double *se_beta_mle,
                   double *beta_remle, double *se_beta_remle);
/* Parameters of the per-SNP tests: a_mode 1 Wald, 2 LRT, 3 score, 4 all, 9 (src/lmm.cpp:1541-1554). */
GB200_API int gb200_lmm_params(gb200_ctx *ctx, int a_mode, double l_min, double l_max, size_t n_region,
                     double l_mle_null, double logl_mle_H0);
/* batch_compute(l) of src/lmm.cpp:1513-1564.  Xb: n x l row-major (ld ldx) raw genotype
 * columns of the analysed individuals, missing values already mean-imputed by the caller
 * (src/lmm.cpp:1611-1618).  l == 0 is a no-op (the reference aborts there, fastblas.cpp:193). */
GB200_API int gb200_lmm_batch(gb200_ctx *ctx, const double *Xb, size_t l, size_t ldx, gb200_sumstat *out);
/* G: l x n SNP-major genotypes/dosages of the analysed individuals, NaN = missing; the
 * mean imputation of src/lmm.cpp:1590-1618 runs on the device. */
GB200_API int gb200_lmm_batch_geno(gb200_ctx *ctx, const double *G, size_t l, size_t ldg, gb200_sumstat *out);
/* bed: l PLINK 2-bit rows over ni_total individuals (bytes_per_snp = ceil(ni_total/4));
 * idv_mask (ni_total bytes, may be NULL = all analysed) is indicator_idv; decode, drop,
 * mean-impute as src/lmm.cpp:1783-1829 on the device. */
GB200_API int gb200_lmm_batch_bed(gb200_ctx *ctx, const unsigned char *bed, const unsigned char *idv_mask,
                        size_t ni_total, size_t l, size_t bytes_per_snp, gb200_sumstat *out);
/* Device-resident input and output (bench `value` leg: inputs already in HBM). */
GB200_API int gb200_lmm_batch_bed_dev(gb200_ctx *ctx, const unsigned char *bed_dev,
                            const unsigned char *idv_mask_dev, size_t ni_total, size_t l,
                            size_t bytes_per_snp, gb200_sumstat *out_dev);

/* Only the per-SNP association kernel on an already rotated batch: UtXt is l x n
 * (SNP-major, U^T x contiguous per SNP, ld ldu).  Exposed for kernel-level parity tests. */
GB200_API int gb200_lmm_assoc_utx(gb200_ctx *ctx, const double *UtXt, size_t l, size_t ldu, gb200_sumstat *out);

/* Projection only: UtXt (l x n, ld n, host) = (U^T Xb)^T for a host batch Xb (n x l). */
GB200_API int gb200_lmm_project(gb200_ctx *ctx, const double *Xb, size_t l, size_t ldx, double *UtXt);

/* ---- G x E: LMM::AnalyzePlinkGXE / AnalyzeBimbamGXE, src/lmm.cpp:2283-2608; wiring src/gemma.cpp:2580-2582, 2809-2828.
 * After gb200_lmm_setup (and, before the batches, gb200_lmm_params): env = the environmental variable of the n analysed individuals (PARAM::CopyGxe,
 * src/param.cpp:2116-2128).  Per SNP the covariates are [W, env, x] (x mean-imputed, flipped to 2 - x when its mean exceeds 1) and
 * the tested variable is x * env; -lmm 2/4 compare with the per-SNP null that contains x (calc_null with n_cvt + 2 covariates);
 * beta changes sign for flipped SNPs.  No NaN rule (the reference's GXE loops have none).  n_cvt + 2 <= GB200_MAX_CVT. */
GB200_API int gb200_lmm_gxe_setup(gb200_ctx *ctx, const double *env);
/* G: l x n SNP-major dosages of the analysed individuals, NaN = missing (BIMBAM); same layout as gb200_lmm_batch_geno */
GB200_API int gb200_lmm_gxe_batch_geno(gb200_ctx *ctx, const double *G, size_t l, size_t ldg, gb200_sumstat *out);
/* raw PLINK rows; same arguments as gb200_lmm_batch_bed */
GB200_API int gb200_lmm_gxe_batch_bed(gb200_ctx *ctx, const unsigned char *bed, const unsigned char *idv_mask, size_t ni_total,
                            size_t l, size_t bytes_per_snp, gb200_sumstat *out);

/* ---- -lm: linear model without random effect, LM::AnalyzeBimbam / AnalyzePlink + CalcvPv + LmCalcP
 * (src/lm.cpp:224-288, 382-640).  W: n x n_cvt (ld ldw, intercept included), y: n, both of the analysed individuals.
 * a_mode 1..4 or 51..54 (Wald / LRT / score / all).  SUMSTAT like the reference's: {beta, se, 0, 0, p_wald, p_lrt, p_score, -0}. */
GB200_API int gb200_lm_setup(gb200_ctx *ctx, size_t n, size_t n_cvt, const double *W, size_t ldw, const double *y);
GB200_API int gb200_lm_batch_geno(gb200_ctx *ctx, const double *G, size_t l, size_t ldg, int a_mode, gb200_sumstat *out);
GB200_API int gb200_lm_batch_bed(gb200_ctx *ctx, const unsigned char *bed, const unsigned char *idv_mask, size_t ni_total, size_t l,
                       size_t bytes_per_snp, int a_mode, gb200_sumstat *out);

/* ---- multivariate LMM (BASELINE config 5): MVLMM::AnalyzeBimbam / AnalyzePlink, src/mvlmm.cpp:2972-3899, -lmm 1/2/3/4.
 * Two phenotypes, 1..3 covariates in this round.  Y: n x n_ph (ld ldy) of the analysed individuals.  setup rotates W and Y, and runs the
 * univariate REML fits of MphInitial (:2786-2796; with the l_min / l_max / n_region of a preceding gb200_lmm_params call, else 1e-5 / 1e5 / 10;
 * on return the univariate state of the context is that of gb200_lmm_setup with the FIRST phenotype); null = EM + Newton-Raphson for REML then ML (:3047-3133; the per-SNP fits start from the
 * ML estimates, :3205-3207); batch = per SNP: REML EM (em_iter/10, em_prec*10), MphCalcP, Newton-Raphson refinement when p < 0.001
 * (:3334-3347); a_mode 2 / 3 / 4 add the likelihood-ratio (ML EM + NR, :3310-3332) and score (:3297-3307) branches.  out: l rows of 8 doubles
 * {beta_1, beta_2, Vbeta_11, Vbeta_12, Vbeta_22, p_wald, p_lrt, p_score} (the columns of MVLMM::WriteFiles, :117-210; tests not run stay 0).
 * V_g / V_e outputs are 2 x 2 row-major, B is 2 x n_cvt. */
GB200_API int gb200_mvlmm_setup(gb200_ctx *ctx, size_t n, size_t n_cvt, size_t n_ph, const double *U, size_t ldu, const double *eval,
                      const double *W, size_t ldw, const double *Y, size_t ldy);
GB200_API int gb200_mvlmm_null(gb200_ctx *ctx, double *Vg_remle, double *Ve_remle, double *B_remle, double *logl_remle, double *Vg_mle,
                     double *Ve_mle, double *B_mle, double *logl_mle);
GB200_API int gb200_mvlmm_batch_geno(gb200_ctx *ctx, const double *G, size_t l, size_t ldg, int a_mode, double *out);
GB200_API int gb200_mvlmm_batch_bed(gb200_ctx *ctx, const unsigned char *bed, const unsigned char *idv_mask, size_t ni_total, size_t l,
                          size_t bytes_per_snp, int a_mode, double *out);

/* Projection only for a PLINK 2-bit batch (same decode / imputation as gb200_lmm_batch_bed),
 * through whichever projection path the options select.  UtXt: l x n host buffer. */
GB200_API int gb200_lmm_project_bed(gb200_ctx *ctx, const unsigned char *bed, const unsigned char *idv_mask,
                          size_t ni_total, size_t l, size_t bytes_per_snp, double *UtXt);

/* Tuning knobs (0 keeps the default): utx_path 0 auto, 1 FP64 tiled, 2 int8 tensor-core
 * (exact int8 digit planes of U, integer genotypes only); n_slices = digit planes of that path (0 = smallest count whose
 * truncation noise sqrt(n) 2^-(6+8(T-1)) stays below 2^-30: 5 up to n = 65 536, else 6); lmm_kernel 0 auto,
 * 1 warp-per-SNP register kernel (n_cvt <= 6), 2 lockstep-CTA pipeline kernel (n_cvt <= 3, n_region <= 64), 3 the
 * any-covariate-count kernel (default for n_cvt >= 7); lmm_hoist 0|1: lockstep kernel with the SNP-independent sums at the
 * lambdas shared by all SNPs computed once per run (default 1); kin_miss_max_permille: chunks with a larger share of missing
 * genotypes take the dense FP64 kinship path (default 200); cta_pair /
 * kin_cta_pair 0|1 run the projection / kinship tensor-core kernel as CTA pairs (cta_group::2);
 * kin_path 0 auto, 1 FP64 only; overlap 0|1 pipelines 2048-SNP sub-batches of the bed entry points on two streams
 * (projection of sub-batch i+1 || tests of sub-batch i; measured slower on B200, default 0); batch_chunk: SNPs per internal
 * sub-batch of the bed entry points (0 = auto: the FP64 U^T X staging buffer stays near 4 GB); stage_mask 1|2|3: measurement
 * runs only -- the bed entry points run just the projection (1) or just the tests on the last projection (2). */
GB200_API int gb200_set_option(gb200_ctx *ctx, const char *name, long value);
/* Current value of a knob; "n_slices" returns the EFFECTIVE plane count for the individuals of the last gb200_lmm_setup*. */
GB200_API int gb200_get_option(gb200_ctx *ctx, const char *name, long *value);

#ifdef __cplusplus
}
#endif
#endif /* GEMMA_B200_H */
