// ref_wrap.cpp -- TEST INFRASTRUCTURE ONLY.  C entry points around the REFERENCE's own functions, compiled together with
// the reference sources (in place under /root/reference/src, never copied) into oracle/_ref/libgemma_ref.so by
// oracle/Makefile.  Used by tests/test_oracle_vs_ref.py to validate the restated oracle (oracle/gemma_oracle.c, refpipe.py)
// against the code it restates, on arbitrary inputs.  GSL is replaced by the API shim in oracle/gsl_shim/.
#include <cmath>
#include <cstring>
#include <functional>
#include <iostream>
#include <map>
#include <set>
#include <sstream>
#include <string>
#include <vector>

#include "gsl/gsl_matrix.h"
#include "gsl/gsl_vector.h"
#include "gemma_io.h"
#include "lmm.h"
#include "mathfunc.h"
#include "param.h"

using namespace std;

// defined (non-static, undeclared in lmm.h) in src/lmm.cpp:484-1125
double LogL_f(double l, void *params);
double LogL_dev1(double l, void *params);
double LogL_dev2(double l, void *params);
double LogRL_f(double l, void *params);
double LogRL_dev1(double l, void *params);
double LogRL_dev2(double l, void *params);
void CalcUab(const gsl_matrix *UtW, const gsl_vector *Uty, gsl_matrix *Uab);
void CalcUab(const gsl_matrix *UtW, const gsl_vector *Uty, const gsl_vector *Utx, gsl_matrix *Uab);

namespace {
struct Quiet {          // the reference prints progress bars and summaries
  streambuf *o, *e; stringstream sink;
  Quiet() { o = cout.rdbuf(sink.rdbuf()); e = cerr.rdbuf(sink.rdbuf()); }
  ~Quiet() { cout.rdbuf(o); cerr.rdbuf(e); }
};
gsl_matrix *mat_from(const double *a, size_t r, size_t c, size_t ld) {
  gsl_matrix *m = gsl_matrix_alloc(r, c);
  for (size_t i = 0; i < r; ++i) for (size_t j = 0; j < c; ++j) gsl_matrix_set(m, i, j, a[i * ld + j]);
  return m;
}
gsl_vector *vec_from(const double *a, size_t n) {
  gsl_vector *v = gsl_vector_alloc(n);
  for (size_t i = 0; i < n; ++i) gsl_vector_set(v, i, a[i]);
  return v;
}
}  // namespace

extern "C" {

size_t ref_getab_index(size_t a, size_t b, size_t n_cvt) { return GetabIndex(a, b, n_cvt); }   // src/param.cpp:1400-1415

// value of one of the six likelihood functions (src/lmm.cpp:484-1125) for the alternative model with x (Utx != NULL) or the null
double ref_eval_fn(char fn /* 'L' | 'R' */, int which /* 0 f, 1 dev1, 2 dev2 */, int calc_null, double l, size_t n, size_t n_cvt,
                   const double *eval, const double *UtW, size_t ldw, const double *Uty, const double *Utx) {
  Quiet q;
  const size_t n_index = (n_cvt + 2 + 1) * (n_cvt + 2) / 2;
  gsl_vector *ev = vec_from(eval, n), *y = vec_from(Uty, n), *ab = gsl_vector_alloc(n_index);
  gsl_matrix *W = mat_from(UtW, n, n_cvt, ldw), *Uab = gsl_matrix_alloc(n, n_index);
  gsl_matrix_set_zero(Uab);
  CalcUab(W, y, Uab);
  if (Utx) { gsl_vector *x = vec_from(Utx, n); CalcUab(W, y, x, Uab); gsl_vector_free(x); }
  FUNC_PARAM p = {calc_null != 0, n, n_cvt, ev, Uab, ab, 0};
  double r;
  if (fn == 'L') r = which == 0 ? LogL_f(l, &p) : which == 1 ? LogL_dev1(l, &p) : LogL_dev2(l, &p);
  else r = which == 0 ? LogRL_f(l, &p) : which == 1 ? LogRL_dev1(l, &p) : LogRL_dev2(l, &p);
  gsl_vector_free(ev); gsl_vector_free(y); gsl_vector_free(ab); gsl_matrix_free(W); gsl_matrix_free(Uab);
  return r;
}

// null model: src/gemma.cpp:2711-2753 -> CalcLambda (src/lmm.cpp:2143-2180), CalcLmmVgVeBeta (:2210-2281), CalcPve (:2183-2205)
int ref_null_model(size_t n, size_t n_cvt, const double *eval, const double *UtW, size_t ldw, const double *Uty, double l_min, double l_max,
                   size_t n_region, double trace_G, double *out8 /* l_mle logl_mle l_remle logl_remle pve pve_se vg_remle ve_remle */,
                   double *beta_remle, double *se_beta_remle, double *beta_mle, double *se_beta_mle, double *vgve_mle) {
  Quiet q;
  gsl_vector *ev = vec_from(eval, n), *y = vec_from(Uty, n);
  gsl_matrix *W = mat_from(UtW, n, n_cvt, ldw);
  double l_mle, logl_mle, l_re, logl_re, pve, pve_se, vg, ve;
  CalcLambda('L', ev, W, y, l_min, l_max, n_region, l_mle, logl_mle);
  CalcLambda('R', ev, W, y, l_min, l_max, n_region, l_re, logl_re);
  gsl_vector *b = gsl_vector_alloc(n_cvt), *se = gsl_vector_alloc(n_cvt);
  CalcLmmVgVeBeta(ev, W, y, l_mle, vg, ve, b, se);
  for (size_t i = 0; i < n_cvt; ++i) { beta_mle[i] = gsl_vector_get(b, i); se_beta_mle[i] = gsl_vector_get(se, i); }
  vgve_mle[0] = vg; vgve_mle[1] = ve;
  CalcLmmVgVeBeta(ev, W, y, l_re, vg, ve, b, se);
  for (size_t i = 0; i < n_cvt; ++i) { beta_remle[i] = gsl_vector_get(b, i); se_beta_remle[i] = gsl_vector_get(se, i); }
  CalcPve(ev, W, y, l_re, trace_G, pve, pve_se);
  out8[0] = l_mle; out8[1] = logl_mle; out8[2] = l_re; out8[3] = logl_re; out8[4] = pve; out8[5] = pve_se; out8[6] = vg; out8[7] = ve;
  gsl_vector_free(ev); gsl_vector_free(y); gsl_matrix_free(W); gsl_vector_free(b); gsl_vector_free(se);
  return 0;
}

// LMM::Analyze (src/lmm.cpp:1474-1658) on in-memory genotypes: G is SNP-major l x ni_total with NaN = missing.
// out: l_kept x 8 doubles in SUMSTAT order (src/param.h:54-66).
int ref_lmm_analyze(size_t ni_total, const int *indicator_idv, size_t n, size_t n_cvt, const double *U, const double *eval, const double *UtW,
                    const double *Uty, const double *W, const double *y, const double *G, size_t l, int a_mode, double l_min, double l_max,
                    size_t n_region, double l_mle_null, double logl_mle_H0, double *out) {
  Quiet q;
  LMM c;
  c.a_mode = a_mode; c.d_pace = 100000; c.l_min = l_min; c.l_max = l_max; c.n_region = n_region; c.l_mle_null = l_mle_null;
  c.logl_mle_H0 = logl_mle_H0; c.ni_total = ni_total; c.ni_test = n; c.ns_total = l; c.ns_test = l; c.n_cvt = n_cvt;
  c.time_UtX = 0; c.time_opt = 0;
  c.indicator_idv.assign(indicator_idv, indicator_idv + ni_total);
  c.indicator_snp.assign(l, 1);
  gsl_matrix *mU = mat_from(U, n, n, n), *mUtW = mat_from(UtW, n, n_cvt, n_cvt), *mW = mat_from(W, n, n_cvt, n_cvt);
  gsl_vector *vev = vec_from(eval, n), *vUty = vec_from(Uty, n), *vy = vec_from(y, n);
  std::function<SnpNameValues(size_t)> fetch = [&](size_t t) {
    std::vector<double> gs(G + t * ni_total, G + (t + 1) * ni_total);
    return std::make_tuple(std::string("snp") + std::to_string(t), gs);
  };
  c.Analyze(fetch, mU, vev, mUtW, vUty, mW, vy, std::set<std::string>());
  for (size_t t = 0; t < c.sumStat.size() && t < l; ++t) {
    const SUMSTAT &s = c.sumStat[t];
    double *o = out + 8 * t;
    o[0] = s.beta; o[1] = s.se; o[2] = s.lambda_remle; o[3] = s.lambda_mle; o[4] = s.p_wald; o[5] = s.p_lrt; o[6] = s.p_score; o[7] = s.logl_H1;
  }
  const int produced = (int)c.sumStat.size();
  gsl_matrix_free(mU); gsl_matrix_free(mUtW); gsl_matrix_free(mW); gsl_vector_free(vev); gsl_vector_free(vUty); gsl_vector_free(vy);
  return produced;
}

// The per-SNP part of batch_compute (src/lmm.cpp:1526-1562) on a given U^T X (n x l, SNP per column): the reference's own
// CalcUab / CalcRLScore / CalcLambda / CalcRLWald and the LRT, called exactly as that closure calls them.  Used as the CPU
// baseline (the U^T X product itself is done by the caller with an optimised BLAS, as the reference does through cblas_dgemm).
int ref_assoc_utx(size_t n, size_t n_cvt, const double *eval, const double *UtW, const double *Uty, const double *UtX, size_t l, size_t ldx,
                  int a_mode, double l_min, double l_max, size_t n_region, double l_mle_null, double logl_mle_H0, double *out) {
  Quiet q;
  LMM c;
  c.a_mode = a_mode; c.ni_test = n; c.n_cvt = n_cvt;
  const size_t n_index = (n_cvt + 2 + 1) * (n_cvt + 2) / 2;
  gsl_vector *ev = vec_from(eval, n), *y = vec_from(Uty, n), *ab = gsl_vector_alloc(n_index), *Utx = gsl_vector_alloc(n);
  gsl_matrix *W = mat_from(UtW, n, n_cvt, n_cvt), *Uab = gsl_matrix_alloc(n, n_index);
  gsl_matrix_set_zero(Uab);
  CalcUab(W, y, Uab);
  for (size_t i = 0; i < l; ++i) {
    for (size_t k = 0; k < n; ++k) gsl_vector_set(Utx, k, UtX[k * ldx + i]);
    CalcUab(W, y, Utx, Uab);
    FUNC_PARAM param1 = {false, n, n_cvt, ev, Uab, ab, 0};
    double lambda_mle = 0.0, lambda_remle = 0.0, beta = 0.0, se = 0.0, p_wald = 0.0, p_lrt = 0.0, p_score = 0.0, logl_H1 = 0.0;
    if (a_mode == 3 || a_mode == 4 || a_mode == 9) c.CalcRLScore(l_mle_null, param1, beta, se, p_score);
    if (a_mode == 1 || a_mode == 4) { CalcLambda('R', param1, l_min, l_max, n_region, lambda_remle, logl_H1); c.CalcRLWald(lambda_remle, param1, beta, se, p_wald); }
    if (a_mode == 2 || a_mode == 4 || a_mode == 9) { CalcLambda('L', param1, l_min, l_max, n_region, lambda_mle, logl_H1); p_lrt = gsl_cdf_chisq_Q(2.0 * (logl_H1 - logl_mle_H0), 1); }
    double *o = out + 8 * i;
    o[0] = beta; o[1] = se; o[2] = lambda_remle; o[3] = lambda_mle; o[4] = p_wald; o[5] = p_lrt; o[6] = p_score; o[7] = logl_H1;
  }
  gsl_vector_free(ev); gsl_vector_free(y); gsl_vector_free(ab); gsl_vector_free(Utx); gsl_matrix_free(W); gsl_matrix_free(Uab);
  return 0;
}

// QC pass of ReadFile_geno (src/gemma_io.cpp:639-873) on a BIMBAM file.  W: ni_test x n_cvt.  Returns ns_total; fills indicator_snp (caller
// provides room for `cap` entries) and the per-SNP n_miss / maf.
long ref_qc_bimbam(const char *file_geno, const int *indicator_idv, size_t ni_total, const double *W, size_t ni_test, size_t n_cvt, double maf_level,
                   double miss_level, double hwe_level, double r2_level, int *indicator_snp, long *n_miss, double *maf, size_t cap, long *ns_test_out) {
  Quiet q;
  set<string> setSnps; map<string, string> chr; map<string, long> bp; map<string, double> cM;
  vector<int> idv(indicator_idv, indicator_idv + ni_total), isnp; vector<SNPINFO> info; size_t ns_test = 0;
  gsl_matrix *mW = mat_from(W, ni_test, n_cvt, n_cvt);
  if (!ReadFile_geno(string(file_geno), setSnps, mW, idv, isnp, maf_level, miss_level, hwe_level, r2_level, chr, bp, cM, info, ns_test)) return -1;
  gsl_matrix_free(mW);
  for (size_t t = 0; t < isnp.size() && t < cap; ++t) { indicator_snp[t] = isnp[t]; n_miss[t] = (long)info[t].n_miss; maf[t] = info[t].maf; }
  *ns_test_out = (long)ns_test;
  return (long)isnp.size();
}

// BimbamKin (src/gemma_io.cpp:1418-1597): K (ni_total x ni_total) from the file and the SNP indicator
int ref_bimbam_kin(const char *file_geno, const int *indicator_snp, size_t ns_total, int k_mode, size_t ni_total, double *K) {
  Quiet q;
  vector<int> isnp(indicator_snp, indicator_snp + ns_total);
  gsl_matrix *mK = gsl_matrix_alloc(ni_total, ni_total);
  gsl_matrix_set_zero(mK);
  const bool ok = BimbamKin(string(file_geno), set<string>(), isnp, k_mode, 100000, mK, false);
  for (size_t i = 0; i < ni_total; ++i) for (size_t j = 0; j < ni_total; ++j) K[i * ni_total + j] = gsl_matrix_get(mK, i, j);
  gsl_matrix_free(mK);
  return ok ? 0 : -1;
}

// ---- PLINK: ReadFile_bim + ReadFile_bed QC (src/gemma_io.cpp:514-556, 876-1064), PlinkKin (:1599-1738), LMM::AnalyzePlink (src/lmm.cpp:1710-1903)
long ref_qc_plink(const char *prefix, const int *indicator_idv, size_t ni_total, const double *W, size_t ni_test, size_t n_cvt, double maf_level,
                  double miss_level, double hwe_level, double r2_level, int *indicator_snp, long *n_miss, double *maf, size_t cap, long *ns_test_out) {
  Quiet q;
  vector<SNPINFO> info;
  if (!ReadFile_bim(string(prefix) + ".bim", info)) return -1;
  set<string> setSnps; vector<int> idv(indicator_idv, indicator_idv + ni_total), isnp; size_t ns_test = 0;
  gsl_matrix *mW = mat_from(W, ni_test, n_cvt, n_cvt);
  if (!ReadFile_bed(string(prefix) + ".bed", setSnps, mW, idv, isnp, info, maf_level, miss_level, hwe_level, r2_level, ns_test)) return -1;
  gsl_matrix_free(mW);
  for (size_t t = 0; t < isnp.size() && t < cap; ++t) { indicator_snp[t] = isnp[t]; n_miss[t] = (long)info[t].n_miss; maf[t] = info[t].maf; }
  *ns_test_out = (long)ns_test;
  return (long)isnp.size();
}

int ref_plink_kin(const char *prefix, const int *indicator_snp, size_t ns_total, int k_mode, size_t ni_total, double *K) {
  Quiet q;
  vector<int> isnp(indicator_snp, indicator_snp + ns_total);
  gsl_matrix *mK = gsl_matrix_alloc(ni_total, ni_total);
  gsl_matrix_set_zero(mK);
  const bool ok = PlinkKin(string(prefix) + ".bed", isnp, k_mode, 100000, mK);
  for (size_t i = 0; i < ni_total; ++i) for (size_t j = 0; j < ni_total; ++j) K[i * ni_total + j] = gsl_matrix_get(mK, i, j);
  gsl_matrix_free(mK);
  return ok ? 0 : -1;
}

int ref_lmm_analyze_plink(const char *prefix, size_t ni_total, const int *indicator_idv, const int *indicator_snp, size_t ns_total, size_t n,
                          size_t n_cvt, const double *U, const double *eval, const double *UtW, const double *Uty, const double *W, const double *y,
                          int a_mode, double l_min, double l_max, size_t n_region, double l_mle_null, double logl_mle_H0, double *out, size_t cap) {
  Quiet q;
  LMM c;
  c.a_mode = a_mode; c.d_pace = 100000; c.l_min = l_min; c.l_max = l_max; c.n_region = n_region; c.l_mle_null = l_mle_null;
  c.logl_mle_H0 = logl_mle_H0; c.ni_total = ni_total; c.ni_test = n; c.ns_total = ns_total; c.n_cvt = n_cvt; c.time_UtX = 0; c.time_opt = 0;
  c.file_bfile = prefix;
  c.indicator_idv.assign(indicator_idv, indicator_idv + ni_total);
  c.indicator_snp.assign(indicator_snp, indicator_snp + ns_total);
  if (!ReadFile_bim(string(prefix) + ".bim", c.snpInfo)) return -1;
  c.ns_test = 0; for (size_t t = 0; t < ns_total; ++t) c.ns_test += indicator_snp[t];
  gsl_matrix *mU = mat_from(U, n, n, n), *mUtW = mat_from(UtW, n, n_cvt, n_cvt), *mW = mat_from(W, n, n_cvt, n_cvt);
  gsl_vector *vev = vec_from(eval, n), *vUty = vec_from(Uty, n), *vy = vec_from(y, n);
  c.AnalyzePlink(mU, vev, mUtW, vUty, mW, vy, std::set<std::string>());
  const size_t got = c.sumStat.size();
  for (size_t t = 0; t < got && t < cap; ++t) {
    const SUMSTAT &s = c.sumStat[t];
    double *o = out + 8 * t;
    o[0] = s.beta; o[1] = s.se; o[2] = s.lambda_remle; o[3] = s.lambda_mle; o[4] = s.p_wald; o[5] = s.p_lrt; o[6] = s.p_score; o[7] = s.logl_H1;
  }
  gsl_matrix_free(mU); gsl_matrix_free(mUtW); gsl_matrix_free(mW); gsl_vector_free(vev); gsl_vector_free(vUty); gsl_vector_free(vy);
  return (int)got;
}

// CenterMatrix (src/mathfunc.cpp:147-177), in place
void ref_center_matrix(double *G, size_t n) {
  gsl_matrix_view v = gsl_matrix_view_array(G, n, n);
  CenterMatrix(&v.matrix);
}

// ReadFile_kin (src/gemma_io.cpp:1186-1294): -km 1 matrix or -km 2 "id1 id2 value" list -> G (ni_test x ni_test); ids (or NULL) fill
// mapID2num the way ReadFile_fam does (individual id -> row of the .fam file)
int ref_read_kin(const char *file_kin, const int *indicator_idv, size_t ni_total, const char *const *ids, int k_mode, double *G, size_t ni_test) {
  Quiet q;
  vector<int> idv(indicator_idv, indicator_idv + ni_total);
  map<string, int> id2num;
  if (ids) for (size_t i = 0; i < ni_total; ++i) id2num[string(ids[i])] = (int)i;
  gsl_matrix *mG = gsl_matrix_alloc(ni_test, ni_test);
  bool error = false;
  ReadFile_kin(string(file_kin), idv, id2num, (size_t)k_mode, error, mG);
  for (size_t i = 0; i < ni_test; ++i) for (size_t j = 0; j < ni_test; ++j) G[i * ni_test + j] = gsl_matrix_get(mG, i, j);
  gsl_matrix_free(mG);
  return error ? -1 : 0;
}

}  // extern "C"
