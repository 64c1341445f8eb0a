// gemma_cli.cpp -- `gemma-b200`: GEMMA's -gk / -eigen / -lmm command line on top of
// libgemma_b200.so.  Host side only: flag parsing, BIMBAM / PLINK readers, SNP QC, text writers.
// All numerics go through the C ABI (include/gemma_b200.h); there is no CPU compute path here.
//
// Drop-in surface re-created (reference file:line, GEMMA tree):
//   flags            GEMMA::Assign            src/gemma.cpp:754-1639 (the -gk/-eigen/-lmm subset)
//   phenotypes       ReadFile_pheno           src/gemma_io.cpp:386-444
//   covariates       ReadFile_cvt, CheckCvt   src/gemma_io.cpp:446-511, src/param.cpp:1937-1990
//   annotation       ReadFile_anno            src/gemma_io.cpp:280-341
//   individuals      ProcessCvtPhen           src/param.cpp:1993-2098
//   BIMBAM QC pass   ReadFile_geno            src/gemma_io.cpp:639-873
//   PLINK            ReadFile_bim/fam/bed     src/gemma_io.cpp:514-636, 876-1064
//   kinship          CalcKin/BimbamKin/PlinkKin src/param.cpp:1300, src/gemma_io.cpp:1418-1738
//   K / U / D files  WriteMatrix/WriteVector, ReadFile_kin/eigenU/eigenD
//                                             src/param.cpp:1886-1935, src/gemma_io.cpp:1186-1415
//   LMM driver       BatchRun LMM branch, AnalyzeBimbam/AnalyzePlink src/gemma.cpp:2556-2871, src/lmm.cpp:1474-1903
//   association file LMM::WriteFiles          src/lmm.cpp:101-225
//   summary / log    CheckData, PrintSummary, WriteLog src/param.cpp:1108-1134,1252-1259, src/gemma.cpp:3148-3597
#include <zlib.h>
#include <sys/stat.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <map>
#include <set>
#include <sstream>
#include <string>
#include <vector>
#include <chrono>
#include <thread>

#include "../../include/gemma_b200.h"
#include "line_pipeline.h"

using std::string;
using std::vector;

static const size_t BATCH = 20000;   // LMM_BATCH_SIZE / K_BATCH_SIZE (src/lmm.h:33, src/param.h:32)

struct SnpInfo {   // SNPINFO, src/param.h:37-51
  string chr, rs; double cM; long bp; string a_minor, a_major;
  long n_miss; double missingness, maf; long n_idv;
};

struct Params {
  string file_geno, file_pheno, file_anno, file_cvt, file_bfile, file_kin, file_ku, file_kd, file_snps, file_ksnps, file_gwasnps, loco, file_gxe, file_weight;
  string path_out = "./output/", file_out = "result";
  vector<size_t> p_column;
  int a_mode = 0;            // 21/22 -gk, 31 -eigen, 1/2/3/4/9 -lmm
  int k_mode = 1;            // -km
  double miss_level = 0.05, maf_level = 0.01, hwe_level = 0.0, r2_level = 0.9999;
  double l_min = 1e-5, l_max = 1e5; size_t n_region = 10;
  long nind = -1;
  bool silence = false, qc_only = false, bin = false;
  int device = -1;
};

// may be called from a worker thread of the line pipeline while the others are still running: flush and leave without
// running static destructors under them
static void die(const string &msg) { std::cout << "error! " << msg << std::endl; std::cout.flush(); std::fflush(nullptr); std::_Exit(1); }

// ---- line readers (plain or gzip, like gzstream's igzstream) --------------------------------------
struct LineReader {
  gzFile f = nullptr; string buf;
  explicit LineReader(const string &path) { f = gzopen(path.c_str(), "rb"); if (f) gzbuffer(f, 1 << 20); }
  ~LineReader() { if (f) gzclose(f); }
  bool ok() const { return f != nullptr; }
  bool next(string &line) {
    line.clear();
    char tmp[65536];
    bool got = false;
    while (gzgets(f, tmp, sizeof(tmp))) {
      got = true;
      size_t len = strlen(tmp);
      if (len && tmp[len - 1] == '\n') { line.append(tmp, len - 1); break; }
      line.append(tmp, len);
    }
    if (!line.empty() && line.back() == '\r') line.pop_back();
    return got;
  }
};

static inline char *tok(char *s, const char *delim = " ,\t") { return strtok(s, delim); }

// ---- small dense helpers for the r2 filter (host, c x c) -------------------------------------------
static bool invert_small(vector<double> A, size_t c, vector<double> &inv) {
  inv.assign(c * c, 0.0);
  for (size_t i = 0; i < c; ++i) inv[i * c + i] = 1.0;
  for (size_t k = 0; k < c; ++k) {
    size_t pr = k;
    for (size_t i = k + 1; i < c; ++i) if (fabs(A[i * c + k]) > fabs(A[pr * c + k])) pr = i;
    if (A[pr * c + k] == 0.0) return false;
    if (pr != k) for (size_t j = 0; j < c; ++j) { std::swap(A[k * c + j], A[pr * c + j]); std::swap(inv[k * c + j], inv[pr * c + j]); }
    const double piv = A[k * c + k];
    for (size_t j = 0; j < c; ++j) { A[k * c + j] /= piv; inv[k * c + j] /= piv; }
    for (size_t i = 0; i < c; ++i) if (i != k) {
      const double f = A[i * c + k];
      if (f != 0.0) for (size_t j = 0; j < c; ++j) { A[i * c + j] -= f * A[k * c + j]; inv[i * c + j] -= f * inv[k * c + j]; }
    }
  }
  return true;
}

// Hardy-Weinberg exact test (Wigginton, Cutler & Abecasis 2005), the statistic CalcHWE returns
// (src/mathfunc.cpp:546-627): two-sided p-value = total probability of heterozygote counts no more
// likely than the observed one, conditional on the allele counts.
static double hwe_exact(size_t n_hom1, size_t n_hom2, size_t n_het) {
  const long total = (long)(n_hom1 + n_hom2 + n_het);
  if (total == 0) return 1.0;
  const long rare_hom = (long)std::min(n_hom1, n_hom2), common_hom = (long)std::max(n_hom1, n_hom2);
  const long rare = 2 * rare_hom + (long)n_het;
  vector<double> pr((size_t)rare + 1, 0.0);
  long mid = rare * (2 * total - rare) / (2 * total);
  if ((rare ^ mid) & 1) mid++;
  pr[mid] = 1.0;
  double norm = 1.0;
  { long het = mid, hr = (rare - mid) / 2, hc = total - het - hr;       // walk down in steps of two heterozygotes
    for (; het > 1; het -= 2) { pr[het - 2] = pr[het] * het * (het - 1.0) / (4.0 * (hr + 1.0) * (hc + 1.0)); norm += pr[het - 2]; hr++; hc++; } }
  { long het = mid, hr = (rare - mid) / 2, hc = total - het - hr;       // walk up
    for (; het <= rare - 2; het += 2) { pr[het + 2] = pr[het] * 4.0 * hr * hc / ((het + 2.0) * (het + 1.0)); norm += pr[het + 2]; hr--; hc--; } }
  (void)common_hom;
  double p = 0.0;
  const double obs = pr[n_het] / norm;
  for (long i = 0; i <= rare; ++i) { const double v = pr[i] / norm; if (v > obs) continue; p += v; }
  return p > 1.0 ? 1.0 : p;
}

// ---- run state -----------------------------------------------------------------------------------
struct Run {
  Params P;
  vector<vector<double>> pheno; vector<vector<int>> ind_pheno;
  vector<vector<double>> cvt; vector<int> ind_cvt; size_t n_cvt = 1;
  bool cvt_from_file = false;                       // ind_cvt still is the covariate file's indicator (not the all-ones default)
  bool cvt_cleared = false;                         // the covariate file held constant columns only: CheckCvt emptied indicator_cvt
  vector<double> weight; vector<int> ind_weight;    // -widv: ReadFile_column(file_weight, indicator_weight, weight, 1), src/param.cpp:241-245
  vector<double> gxe; vector<int> ind_gxe;          // -gxe: ReadFile_column(file_gxe, indicator_gxe, gxe, 1), src/param.cpp:236-240
  vector<int> indicator_idv; size_t ni_total = 0, ni_test = 0;
  std::map<string, std::tuple<string, long, double>> anno;
  std::set<string> setSnps, setKSnps, setGWASnps;
  vector<SnpInfo> snpInfo; vector<int> indicator_snp; size_t ns_total = 0, ns_test = 0;
  std::map<string, int> mapID2num;
  // results
  gb200_nullmodel nm{}; vector<double> beta_mle, se_mle, beta_remle, se_remle; double trace_G = 0.0;
  vector<gb200_sumstat> sumStat;
  double t_total = 0, t_kin = 0, t_eigen = 0, t_lmm = 0;
  string cmdline;
};

static void read_pheno(Run &R) {                       // ReadFile_pheno, src/gemma_io.cpp:386-444
  LineReader in(R.P.file_pheno);
  if (!in.ok()) die("fail to open phenotype file: " + R.P.file_pheno);
  const size_t p_max = *std::max_element(R.P.p_column.begin(), R.P.p_column.end());
  std::map<size_t, size_t> col; for (size_t i = 0; i < R.P.p_column.size(); ++i) col[R.P.p_column[i]] = i;
  string line;
  while (in.next(line)) {
    vector<double> row(R.P.p_column.size(), -9); vector<int> ind(R.P.p_column.size(), 0);
    char *p = tok(&line[0]);
    for (size_t i = 0; i < p_max; ++i) {
      if (!p) die("Number of phenotypes in pheno file do not match phenotypes in geno file");
      auto it = col.find(i + 1);
      if (it != col.end()) { if (strcmp(p, "NA") == 0) { ind[it->second] = 0; row[it->second] = -9; } else { ind[it->second] = 1; row[it->second] = atof(p); } }
      p = tok(nullptr);
    }
    R.pheno.push_back(row); R.ind_pheno.push_back(ind);
  }
}

static void read_cvt(Run &R) {                         // ReadFile_cvt, src/gemma_io.cpp:446-511
  LineReader in(R.P.file_cvt);
  if (!in.ok()) die("fail to open covariates file: " + R.P.file_cvt);
  string line;
  while (in.next(line)) {
    vector<double> v; int na = 0;
    for (char *p = tok(&line[0]); p; p = tok(nullptr)) { if (strcmp(p, "NA") == 0) { na = 1; v.push_back(-9); } else v.push_back(atof(p)); }
    R.ind_cvt.push_back(na ? 0 : 1); R.cvt.push_back(v);
  }
  bool first = true;
  for (size_t i = 0; i < R.ind_cvt.size(); ++i) {
    if (!R.ind_cvt[i]) continue;
    if (first) { R.n_cvt = R.cvt[i].size(); first = false; }
    else if (R.cvt[i].size() != R.n_cvt) die("number of covariates in row " + std::to_string(i) + " do not match other rows.");
  }
}

static void read_column(const string &file, vector<int> &ind, vector<double> &val) {   // ReadFile_column(.., 1), src/gemma_io.cpp:344-383
  LineReader in(file);
  if (!in.ok()) die("fail to open phenotype file: " + file);
  string line;
  while (in.next(line)) {
    char *p = tok(&line[0]);
    if (!p) die("Problem reading PHENO column");
    if (strcmp(p, "NA") == 0) { ind.push_back(0); val.push_back(-9); }
    else { ind.push_back(1); val.push_back(atof(p)); }
  }
}
static void read_gxe(Run &R) { read_column(R.P.file_gxe, R.ind_gxe, R.gxe); }
static void read_weight(Run &R) { read_column(R.P.file_weight, R.ind_weight, R.weight); }

static void read_anno(Run &R) {                        // ReadFile_anno, src/gemma_io.cpp:280-341
  LineReader in(R.P.file_anno);
  if (!in.ok()) die("fail to open annotation file: " + R.P.file_anno);
  string line;
  while (in.next(line)) {
    char *p = tok(&line[0]); if (!p) continue;
    string rs = p; long bp = -9; string chr = "-9"; double cM = -9;
    p = tok(nullptr); if (!p) die("Problem reading annotation file " + R.P.file_anno);
    bp = strcmp(p, "NA") == 0 ? -9 : atol(p);
    p = tok(nullptr); if (p && strcmp(p, "NA") != 0) chr = p;
    p = p ? tok(nullptr) : nullptr; if (p && strcmp(p, "NA") != 0) cM = atof(p);
    R.anno[rs] = std::make_tuple(chr, bp, cM);
  }
}

static void read_snp_set(const string &file, std::set<string> &out) {      // ReadFile_snps, src/gemma_io.cpp:153-176
  LineReader in(file);
  if (!in.ok()) die("fail to open snps file: " + file);
  string line;
  while (in.next(line)) { char *p = tok(&line[0]); if (p) out.insert(p); }
}

static void read_bim(Run &R) {                         // ReadFile_bim, src/gemma_io.cpp:514-556
  std::ifstream in(R.P.file_bfile + ".bim");
  if (!in) die("error opening .bim file: " + R.P.file_bfile + ".bim");
  string line;
  while (std::getline(in, line)) {
    SnpInfo s{};
    char *p = tok(&line[0], " \t"); if (!p) continue; s.chr = p;
    p = tok(nullptr, " \t"); if (!p) die("bad .bim line"); s.rs = p;
    p = tok(nullptr, " \t"); if (!p) die("bad .bim line"); s.cM = atof(p);
    p = tok(nullptr, " \t"); if (!p) die("bad .bim line"); s.bp = atol(p);
    p = tok(nullptr, " \t"); if (!p) die("bad .bim line"); s.a_minor = p;
    p = tok(nullptr, " \t"); if (!p) die("bad .bim line"); s.a_major = p;
    s.n_miss = 0; s.missingness = -9; s.maf = -9; s.n_idv = 0;
    R.snpInfo.push_back(s);
  }
}

static void read_fam(Run &R) {                         // ReadFile_fam, src/gemma_io.cpp:559-635
  LineReader in(R.P.file_bfile + ".fam");
  if (!in.ok()) die("error opening .fam file: " + R.P.file_bfile + ".fam");
  const size_t p_max = *std::max_element(R.P.p_column.begin(), R.P.p_column.end());
  std::map<size_t, size_t> col; for (size_t i = 0; i < R.P.p_column.size(); ++i) col[R.P.p_column[i]] = i;
  string line; int c = 0;
  while (in.next(line)) {
    vector<double> row(R.P.p_column.size(), -9); vector<int> ind(R.P.p_column.size(), 0);
    char *p = tok(&line[0], " \t"); if (!p) continue;
    p = tok(nullptr, " \t"); if (!p) die("bad .fam line"); string id = p;
    for (int k = 0; k < 4; ++k) { p = tok(nullptr, " \t"); if (!p && k < 3) die("bad .fam line"); }
    for (size_t i = 0; i < p_max; ++i) {
      auto it = col.find(i + 1);
      if (it != col.end()) {
        if (!p) die("Problem reading FAM file (phenotypes do not match geno file)");
        if (strcmp(p, "NA") == 0) { ind[it->second] = 0; row[it->second] = -9; }
        else { double v = atof(p); if (v == -9) { ind[it->second] = 0; row[it->second] = -9; } else { ind[it->second] = 1; row[it->second] = v; } }
      }
      p = tok(nullptr);
    }
    R.pheno.push_back(row); R.ind_pheno.push_back(ind); R.mapID2num[id] = c++;
  }
}

static void process_cvt_phen(Run &R) {                 // ProcessCvtPhen + CheckCvt, src/param.cpp:1993-2098, 1937-1990
  R.ni_total = R.ind_pheno.size();
  R.indicator_idv.assign(R.ni_total, 1);
  for (size_t i = 0; i < R.ni_total; ++i) for (int v : R.ind_pheno[i]) if (!v) R.indicator_idv[i] = 0;
  if (!R.ind_cvt.empty()) {
    if (R.ind_cvt.size() != R.ni_total) die("number of rows in the covariates file do not match the number of individuals");
    for (size_t i = 0; i < R.ni_total; ++i) R.indicator_idv[i] *= R.ind_cvt[i];
  }
  if (!R.ind_gxe.empty()) {                            // src/param.cpp:1008-1014, 2016-2020
    if (R.ind_gxe.size() != R.ni_total) die("number of rows in the gxe file do not match the number of individuals. ");
    for (size_t i = 0; i < R.ni_total; ++i) R.indicator_idv[i] *= R.ind_gxe[i];
  }
  if (!R.ind_weight.empty()) {                         // residual weights, src/param.cpp:1015-1021, 2023-2027
    if (R.ind_weight.size() != R.ni_total) die("number of rows in the weight file do not match the number of individuals. ");
    for (size_t i = 0; i < R.ni_total; ++i) R.indicator_idv[i] *= R.ind_weight[i];
  }
  R.ni_test = 0; for (int v : R.indicator_idv) R.ni_test += v;
  if (R.ni_test == 0) die("number of analyzed individuals equals 0. ");
  if (!R.ind_cvt.empty()) {
    vector<size_t> rows; for (size_t i = 0; i < R.ni_total; ++i) if (R.indicator_idv[i] && R.ind_cvt[i]) rows.push_back(i);
    size_t n_const = 0;
    for (size_t j = 0; j < R.n_cvt; ++j) {
      double mn = R.cvt[rows[0]][j], mx = mn;
      for (size_t r : rows) { mn = std::min(mn, R.cvt[r][j]); mx = std::max(mx, R.cvt[r][j]); }
      if (mn == mx) n_const++;
    }
    if (n_const == R.n_cvt) { R.ind_cvt.clear(); R.cvt.clear(); R.n_cvt = 1; R.cvt_cleared = true; }
    else if (n_const == 0) {
      std::cout << "no intercept term is found in the cvt file: a column of 1s is added" << std::endl;
      for (size_t r : rows) R.cvt[r].push_back(1.0);
      R.n_cvt++;
    }
  }
  R.cvt_from_file = !R.ind_cvt.empty();
  if (R.ind_cvt.empty()) { R.cvt.assign(R.ni_total, vector<double>(1, 1.0)); R.ind_cvt.assign(R.ni_total, 1); R.n_cvt = 1; }
}

// Size trim_individuals (src/param.cpp:74-90) leaves a flag vector of `size` entries with: the NUMBER of set flags seen when the scan
// stops (not the index it stopped at), i.e. min(#set, nind).
static size_t trim_count(const vector<int> &v, size_t size, size_t ni_max) {
  size_t count = 0;
  for (size_t i = 0; i < size; ++i) { if (v[i]) count++; if (count >= ni_max) break; }
  return count;
}

// -nind: the reference trims indicator_cvt right after reading the covariates (src/param.cpp:234), indicator_idv and indicator_cvt
// again after ProcessCvtPhen on the BIMBAM branch only (:315-316; the PLINK branch rebuilds indicator_idv from the .fam and never
// trims it), never indicator_gxe, and then CheckData (:1001-1014) refuses the run when the sizes disagree.  Without a covariate
// file indicator_cvt is the all-ones vector ProcessCvtPhen builds (:2085-2092), trimmed like the rest.  Consequences reproduced here:
// PLINK input ignores -nind unless a covariate file is given, in which case the run is refused; BIMBAM input is refused when a kept
// covariate row is missing (the second trim shrinks the vector again) or when -nind exceeds the number of usable individuals while
// some are unusable.  Otherwise the first min(#set, nind) individuals of the file stay and ni_total / ni_test follow (:1033-1041).
static void trim_individuals(Run &R) {
  if (R.P.nind <= 0) return;
  const size_t ni_max = (size_t)R.P.nind;
  const string cvt_msg = "number of rows in the covariates file do not match the number of individuals. ";
  size_t cvt_size = R.ind_cvt.size();                                                  // all ones unless it came from the file
  if (R.cvt_from_file) cvt_size = trim_count(R.ind_cvt, cvt_size, ni_max);            // :234
  if (!R.P.file_bfile.empty()) {
    if (R.cvt_from_file && cvt_size != R.indicator_idv.size()) die(cvt_msg + std::to_string(cvt_size));
    return;
  }
  const size_t count = trim_count(R.indicator_idv, R.indicator_idv.size(), ni_max);   // :315
  if (!R.cvt_cleared) {
    cvt_size = trim_count(R.ind_cvt, cvt_size, ni_max);                                // :316
    if (cvt_size != count) die(cvt_msg + std::to_string(cvt_size));
  }
  if (!R.ind_gxe.empty() && R.ind_gxe.size() != count) die("number of rows in the gxe file do not match the number of individuals. ");
  if (!R.ind_weight.empty() && R.ind_weight.size() != count) die("number of rows in the weight file do not match the number of individuals. ");
  if (count == R.indicator_idv.size()) return;
  R.indicator_idv.resize(count); R.ni_total = count;
  R.ni_test = 0; for (int v : R.indicator_idv) R.ni_test += v;
}

// W (ni_test x n_cvt) and y (ni_test) of the analysed individuals: CopyCvtPhen, src/param.cpp:2146-2198
static void copy_cvt_phen(const Run &R, vector<double> &W, vector<double> &y) {
  W.assign(R.ni_test * R.n_cvt, 0.0); y.assign(R.ni_test, 0.0);
  size_t k = 0;
  for (size_t i = 0; i < R.ni_total; ++i) {
    if (!R.indicator_idv[i]) continue;
    y[k] = R.pheno[i][0];
    for (size_t j = 0; j < R.n_cvt; ++j) W[k * R.n_cvt + j] = R.cvt[i][j];
    k++;
  }
}

struct R2Filter {            // the -r2 filter state (src/gemma_io.cpp:672-688, 826-850)
  size_t c = 1; vector<double> W, WtWi;
  void init(const Run &R) {
    vector<double> y; copy_cvt_phen(R, W, y); c = R.n_cvt;
    vector<double> WtW(c * c, 0.0);
    for (size_t i = 0; i < R.ni_test; ++i) for (size_t a = 0; a < c; ++a) for (size_t b = 0; b < c; ++b) WtW[a * c + b] += W[i * c + a] * W[i * c + b];
    if (!invert_small(WtW, c, WtWi)) WtWi.assign(c * c, 0.0);
  }
  bool correlated(const vector<double> &x, double r2_level) const {
    if (c == 1) return false;
    vector<double> Wtx(c, 0.0), t(c, 0.0);
    double v_x = 0.0, v_w = 0.0;
    for (size_t i = 0; i < x.size(); ++i) { v_x += x[i] * x[i]; for (size_t a = 0; a < c; ++a) Wtx[a] += W[i * c + a] * x[i]; }
    for (size_t a = 0; a < c; ++a) for (size_t b = 0; b < c; ++b) t[a] += WtWi[a * c + b] * Wtx[b];
    for (size_t a = 0; a < c; ++a) v_w += Wtx[a] * t[a];
    return v_w / v_x > r2_level;
  }
};

// QC pass over a BIMBAM mean-genotype file: ReadFile_geno, src/gemma_io.cpp:639-873.  Lines are tokenised and tested on the
// host worker pool (line_pipeline.h); the consumer appends the per-SNP records in file order.
struct RowBlock { size_t rows = 0, ncol = 0; vector<double> v; };     // parsed genotype rows of one block of lines (SNP-major)
struct QcBlock { vector<SnpInfo> info; vector<int> keep; double min_g = 1e300, max_g = -1e300; };

static void qc_bimbam(Run &R) {
  R2Filter r2; r2.init(R);
  const Run *Rc = &R; const R2Filter *r2c = &r2;
  LinePipeline<QcBlock> pipe(R.P.file_geno, [Rc, r2c](LineBlock &blk, QcBlock &out) {
    const Run &R = *Rc;
    vector<double> geno(R.ni_test); vector<char> miss(R.ni_test);
    for (char *line : blk.lines) {
      char *cur = line;
      char *p = next_token(cur); if (!p) continue;
      SnpInfo s{}; s.rs = p;
      p = next_token(cur); if (!p) die("Parsing input file '" + R.P.file_geno + "' failed"); s.a_minor = p;
      p = next_token(cur); if (!p) die("Parsing input file '" + R.P.file_geno + "' failed"); s.a_major = p;
      if (!R.setSnps.empty() && !R.setSnps.count(s.rs)) {
        s.chr = "-9"; s.bp = -9; s.cM = -9; s.n_miss = 0; s.missingness = -9; s.maf = -9; s.n_idv = 0;
        out.info.push_back(s); out.keep.push_back(0); continue;
      }
      auto it = R.anno.find(s.rs);
      if (it == R.anno.end()) { s.chr = "-9"; s.bp = -9; s.cM = -9; } else { s.chr = std::get<0>(it->second); s.bp = std::get<1>(it->second); s.cM = std::get<2>(it->second); }
      double maf = 0.0; size_t n_miss = 0, n_0 = 0, n_1 = 0, n_2 = 0, c_idv = 0; int flag_poly = 0; double geno_old = -9;
      std::fill(miss.begin(), miss.end(), 0);
      for (size_t i = 0; i < R.ni_total; ++i) {
        p = next_token(cur);
        if (!p) die("Problem reading geno file (not enough genotypes in line)");
        if (!R.indicator_idv[i]) continue;
        if (p[0] == 'N' && p[1] == 'A' && p[2] == 0) { miss[c_idv] = 1; n_miss++; c_idv++; continue; }
        const double g = token_to_double(p);
        if (g >= 0 && g <= 0.5) n_0++;
        if (g > 0.5 && g < 1.5) n_1++;
        if (g >= 1.5 && g <= 2.0) n_2++;
        geno[c_idv] = g;
        if (g < out.min_g) out.min_g = g;
        if (g > out.max_g) out.max_g = g;
        if (flag_poly == 0) { geno_old = g; flag_poly = 2; }
        if (flag_poly == 2 && g != geno_old) flag_poly = 1;
        maf += g; c_idv++;
      }
      maf /= 2.0 * (double)(R.ni_test - n_miss);
      s.n_miss = (long)n_miss; s.missingness = (double)n_miss / (double)R.ni_test; s.maf = maf; s.n_idv = (long)(R.ni_test - n_miss);
      int keep = 1;
      if ((double)n_miss / (double)R.ni_test > R.P.miss_level) keep = 0;
      else if ((maf < R.P.maf_level || maf > (1.0 - R.P.maf_level)) && R.P.maf_level != -1) keep = 0;
      else if (flag_poly != 1) keep = 0;
      else if (R.P.hwe_level != 0 && R.P.maf_level != -1 && hwe_exact(n_0, n_2, n_1) < R.P.hwe_level) keep = 0;
      else {
        for (size_t i = 0; i < R.ni_test; ++i) if (miss[i]) geno[i] = maf * 2.0;
        if (r2c->correlated(geno, R.P.r2_level)) keep = 0;
      }
      out.info.push_back(std::move(s)); out.keep.push_back(keep);
    }
  });
  if (!pipe.ok()) die("error reading genotype file:" + R.P.file_geno);
  double min_g = 1e300, max_g = -1e300;
  QcBlock blk;
  while (pipe.next(blk)) {
    for (size_t k = 0; k < blk.info.size(); ++k) { R.snpInfo.push_back(std::move(blk.info[k])); R.indicator_snp.push_back(blk.keep[k]); R.ns_test += blk.keep[k]; }
    min_g = std::min(min_g, blk.min_g); max_g = std::max(max_g, blk.max_g);
  }
  R.ns_total = R.indicator_snp.size();
  if (min_g != 0.0) std::cout << "**** WARNING: The minimum genotype value is not 0.0 - this is not the BIMBAM standard and will skew l_lme and effect sizes" << std::endl;
  if (max_g != 2.0) std::cout << "**** WARNING: The maximum genotype value is not 2.0 - this is not the BIMBAM standard and will skew l_lme and effect sizes" << std::endl;
}

static vector<unsigned char> g_bed;      // whole .bed payload (without the 3 magic bytes)
static size_t g_nbit = 0;

// QC pass over a PLINK .bed: ReadFile_bed, src/gemma_io.cpp:876-1064.  With a context the per-SNP counting and the
// r2 terms run on the device (gb200_qc_bed); the thresholds are applied here in the reference's order.
static void qc_plink(Run &R, gb200_ctx *ctx) {
  std::ifstream in(R.P.file_bfile + ".bed", std::ios::binary);
  if (!in) die("error reading bed file:" + R.P.file_bfile + ".bed");
  g_nbit = (R.ni_total + 3) / 4;
  R.ns_total = R.snpInfo.size();
  g_bed.resize(g_nbit * R.ns_total);
  // The reference skips the three magic bytes unread (seekg(3), src/gemma_io.cpp:951-953) and would decode an individual-major
  // or truncated file into garbage; refuse both loudly instead.
  unsigned char magic[3] = {0, 0, 0};
  in.read(reinterpret_cast<char *>(magic), 3);
  if (in.gcount() != 3 || magic[0] != 0x6c || magic[1] != 0x1b) die("not a PLINK .bed file: " + R.P.file_bfile + ".bed");
  if (magic[2] != 0x01) die("individual-major .bed files are not supported (the reference assumes SNP-major): " + R.P.file_bfile + ".bed");
  in.read(reinterpret_cast<char *>(g_bed.data()), (std::streamsize)g_bed.size());
  if ((size_t)in.gcount() != g_bed.size()) die("truncated .bed file: expected " + std::to_string(g_bed.size()) + " genotype bytes for " +
                                                std::to_string(R.ns_total) + " SNPs x " + std::to_string(R.ni_total) + " individuals");
  R2Filter r2; r2.init(R);
  if (ctx) {
    vector<unsigned char> mask(R.ni_total); for (size_t i = 0; i < R.ni_total; ++i) mask[i] = (unsigned char)R.indicator_idv[i];
    vector<gb200_snpqc> st(R.ns_total);
    const bool with_w = r2.c != 1;
    if (gb200_qc_bed(ctx, g_bed.data(), mask.data(), R.ni_total, R.ns_total, g_nbit, with_w ? r2.W.data() : nullptr,
                     with_w ? r2.WtWi.data() : nullptr, with_w ? r2.c : 0, st.data()) != 0)
      die(string("gb200_qc_bed: ") + gb200_last_error(ctx));
    for (size_t t = 0; t < R.ns_total; ++t) {
      SnpInfo &s = R.snpInfo[t];
      if (!R.setSnps.empty() && !R.setSnps.count(s.rs)) { s.n_miss = -9; s.missingness = -9; s.maf = -9; R.indicator_snp.push_back(0); continue; }
      const gb200_snpqc &q = st[t];
      const size_t n_miss = (size_t)q.n_miss, n_0 = (size_t)q.n_0, n_1 = (size_t)q.n_1, n_2 = (size_t)q.n_2;
      const double maf = q.maf;
      s.n_miss = (long)n_miss; s.missingness = (double)n_miss / (double)R.ni_test; s.maf = maf; s.n_idv = (long)(R.ni_test - n_miss);
      int keep = 1;
      if ((double)n_miss / (double)R.ni_test > R.P.miss_level) keep = 0;
      else if ((maf < R.P.maf_level || maf > (1.0 - R.P.maf_level)) && R.P.maf_level != -1) keep = 0;
      else if ((n_0 + n_1) == 0 || (n_1 + n_2) == 0 || (n_2 + n_0) == 0) keep = 0;
      else if (R.P.hwe_level != 0 && R.P.maf_level != -1 && hwe_exact(n_0, n_2, n_1) < R.P.hwe_level) keep = 0;
      else if (with_w && q.v_w / q.v_x > R.P.r2_level) keep = 0;
      R.indicator_snp.push_back(keep); R.ns_test += keep;
    }
    return;
  }
  vector<double> geno(R.ni_test); vector<char> miss(R.ni_test);
  for (size_t t = 0; t < R.ns_total; ++t) {
    SnpInfo &s = R.snpInfo[t];
    if (!R.setSnps.empty() && !R.setSnps.count(s.rs)) { s.n_miss = -9; s.missingness = -9; s.maf = -9; R.indicator_snp.push_back(0); continue; }
    const unsigned char *row = g_bed.data() + t * g_nbit;
    double maf = 0.0; size_t n_miss = 0, n_0 = 0, n_1 = 0, n_2 = 0, c_idv = 0;
    std::fill(miss.begin(), miss.end(), 0);
    for (size_t j = 0; j < R.ni_total; ++j) {
      if (!R.indicator_idv[j]) continue;
      const unsigned b = (unsigned)row[j >> 2] >> (2 * (j & 3));
      const unsigned lo = b & 1u, hi = (b >> 1) & 1u;
      if (lo == 0) { if (hi == 0) { geno[c_idv] = 2.0; maf += 2.0; n_2++; } else { geno[c_idv] = 1.0; maf += 1.0; n_1++; } }
      else { if (hi == 1) { geno[c_idv] = 0.0; n_0++; } else { miss[c_idv] = 1; n_miss++; } }
      c_idv++;
    }
    maf /= 2.0 * (double)(R.ni_test - n_miss);
    s.n_miss = (long)n_miss; s.missingness = (double)n_miss / (double)R.ni_test; s.maf = maf; s.n_idv = (long)(R.ni_test - n_miss);
    int keep = 1;
    if ((double)n_miss / (double)R.ni_test > R.P.miss_level) keep = 0;
    else if ((maf < R.P.maf_level || maf > (1.0 - R.P.maf_level)) && R.P.maf_level != -1) keep = 0;
    else if ((n_0 + n_1) == 0 || (n_1 + n_2) == 0 || (n_2 + n_0) == 0) keep = 0;
    else if (R.P.hwe_level != 0 && R.P.maf_level != -1 && hwe_exact(n_0, n_2, n_1) < R.P.hwe_level) keep = 0;
    else {
      for (size_t i = 0; i < R.ni_test; ++i) if (miss[i]) geno[i] = maf * 2.0;
      if (r2.correlated(geno, R.P.r2_level)) keep = 0;
    }
    R.indicator_snp.push_back(keep); R.ns_test += keep;
  }
}

static void print_counts(const Run &R) {               // CheckData, src/param.cpp:1108-1134
  std::cout << "## number of total individuals = " << R.ni_total << std::endl;
  std::cout << "## number of analyzed individuals = " << R.ni_test << std::endl;
  std::cout << "## number of covariates = " << R.n_cvt << std::endl;
  std::cout << "## number of phenotypes = " << R.P.p_column.size() << std::endl;
  std::cout << "## number of total SNPs/var        = " << std::setw(8) << R.ns_total << std::endl;
  if (!R.setSnps.empty()) std::cout << "## number of considered SNPS       = " << std::setw(8) << R.setSnps.size() << std::endl;
  if (!R.setKSnps.empty()) std::cout << "## number of SNPS for K            = " << std::setw(8) << R.setKSnps.size() << std::endl;
  if (!R.setGWASnps.empty()) std::cout << "## number of SNPS for GWAS         = " << std::setw(8) << R.setGWASnps.size() << std::endl;
  std::cout << "## number of analyzed SNPs         = " << std::setw(8) << R.ns_test << std::endl;
}

static string out_path(const Run &R, const string &suffix) { return R.P.path_out + "/" + R.P.file_out + "." + suffix + ".txt"; }

// Binary side channel for K / U / D (the reference's own design notes ask for a new kinship format,
// doc/developers/design.org): the 10-digit text of a 50k x 50k matrix is ~35 GB and its rounding is visible in the results.
// With -bin the writers add "<file>.bin" next to the text file; every matrix reader accepts a path ending in ".bin".
// Layout: 8-byte magic "GB2MAT01", uint64 rows, uint64 cols, rows*cols little-endian doubles (row-major).
static bool is_bin(const string &path) { return path.size() > 4 && path.compare(path.size() - 4, 4, ".bin") == 0; }
static void write_bin(const string &path, const double *M, size_t rows, size_t cols) {
  FILE *f = fopen(path.c_str(), "wb");
  if (!f) { std::cout << "error writing file: " << path << std::endl; return; }
  const uint64_t hdr[2] = {rows, cols};
  fwrite("GB2MAT01", 1, 8, f); fwrite(hdr, sizeof(uint64_t), 2, f); fwrite(M, sizeof(double), rows * cols, f);
  fclose(f);
}
static void read_bin(const string &path, vector<double> &M, size_t &rows, size_t &cols) {
  FILE *f = fopen(path.c_str(), "rb");
  if (!f) die("fail to open binary matrix file: " + path);
  char magic[8]; uint64_t hdr[2];
  if (fread(magic, 1, 8, f) != 8 || memcmp(magic, "GB2MAT01", 8) != 0 || fread(hdr, sizeof(uint64_t), 2, f) != 2) die("not a gemma-b200 binary matrix: " + path);
  rows = hdr[0]; cols = hdr[1];
  M.resize(rows * cols);
  if (fread(M.data(), sizeof(double), rows * cols, f) != rows * cols) die("truncated binary matrix: " + path);
  fclose(f);
}

static void write_matrix(const Run &R, const double *M, size_t rows, size_t cols, const string &suffix) {   // WriteMatrix, src/param.cpp:1886-1910
  if (R.P.bin) write_bin(out_path(R, suffix) + ".bin", M, rows, cols);
  // precision(10) with the default float field == "%.10g"; rows are formatted on the host worker threads, written in order
  FILE *f = fopen(out_path(R, suffix).c_str(), "w");
  if (!f) { std::cout << "error writing file: " << out_path(R, suffix) << std::endl; return; }
  const int nt = host_threads();
  const size_t group = std::max<size_t>(1, std::min<size_t>(rows, (size_t(1) << 22) / std::max<size_t>(cols, 1) + 1)) * (size_t)nt;
  vector<string> text(std::min(group, rows));
  for (size_t r0 = 0; r0 < rows; r0 += group) {
    const size_t r1 = std::min(rows, r0 + group);
    vector<std::thread> th;
    for (int t = 0; t < nt; ++t)
      th.emplace_back([&, t]() {
        char buf[40];
        for (size_t i = r0 + (size_t)t; i < r1; i += (size_t)nt) {
          string &line = text[i - r0];
          line.clear(); line.reserve(cols * 14);
          for (size_t j = 0; j < cols; ++j) {
            if (j) line.push_back('\t');
            line.append(buf, (size_t)snprintf(buf, sizeof buf, "%.10g", M[i * cols + j]));
          }
          line.push_back('\n');
        }
      });
    for (auto &x : th) x.join();
    for (size_t i = r0; i < r1; ++i) fwrite(text[i - r0].data(), 1, text[i - r0].size(), f);
  }
  fclose(f);
}
static void write_vector(const Run &R, const double *v, size_t n, const string &suffix) {                    // WriteVector, src/param.cpp:1912-1935
  if (R.P.bin) write_bin(out_path(R, suffix) + ".bin", v, n, 1);
  std::ofstream out(out_path(R, suffix));
  if (!out) { std::cout << "error writing file: " << out_path(R, suffix) << std::endl; return; }
  out.precision(10);
  for (size_t i = 0; i < n; ++i) out << v[i] << std::endl;
}

#define GB(call) do { int _rc = (call); if (_rc != 0) die(string(#call) + ": " + gb200_last_error(ctx)); } while (0)

static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// -gk: PARAM::CalcKin -> BimbamKin / PlinkKin
static void run_kinship(Run &R, gb200_ctx *ctx) {
  const int k_mode = R.P.a_mode - 20;
  std::cout << "Calculating Relatedness Matrix ... " << std::endl;
  const double t0 = now_s();
  GB(gb200_kin_begin(ctx, R.ni_total, k_mode));
  if (!R.P.file_bfile.empty()) {
    vector<unsigned char> rows; rows.reserve(BATCH * g_nbit);
    size_t l = 0;
    for (size_t t = 0; t < R.ns_total; ++t) {
      if (!R.indicator_snp[t]) continue;            // PlinkKin ignores -ksnps
      rows.insert(rows.end(), g_bed.begin() + t * g_nbit, g_bed.begin() + (t + 1) * g_nbit);
      if (++l == BATCH) { GB(gb200_kin_add_bed(ctx, rows.data(), l, g_nbit)); rows.clear(); l = 0; }
    }
    if (l) GB(gb200_kin_add_bed(ctx, rows.data(), l, g_nbit));
  } else {
    // BimbamKin, src/gemma_io.cpp:1418-1597: rows of all ni_total individuals, "NA" -> NaN; parsed on the worker pool
    const Run *Rc = &R;
    LinePipeline<RowBlock> pipe(R.P.file_geno, [Rc](LineBlock &blk, RowBlock &out) {
      const Run &R = *Rc;
      out.ncol = R.ni_total;
      for (size_t k = 0; k < blk.lines.size(); ++k) {
        const size_t cur_line = blk.first_line + k;
        if (cur_line >= R.indicator_snp.size() || !R.indicator_snp[cur_line]) continue;
        char *cur = blk.lines[k];
        char *p = next_token(cur); if (!p) continue;
        if (!R.setKSnps.empty() && !R.setKSnps.count(p)) continue;       // -ksnps / -loco, src/gemma_io.cpp:1479
        p = next_token(cur); p = next_token(cur);
        const size_t off = out.v.size();
        out.v.resize(off + R.ni_total);
        double *g = out.v.data() + off;
        for (size_t i = 0; i < R.ni_total; ++i) {
          p = next_token(cur);
          if (!p) die("line " + std::to_string(cur_line + 1) + " of " + R.P.file_geno + ": number of fields");
          g[i] = (p[0] == 'N' && p[1] == 'A') ? NAN : token_to_double(p);
        }
        out.rows++;
      }
    });
    if (!pipe.ok()) die("error reading genotype file:" + R.P.file_geno);
    const size_t chunk = std::max<size_t>(1, std::min<size_t>(BATCH, (size_t(1) << 28) / R.ni_total));
    vector<double> G(chunk * R.ni_total);
    size_t l = 0;
    RowBlock blk;
    while (pipe.next(blk)) {
      for (size_t r = 0; r < blk.rows; ++r) {
        std::memcpy(G.data() + l * R.ni_total, blk.v.data() + r * R.ni_total, R.ni_total * sizeof(double));
        if (++l == chunk) { GB(gb200_kin_add_geno(ctx, G.data(), l, R.ni_total, R.ni_total)); l = 0; }
      }
    }
    if (l) GB(gb200_kin_add_geno(ctx, G.data(), l, R.ni_total, R.ni_total));
  }
  vector<double> K(R.ni_total * R.ni_total);
  size_t used = 0;
  GB(gb200_kin_finish(ctx, K.data(), R.ni_total, &used));
  R.t_kin = now_s() - t0;
  write_matrix(R, K.data(), R.ni_total, R.ni_total, k_mode == 1 ? "cXX" : "sXX");
}

static void read_kin(Run &R, vector<double> &G) {      // ReadFile_kin -km 1/2, src/gemma_io.cpp:1186-1294
  const size_t n = R.ni_test;
  if (is_bin(R.P.file_kin)) {
    vector<double> full; size_t rows = 0, cols = 0;
    read_bin(R.P.file_kin, full, rows, cols);
    if (rows != R.ni_total || cols != R.ni_total) die("number of rows in the kinship file does not match the number of individuals.");
    G.assign(n * n, 0.0);
    size_t it = 0;
    for (size_t i = 0; i < R.ni_total; ++i) {
      if (!R.indicator_idv[i]) continue;
      size_t jt = 0;
      for (size_t j = 0; j < R.ni_total; ++j) if (R.indicator_idv[j]) G[it * n + jt++] = full[i * R.ni_total + j];
      ++it;
    }
    return;
  }
  LineReader in(R.P.file_kin);
  if (!in.ok()) die("fail to open kinship file: " + R.P.file_kin);
  G.assign(n * n, 0.0);
  string line;
  if (R.P.k_mode == 1) {
    // n x n text (ReadFile_kin, -km 1): rows parsed on the worker pool, kept rows / columns copied in file order
    struct KinBlock { size_t first = 0, nlines = 0; vector<double> v; };
    const Run *Rc = &R;
    LinePipeline<KinBlock> pipe(R.P.file_kin, [Rc, n](LineBlock &blk, KinBlock &out) {
      const Run &R = *Rc;
      out.first = blk.first_line; out.nlines = blk.lines.size();
      for (size_t k = 0; k < blk.lines.size(); ++k) {
        const size_t i_total = blk.first_line + k;
        if (i_total >= R.ni_total) die("number of rows in the kinship file is larger than the number of phenotypes");
        if (!R.indicator_idv[i_total]) continue;
        const size_t off = out.v.size();
        out.v.resize(off + n);
        size_t j_total = 0, j_test = 0;
        char *cur = blk.lines[k];
        for (char *p = next_token(cur); p; p = next_token(cur)) {
          if (j_total == R.ni_total) die("number of columns in the kinship file is larger than the number of individuals for row = " + std::to_string(i_total));
          if (R.indicator_idv[j_total]) { out.v[off + j_test] = token_to_double(p); j_test++; }
          j_total++;
        }
        if (j_total != R.ni_total) die("number of columns in the kinship file does not match the number of individuals for row = " + std::to_string(i_total));
      }
    });
    if (!pipe.ok()) die("fail to open kinship file: " + R.P.file_kin);
    size_t i_test = 0, i_total = 0;
    KinBlock blk;
    while (pipe.next(blk)) {
      std::memcpy(G.data() + i_test * n, blk.v.data(), blk.v.size() * sizeof(double));
      i_test += blk.v.size() / n; i_total += blk.nlines;
    }
    if (i_total != R.ni_total) die("number of rows in the kinship file does not match the number of individuals.");
  } else {
    std::map<size_t, size_t> id2id; size_t c = 0;
    for (size_t i = 0; i < R.ni_total; ++i) if (R.indicator_idv[i]) id2id[i] = c++;
    while (in.next(line)) {
      char *p = tok(&line[0]); if (!p) continue; string id1 = p;
      p = tok(nullptr); if (!p) die("bad kinship line"); string id2 = p;
      p = tok(nullptr); if (!p) die("bad kinship line"); const double d = atof(p);
      auto a = R.mapID2num.find(id1), b = R.mapID2num.find(id2);
      if (a == R.mapID2num.end() || b == R.mapID2num.end()) continue;
      if (!R.indicator_idv[a->second] || !R.indicator_idv[b->second]) continue;
      const size_t i = id2id[a->second], j = id2id[b->second];
      const double old = G[i * n + j];
      if (old != 0 && old != d) die("redundant and unequal terms in the kinship file, for id1 = " + id1 + " and id2 = " + id2);
      G[i * n + j] = d; G[j * n + i] = d;
    }
  }
}

static void read_dense_rows(const string &file, double *dst, size_t rows, size_t cols, const char *what) {   // ReadFile_eigenU/D, src/gemma_io.cpp:1323-1415
  if (is_bin(file)) {
    vector<double> M; size_t r = 0, c = 0;
    read_bin(file, M, r, c);
    if (r != rows || c != cols) die(string("shape of the binary ") + what + " file does not match the analysed individuals");
    std::copy(M.begin(), M.end(), dst);
    return;
  }
  LineReader in(file);
  if (!in.ok()) die(string("fail to open the ") + what + " file: " + file);
  std::fill(dst, dst + rows * cols, 0.0);
  string line; size_t i = 0;
  while (in.next(line)) {
    if (i == rows) die(string("number of rows in the ") + what + " file is larger than expected.");
    size_t j = 0;
    for (char *p = tok(&line[0]); p; p = tok(nullptr)) {
      if (j == cols) die(string("number of columns in the ") + what + " file is larger than expected, for row = " + std::to_string(i));
      dst[i * cols + j++] = atof(p);
    }
    i++;
  }
}

static void write_assoc(const Run &R) {                // LMM::WriteFiles, src/lmm.cpp:101-225
  std::ofstream out(out_path(R, "assoc"));
  if (!out) { std::cout << "error writing file: " << out_path(R, "assoc") << std::endl; return; }
  const int m = R.P.a_mode;
  out << "chr\trs\tps\tn_miss\tallele1\tallele0\taf\t";
  if (m != 2) out << "beta\tse\t";
  if (m != 3 && m != 9) out << "logl_H1\t";
  if (m == 1) out << "l_remle\tp_wald" << std::endl;
  else if (m == 2 || m == 9) out << "l_mle\tp_lrt" << std::endl;
  else if (m == 3) out << "p_score" << std::endl;
  else out << "l_remle\tl_mle\tp_wald\tp_lrt\tp_score" << std::endl;
  size_t t = 0;
  for (size_t i = 0; i < R.snpInfo.size(); ++i) {
    if (!R.indicator_snp[i]) continue;
    if (!R.setGWASnps.empty() && !R.setGWASnps.count(R.snpInfo[i].rs)) continue;     // src/lmm.cpp:209-210
    const SnpInfo &s = R.snpInfo[i]; const gb200_sumstat &st = R.sumStat[t++];
    out << s.chr << "\t" << s.rs << "\t" << s.bp << "\t" << s.n_miss << "\t" << s.a_minor << "\t" << s.a_major << "\t"
        << std::fixed << std::setprecision(3) << s.maf << "\t";
    out << std::scientific << std::setprecision(6);
    if (m != 2) out << st.beta << "\t" << st.se << "\t";
    if (m != 3 && m != 9) out << st.logl_H1 << "\t";
    if (m == 1) out << st.lambda_remle << "\t" << st.p_wald << std::endl;
    else if (m == 2 || m == 9) out << st.lambda_mle << "\t" << st.p_lrt << std::endl;
    else if (m == 3) out << st.p_score << std::endl;
    else out << st.lambda_remle << "\t" << st.lambda_mle << "\t" << st.p_wald << "\t" << st.p_lrt << "\t" << st.p_score << std::endl;
  }
}

// multivariate LMM, two phenotypes: MVLMM::AnalyzeBimbam / AnalyzePlink (src/mvlmm.cpp:2972-3899), MVLMM::WriteFiles (:117-210)
static void run_mvlmm(Run &R, gb200_ctx *ctx, const vector<double> &U, const vector<double> &eval, const vector<double> &W) {
  const size_t n = R.ni_test, d = 2;
  vector<double> Y(n * d);
  { size_t k = 0; for (size_t i = 0; i < R.ni_total; ++i) { if (!R.indicator_idv[i]) continue; Y[k * d] = R.pheno[i][0]; Y[k * d + 1] = R.pheno[i][1]; k++; } }
  const double t0 = now_s();
  GB(gb200_lmm_params(ctx, 1, R.P.l_min, R.P.l_max, R.P.n_region, 0.0, 0.0));     // -lmin / -lmax / -region reach MphInitial's univariate fits (src/mvlmm.cpp:2786-2796)
  GB(gb200_mvlmm_setup(ctx, n, R.n_cvt, d, U.data(), n, eval.data(), W.data(), R.n_cvt, Y.data(), d));
  double Vg[4], Ve[4], Vgm[4], Vem[4], lr, lm; vector<double> Br(d * R.n_cvt), Bm(d * R.n_cvt);
  GB(gb200_mvlmm_null(ctx, Vg, Ve, Br.data(), &lr, Vgm, Vem, Bm.data(), &lm));
  std::cout.setf(std::ios_base::fixed, std::ios_base::floatfield); std::cout.precision(4);                  // src/mvlmm.cpp:3087-3088
  std::cout << "REMLE estimate for Vg in the null model: " << std::endl << Vg[0] << "\t" << std::endl << Vg[2] << "\t" << Vg[3] << "\t" << std::endl;
  std::cout << "REMLE estimate for Ve in the null model: " << std::endl << Ve[0] << "\t" << std::endl << Ve[2] << "\t" << Ve[3] << "\t" << std::endl;
  std::cout << "REMLE likelihood = " << lr << std::endl;
  std::cout << "MLE estimate for Vg in the null model: " << std::endl << Vgm[0] << "\t" << std::endl << Vgm[2] << "\t" << Vgm[3] << "\t" << std::endl;
  std::cout << "MLE estimate for Ve in the null model: " << std::endl << Vem[0] << "\t" << std::endl << Vem[2] << "\t" << Vem[3] << "\t" << std::endl;
  std::cout << "MLE likelihood = " << lm << std::endl;
  std::cout.unsetf(std::ios_base::floatfield); std::cout.precision(6);
  vector<double> stat; stat.reserve(R.ns_test * 8);
  vector<double> out;
  if (!R.P.file_bfile.empty()) {
    vector<unsigned char> mask(R.ni_total); for (size_t i = 0; i < R.ni_total; ++i) mask[i] = (unsigned char)R.indicator_idv[i];
    vector<unsigned char> rows; rows.reserve(BATCH * g_nbit);
    size_t l = 0;
    auto flush = [&]() {
      if (!l) return;
      out.resize(l * 8);
      GB(gb200_mvlmm_batch_bed(ctx, rows.data(), mask.data(), R.ni_total, l, g_nbit, R.P.a_mode, out.data()));
      stat.insert(stat.end(), out.begin(), out.end());
      rows.clear(); l = 0;
    };
    for (size_t t = 0; t < R.ns_total; ++t) {
      if (!R.indicator_snp[t]) continue;
      rows.insert(rows.end(), g_bed.begin() + t * g_nbit, g_bed.begin() + (t + 1) * g_nbit);
      if (++l == BATCH) flush();
    }
    flush();
  } else {
    const Run *Rc = &R;
    LinePipeline<RowBlock> pipe(R.P.file_geno, [Rc, n](LineBlock &blk, RowBlock &o) {
      const Run &R = *Rc;
      o.ncol = n;
      for (size_t k = 0; k < blk.lines.size(); ++k) {
        const size_t cur_line = blk.first_line + k;
        if (cur_line >= R.indicator_snp.size() || !R.indicator_snp[cur_line]) continue;
        char *cur = blk.lines[k];
        char *p = next_token(cur); p = next_token(cur); p = next_token(cur);
        const size_t off = o.v.size();
        o.v.resize(off + n);
        double *g = o.v.data() + off; size_t pos = 0;
        for (size_t i = 0; i < R.ni_total; ++i) {
          p = next_token(cur);
          if (!p) die("Problem reading geno file (not enough genotypes in line)");
          if (!R.indicator_idv[i]) continue;
          g[pos++] = (p[0] == 'N' && p[1] == 'A' && p[2] == 0) ? NAN : token_to_double(p);
        }
        o.rows++;
      }
    });
    if (!pipe.ok()) die("error reading genotype file:" + R.P.file_geno);
    const size_t chunk = std::max<size_t>(1, std::min<size_t>(BATCH, (size_t(1) << 28) / n));
    vector<double> G(chunk * n);
    size_t l = 0;
    auto flush = [&]() {
      if (!l) return;
      out.resize(l * 8);
      GB(gb200_mvlmm_batch_geno(ctx, G.data(), l, n, R.P.a_mode, out.data()));
      stat.insert(stat.end(), out.begin(), out.end());
      l = 0;
    };
    RowBlock blk;
    while (pipe.next(blk)) {
      for (size_t r = 0; r < blk.rows; ++r) {
        std::memcpy(G.data() + l * n, blk.v.data() + r * n, n * sizeof(double));
        if (++l == chunk) flush();
      }
    }
    flush();
  }
  R.t_lmm = now_s() - t0;
  std::ofstream o(out_path(R, "assoc"));
  if (!o) { std::cout << "error writing file: " << out_path(R, "assoc") << std::endl; return; }
  const int m = R.P.a_mode;
  o << "chr\trs\tps\tn_miss\tallele1\tallele0\taf\tbeta_1\tbeta_2\tVbeta_1_1\tVbeta_1_2\tVbeta_2_2\t"
    << (m == 1 ? "p_wald" : m == 2 ? "p_lrt" : m == 3 ? "p_score" : "p_wald\tp_lrt\tp_score") << std::endl;
  size_t t = 0;
  for (size_t i = 0; i < R.snpInfo.size(); ++i) {
    if (!R.indicator_snp[i]) continue;
    const SnpInfo &s = R.snpInfo[i]; const double *st = stat.data() + 8 * (t++);
    o << s.chr << "\t" << s.rs << "\t" << s.bp << "\t" << s.n_miss << "\t" << s.a_minor << "\t" << s.a_major << "\t" << std::fixed << std::setprecision(3) << s.maf
      << "\t" << std::scientific << std::setprecision(6) << st[0] << "\t" << st[1] << "\t" << st[2] << "\t" << st[3] << "\t" << st[4] << "\t";
    if (m == 1) o << st[5] << std::endl; else if (m == 2) o << st[6] << std::endl; else if (m == 3) o << st[7] << std::endl;
    else o << st[5] << "\t" << st[6] << "\t" << st[7] << std::endl;
  }
}

// CenterMatrix (src/mathfunc.cpp:147-177) on the host, only for the -widv route (the default route centres on the device):
// G <- G - (Gw 1' + 1 Gw')/n + (1'Gw)/n^2, upper triangle mirrored down.
static void center_matrix_host(vector<double> &G, size_t n) {
  vector<double> Gw(n, 0.0);
  for (size_t i = 0; i < n; ++i) { double a = 0.0; for (size_t j = 0; j < n; ++j) a += G[i * n + j]; Gw[i] = a; }
  double d = 0.0; for (size_t i = 0; i < n; ++i) d += Gw[i];
  const double inv = 1.0 / (double)n, dd = d / ((double)n * (double)n);
  for (size_t i = 0; i < n; ++i)
    for (size_t j = i; j < n; ++j) {
      const double v = G[i * n + j] - inv * (Gw[i] + Gw[j]) + dd;
      G[i * n + j] = v; G[j * n + i] = v;
    }
}
static double safe_sqrt_cli(double d) {              // src/mathfunc.cpp:122-131 (every d < 0.001 becomes |d|)
  if (d < 0.001) d = std::fabs(d);
  return std::sqrt(d);
}

// -widv residual weights (src/gemma.cpp:2594-2644): G (ni_test x ni_test, as read) is centred and G_ij /= sqrt(w_i w_j) (0 where a weight
// is not positive); the caller eigendecomposes WITHOUT centring again and scales row i of U by sqrt(w_i).  Returns the weights of the
// analysed individuals (CopyWeight, src/param.cpp:2130-2142).  Host loops, O(n^2).
static vector<double> weighted_kinship(const Run &R, vector<double> &G) {
  const size_t n = R.ni_test;
  vector<double> w; w.reserve(n);
  for (size_t i = 0; i < R.ni_total; ++i) if (R.indicator_idv[i] && R.ind_weight[i]) w.push_back(R.weight[i]);
  if (w.size() != n) die("internal: weights of the analysed individuals");
  center_matrix_host(G, n);
  for (size_t i = 0; i < n; ++i)
    for (size_t j = i; j < n; ++j) {
      double d = G[i * n + j];
      d = (w[i] <= 0 || w[j] <= 0) ? 0.0 : d / safe_sqrt_cli(w[i] * w[j]);
      G[i * n + j] = d; G[j * n + i] = d;
    }
  return w;
}

// LMM branch of BatchRun (src/gemma.cpp:2556-2871) incl. -eigen
static void run_lmm(Run &R, gb200_ctx *ctx) {
  const size_t n = R.ni_test;
  vector<double> W, y; copy_cvt_phen(R, W, y);
  vector<double> U(n * n), eval(n);
  double t0 = now_s();
  if (!R.P.file_kin.empty()) {
    vector<double> G; read_kin(R, G);
    std::cout << "Start Eigen-Decomposition..." << std::endl;
    int n_zero = 0, n_neg = 0;
    if (R.weight.empty()) {
      GB(gb200_eigh(ctx, G.data(), n, n, /*center=*/1, U.data(), n, eval.data(), &R.trace_G, &n_zero, &n_neg));
    } else {
      const vector<double> w = weighted_kinship(R, G);
      GB(gb200_eigh(ctx, G.data(), n, n, /*center=*/0, U.data(), n, eval.data(), &R.trace_G, &n_zero, &n_neg));
      for (size_t i = 0; i < n; ++i) {
        const double wi = w[i] <= 0 ? 0.0 : safe_sqrt_cli(w[i]);
        for (size_t j = 0; j < n; ++j) U[i * n + j] *= wi;
      }
    }
    if (n_zero > 1) std::cout << "**** WARNING: Matrix G has " << n_zero << " eigenvalues close to zero" << std::endl;
  } else {
    read_dense_rows(R.P.file_ku, U.data(), n, n, "U");
    read_dense_rows(R.P.file_kd, eval.data(), n, 1, "D");
    R.trace_G = 0.0;
    for (size_t i = 0; i < n; ++i) { if (eval[i] < 1e-10) eval[i] = 0; R.trace_G += eval[i]; }     // src/gemma.cpp:2662-2667
    R.trace_G /= (double)n;
  }
  R.t_eigen = now_s() - t0;
  if (R.P.a_mode == 31) { write_matrix(R, U.data(), n, n, "eigenU"); write_vector(R, eval.data(), n, "eigenD"); return; }

  if (R.P.p_column.size() == 2) { run_mvlmm(R, ctx, U, eval, W); return; }
  t0 = now_s();
  GB(gb200_lmm_setup(ctx, n, R.n_cvt, U.data(), n, eval.data(), W.data(), R.n_cvt, y.data(), nullptr, nullptr));
  R.beta_mle.resize(R.n_cvt); R.se_mle.resize(R.n_cvt); R.beta_remle.resize(R.n_cvt); R.se_remle.resize(R.n_cvt);
  GB(gb200_lmm_null(ctx, R.P.l_min, R.P.l_max, R.P.n_region, R.trace_G, &R.nm, R.beta_mle.data(), R.se_mle.data(),
                    R.beta_remle.data(), R.se_remle.data()));
  std::cout << "pve estimate =" << R.nm.pve_null << std::endl;            // PrintSummary, src/param.cpp:1252-1259
  std::cout << "se(pve) =" << R.nm.pve_se_null << std::endl;
  GB(gb200_lmm_params(ctx, R.P.a_mode, R.P.l_min, R.P.l_max, R.P.n_region, R.nm.l_mle_null, R.nm.logl_mle_H0));
  const bool gxe = !R.P.file_gxe.empty();
  if (gxe) {                                            // PARAM::CopyGxe, src/param.cpp:2116-2128
    vector<double> env; env.reserve(n);
    for (size_t i = 0; i < R.ni_total; ++i) if (R.indicator_idv[i]) env.push_back(R.gxe[i]);
    GB(gb200_lmm_gxe_setup(ctx, env.data()));
  }
  R.sumStat.clear(); R.sumStat.reserve(R.ns_test);
  vector<gb200_sumstat> out;
  if (!R.P.file_bfile.empty()) {                       // AnalyzePlink, src/lmm.cpp:1710-1903
    vector<unsigned char> mask(R.ni_total); for (size_t i = 0; i < R.ni_total; ++i) mask[i] = (unsigned char)R.indicator_idv[i];
    vector<unsigned char> rows; rows.reserve(BATCH * g_nbit);
    size_t l = 0;
    auto flush = [&]() {
      if (!l) return;
      out.resize(l);
      if (gxe) GB(gb200_lmm_gxe_batch_bed(ctx, rows.data(), mask.data(), R.ni_total, l, g_nbit, out.data()));      // AnalyzePlinkGXE
      else GB(gb200_lmm_batch_bed(ctx, rows.data(), mask.data(), R.ni_total, l, g_nbit, out.data()));
      R.sumStat.insert(R.sumStat.end(), out.begin(), out.end());
      rows.clear(); l = 0;
    };
    for (size_t t = 0; t < R.ns_total; ++t) {
      if (!R.indicator_snp[t]) continue;
      // -gwasnps: the reference's AnalyzePlink ignores the set while WriteFiles applies it (src/lmm.cpp:209-210), which shifts
      // every later row onto another SNP's statistics; here the set filters the tested SNPs, as in the BIMBAM loop below
      if (!R.setGWASnps.empty() && !R.setGWASnps.count(R.snpInfo[t].rs)) continue;
      rows.insert(rows.end(), g_bed.begin() + t * g_nbit, g_bed.begin() + (t + 1) * g_nbit);
      if (++l == BATCH) flush();
    }
    flush();
  } else {                                             // AnalyzeBimbam, src/lmm.cpp:1660-1706 + Analyze :1474-1658
    // analysed individuals only, "NA" -> NaN (src/lmm.cpp:1675-1700); parsed on the worker pool, tested in file order
    const Run *Rc = &R;
    LinePipeline<RowBlock> pipe(R.P.file_geno, [Rc, n](LineBlock &blk, RowBlock &out) {
      const Run &R = *Rc;
      out.ncol = n;
      for (size_t k = 0; k < blk.lines.size(); ++k) {
        const size_t cur_line = blk.first_line + k;
        if (cur_line >= R.indicator_snp.size() || !R.indicator_snp[cur_line]) continue;
        char *cur = blk.lines[k];
        char *p = next_token(cur);
        if (!R.setGWASnps.empty() && (!p || !R.setGWASnps.count(p))) continue;       // -gwasnps / -loco, src/lmm.cpp:1585-1587
        p = next_token(cur); p = next_token(cur);
        const size_t off = out.v.size();
        out.v.resize(off + n);
        double *g = out.v.data() + off; size_t pos = 0;
        for (size_t i = 0; i < R.ni_total; ++i) {
          p = next_token(cur);
          if (!p) die("Problem reading geno file (not enough genotypes in line)");
          if (!R.indicator_idv[i]) continue;
          g[pos++] = (p[0] == 'N' && p[1] == 'A' && p[2] == 0) ? NAN : token_to_double(p);      // "NA" -> NaN, src/lmm.cpp:1690-1694
        }
        out.rows++;
      }
    });
    if (!pipe.ok()) die("error reading genotype file:" + R.P.file_geno);
    const size_t chunk = std::max<size_t>(1, std::min<size_t>(BATCH, (size_t(1) << 28) / n));
    vector<double> G(chunk * n);
    size_t l = 0;
    auto flush = [&]() {
      if (!l) return;
      out.resize(l);
      // (the reference's AnalyzeBimbamGXE opens file_gene instead of file_geno, src/lmm.cpp:2288: BIMBAM + -gxe fails there)
      if (gxe) GB(gb200_lmm_gxe_batch_geno(ctx, G.data(), l, n, out.data()));
      else GB(gb200_lmm_batch_geno(ctx, G.data(), l, n, out.data()));
      R.sumStat.insert(R.sumStat.end(), out.begin(), out.end());
      l = 0;
    };
    RowBlock blk;
    while (pipe.next(blk)) {
      for (size_t r = 0; r < blk.rows; ++r) {
        std::memcpy(G.data() + l * n, blk.v.data() + r * n, n * sizeof(double));
        if (++l == chunk) flush();
      }
    }
    flush();
  }
  R.t_lmm = now_s() - t0;
  write_assoc(R);
}

// -lm: LM branch of BatchRun (src/gemma.cpp:2062-2107), LM::AnalyzeBimbam / AnalyzePlink (src/lm.cpp:382-640), LM::WriteFiles (:83-222)
static void run_lm(Run &R, gb200_ctx *ctx) {
  const size_t n = R.ni_test;
  vector<double> W, y; copy_cvt_phen(R, W, y);
  const double t0 = now_s();
  GB(gb200_lm_setup(ctx, n, R.n_cvt, W.data(), R.n_cvt, y.data()));
  R.sumStat.clear(); R.sumStat.reserve(R.ns_test);
  vector<gb200_sumstat> out;
  if (!R.P.file_bfile.empty()) {
    vector<unsigned char> mask(R.ni_total); for (size_t i = 0; i < R.ni_total; ++i) mask[i] = (unsigned char)R.indicator_idv[i];
    vector<unsigned char> rows; rows.reserve(BATCH * g_nbit);
    size_t l = 0;
    auto flush = [&]() {
      if (!l) return;
      out.resize(l);
      GB(gb200_lm_batch_bed(ctx, rows.data(), mask.data(), R.ni_total, l, g_nbit, R.P.a_mode, out.data()));
      R.sumStat.insert(R.sumStat.end(), out.begin(), out.end());
      rows.clear(); l = 0;
    };
    for (size_t t = 0; t < R.ns_total; ++t) {
      if (!R.indicator_snp[t]) continue;
      rows.insert(rows.end(), g_bed.begin() + t * g_nbit, g_bed.begin() + (t + 1) * g_nbit);
      if (++l == BATCH) flush();
    }
    flush();
  } else {
    const Run *Rc = &R;
    LinePipeline<RowBlock> pipe(R.P.file_geno, [Rc, n](LineBlock &blk, RowBlock &o) {
      const Run &R = *Rc;
      o.ncol = n;
      for (size_t k = 0; k < blk.lines.size(); ++k) {
        const size_t cur_line = blk.first_line + k;
        if (cur_line >= R.indicator_snp.size() || !R.indicator_snp[cur_line]) continue;
        char *cur = blk.lines[k];
        char *p = next_token(cur); p = next_token(cur); p = next_token(cur);
        const size_t off = o.v.size();
        o.v.resize(off + n);
        double *g = o.v.data() + off; size_t pos = 0;
        for (size_t i = 0; i < R.ni_total; ++i) {
          p = next_token(cur);
          if (!p) die("Problem reading geno file (not enough genotypes in line)");
          if (!R.indicator_idv[i]) continue;
          g[pos++] = (p[0] == 'N' && p[1] == 'A' && p[2] == 0) ? NAN : token_to_double(p);
        }
        o.rows++;
      }
    });
    if (!pipe.ok()) die("error reading genotype file:" + R.P.file_geno);
    const size_t chunk = std::max<size_t>(1, std::min<size_t>(BATCH, (size_t(1) << 28) / n));
    vector<double> G(chunk * n);
    size_t l = 0;
    auto flush = [&]() {
      if (!l) return;
      out.resize(l);
      GB(gb200_lm_batch_geno(ctx, G.data(), l, n, R.P.a_mode, out.data()));
      R.sumStat.insert(R.sumStat.end(), out.begin(), out.end());
      l = 0;
    };
    RowBlock blk;
    while (pipe.next(blk)) {
      for (size_t r = 0; r < blk.rows; ++r) {
        std::memcpy(G.data() + l * n, blk.v.data() + r * n, n * sizeof(double));
        if (++l == chunk) flush();
      }
    }
    flush();
  }
  R.t_lmm = now_s() - t0;
  std::ofstream o(out_path(R, "assoc"));
  if (!o) { std::cout << "error writing file: " << out_path(R, "assoc") << std::endl; return; }
  const int m = R.P.a_mode;
  o << "chr\trs\tps\tn_mis\tn_obs\tallele1\tallele0\taf\t";
  if (m == 51) o << "beta\tse\tp_wald" << std::endl;
  else if (m == 52) o << "p_lrt" << std::endl;
  else if (m == 53) o << "beta\tse\tp_score" << std::endl;
  else o << "beta\tse\tp_wald\tp_lrt\tp_score" << std::endl;
  size_t t = 0;
  for (size_t i = 0; i < R.snpInfo.size(); ++i) {
    if (!R.indicator_snp[i]) continue;
    const SnpInfo &s = R.snpInfo[i]; const gb200_sumstat &st = R.sumStat[t++];
    o << s.chr << "\t" << s.rs << "\t" << s.bp << "\t" << s.n_miss << "\t" << (long)R.ni_test - s.n_miss << "\t" << s.a_minor << "\t" << s.a_major << "\t"
      << std::fixed << std::setprecision(3) << s.maf << "\t" << std::scientific << std::setprecision(6);
    if (m == 51) o << st.beta << "\t" << st.se << "\t" << st.p_wald << std::endl;
    else if (m == 52) o << st.p_lrt << std::endl;
    else if (m == 53) o << st.beta << "\t" << st.se << "\t" << st.p_score << std::endl;
    else o << st.beta << "\t" << st.se << "\t" << st.p_wald << "\t" << st.p_lrt << "\t" << st.p_score << std::endl;
  }
}

static void write_log(const Run &R) {                  // WriteLog, src/gemma.cpp:3148-3597 (the -gk/-lmm lines)
  std::ofstream out(out_path(R, "log"));
  if (!out) return;
  out << "##" << std::endl << "## gemma-b200 (GEMMA-compatible -gk/-eigen/-lmm on libgemma_b200)" << std::endl << "##" << std::endl;
  out << "## Command Line Input = " << R.cmdline << std::endl << "##" << std::endl;
  out << "## Summary Statistics:" << std::endl;
  out << "## number of total individuals = " << R.ni_total << std::endl;
  out << "## number of analyzed individuals = " << R.ni_test << std::endl;
  out << "## number of covariates = " << R.n_cvt << std::endl;
  out << "## number of phenotypes = " << R.P.p_column.size() << std::endl;
  out << "## number of total SNPs/var = " << R.ns_total << std::endl;
  out << "## number of analyzed SNPs/var = " << R.ns_test << std::endl;
  const int m = R.P.a_mode;
  if (m == 1 || m == 2 || m == 3 || m == 4 || m == 9) {
    out << "## REMLE log-likelihood in the null model = " << R.nm.logl_remle_H0 << std::endl;
    out << "## MLE log-likelihood in the null model = " << R.nm.logl_mle_H0 << std::endl;
    out << "## pve estimate in the null model = " << R.nm.pve_null << std::endl;
    out << "## se(pve) in the null model = " << R.nm.pve_se_null << std::endl;
    out << "## vg estimate in the null model = " << R.nm.vg_remle << std::endl;
    out << "## ve estimate in the null model = " << R.nm.ve_remle << std::endl;
    out << "## beta estimate in the null model = "; for (double b : R.beta_remle) out << "  " << b; out << std::endl;
    out << "## se(beta) = "; for (double b : R.se_remle) out << "  " << b; out << std::endl;
  }
  out << "##" << std::endl << "## Computation Time (wall clock, seconds):" << std::endl;
  out << "## total computation time = " << R.t_total << std::endl;
  if (m == 21 || m == 22) out << "##      time on calculating relatedness matrix = " << R.t_kin << std::endl;
  if (m < 20 || m == 31) out << "##      time on eigen-decomposition = " << R.t_eigen << std::endl;
  if (m < 20) out << "##      time on UtX + optimization = " << R.t_lmm << std::endl;
  out << "##" << std::endl;
}

static void usage() {
  std::cout << "gemma-b200: GEMMA-compatible -gk / -eigen / -lmm on a B200\n"
               " -g/-p/-a/-c files (BIMBAM)  |  -bfile prefix (PLINK)   -n col...   -o prefix  -outdir dir\n"
               " -gk [1|2]   -eigen   -lmm [1|2|3|4|9]   -lm [1|2|3|4]   -k K.txt [-km 1|2]   -d D.txt -u U.txt\n"
               " -miss x -maf x -hwe x -r2 x -notsnp -snps file -ksnps file -gwasnps file -loco chr -gxe file -widv file -lmin x -lmax x -region n -nind n -silence\n"
               " -bin  also write K / U / D as <file>.bin (exact doubles); -k/-u/-d accept such .bin files\n";
}

int main(int argc, char **argv) {
  Run R; Params &P = R.P;
  for (int i = 0; i < argc; ++i) { if (i) R.cmdline += " "; R.cmdline += argv[i]; }
  if (argc <= 1) { usage(); return 0; }
  auto need = [&](int &i) -> const char * { if (i + 1 >= argc || (argv[i + 1][0] == '-' && !isdigit((unsigned char)argv[i + 1][1]) && argv[i + 1][1] != '.')) die(string("missing value for ") + argv[i]); return argv[++i]; };
  auto optnum = [&](int &i, int dflt) -> int { if (i + 1 < argc && argv[i + 1][0] != '-') return atoi(argv[++i]); return dflt; };
  int n_modes = 0;
  for (int i = 1; i < argc; ++i) {
    const string a = argv[i];
    if (a == "-g" || a == "-geno") P.file_geno = need(i);
    else if (a == "-p" || a == "-pheno") P.file_pheno = need(i);
    else if (a == "-a" || a == "-anno") P.file_anno = need(i);
    else if (a == "-c" || a == "-cvt") P.file_cvt = need(i);
    else if (a == "-bfile" || a == "-bf") P.file_bfile = need(i);
    else if (a == "-k" || a == "-kin") P.file_kin = need(i);
    else if (a == "-km") P.k_mode = atoi(need(i));
    else if (a == "-d") P.file_kd = need(i);
    else if (a == "-u") P.file_ku = need(i);
    else if (a == "-snps") P.file_snps = need(i);
    else if (a == "-ksnps") P.file_ksnps = need(i);
    else if (a == "-gwasnps") P.file_gwasnps = need(i);
    else if (a == "-loco") P.loco = need(i);
    else if (a == "-gxe") P.file_gxe = need(i);
    else if (a == "-o") P.file_out = need(i);
    else if (a == "-outdir") P.path_out = need(i);
    else if (a == "-n") { while (i + 1 < argc && argv[i + 1][0] != '-') P.p_column.push_back((size_t)atoi(argv[++i])); }
    else if (a == "-miss") P.miss_level = atof(need(i));
    else if (a == "-maf") { P.maf_level = atof(need(i)); }
    else if (a == "-hwe") P.hwe_level = atof(need(i));
    else if (a == "-r2") P.r2_level = atof(need(i));
    else if (a == "-notsnp") P.maf_level = -1;
    else if (a == "-lmin") P.l_min = atof(need(i));
    else if (a == "-lmax") P.l_max = atof(need(i));
    else if (a == "-region") P.n_region = (size_t)atoi(need(i));
    else if (a == "-nind") P.nind = atol(need(i));
    else if (a == "-widv") P.file_weight = need(i);                              // src/gemma.cpp:819-826
    else if (a == "-device") P.device = atoi(need(i));
    else if (a == "-gk") { P.a_mode = 20 + optnum(i, 1); n_modes++; }            // src/gemma.cpp:1124-1139
    else if (a == "-eigen") { P.a_mode = 31; n_modes++; }
    else if (a == "-lmm") { P.a_mode = optnum(i, 1); n_modes++; }                 // src/gemma.cpp:1299-1314
    else if (a == "-lm") { P.a_mode = 50 + optnum(i, 1); n_modes++; }             // src/gemma.cpp:1283-1298
    else if (a == "-silence" || a == "--quiet") P.silence = true;                // src/gemma.cpp:777
    else if (a == "-qc-only") P.qc_only = true;
    else if (a == "-bin") P.bin = true;
    else if (a == "-no-check" || a == "-check" || a == "-debug" || a == "-strict" || a == "-legacy" || a == "-nocheck" ||
             a == "-no-fpe-check" || a == "-debug-data" || a == "-debug-dump") {}   // debug / checking switches of the reference: no effect here
    else if (a == "-pace" || a == "-seed" || a == "-issue") {                      // accepted like the reference (progress pace, RNG seed of the
      if (i + 1 < argc && argv[i + 1][0] != '-') ++i;                              // sampling modes, test hook): nothing on this path uses them
    }
    else if (a == "-h" || a == "-help") { usage(); return 0; }
    else die("unrecognized option " + a);                                          // src/gemma.cpp:1626-1629
  }
  if (n_modes > 1) die("only one of -gk -eigen -lmm is allowed");                  // src/gemma.cpp:1125-1131
  if (n_modes == 0 && !P.qc_only) die("no analysis selected (use -gk, -eigen or -lmm)");
  if (!(P.a_mode == 21 || P.a_mode == 22 || P.a_mode == 31 || P.a_mode == 1 || P.a_mode == 2 || P.a_mode == 3 || P.a_mode == 4 || P.a_mode == 9 ||
        (P.a_mode >= 51 && P.a_mode <= 54)) && !P.qc_only)
    die("analysis mode not supported by gemma-b200 (only -gk 1/2, -eigen, -lmm 1/2/3/4/9, -lm 1/2/3/4)");
  if (P.p_column.empty()) P.p_column.push_back(1);                                 // src/param.cpp:635-636
  if (P.p_column.size() > 1 && (P.a_mode < 20 || P.a_mode >= 51) && !P.qc_only) {
    if (!(P.a_mode >= 1 && P.a_mode <= 4 && P.p_column.size() == 2)) die("multivariate analysis: only -lmm 1/2/3/4 with two phenotypes (-n a b) is supported");
    if (!P.file_gxe.empty()) die("multivariate G x E is not supported");
  }
  if (P.file_bfile.empty() && (P.file_geno.empty() || P.file_pheno.empty())) die("need -g and -p, or -bfile");
  const bool is_lmm = (P.a_mode < 20 && P.a_mode > 0) || P.a_mode == 31;
  if (is_lmm && P.file_kin.empty() && (P.file_kd.empty() || P.file_ku.empty())) die("missing relatedness file (-k) or eigen files (-d and -u)");   // src/param.cpp:951-956
  std::stringstream sink; std::streambuf *old = nullptr;
  if (P.silence) old = std::cout.rdbuf(sink.rdbuf());
  mkdir(P.path_out.c_str(), 0755);                                                 // src/main.cpp:59-66
  const double t_start = now_s();

  if (!P.loco.empty()) {                                                           // CheckParam, src/param.cpp:923-933
    if (!((P.a_mode >= 1 && P.a_mode <= 4) || P.a_mode == 9 || P.a_mode == 21 || P.a_mode == 22 || (P.qc_only && P.a_mode == 0)))
      die("LOCO only works with LMM and K");
    if (!P.file_gxe.empty()) die("LOCO does not support GXE (yet)");
    if (P.file_anno.empty()) die("LOCO requires annotation file (-a switch)");
    if (!P.file_ksnps.empty()) die("LOCO does not allow -ksnps switch");
    if (!P.file_gwasnps.empty()) die("LOCO does not allow -gwasnps switch");
    if (!P.file_bfile.empty()) die("LOCO with PLINK input mis-aligns rows in the reference (its own tests are disabled); use BIMBAM input");
  }
  std::cout << "Reading Files ... " << std::endl;
  if (!P.file_snps.empty()) read_snp_set(P.file_snps, R.setSnps);
  if (!P.file_ksnps.empty()) read_snp_set(P.file_ksnps, R.setKSnps);
  if (!P.file_gwasnps.empty()) read_snp_set(P.file_gwasnps, R.setGWASnps);
  if (!P.file_anno.empty()) read_anno(R);
  if (!P.loco.empty())                                                             // src/param.cpp:52-66, 497-500
    for (const auto &kv : R.anno) (std::get<0>(kv.second) != P.loco ? R.setKSnps : R.setGWASnps).insert(kv.first);
  if (!P.file_bfile.empty()) { read_bim(R); read_fam(R); if (!P.file_pheno.empty()) { R.pheno.clear(); R.ind_pheno.clear(); read_pheno(R); } }
  else read_pheno(R);
  if (!P.file_cvt.empty()) read_cvt(R);
  if (!P.file_gxe.empty()) read_gxe(R);
  if (!P.file_weight.empty()) read_weight(R);
  process_cvt_phen(R);
  trim_individuals(R);
  gb200_ctx *ctx = nullptr;
  if (!P.qc_only && gb200_create(&ctx, P.device, nullptr) != GB200_OK) die("no CUDA device: gemma-b200 has no CPU fallback");
  if (!P.file_bfile.empty()) qc_plink(R, ctx); else qc_bimbam(R);
  print_counts(R);
  if (P.qc_only) {
    std::ofstream out(out_path(R, "qc"));
    for (size_t i = 0; i < R.snpInfo.size(); ++i)
      out << R.snpInfo[i].rs << "\t" << R.indicator_snp[i] << "\t" << R.snpInfo[i].n_miss << "\t" << std::setprecision(17) << R.snpInfo[i].maf
          << "\t" << R.snpInfo[i].chr << "\t" << R.snpInfo[i].bp << "\t" << R.snpInfo[i].a_minor << "\t" << R.snpInfo[i].a_major << "\n";
    if (!P.file_kin.empty()) {                       // the kinship matrix of the analysed individuals exactly as read_kin hands it on
      vector<double> G;
      read_kin(R, G);
      write_bin(out_path(R, "kin") + ".bin", G.data(), R.ni_test, R.ni_test);
      write_matrix(R, G.data(), R.ni_test, R.ni_test, "kin");            // and through the production text writer (WriteMatrix)
      if (!R.weight.empty()) {                        // -widv: the matrix that goes to the eigensolver (centred, weighted) and the weights
        const vector<double> w = weighted_kinship(R, G);
        write_bin(out_path(R, "wkin") + ".bin", G.data(), R.ni_test, R.ni_test);
        write_bin(out_path(R, "widv") + ".bin", w.data(), R.ni_test, 1);
      }
    }
    if (!P.file_ku.empty() && !P.file_kd.empty()) {   // -u / -d: ReadFile_eigenU / ReadFile_eigenD, then the production writers
      vector<double> U(R.ni_test * R.ni_test), D(R.ni_test);
      read_dense_rows(P.file_ku, U.data(), R.ni_test, R.ni_test, "U");
      read_dense_rows(P.file_kd, D.data(), R.ni_test, 1, "D");
      write_matrix(R, U.data(), R.ni_test, R.ni_test, "eigU");
      write_vector(R, D.data(), R.ni_test, "eigD");
    }
    if (old) std::cout.rdbuf(old);
    return 0;
  }
  if (R.ns_test == 0) die("number of analyzed SNPs equals 0");

  if (P.a_mode == 21 || P.a_mode == 22) run_kinship(R, ctx); else if (P.a_mode >= 51) run_lm(R, ctx); else run_lmm(R, ctx);
  R.t_total = now_s() - t_start;
  write_log(R);
  gb200_destroy(ctx);
  if (old) std::cout.rdbuf(old);
  return 0;
}
