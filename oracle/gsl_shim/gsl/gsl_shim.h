/* gsl_shim.h -- TEST INFRASTRUCTURE ONLY (oracle/_ref).
 *
 * The reference links the system GSL 2.x (Makefile:163-165), which this image does not have.  To compile the REFERENCE's own
 * src/lmm.cpp (in place, never copied) into oracle/_ref/ this header declares the subset of the GSL API that file and the
 * headers it includes use: data structures with GSL's public field layout (size/stride/data, size1/size2/tda/data), the
 * element accessors, views, the BLAS level-1/2 calls, the root-solver front end and the cdf tails.  The implementations
 * (oracle/gsl_shim/gsl_shim.cpp) are plain restatements: vector/matrix/BLAS routines are their textbook loops, the Brent /
 * Newton solvers and the cdf tails reuse the restated GSL algorithms of oracle/gemma_oracle.c (roots/brent.c, roots/newton.c,
 * roots/convergence.c, cdf/fdist.c, cdf/beta_inc.c, cdf/gamma.c).  What the resulting library validates is therefore the
 * reference's OWN code: CalcUab, CalcPab/PPab/PPPab, LogL_* / LogRL_*, CalcLambda's control flow, CalcRLWald / CalcRLScore. */
#ifndef GB_GSL_SHIM_H
#define GB_GSL_SHIM_H
#include <stddef.h>
#include <stdlib.h>
#include <math.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { GSL_SUCCESS = 0, GSL_FAILURE = -1, GSL_CONTINUE = -2, GSL_EDOM = 1, GSL_ERANGE = 2, GSL_EFAULT = 3, GSL_EINVAL = 4, GSL_EFAILED = 5,
       GSL_EFACTOR = 6, GSL_ESANITY = 7, GSL_ENOMEM = 8, GSL_EBADFUNC = 9, GSL_ERUNAWAY = 10, GSL_EMAXITER = 11, GSL_EZERODIV = 12,
       GSL_EBADTOL = 13, GSL_ETOL = 14, GSL_EUNDRFLW = 15, GSL_EOVRFLW = 16, GSL_ELOSS = 17, GSL_EROUND = 18, GSL_EBADLEN = 19,
       GSL_ENOTSQR = 20, GSL_ESING = 21, GSL_EDIVERGE = 22 };
#define GSL_VERSION "2.x API shim (oracle/gsl_shim)"
#define GSL_DBL_EPSILON 2.2204460492503131e-16
#define GSL_NAN (NAN)
#define GSL_POSINF (INFINITY)
#define GSL_NEGINF (-INFINITY)
#define GSL_MAX(a, b) ((a) > (b) ? (a) : (b))
#define GSL_MIN(a, b) ((a) < (b) ? (a) : (b))

typedef void gsl_error_handler_t(const char *reason, const char *file, int line, int gsl_errno);
gsl_error_handler_t *gsl_set_error_handler(gsl_error_handler_t *new_handler);
gsl_error_handler_t *gsl_set_error_handler_off(void);
const char *gsl_strerror(const int gsl_errno);
int gsl_isnan(const double x);
int gsl_isinf(const double x);
int gsl_finite(const double x);

typedef struct { size_t size; double *data; } gsl_block;
typedef struct { size_t size; size_t stride; double *data; gsl_block *block; int owner; } gsl_vector;
typedef struct { gsl_vector vector; } gsl_vector_view;
typedef struct { gsl_vector vector; } gsl_vector_const_view;
typedef struct { size_t size1; size_t size2; size_t tda; double *data; gsl_block *block; int owner; } gsl_matrix;
typedef struct { gsl_matrix matrix; } gsl_matrix_view;
typedef struct { gsl_matrix matrix; } gsl_matrix_const_view;
typedef struct { size_t size; size_t *data; } gsl_permutation;
typedef struct { int dummy; } gsl_rng;
typedef struct { const char *name; } gsl_rng_type;

gsl_vector *gsl_vector_alloc(size_t n);
gsl_vector *gsl_vector_calloc(size_t n);
void gsl_vector_free(gsl_vector *v);
static inline double gsl_vector_get(const gsl_vector *v, const size_t i) { return v->data[i * v->stride]; }
static inline void gsl_vector_set(gsl_vector *v, const size_t i, double x) { v->data[i * v->stride] = x; }
static inline double *gsl_vector_ptr(gsl_vector *v, const size_t i) { return v->data + i * v->stride; }
void gsl_vector_set_all(gsl_vector *v, double x);
void gsl_vector_set_zero(gsl_vector *v);
int gsl_vector_memcpy(gsl_vector *dest, const gsl_vector *src);
int gsl_vector_mul(gsl_vector *a, const gsl_vector *b);
int gsl_vector_div(gsl_vector *a, const gsl_vector *b);
int gsl_vector_add(gsl_vector *a, const gsl_vector *b);
int gsl_vector_sub(gsl_vector *a, const gsl_vector *b);
int gsl_vector_scale(gsl_vector *a, const double x);
int gsl_vector_add_constant(gsl_vector *a, const double x);
double gsl_vector_max(const gsl_vector *v);
void gsl_vector_minmax(const gsl_vector *v, double *min_out, double *max_out);
double gsl_vector_min(const gsl_vector *v);
gsl_vector_view gsl_vector_subvector(gsl_vector *v, size_t i, size_t n);
gsl_vector_const_view gsl_vector_const_subvector(const gsl_vector *v, size_t i, size_t n);
gsl_vector_view gsl_vector_view_array(double *v, size_t n);
gsl_vector_const_view gsl_vector_const_view_array(const double *v, size_t n);

gsl_matrix *gsl_matrix_alloc(size_t n1, size_t n2);
gsl_matrix *gsl_matrix_calloc(size_t n1, size_t n2);
void gsl_matrix_free(gsl_matrix *m);
static inline double gsl_matrix_get(const gsl_matrix *m, const size_t i, const size_t j) { return m->data[i * m->tda + j]; }
static inline void gsl_matrix_set(gsl_matrix *m, const size_t i, const size_t j, double x) { m->data[i * m->tda + j] = x; }
static inline double *gsl_matrix_ptr(gsl_matrix *m, const size_t i, const size_t j) { return m->data + i * m->tda + j; }
void gsl_matrix_set_zero(gsl_matrix *m);
void gsl_matrix_set_all(gsl_matrix *m, double x);
void gsl_matrix_set_identity(gsl_matrix *m);
int gsl_matrix_memcpy(gsl_matrix *dest, const gsl_matrix *src);
int gsl_matrix_scale(gsl_matrix *a, const double x);
int gsl_matrix_add(gsl_matrix *a, const gsl_matrix *b);
int gsl_matrix_sub(gsl_matrix *a, const gsl_matrix *b);
int gsl_matrix_add_constant(gsl_matrix *a, const double x);
int gsl_matrix_transpose_memcpy(gsl_matrix *dest, const gsl_matrix *src);
int gsl_matrix_transpose(gsl_matrix *m);
int gsl_matrix_mul_elements(gsl_matrix *a, const gsl_matrix *b);
double gsl_matrix_max(const gsl_matrix *m);
double gsl_matrix_min(const gsl_matrix *m);
gsl_vector_view gsl_matrix_row(gsl_matrix *m, size_t i);
gsl_vector_view gsl_matrix_column(gsl_matrix *m, size_t j);
gsl_vector_view gsl_matrix_diagonal(gsl_matrix *m);
gsl_vector_const_view gsl_matrix_const_row(const gsl_matrix *m, size_t i);
gsl_vector_view gsl_matrix_subrow(gsl_matrix *m, size_t i, size_t offset, size_t n);
gsl_vector_const_view gsl_matrix_const_subrow(const gsl_matrix *m, size_t i, size_t offset, size_t n);
gsl_vector_view gsl_matrix_subcolumn(gsl_matrix *m, size_t j, size_t offset, size_t n);
gsl_vector_const_view gsl_matrix_const_column(const gsl_matrix *m, size_t j);
gsl_matrix_view gsl_matrix_submatrix(gsl_matrix *m, size_t i, size_t j, size_t n1, size_t n2);
gsl_matrix_const_view gsl_matrix_const_submatrix(const gsl_matrix *m, size_t i, size_t j, size_t n1, size_t n2);
gsl_matrix_view gsl_matrix_view_array(double *base, size_t n1, size_t n2);
int gsl_matrix_get_row(gsl_vector *v, const gsl_matrix *m, const size_t i);
int gsl_matrix_get_col(gsl_vector *v, const gsl_matrix *m, const size_t j);
int gsl_matrix_set_row(gsl_matrix *m, const size_t i, const gsl_vector *v);
int gsl_matrix_set_col(gsl_matrix *m, const size_t j, const gsl_vector *v);

typedef struct { size_t size; size_t stride; int *data; void *block; int owner; } gsl_vector_int;
typedef struct { size_t size1; size_t size2; size_t tda; int *data; void *block; int owner; } gsl_matrix_int;
gsl_vector_int *gsl_vector_int_alloc(size_t n);
void gsl_vector_int_free(gsl_vector_int *v);
static inline int gsl_vector_int_get(const gsl_vector_int *v, const size_t i) { return v->data[i * v->stride]; }
static inline void gsl_vector_int_set(gsl_vector_int *v, const size_t i, int x) { v->data[i * v->stride] = x; }
gsl_matrix_int *gsl_matrix_int_alloc(size_t n1, size_t n2);
void gsl_matrix_int_free(gsl_matrix_int *m);
static inline int gsl_matrix_int_get(const gsl_matrix_int *m, const size_t i, const size_t j) { return m->data[i * m->tda + j]; }
static inline void gsl_matrix_int_set(gsl_matrix_int *m, const size_t i, const size_t j, int x) { m->data[i * m->tda + j] = x; }

gsl_permutation *gsl_permutation_alloc(size_t n);
gsl_permutation *gsl_permutation_calloc(size_t n);
void gsl_permutation_init(gsl_permutation *p);
static inline size_t gsl_permutation_get(const gsl_permutation *p, const size_t i) { return p->data[i]; }
void gsl_permutation_free(gsl_permutation *p);

/* cblas enums (gsl_cblas.h) */
enum CBLAS_ORDER { CblasRowMajor = 101, CblasColMajor = 102 };
enum CBLAS_TRANSPOSE { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 };
enum CBLAS_UPLO { CblasUpper = 121, CblasLower = 122 };
enum CBLAS_DIAG { CblasNonUnit = 131, CblasUnit = 132 };
enum CBLAS_SIDE { CblasLeft = 141, CblasRight = 142 };
typedef enum CBLAS_TRANSPOSE CBLAS_TRANSPOSE_t;
typedef enum CBLAS_UPLO CBLAS_UPLO_t;
typedef enum CBLAS_DIAG CBLAS_DIAG_t;
typedef enum CBLAS_SIDE CBLAS_SIDE_t;
void cblas_dgemm(const enum CBLAS_ORDER Order, const enum CBLAS_TRANSPOSE TransA, const enum CBLAS_TRANSPOSE TransB, const int M, const int N,
                 const int K, const double alpha, const double *A, const int lda, const double *B, const int ldb, const double beta, double *C,
                 const int ldc);

int gsl_blas_ddot(const gsl_vector *x, const gsl_vector *y, double *result);
double gsl_blas_dnrm2(const gsl_vector *x);
int gsl_blas_daxpy(double alpha, const gsl_vector *x, gsl_vector *y);
void gsl_blas_dscal(double alpha, gsl_vector *x);
int gsl_blas_dgemv(CBLAS_TRANSPOSE_t TransA, double alpha, const gsl_matrix *A, const gsl_vector *x, double beta, gsl_vector *y);
int gsl_blas_dsyr(CBLAS_UPLO_t Uplo, double alpha, const gsl_vector *x, gsl_matrix *A);
int gsl_blas_dsyr2(CBLAS_UPLO_t Uplo, double alpha, const gsl_vector *x, const gsl_vector *y, gsl_matrix *A);
int gsl_blas_dger(double alpha, const gsl_vector *x, const gsl_vector *y, gsl_matrix *A);
int gsl_blas_dgemm(CBLAS_TRANSPOSE_t TransA, CBLAS_TRANSPOSE_t TransB, double alpha, const gsl_matrix *A, const gsl_matrix *B, double beta,
                   gsl_matrix *C);
int gsl_blas_dtrsv(CBLAS_UPLO_t Uplo, CBLAS_TRANSPOSE_t TransA, CBLAS_DIAG_t Diag, const gsl_matrix *A, gsl_vector *x);
int gsl_blas_dsyrk(CBLAS_UPLO_t Uplo, CBLAS_TRANSPOSE_t Trans, double alpha, const gsl_matrix *A, double beta, gsl_matrix *C);

int gsl_linalg_LU_decomp(gsl_matrix *A, gsl_permutation *p, int *signum);
int gsl_linalg_LU_solve(const gsl_matrix *LU, const gsl_permutation *p, const gsl_vector *b, gsl_vector *x);
int gsl_linalg_LU_invert(const gsl_matrix *LU, const gsl_permutation *p, gsl_matrix *inverse);
double gsl_linalg_LU_det(gsl_matrix *LU, int signum);
double gsl_linalg_LU_lndet(gsl_matrix *LU);

int gsl_linalg_QR_decomp(gsl_matrix *A, gsl_vector *tau);
int gsl_linalg_QR_solve(const gsl_matrix *QR, const gsl_vector *tau, const gsl_vector *b, gsl_vector *x);
double gsl_sf_exp(const double x);
double gsl_sf_log_1plusx(const double x);
double gsl_sf_lngamma(double x);
double gsl_sf_gamma(double x);
double gsl_sf_lnbeta(const double a, const double b);
int gsl_linalg_cholesky_decomp(gsl_matrix *A);
int gsl_linalg_cholesky_decomp1(gsl_matrix *A);
int gsl_linalg_cholesky_solve(const gsl_matrix *cholesky, const gsl_vector *b, gsl_vector *x);
int gsl_linalg_cholesky_invert(gsl_matrix *cholesky);
typedef struct { size_t size; double *d; double *sd; } gsl_eigen_symm_workspace;
typedef struct { size_t size; double *d; double *sd; double *gc; double *gs; } gsl_eigen_symmv_workspace;
gsl_eigen_symm_workspace *gsl_eigen_symm_alloc(const size_t n);
void gsl_eigen_symm_free(gsl_eigen_symm_workspace *w);
int gsl_eigen_symm(gsl_matrix *A, gsl_vector *eval, gsl_eigen_symm_workspace *w);
gsl_eigen_symmv_workspace *gsl_eigen_symmv_alloc(const size_t n);
void gsl_eigen_symmv_free(gsl_eigen_symmv_workspace *w);
int gsl_eigen_symmv(gsl_matrix *A, gsl_vector *eval, gsl_matrix *evec, gsl_eigen_symmv_workspace *w);

extern const gsl_rng_type *gsl_rng_default;
extern unsigned long int gsl_rng_default_seed;
extern const gsl_rng_type *gsl_rng_mt19937;
const gsl_rng_type *gsl_rng_env_setup(void);
gsl_rng *gsl_rng_alloc(const gsl_rng_type *T);
void gsl_rng_free(gsl_rng *r);
void gsl_rng_set(const gsl_rng *r, unsigned long int seed);
const char *gsl_rng_name(const gsl_rng *r);
unsigned long int gsl_rng_get(const gsl_rng *r);
double gsl_rng_uniform(const gsl_rng *r);
unsigned long int gsl_rng_uniform_int(const gsl_rng *r, unsigned long int n);
double gsl_ran_gaussian(const gsl_rng *r, const double sigma);
double gsl_ran_ugaussian(const gsl_rng *r);
double gsl_ran_gamma(const gsl_rng *r, const double a, const double b);
double gsl_ran_beta(const gsl_rng *r, const double a, const double b);
double gsl_ran_chisq(const gsl_rng *r, const double nu);
double gsl_ran_exponential(const gsl_rng *r, const double mu);
unsigned int gsl_ran_binomial(const gsl_rng *r, double p, unsigned int n);
unsigned int gsl_ran_geometric(const gsl_rng *r, const double p);
int gsl_ran_choose(const gsl_rng *r, void *dest, size_t k, void *src, size_t n, size_t size);
void gsl_ran_shuffle(const gsl_rng *r, void *base, size_t nmembm, size_t size);
typedef struct { size_t K; size_t *A; double *F; } gsl_ran_discrete_t;
gsl_ran_discrete_t *gsl_ran_discrete_preproc(size_t K, const double *P);
void gsl_ran_discrete_free(gsl_ran_discrete_t *g);
size_t gsl_ran_discrete(const gsl_rng *r, const gsl_ran_discrete_t *g);
double gsl_ran_gaussian_pdf(const double x, const double sigma);
double gsl_ran_binomial_pdf(const unsigned int k, const double p, const unsigned int n);
double gsl_ran_geometric_pdf(const unsigned int k, const double p);
double gsl_cdf_gaussian_P(const double x, const double sigma);
double gsl_cdf_gaussian_Q(const double x, const double sigma);
double gsl_cdf_ugaussian_P(const double x);
double gsl_cdf_chisq_P(const double x, const double nu);
double gsl_cdf_tdist_P(const double x, const double nu);
double gsl_stats_mean(const double data[], const size_t stride, const size_t n);
double gsl_stats_variance(const double data[], const size_t stride, const size_t n);
double gsl_stats_sd(const double data[], const size_t stride, const size_t n);
void gsl_sort(double *data, const size_t stride, const size_t n);
void gsl_sort_vector(gsl_vector *v);
int gsl_sort_vector_index(gsl_permutation *p, const gsl_vector *v);

double gsl_cdf_chisq_Q(const double x, const double nu);
double gsl_cdf_chisq_Qinv(const double Q, const double nu);
double gsl_cdf_chisq_Pinv(const double P, const double nu);
double gsl_cdf_fdist_Q(const double x, const double nu1, const double nu2);
double gsl_cdf_tdist_Q(const double x, const double nu);
double gsl_cdf_ugaussian_Q(const double x);

/* roots */
typedef struct { double (*function)(double x, void *params); void *params; } gsl_function;
typedef struct { double (*f)(double x, void *params); double (*df)(double x, void *params); void (*fdf)(double x, void *params, double *f, double *df);
                 void *params; } gsl_function_fdf;
#define GSL_FN_EVAL(F, x) (*((F)->function))(x, (F)->params)
typedef struct { const char *name; int kind; } gsl_root_fsolver_type;
typedef struct { const gsl_root_fsolver_type *type; gsl_function *function; double root, x_lower, x_upper; void *state; } gsl_root_fsolver;
typedef struct { const char *name; int kind; } gsl_root_fdfsolver_type;
typedef struct { const gsl_root_fdfsolver_type *type; gsl_function_fdf *fdf; double root; void *state; } gsl_root_fdfsolver;
extern const gsl_root_fsolver_type *gsl_root_fsolver_brent;
extern const gsl_root_fsolver_type *gsl_root_fsolver_bisection;
extern const gsl_root_fdfsolver_type *gsl_root_fdfsolver_newton;
gsl_root_fsolver *gsl_root_fsolver_alloc(const gsl_root_fsolver_type *T);
void gsl_root_fsolver_free(gsl_root_fsolver *s);
int gsl_root_fsolver_set(gsl_root_fsolver *s, gsl_function *f, double x_lower, double x_upper);
int gsl_root_fsolver_iterate(gsl_root_fsolver *s);
double gsl_root_fsolver_root(const gsl_root_fsolver *s);
double gsl_root_fsolver_x_lower(const gsl_root_fsolver *s);
double gsl_root_fsolver_x_upper(const gsl_root_fsolver *s);
gsl_root_fdfsolver *gsl_root_fdfsolver_alloc(const gsl_root_fdfsolver_type *T);
void gsl_root_fdfsolver_free(gsl_root_fdfsolver *s);
int gsl_root_fdfsolver_set(gsl_root_fdfsolver *s, gsl_function_fdf *fdf, double root);
int gsl_root_fdfsolver_iterate(gsl_root_fdfsolver *s);
double gsl_root_fdfsolver_root(const gsl_root_fdfsolver *s);
int gsl_root_test_interval(double x_lower, double x_upper, double epsabs, double epsrel);
int gsl_root_test_delta(double x1, double x0, double epsabs, double epsrel);

/* multiroots (logistic.cpp; off the validated path: declarations only) */
typedef struct { int (*f)(const gsl_vector *x, void *params, gsl_vector *f); int (*df)(const gsl_vector *x, void *params, gsl_matrix *df);
                 int (*fdf)(const gsl_vector *x, void *params, gsl_vector *f, gsl_matrix *df); size_t n; void *params; } gsl_multiroot_function_fdf;
typedef struct { const char *name; } gsl_multiroot_fdfsolver_type;
typedef struct { const gsl_multiroot_fdfsolver_type *type; gsl_multiroot_function_fdf *fdf; gsl_vector *x; gsl_vector *f; gsl_matrix *J; gsl_vector *dx;
                 void *state; } gsl_multiroot_fdfsolver;
extern const gsl_multiroot_fdfsolver_type *gsl_multiroot_fdfsolver_hybridsj;
gsl_multiroot_fdfsolver *gsl_multiroot_fdfsolver_alloc(const gsl_multiroot_fdfsolver_type *T, size_t n);
void gsl_multiroot_fdfsolver_free(gsl_multiroot_fdfsolver *s);
int gsl_multiroot_fdfsolver_set(gsl_multiroot_fdfsolver *s, gsl_multiroot_function_fdf *fdf, const gsl_vector *x);
int gsl_multiroot_fdfsolver_iterate(gsl_multiroot_fdfsolver *s);
int gsl_multiroot_test_residual(const gsl_vector *f, double epsabs);

#ifdef __cplusplus
}
#endif
#endif
