/* TEST INFRASTRUCTURE ONLY (oracle/_ref): symbols referenced by functions of the compiled reference files that are NOT on the
 * validated path (LAPACK eigen / Cholesky, GSL rng, eigen, chisq quantile).  They only have to exist for the loader; reaching
 * one aborts loudly. */
#include <stdio.h>
#include <stdlib.h>
#define GB_STUB(name) void name(void) { fprintf(stderr, "gsl/lapack shim: %s is not restated (not on the validated path)\n", #name); abort(); }
#ifndef GB_HAVE_OPENBLAS
GB_STUB(ddot_) GB_STUB(dgemm_) GB_STUB(dpotrf_) GB_STUB(dpotrs_) GB_STUB(dsyev_) GB_STUB(dsyevr_)
#endif
GB_STUB(gsl_cdf_chisq_Qinv) GB_STUB(gsl_eigen_symm) GB_STUB(gsl_eigen_symm_alloc) GB_STUB(gsl_eigen_symm_free)
GB_STUB(gsl_linalg_cholesky_decomp) GB_STUB(gsl_ran_choose) GB_STUB(gsl_rng_get)
