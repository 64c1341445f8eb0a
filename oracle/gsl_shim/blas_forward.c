/* TEST INFRASTRUCTURE ONLY (oracle/_ref).  The reference calls CBLAS / LAPACK by their standard names (cblas_dgemm through
 * src/fastblas.cpp, dsyevr_ / dgemm_ / dpotrf_ ... through src/lapack.cpp); the only BLAS/LAPACK in this image is the OpenBLAS
 * bundled with scipy, which exports every symbol with a "scipy_" prefix.  These thunks forward the standard names (a bare jump
 * keeps all register and stack arguments in place, so one thunk per name works for any signature). */
#define GB_FWD(name) __asm__(".text\n.globl " #name "\n.type " #name ", @function\n" #name ":\n\tjmp scipy_" #name "@PLT\n");
GB_FWD(cblas_dgemm) GB_FWD(dgemm_) GB_FWD(dsyevr_) GB_FWD(dsyev_) GB_FWD(dpotrf_) GB_FWD(dpotrs_) GB_FWD(ddot_)
