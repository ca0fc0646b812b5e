/* oracle/shim/cblas.h -- TEST INFRASTRUCTURE, not product code.
 *
 * Minimal CBLAS surface needed to compile the unmodified reference sources
 * (/root/reference/src/vector.hpp:30,157, vector.cpp:95,161,176,
 * symmmatrix.cpp:95,237, symmmatrix.hpp:30-38).  The reference depends on an
 * un-vendored, unpinned BLAS ("-lblas -llapack", makefile:9); this header maps
 * the six routines it calls either onto SciPy's bundled OpenBLAS (symbols are
 * prefixed scipy_, no header ships with it) when BMAGWA_SHIM_OPENBLAS is
 * defined, or onto the plain sequential C implementations in blas_ref.c.
 */
#ifndef BMAGWA_ORACLE_SHIM_CBLAS_H
#define BMAGWA_ORACLE_SHIM_CBLAS_H

#ifdef __cplusplus
extern "C" {
#endif

typedef enum { CblasRowMajor = 101, CblasColMajor = 102 } CBLAS_ORDER;
typedef enum { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 } CBLAS_TRANSPOSE;
typedef enum { CblasUpper = 121, CblasLower = 122 } CBLAS_UPLO;
typedef enum { CblasNonUnit = 131, CblasUnit = 132 } CBLAS_DIAG;

#ifdef BMAGWA_SHIM_OPENBLAS
#define SHIM_BLAS(name) scipy_##name
#else
#define SHIM_BLAS(name) shim_##name
#endif

double SHIM_BLAS(cblas_ddot)(int n, const double* x, int incx, const double* y, int incy);
void SHIM_BLAS(cblas_dgemv)(CBLAS_ORDER order, CBLAS_TRANSPOSE trans, int m, int n,
                            double alpha, const double* a, int lda, const double* x, int incx,
                            double beta, double* y, int incy);
void SHIM_BLAS(cblas_dtrmv)(CBLAS_ORDER order, CBLAS_UPLO uplo, CBLAS_TRANSPOSE trans,
                            CBLAS_DIAG diag, int n, const double* a, int lda, double* x, int incx);
void SHIM_BLAS(cblas_dtrsv)(CBLAS_ORDER order, CBLAS_UPLO uplo, CBLAS_TRANSPOSE trans,
                            CBLAS_DIAG diag, int n, const double* a, int lda, double* x, int incx);
void SHIM_BLAS(cblas_dsyrk)(CBLAS_ORDER order, CBLAS_UPLO uplo, CBLAS_TRANSPOSE trans, int n, int k,
                            double alpha, const double* a, int lda, double beta, double* c, int ldc);
void SHIM_BLAS(cblas_drotg)(double* a, double* b, double* c, double* s);

#ifdef BMAGWA_SHIM_OPENBLAS
void scipy_dpotrf_(char* uplo, int* n, double* a, int* lda, int* info);
void scipy_openblas_set_num_threads(int n);
#define dpotrf_ scipy_dpotrf_
#endif

#define cblas_ddot SHIM_BLAS(cblas_ddot)
#define cblas_dgemv SHIM_BLAS(cblas_dgemv)
#define cblas_dtrmv SHIM_BLAS(cblas_dtrmv)
#define cblas_dtrsv SHIM_BLAS(cblas_dtrsv)
#define cblas_dsyrk SHIM_BLAS(cblas_dsyrk)
#define cblas_drotg SHIM_BLAS(cblas_drotg)

#ifdef __cplusplus
}
#endif

#endif
