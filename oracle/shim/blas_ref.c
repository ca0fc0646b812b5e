/* oracle/shim/blas_ref.c -- TEST INFRASTRUCTURE, not product code.
 *
 * Third-party arithmetic the reference links but does not vendor, restated in
 * plain sequential C from the published (netlib) algorithms so the unmodified
 * reference sources can be built in an image without BLAS/LAPACK/gfortran:
 *
 *   - CBLAS level 1/2/3 routines called at /root/reference/src/vector.hpp:157,
 *     vector.cpp:95,161,176, symmmatrix.cpp:95,237 (only the ColMajor/Upper/
 *     NonUnit variants the reference uses are implemented);
 *   - LAPACK dpotrf_ ('U'), called at symmmatrix.hpp:145;
 *   - LINPACK dchex (vendored by the reference as Fortran 77,
 *     /root/reference/src/dchex.f; no Fortran compiler in this image), both the
 *     job=1 and job=2 branches, translated statement by statement
 *     (dchex.f:125-246); the reference only uses job=2 (symmmatrix.cpp:178).
 *
 * When the shim is built with -DBMAGWA_SHIM_OPENBLAS the CBLAS/LAPACK entry
 * points come from SciPy's bundled OpenBLAS instead and only dchex_ (and the
 * drotg it needs) is taken from this file.
 */
#include <math.h>
#include <stdlib.h>
#include "cblas.h"

static void ref_drotg(double* da, double* db, double* c, double* s)
{
  /* netlib reference BLAS drotg (1978 version, as used by LINPACK) */
  double roe = *db, scale, r, z;
  if (fabs(*da) > fabs(*db)) roe = *da;
  scale = fabs(*da) + fabs(*db);
  if (scale == 0.0) {
    *c = 1.0; *s = 0.0; r = 0.0; z = 0.0;
  } else {
    double ta = *da / scale, tb = *db / scale;
    r = scale * sqrt(ta * ta + tb * tb);
    r = (roe < 0.0 ? -1.0 : 1.0) * r;
    *c = *da / r;
    *s = *db / r;
    z = 1.0;
    if (fabs(*da) > fabs(*db)) z = *s;
    if (fabs(*db) >= fabs(*da) && *c != 0.0) z = 1.0 / *c;
  }
  *da = r;
  *db = z;
}

#ifndef BMAGWA_SHIM_OPENBLAS

double shim_cblas_ddot(int n, const double* x, int incx, const double* y, int incy)
{
  /* eight partial sums, the way optimised BLAS kernels unroll the reduction */
  if (incx == 1 && incy == 1) {
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0, a5 = 0, a6 = 0, a7 = 0;
    int i = 0;
    for (; i + 8 <= n; i += 8) {
      a0 += x[i] * y[i];         a1 += x[i + 1] * y[i + 1];
      a2 += x[i + 2] * y[i + 2]; a3 += x[i + 3] * y[i + 3];
      a4 += x[i + 4] * y[i + 4]; a5 += x[i + 5] * y[i + 5];
      a6 += x[i + 6] * y[i + 6]; a7 += x[i + 7] * y[i + 7];
    }
    for (; i < n; ++i) a0 += x[i] * y[i];
    return ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
  }
  double acc = 0;
  for (int i = 0; i < n; ++i) acc += x[(long)i * incx] * y[(long)i * incy];
  return acc;
}

void shim_cblas_dgemv(CBLAS_ORDER order, CBLAS_TRANSPOSE trans, int m, int n, double alpha,
                      const double* a, int lda, const double* x, int incx, double beta,
                      double* y, int incy)
{
  (void)order; /* ColMajor only */
  if (trans == CblasNoTrans) {
    for (int i = 0; i < m; ++i) y[(long)i * incy] = (beta == 0.0) ? 0.0 : beta * y[(long)i * incy];
    for (int j = 0; j < n; ++j) {
      const double t = alpha * x[(long)j * incx];
      const double* col = a + (long)j * lda;
      if (incy == 1)
        for (int i = 0; i < m; ++i) y[i] += t * col[i];
      else
        for (int i = 0; i < m; ++i) y[(long)i * incy] += t * col[i];
    }
  } else {
    for (int j = 0; j < n; ++j) {
      const double* col = a + (long)j * lda;
      double d = shim_cblas_ddot(m, col, 1, x, incx);
      double* yj = y + (long)j * incy;
      *yj = alpha * d + ((beta == 0.0) ? 0.0 : beta * *yj);
    }
  }
}

void shim_cblas_dtrmv(CBLAS_ORDER order, CBLAS_UPLO uplo, CBLAS_TRANSPOSE trans, CBLAS_DIAG diag,
                      int n, const double* a, int lda, double* x, int incx)
{
  (void)order; (void)uplo; (void)diag; (void)incx; /* ColMajor, Upper, NonUnit, incx=1 */
  if (trans == CblasNoTrans) {
    /* x := U x */
    for (int j = 0; j < n; ++j) {
      const double t = x[j];
      const double* col = a + (long)j * lda;
      for (int i = 0; i < j; ++i) x[i] += t * col[i];
      x[j] = t * col[j];
    }
  } else {
    /* x := U' x */
    for (int j = n - 1; j >= 0; --j) {
      const double* col = a + (long)j * lda;
      double t = x[j] * col[j];
      for (int i = j - 1; i >= 0; --i) t += col[i] * x[i];
      x[j] = t;
    }
  }
}

void shim_cblas_dtrsv(CBLAS_ORDER order, CBLAS_UPLO uplo, CBLAS_TRANSPOSE trans, CBLAS_DIAG diag,
                      int n, const double* a, int lda, double* x, int incx)
{
  (void)order; (void)uplo; (void)diag; (void)incx;
  if (trans == CblasNoTrans) {
    /* solve U x = b */
    for (int j = n - 1; j >= 0; --j) {
      const double* col = a + (long)j * lda;
      x[j] /= col[j];
      const double t = x[j];
      for (int i = j - 1; i >= 0; --i) x[i] -= t * col[i];
    }
  } else {
    /* solve U' x = b */
    for (int j = 0; j < n; ++j) {
      const double* col = a + (long)j * lda;
      double t = x[j];
      for (int i = 0; i < j; ++i) t -= col[i] * x[i];
      x[j] = t / col[j];
    }
  }
}

void shim_cblas_dsyrk(CBLAS_ORDER order, CBLAS_UPLO uplo, CBLAS_TRANSPOSE trans, int n, int k,
                      double alpha, const double* a, int lda, double beta, double* c, int ldc)
{
  (void)order; (void)uplo; (void)trans; /* ColMajor, Upper, Trans: C := alpha A'A + beta C, A is k x n */
  for (int j = 0; j < n; ++j)
    for (int i = 0; i <= j; ++i) {
      double d = shim_cblas_ddot(k, a + (long)i * lda, 1, a + (long)j * lda, 1);
      double* cij = c + (long)j * ldc + i;
      *cij = alpha * d + ((beta == 0.0) ? 0.0 : beta * *cij);
    }
}

void shim_cblas_drotg(double* a, double* b, double* c, double* s) { ref_drotg(a, b, c, s); }

/* LAPACK dpotrf, uplo='U', unblocked (dpotf2 algorithm): A = U'U */
void dpotrf_(char* uplo, int* n_, double* a, int* lda_, int* info)
{
  (void)uplo;
  const int n = *n_;
  const long lda = *lda_;
  *info = 0;
  for (int j = 0; j < n; ++j) {
    double* colj = a + j * lda;
    double ajj = colj[j] - shim_cblas_ddot(j, colj, 1, colj, 1);
    if (ajj <= 0.0 || ajj != ajj) {
      colj[j] = ajj;
      *info = j + 1;
      return;
    }
    ajj = sqrt(ajj);
    colj[j] = ajj;
    for (int c = j + 1; c < n; ++c) {
      double* colc = a + c * lda;
      colc[j] = (colc[j] - shim_cblas_ddot(j, colj, 1, colc, 1)) / ajj;
    }
  }
}

#endif /* !BMAGWA_SHIM_OPENBLAS */

/* LINPACK dchex, translated from /root/reference/src/dchex.f:125-246.
 * Fortran arrays are 1-based and column-major: r(i,j) -> R(i,j) below. */
void dchex_(double* r, int* ldr_, int* p_, int* k_, int* l_, double* z, int* ldz_, int* nz_,
            double* c, double* s, int* job_)
{
  const long ldr = *ldr_;
  const int p = *p_, k = *k_, l = *l_, nz = *nz_, job = *job_;
  const long ldz = ldz_ ? *ldz_ : 0;
#define R(i, j) r[((long)(j) - 1) * ldr + ((i) - 1)]
#define Z(i, j) z[((long)(j) - 1) * ldz + ((i) - 1)]
#define C(i) c[(i) - 1]
#define S(i) s[(i) - 1]
  const int km1 = k - 1, kp1 = k + 1, lmk = l - k, lm1 = l - 1;
  int i, ii, il, iu, j, jj;
  double t;

  if (job == 1) {
    /* right circular shift (dchex.f:132-181) */
    for (i = 1; i <= l; ++i) {
      ii = l - i + 1;
      S(i) = R(ii, l);
    }
    for (jj = k; jj <= lm1; ++jj) {
      j = lm1 - jj + k;
      for (i = 1; i <= j; ++i) R(i, j + 1) = R(i, j);
      R(j + 1, j + 1) = 0.0;
    }
    if (k != 1) {
      for (i = 1; i <= km1; ++i) {
        ii = l - i + 1;
        R(i, k) = S(ii);
      }
    }
    t = S(1);
    for (i = 1; i <= lmk; ++i) {
      ref_drotg(&S(i + 1), &t, &C(i), &S(i));
      t = S(i + 1);
    }
    R(k, k) = t;
    for (j = kp1; j <= p; ++j) {
      il = (l - j + 1 > 1) ? l - j + 1 : 1;
      for (ii = il; ii <= lmk; ++ii) {
        i = l - ii;
        t = C(ii) * R(i, j) + S(ii) * R(i + 1, j);
        R(i + 1, j) = C(ii) * R(i + 1, j) - S(ii) * R(i, j);
        R(i, j) = t;
      }
    }
    if (nz >= 1) {
      for (j = 1; j <= nz; ++j)
        for (ii = 1; ii <= lmk; ++ii) {
          i = l - ii;
          t = C(ii) * Z(i, j) + S(ii) * Z(i + 1, j);
          Z(i + 1, j) = C(ii) * Z(i + 1, j) - S(ii) * Z(i, j);
          Z(i, j) = t;
        }
    }
    return;
  }

  /* left circular shift (dchex.f:186-246) */
  for (i = 1; i <= k; ++i) {
    ii = lmk + i;
    S(ii) = R(i, k);
  }
  for (j = k; j <= lm1; ++j) {
    for (i = 1; i <= j; ++i) R(i, j) = R(i, j + 1);
    jj = j - km1;
    S(jj) = R(j + 1, j + 1);
  }
  for (i = 1; i <= k; ++i) {
    ii = lmk + i;
    R(i, l) = S(ii);
  }
  for (i = kp1; i <= l; ++i) R(i, l) = 0.0;

  for (j = k; j <= p; ++j) {
    if (j != k) {
      iu = (j - 1 < l - 1) ? j - 1 : l - 1;
      for (i = k; i <= iu; ++i) {
        ii = i - k + 1;
        t = C(ii) * R(i, j) + S(ii) * R(i + 1, j);
        R(i + 1, j) = C(ii) * R(i + 1, j) - S(ii) * R(i, j);
        R(i, j) = t;
      }
    }
    if (j < l) {
      jj = j - k + 1;
      t = S(jj);
      ref_drotg(&R(j, j), &t, &C(jj), &S(jj));
    }
  }
  if (nz >= 1) {
    for (j = 1; j <= nz; ++j)
      for (i = k; i <= lm1; ++i) {
        ii = i - km1;
        t = C(ii) * Z(i, j) + S(ii) * Z(i + 1, j);
        Z(i + 1, j) = C(ii) * Z(i + 1, j) - S(ii) * Z(i, j);
        Z(i, j) = t;
      }
  }
#undef R
#undef Z
#undef C
#undef S
}
