/* TEST INFRASTRUCTURE ONLY (oracle build): minimal CBLAS prototypes that bind the
 * unmodified reference sources to the LP64 OpenBLAS bundled inside the scipy wheel
 * (symbols carry a "scipy_" prefix there).  Included by the reference inside
 * extern "C" (reference: src/xerus/blasLapackWrapper.cpp:33-36). */
#pragma once
enum CBLAS_ORDER {CblasRowMajor=101, CblasColMajor=102};
enum CBLAS_TRANSPOSE {CblasNoTrans=111, CblasTrans=112, CblasConjTrans=113};
#define XB_SYM(n) __asm__("scipy_" #n)
double cblas_dasum(int n, const double* x, int incx) XB_SYM(cblas_dasum);
double cblas_dnrm2(int n, const double* x, int incx) XB_SYM(cblas_dnrm2);
double cblas_ddot(int n, const double* x, int incx, const double* y, int incy) XB_SYM(cblas_ddot);
void cblas_dgemv(enum CBLAS_ORDER, enum CBLAS_TRANSPOSE, int m, int n, double alpha, const double* a, int lda,
                 const double* x, int incx, double beta, double* y, int incy) XB_SYM(cblas_dgemv);
void cblas_dger(enum CBLAS_ORDER, int m, int n, double alpha, const double* x, int incx, const double* y, int incy,
                double* a, int lda) XB_SYM(cblas_dger);
void cblas_dgemm(enum CBLAS_ORDER, enum CBLAS_TRANSPOSE, enum CBLAS_TRANSPOSE, int m, int n, int k, double alpha,
                 const double* a, int lda, const double* b, int ldb, double beta, double* c, int ldc) XB_SYM(cblas_dgemm);
#undef XB_SYM
