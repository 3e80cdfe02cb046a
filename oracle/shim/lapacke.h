/* TEST INFRASTRUCTURE ONLY (oracle build): the 11 LAPACKE entry points the reference
 * calls (src/xerus/blasLapackWrapper.cpp:218-701), bound to scipy's bundled OpenBLAS. */
#pragma once
#define LAPACK_ROW_MAJOR 101
#define LAPACK_COL_MAJOR 102
typedef int lapack_int;
#define XB_SYM(n) __asm__("scipy_" #n)
#ifdef __cplusplus
extern "C" {
#endif
lapack_int LAPACKE_dgesdd(int layout, char jobz, lapack_int m, lapack_int n, double* a, lapack_int lda, double* s,
                          double* u, lapack_int ldu, double* vt, lapack_int ldvt) XB_SYM(LAPACKE_dgesdd);
lapack_int LAPACKE_dgeqp3(int layout, lapack_int m, lapack_int n, double* a, lapack_int lda, lapack_int* jpvt,
                          double* tau) XB_SYM(LAPACKE_dgeqp3);
lapack_int LAPACKE_dorgqr(int layout, lapack_int m, lapack_int n, lapack_int k, double* a, lapack_int lda,
                          const double* tau) XB_SYM(LAPACKE_dorgqr);
lapack_int LAPACKE_dgeqrf(int layout, lapack_int m, lapack_int n, double* a, lapack_int lda, double* tau) XB_SYM(LAPACKE_dgeqrf);
lapack_int LAPACKE_dgerqf(int layout, lapack_int m, lapack_int n, double* a, lapack_int lda, double* tau) XB_SYM(LAPACKE_dgerqf);
lapack_int LAPACKE_dorgrq(int layout, lapack_int m, lapack_int n, lapack_int k, double* a, lapack_int lda,
                          const double* tau) XB_SYM(LAPACKE_dorgrq);
lapack_int LAPACKE_dgesv(int layout, lapack_int n, lapack_int nrhs, double* a, lapack_int lda, lapack_int* ipiv,
                         double* b, lapack_int ldb) XB_SYM(LAPACKE_dgesv);
lapack_int LAPACKE_dpotrf2(int layout, char uplo, lapack_int n, double* a, lapack_int lda) XB_SYM(LAPACKE_dpotrf2);
lapack_int LAPACKE_dpotrs(int layout, char uplo, lapack_int n, lapack_int nrhs, const double* a, lapack_int lda,
                          double* b, lapack_int ldb) XB_SYM(LAPACKE_dpotrs);
lapack_int LAPACKE_dsysv(int layout, char uplo, lapack_int n, lapack_int nrhs, double* a, lapack_int lda,
                         lapack_int* ipiv, double* b, lapack_int ldb) XB_SYM(LAPACKE_dsysv);
lapack_int LAPACKE_dgelsd(int layout, lapack_int m, lapack_int n, lapack_int nrhs, double* a, lapack_int lda,
                          double* b, lapack_int ldb, double* s, double rcond, lapack_int* rank) XB_SYM(LAPACKE_dgelsd);
#ifdef __cplusplus
}
#endif
#undef XB_SYM
