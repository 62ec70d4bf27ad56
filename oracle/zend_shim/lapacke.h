/* TEST INFRASTRUCTURE — LAPACKE prototypes so linalg.c compiles; the
 * factorizations are out of scope and are never called by the
 * oracle entry points. */
#ifndef NB200_ORACLE_LAPACKE_H
#define NB200_ORACLE_LAPACKE_H
#define LAPACK_ROW_MAJOR 101
#define LAPACK_COL_MAJOR 102
typedef int lapack_int;
lapack_int LAPACKE_sgesdd(int layout, char jobz, lapack_int m, lapack_int n, float *a, lapack_int lda, float *s, float *u, lapack_int ldu, float *vt, lapack_int ldvt);
lapack_int LAPACKE_sgetrf(int layout, lapack_int m, lapack_int n, float *a, lapack_int lda, lapack_int *ipiv);
lapack_int LAPACKE_sgetri(int layout, lapack_int n, float *a, lapack_int lda, const lapack_int *ipiv);
lapack_int LAPACKE_sgeqrf(int layout, lapack_int m, lapack_int n, float *a, lapack_int lda, float *tau);
lapack_int LAPACKE_sorgqr(int layout, lapack_int m, lapack_int n, lapack_int k, float *a, lapack_int lda, const float *tau);
lapack_int LAPACKE_sgeev(int layout, char jobvl, char jobvr, lapack_int n, float *a, lapack_int lda, float *wr, float *wi, float *vl, lapack_int ldvl, float *vr, lapack_int ldvr);
lapack_int LAPACKE_sgels(int layout, char trans, lapack_int m, lapack_int n, lapack_int nrhs, float *a, lapack_int lda, float *b, lapack_int ldb);
lapack_int LAPACKE_sgelsd(int layout, lapack_int m, lapack_int n, lapack_int nrhs, float *a, lapack_int lda, float *b, lapack_int ldb, float *s, float rcond, lapack_int *rank);
lapack_int LAPACKE_sgesv(int layout, lapack_int n, lapack_int nrhs, float *a, lapack_int lda, lapack_int *ipiv, float *b, lapack_int ldb);
lapack_int LAPACKE_spotrf(int layout, char uplo, lapack_int n, float *a, lapack_int lda);
lapack_int LAPACKE_sgesvd(int layout, char jobu, char jobvt, lapack_int m, lapack_int n, float *a, lapack_int lda, float *s, float *u, lapack_int ldu, float *vt, lapack_int ldvt, float *superb);
void sgetrf_(int *m, int *n, float *a, int *lda, int *ipiv, int *info);
void sgetri_(int *n, float *a, int *lda, int *ipiv, float *work, int *lwork, int *info);
#endif
