/* TEST INFRASTRUCTURE — hand-declared prototypes for the four cblas entry
 * points the reference calls (no cblas.h in this image).  Symbols resolve to
 * the scipy-bundled OpenBLAS (prefix scipy_) through the -D renames in
 * oracle/build_ref.sh. */
#ifndef NB200_ORACLE_CBLAS_H
#define NB200_ORACLE_CBLAS_H
typedef enum CBLAS_ORDER { CblasRowMajor = 101, CblasColMajor = 102 } CBLAS_ORDER;
typedef enum CBLAS_TRANSPOSE { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 } CBLAS_TRANSPOSE;
typedef CBLAS_ORDER CBLAS_LAYOUT;
void cblas_sgemm(CBLAS_ORDER order, CBLAS_TRANSPOSE ta, CBLAS_TRANSPOSE tb, int m, int n, int k,
                 float alpha, const float *a, int lda, const float *b, int ldb, float beta, float *c, int ldc);
void cblas_sgemv(CBLAS_ORDER order, CBLAS_TRANSPOSE ta, int m, int n, float alpha, const float *a, int lda,
                 const float *x, int incx, float beta, float *y, int incy);
float cblas_sasum(int n, const float *x, int incx);
float cblas_sdot(int n, const float *x, int incx, const float *y, int incy);
float cblas_snrm2(int n, const float *x, int incx);
void cblas_sger(CBLAS_ORDER order, int m, int n, float alpha, const float *x, int incx, const float *y, int incy, float *a, int lda);
void cblas_scopy(int n, const float *x, int incx, float *y, int incy);
void cblas_sscal(int n, float alpha, float *x, int incx);
void cblas_saxpy(int n, float alpha, const float *x, int incx, float *y, int incy);
#endif
