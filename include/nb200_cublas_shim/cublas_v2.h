/*
 * cublas_v2.h shim — put this directory FIRST on the include path when compiling NumPower's
 * host sources with -DHAVE_CUBLAS against libnb200.so.  src/ndmath/linalg.c:54-72 calls
 * cublasCreate / cublasSgemm / cublasDestroy directly from host code; this header maps those
 * three calls onto the B200 backend (numpower_b200/csrc/legacy.cu: nb200_shim_cublas*), so no
 * cuBLAS is linked or called on the matmul path.  Other files only need the types to exist
 * (src/gpu_alloc.c:14 uses cublasStatus_t for a cudaMalloc result).
 */
#ifndef NB200_CUBLAS_SHIM_H
#define NB200_CUBLAS_SHIM_H
#ifdef __cplusplus
extern "C" {
#endif
typedef void *cublasHandle_t;
typedef int cublasStatus_t;
typedef enum { CUBLAS_OP_N = 0, CUBLAS_OP_T = 1, CUBLAS_OP_C = 2 } cublasOperation_t;
#define CUBLAS_STATUS_SUCCESS 0
int nb200_shim_cublasCreate(void **handle);
int nb200_shim_cublasDestroy(void *handle);
int nb200_shim_cublasSgemm(void *handle, int transa, int transb, int m, int n, int k, const float *alpha,
                           const float *A, int lda, const float *B, int ldb, const float *beta, float *C, int ldc);
#define cublasCreate(h) nb200_shim_cublasCreate((void **)(h))
#define cublasDestroy(h) nb200_shim_cublasDestroy((void *)(h))
#define cublasSgemm(h, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc) \
    nb200_shim_cublasSgemm((void *)(h), (int)(ta), (int)(tb), (m), (n), (k), (alpha), (A), (lda), (B), (ldb), (beta), (C), (ldc))
#ifdef __cplusplus
}
#endif
#endif
