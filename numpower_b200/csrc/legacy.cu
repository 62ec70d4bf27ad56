// Legacy device-backend symbols of NumPower (include/nb200_legacy.h) on top of the nb200 C-ABI.
// Each function replaces the same-named wrapper of /root/reference/src/ndmath/cuda/cuda_math.cu
// (extern "C" block at :836-1681) or src/gpu_alloc.c (:11-54).  Reference contract kept:
// blocking, `void` returns, failures raised through zend_throw_error (a weak symbol: resolved
// from the PHP runtime / host when present, otherwise printed to stderr).
#include "common.cuh"
#include "../../include/nb200_legacy.h"

extern "C" {
// provided by the host process (PHP's Zend runtime); weak so libnb200.so also loads stand-alone
void zend_throw_error(void *exception_ce, const char *format, ...) __attribute__((weak));
// src/initializers.h:31 — provided by the host objects
struct NDArray *NDArray_Copy(struct NDArray *a, int device) __attribute__((weak));
// src/initializers.h:27 — uninitialised array of the same shape / device (the out-of-place unary route)
struct NDArray *NDArray_EmptyLike(struct NDArray *a) __attribute__((weak));
}

namespace {

// Layout of the host's struct NDArray / NDArrayDescriptor, src/ndarray.h:52-74 (interface only).
struct HostDescriptor { const char *type; int elsize; long numElements; };
struct HostNDArray {
    int uuid; int *strides; int *dimensions; int ndim; char *data; struct HostNDArray *base; int flags;
    HostDescriptor *descriptor; void *iterator; void *php_iterator; int refcount; int device;
};

void raise(const char *what) {
    const char *msg = nb200_last_error();
    if (zend_throw_error) zend_throw_error(nullptr, "%s: %s", what, msg);
    else fprintf(stderr, "libnb200 (legacy ABI): %s: %s\n", what, msg);
}
// The reference host mixes its own default-stream calls (cudaMemset in NDArray_Zeros,
// initializers.c:438-442) with backend calls, and every reference wrapper ended in
// cudaDeviceSynchronize: fence on both sides so ordering matches what the host assumes.
void enter() { cudaDeviceSynchronize(); }
void leave(int rc, const char *what) {
    if (rc != NB200_OK) { raise(what); return; }
    if (nb200_synchronize() != NB200_OK) raise(what);
}
void binary(int op, float *a, float *b, float *rtn, int n, const char *what) {
    enter();
    int64_t shape[1] = {n}, st[1] = {1};
    leave(nb200_ew_binary(op, rtn, a, b, 1, shape, st, st), what);
}
void domain_check(int op, const char *what) {
    if (op == NB200_UN_ARCCOS || op == NB200_UN_ARCCOSH || op == NB200_UN_ARCTANH) {
        int flag = 0;
        if (nb200_poll_domain_error(&flag) == NB200_OK && flag) {
            // the CPU functors exit(1) here (double_math.c:145-148); the backend raises instead
            if (zend_throw_error) zend_throw_error(nullptr, "RuntimeError: Invalid argument provided for %s", what);
            else fprintf(stderr, "libnb200: RuntimeError: Invalid argument provided for %s\n", what);
        }
    }
}
void unary(int op, float *d, int n, float p0, float p1, const char *what) {
    enter();
    int rc = nb200_ew_unary(op, d, d, n, p0, p1);
    leave(rc, what);
    if (rc == NB200_OK) domain_check(op, what);
}
void not_implemented(const char *what) {
    if (zend_throw_error) zend_throw_error(nullptr, "%s is not implemented in the B200 backend (out of scope: SURVEY.md section 2)", what);
    else fprintf(stderr, "libnb200: %s is not implemented in the B200 backend\n", what);
}

}  // namespace

extern "C" {

// ---- gpu_alloc.c:11-54
void vmalloc(void **target, unsigned int size) {
    if (nb200_alloc(target, (int64_t)size) != NB200_OK) raise("vmalloc");   // "device memory allocation failed"
}
void vfree(void *target) { if (nb200_free(target) != NB200_OK) raise("vfree"); }
void vmemcheck(void) {
    int64_t live = 0;
    nb200_mem_stats(&live, nullptr);
    if (live != 0) printf("\nVRAM MEMORY LEAK: leaked %d array(s)\n", (int)live);   // gpu_alloc.c:38
}
void vmemcpyd2d(char *src, char *dst, unsigned int size) {
    enter();
    leave(nb200_copy_d2d(dst, src, (int64_t)size), "vmemcpyd2d");
}
void vmemcpyh2d(char *src, char *dst, unsigned int size) {
    enter();
    if (nb200_copy_h2d(dst, src, (int64_t)size) != NB200_OK) raise("vmemcpyh2d");
}
float NDArray_VFLOAT(char *target) {
    float v = 0.f;
    enter();
    if (nb200_copy_d2h(&v, target, sizeof(float)) != NB200_OK) raise("NDArray_VFLOAT");
    return v;
}
float NDArray_VFLOATF_I(float *target, int index) { return NDArray_VFLOAT(reinterpret_cast<char *>(target + index)); }

// ---- binary ops (cuda_math.cu:1063-1109): equal-length contiguous device arrays
void cuda_add_float(int, float *a, float *b, float *rtn, int n) { binary(NB200_ADD, a, b, rtn, n, "cuda_add_float"); }
void cuda_subtract_float(int, float *a, float *b, float *rtn, int n) { binary(NB200_SUB, a, b, rtn, n, "cuda_subtract_float"); }
void cuda_multiply_float(int, float *a, float *b, float *rtn, int n) { binary(NB200_MUL, a, b, rtn, n, "cuda_multiply_float"); }
void cuda_divide_float(int, float *a, float *b, float *rtn, int n) { binary(NB200_DIV, a, b, rtn, n, "cuda_divide_float"); }
// CPU-path semantics (floored, fused) rather than the old kernel's fmodf: SURVEY.md §8 a-2
void cuda_mod_float(int, float *a, float *b, float *rtn, int n) { binary(NB200_MOD, a, b, rtn, n, "cuda_mod_float"); }
void cuda_pow_float(int, float *a, float *b, float *rtn, int n) { binary(NB200_POW, a, b, rtn, n, "cuda_pow_float"); }
void cuda_float_compare_equal(int, float *a, float *b, float *r, int n) { binary(NB200_CMP_EQ, a, b, r, n, "cuda_float_compare_equal"); }
void cuda_float_compare_not_equal(int, float *a, float *b, float *r, int n) { binary(NB200_CMP_NE, a, b, r, n, "cuda_float_compare_not_equal"); }
void cuda_float_compare_greater(int, float *a, float *b, float *r, int n) { binary(NB200_CMP_GT, a, b, r, n, "cuda_float_compare_greater"); }
void cuda_float_compare_greater_equal(int, float *a, float *b, float *r, int n) { binary(NB200_CMP_GE, a, b, r, n, "cuda_float_compare_greater_equal"); }
void cuda_float_compare_less(int, float *a, float *b, float *r, int n) { binary(NB200_CMP_LT, a, b, r, n, "cuda_float_compare_less"); }
void cuda_float_compare_less_equal(int, float *a, float *b, float *r, int n) { binary(NB200_CMP_LE, a, b, r, n, "cuda_float_compare_less_equal"); }
void cuda_float_arctan2(int, float *d_array, float *y_array) { (void)d_array; (void)y_array; not_implemented("cuda_float_arctan2 (no element count in the legacy signature; use nb200_ew_binary(NB200_ARCTAN2))"); }

// ---- reductions (cuda_math.cu:920-944, 1015-1045). rtn is a HOST pointer holding the identity.
void cuda_sum_float(int, float *a, float *rtn, int n) {
    enter();
    float v = 0.f;
    int rc = nb200_reduce_full_host(NB200_SUM, &v, a, n);
    if (rc != NB200_OK) { raise("cuda_sum_float"); return; }
    *rtn += v;
}
void cuda_prod_float(int, float *a, float *rtn, int n) {
    enter();
    float v = 1.f;
    int rc = nb200_reduce_full_host(NB200_PROD, &v, a, n);
    if (rc != NB200_OK) { raise("cuda_prod_float"); return; }
    *rtn *= v;
}
float cuda_max_float(float *a, int n) {
    enter();
    float v = 0.f;
    if (nb200_reduce_full_host(NB200_MAX, &v, a, n) != NB200_OK) raise("cuda_max_float");
    return v;
}
float cuda_min_float(float *a, int n) {
    enter();
    float v = 0.f;
    if (nb200_reduce_full_host(NB200_MIN, &v, a, n) != NB200_OK) raise("cuda_min_float");
    return v;
}
// array_equal: 1 iff every pair compares equal (exact; the old kernel's |a-b|<=1e-7 is the same test for fp32 magnitudes >= 1)
int cuda_equal_float(int, float *a, float *b, int n) {
    enter();
    if (n <= 0) return 1;
    float *tmp = nullptr;
    if (nb200_alloc(reinterpret_cast<void **>(&tmp), (int64_t)n * 4) != NB200_OK) { raise("cuda_equal_float"); return 0; }
    int64_t shape[1] = {n}, st[1] = {1};
    float mn = 0.f;
    int rc = nb200_ew_binary(NB200_CMP_EQ, tmp, a, b, 1, shape, st, st);
    if (rc == NB200_OK) rc = nb200_reduce_full_host(NB200_MIN, &mn, tmp, n);
    nb200_free(tmp);
    if (rc != NB200_OK) { raise("cuda_equal_float"); return 0; }
    return mn == 1.0f;
}
void cuda_fill_float(float *a, float value, int n) { enter(); leave(nb200_fill(a, value, n), "cuda_fill_float"); }

// ---- in-place unaries (cuda_math.cu:1111-1414)
#define NB_LEGACY_UNARY(name, op) void name(int n, float *d) { unary(op, d, n, 0.f, 0.f, #name); }
NB_LEGACY_UNARY(cuda_float_abs, NB200_UN_ABS)         NB_LEGACY_UNARY(cuda_float_expm1, NB200_UN_EXPM1)
NB_LEGACY_UNARY(cuda_float_exp, NB200_UN_EXP)         NB_LEGACY_UNARY(cuda_float_sqrt, NB200_UN_SQRT)
NB_LEGACY_UNARY(cuda_float_log, NB200_UN_LOG)         NB_LEGACY_UNARY(cuda_float_logb, NB200_UN_LOGB)
NB_LEGACY_UNARY(cuda_float_log2, NB200_UN_LOG2)       NB_LEGACY_UNARY(cuda_float_log1p, NB200_UN_LOG1P)
NB_LEGACY_UNARY(cuda_float_log10, NB200_UN_LOG10)     NB_LEGACY_UNARY(cuda_float_sin, NB200_UN_SIN)
NB_LEGACY_UNARY(cuda_float_cos, NB200_UN_COS)         NB_LEGACY_UNARY(cuda_float_tan, NB200_UN_TAN)
NB_LEGACY_UNARY(cuda_float_arcsin, NB200_UN_ARCSIN)   NB_LEGACY_UNARY(cuda_float_arccos, NB200_UN_ARCCOS)
NB_LEGACY_UNARY(cuda_float_arctan, NB200_UN_ARCTAN)   NB_LEGACY_UNARY(cuda_float_degrees, NB200_UN_DEGREES)
NB_LEGACY_UNARY(cuda_float_radians, NB200_UN_RADIANS) NB_LEGACY_UNARY(cuda_float_sinh, NB200_UN_SINH)
NB_LEGACY_UNARY(cuda_float_cosh, NB200_UN_COSH)       NB_LEGACY_UNARY(cuda_float_tanh, NB200_UN_TANH)
NB_LEGACY_UNARY(cuda_float_arcsinh, NB200_UN_ARCSINH) NB_LEGACY_UNARY(cuda_float_arccosh, NB200_UN_ARCCOSH)
NB_LEGACY_UNARY(cuda_float_arctanh, NB200_UN_ARCTANH) NB_LEGACY_UNARY(cuda_float_rint, NB200_UN_RINT)
NB_LEGACY_UNARY(cuda_float_fix, NB200_UN_FIX)         NB_LEGACY_UNARY(cuda_float_ceil, NB200_UN_CEIL)
NB_LEGACY_UNARY(cuda_float_floor, NB200_UN_FLOOR)     NB_LEGACY_UNARY(cuda_float_sinc, NB200_UN_SINC)
NB_LEGACY_UNARY(cuda_float_trunc, NB200_UN_TRUNC)     NB_LEGACY_UNARY(cuda_float_negate, NB200_UN_NEGATIVE)
NB_LEGACY_UNARY(cuda_float_sign, NB200_UN_SIGN)       NB_LEGACY_UNARY(cuda_float_positive, NB200_UN_POSITIVE)
NB_LEGACY_UNARY(cuda_float_reciprocal, NB200_UN_RECIPROCAL)
void cuda_float_clip(int n, float *d, float lo, float hi) { unary(NB200_UN_CLIP, d, n, lo, hi, "cuda_float_clip"); }
void cuda_float_round(int n, float *d, float decimals) { unary(NB200_UN_ROUND, d, n, decimals, 0.f, "cuda_float_round"); }

// ---- NDArray-level unary drivers (cuda_math.cu:1532-1558).  The reference copies the operand (NDArray_Copy: 8 B per element) and
// then runs the in-place op on the copy (another 8 B per element).  Here the PHP method's `cuda_float_<op>` argument is recognised
// by its address and the op runs OUT OF PLACE into a fresh array from the host's own NDArray_EmptyLike (initializers.c:406): one
// kernel, 8 B per element, no host patch needed (numpower.c:1648-3348 stays as it is).  An op pointer that is not one of this
// library's wrappers takes the reference's copy-then-in-place route.
static int unary_op_of(const void *fn) {
    static const struct { const void *fn; int op; } table[] = {
        {(const void *)cuda_float_abs, NB200_UN_ABS}, {(const void *)cuda_float_expm1, NB200_UN_EXPM1}, {(const void *)cuda_float_exp, NB200_UN_EXP},
        {(const void *)cuda_float_sqrt, NB200_UN_SQRT}, {(const void *)cuda_float_log, NB200_UN_LOG}, {(const void *)cuda_float_logb, NB200_UN_LOGB},
        {(const void *)cuda_float_log2, NB200_UN_LOG2}, {(const void *)cuda_float_log1p, NB200_UN_LOG1P}, {(const void *)cuda_float_log10, NB200_UN_LOG10},
        {(const void *)cuda_float_sin, NB200_UN_SIN}, {(const void *)cuda_float_cos, NB200_UN_COS}, {(const void *)cuda_float_tan, NB200_UN_TAN},
        {(const void *)cuda_float_arcsin, NB200_UN_ARCSIN}, {(const void *)cuda_float_arccos, NB200_UN_ARCCOS}, {(const void *)cuda_float_arctan, NB200_UN_ARCTAN},
        {(const void *)cuda_float_degrees, NB200_UN_DEGREES}, {(const void *)cuda_float_radians, NB200_UN_RADIANS}, {(const void *)cuda_float_sinh, NB200_UN_SINH},
        {(const void *)cuda_float_cosh, NB200_UN_COSH}, {(const void *)cuda_float_tanh, NB200_UN_TANH}, {(const void *)cuda_float_arcsinh, NB200_UN_ARCSINH},
        {(const void *)cuda_float_arccosh, NB200_UN_ARCCOSH}, {(const void *)cuda_float_arctanh, NB200_UN_ARCTANH}, {(const void *)cuda_float_rint, NB200_UN_RINT},
        {(const void *)cuda_float_fix, NB200_UN_FIX}, {(const void *)cuda_float_ceil, NB200_UN_CEIL}, {(const void *)cuda_float_floor, NB200_UN_FLOOR},
        {(const void *)cuda_float_sinc, NB200_UN_SINC}, {(const void *)cuda_float_trunc, NB200_UN_TRUNC}, {(const void *)cuda_float_negate, NB200_UN_NEGATIVE},
        {(const void *)cuda_float_sign, NB200_UN_SIGN}, {(const void *)cuda_float_positive, NB200_UN_POSITIVE}, {(const void *)cuda_float_reciprocal, NB200_UN_RECIPROCAL},
        {(const void *)cuda_float_clip, NB200_UN_CLIP}, {(const void *)cuda_float_round, NB200_UN_ROUND},
    };
    for (const auto &e : table) if (e.fn == fn) return e.op;
    return -1;
}
// out-of-place route; nullptr = not served (unknown op pointer or a host without NDArray_EmptyLike)
static struct NDArray *unary_out_of_place(struct NDArray *nd, const void *fn, float p0, float p1, const char *what) {
    const int op = unary_op_of(fn);
    if (op < 0 || !NDArray_EmptyLike) return nullptr;
    HostNDArray *h = reinterpret_cast<HostNDArray *>(nd);
    HostNDArray *r = reinterpret_cast<HostNDArray *>(NDArray_EmptyLike(nd));
    if (!r) return nullptr;
    enter();
    int rc = nb200_ew_unary(op, reinterpret_cast<float *>(r->data), reinterpret_cast<const float *>(h->data), (int64_t)h->descriptor->numElements, p0, p1);
    leave(rc, what);
    if (rc == NB200_OK) domain_check(op, what);
    return reinterpret_cast<struct NDArray *>(r);
}
struct NDArray *NDArrayMathGPU_ElementWise(struct NDArray *nd, ElementWiseFloatGPUOperation op) {
    if (struct NDArray *r = unary_out_of_place(nd, (const void *)op, 0.f, 0.f, "NDArrayMathGPU_ElementWise")) return r;
    if (!NDArray_Copy) { not_implemented("NDArrayMathGPU_ElementWise without the host's NDArray_Copy"); return nullptr; }
    HostNDArray *h = reinterpret_cast<HostNDArray *>(nd);
    HostNDArray *r = reinterpret_cast<HostNDArray *>(NDArray_Copy(nd, h->device));
    op((int)r->descriptor->numElements, reinterpret_cast<float *>(r->data));
    return reinterpret_cast<struct NDArray *>(r);
}
struct NDArray *NDArrayMathGPU_ElementWise1F(struct NDArray *nd, ElementWiseFloatGPUOperation1F op, float v1) {
    if (struct NDArray *r = unary_out_of_place(nd, (const void *)op, v1, 0.f, "NDArrayMathGPU_ElementWise1F")) return r;
    if (!NDArray_Copy) { not_implemented("NDArrayMathGPU_ElementWise1F without the host's NDArray_Copy"); return nullptr; }
    HostNDArray *h = reinterpret_cast<HostNDArray *>(nd);
    HostNDArray *r = reinterpret_cast<HostNDArray *>(NDArray_Copy(nd, h->device));
    op((int)r->descriptor->numElements, reinterpret_cast<float *>(r->data), v1);
    return reinterpret_cast<struct NDArray *>(r);
}
struct NDArray *NDArrayMathGPU_ElementWise2F(struct NDArray *nd, ElementWiseFloatGPUOperation2F op, float v1, float v2) {
    if (struct NDArray *r = unary_out_of_place(nd, (const void *)op, v1, v2, "NDArrayMathGPU_ElementWise2F")) return r;
    if (!NDArray_Copy) { not_implemented("NDArrayMathGPU_ElementWise2F without the host's NDArray_Copy"); return nullptr; }
    HostNDArray *h = reinterpret_cast<HostNDArray *>(nd);
    HostNDArray *r = reinterpret_cast<HostNDArray *>(NDArray_Copy(nd, h->device));
    op((int)r->descriptor->numElements, reinterpret_cast<float *>(r->data), v1, v2);
    return reinterpret_cast<struct NDArray *>(r);
}
struct NDArray *NDArrayMathGPU_ElementWise1N(struct NDArray *nd, ElementWiseFloatGPUOperation1N op, struct NDArray *v1) {
    if (!NDArray_Copy) { not_implemented("NDArrayMathGPU_ElementWise1N without the host's NDArray_Copy"); return nullptr; }
    HostNDArray *h = reinterpret_cast<HostNDArray *>(nd);
    HostNDArray *r = reinterpret_cast<HostNDArray *>(NDArray_Copy(nd, h->device));
    op((int)r->descriptor->numElements, reinterpret_cast<float *>(r->data),
       reinterpret_cast<float *>(reinterpret_cast<HostNDArray *>(v1)->data));
    return reinterpret_cast<struct NDArray *>(r);
}

// ---- gemv / transpose (cuda_math.cu:1416-1422, 1287-1294)
void cuda_float_multiply_matrix_vector(int, float *A, float *x, float *y, int rows, int cols) {
    enter();
    leave(nb200_gemv(y, A, x, rows, cols), "cuda_float_multiply_matrix_vector");
}
// The host calls this with d_in == d_out (manipulation.c:124); transpose through a temporary.
void cuda_float_transpose(int, int, const float *d_in, float *d_out, int width, int height) {
    enter();
    const int64_t n = (int64_t)width * height;
    if (d_in != d_out) { leave(nb200_transpose2d(d_out, d_in, height, width), "cuda_float_transpose"); return; }
    float *tmp = nullptr;
    if (nb200_alloc(reinterpret_cast<void **>(&tmp), n * 4) != NB200_OK) { raise("cuda_float_transpose"); return; }
    int rc = nb200_transpose2d(tmp, d_in, height, width);
    if (rc == NB200_OK) rc = nb200_copy_d2d(d_out, tmp, n * 4);
    leave(rc, "cuda_float_transpose");
    nb200_free(tmp);
}

// ---- cuBLAS entry points NDArray_FMatmul calls directly (linalg.c:55-71); include/nb200_cublas_shim/cublas_v2.h
// declares them for the host build.  Column-major C'(m x n) = A'(m x k) B'(k x n) is row-major
// C'^T = B'^T A'^T, i.e. nb200_sgemm(C', A_row = B', B_row = A', M = n, N = m, K = k).
int nb200_shim_cublasCreate(void **handle) { if (handle) *handle = reinterpret_cast<void *>(0x1); return nb200_init(0) == NB200_OK ? 0 : 1; }
int nb200_shim_cublasDestroy(void *) { return 0; }
int nb200_shim_cublasSgemm(void *, int transa, int transb, int m, int n, int k, const float *alpha, const float *A, int lda,
                           const float *B, int ldb, const float *beta, float *C, int ldc) {
    if (transa != 0 || transb != 0 || !alpha || !beta || *alpha != 1.0f || *beta != 0.0f) {
        not_implemented("cublasSgemm shim with transposes or alpha != 1 / beta != 0");
        return 1;
    }
    enter();
    int rc = nb200_sgemm(C, B, A, n, m, k, ldb, lda, ldc, NB200_GEMM_AUTO);
    leave(rc, "cublasSgemm (nb200_sgemm)");
    return rc == NB200_OK ? 0 : 1;
}

// ---- NDArray_Outer's device branch (linalg.c:745-748; calculateOuterProductFloat, cuda_math.cu:70-77: r[i*n+j] = a[i]*b[j], one
// thread per element on 16x16 blocks).  Here: the broadcast multiply of a column (stride 1, 0) by a row (stride 0, 1) - one
// HBM-bound launch of ew_bcast2d that reads m + n floats and writes m*n.
void cuda_calculate_outer_product(int m, int n, float *a_array, float *b_array, float *r_array) {
    enter();
    const int64_t shape[2] = {m, n}, sa[2] = {1, 0}, sb[2] = {0, 1};
    leave(nb200_ew_binary(NB200_MUL, r_array, a_array, b_array, 2, shape, sa, sb), "cuda_calculate_outer_product");
}

// ---- out-of-scope exports: link, then raise
int cuda_svd_float(float *, float *, float *, float *, int, int) { not_implemented("cuda_svd_float"); return -1; }
int cuda_det_float(float *, float *, int) { not_implemented("cuda_det_float"); return -1; }
void cuda_matrix_float_inverse(float *, int) { not_implemented("cuda_matrix_float_inverse"); }
void cuda_float_lu(float *, float *, float *, float *, int) { not_implemented("cuda_float_lu"); }
void cuda_lstsq_float(float *, int, int, float *, int, float *) { not_implemented("cuda_lstsq_float"); }
void cuda_convolve2d_same_float(const float *, const float *, const int *, const int *, const int *, const int *, char,
                                float *, float) { not_implemented("cuda_convolve2d_same_float"); }
void cuda_matrix_float_l1norm(float *, float *, int, int) { not_implemented("cuda_matrix_float_l1norm"); }
int cuda_matrix_float_l2norm(float *, float *, int, int) { not_implemented("cuda_matrix_float_l2norm"); return -1; }
void cuda_matrix_eig_float(float *, int, float *) { not_implemented("cuda_matrix_eig_float"); }
float cuda_float_median_float(int, float *, int) { not_implemented("cuda_float_median_float"); return 0.f; }
float *cuda_dnn_conv2d_float32(float *, int, int, int, int, int, int *, int, char) { not_implemented("cuda_dnn_conv2d_float32"); return nullptr; }
float *cuda_dnn_conv2d_float32_backward(float *, float *, float *, float, float, int, int, int, int, int, char) {
    not_implemented("cuda_dnn_conv2d_float32_backward");
    return nullptr;
}

}  // extern "C"
