/*
 * nb200.h — C-ABI of the B200-native device backend for NumPower's NDArray.
 *
 * This is the drop-in boundary (SURVEY.md §8 b).  In the reference the device
 * backend is a link-time interface: two objects (`cuda_math.o`, `gpu_alloc.o`)
 * exporting the `extern "C"` symbols of src/ndmath/cuda/cuda_math.h:14-79 and
 * src/gpu_alloc.h:8-15, selected by `#ifdef HAVE_CUBLAS` at the call sites in
 * src/ndmath/arithmetics.c, src/ndarray.c, src/ndmath/linalg.c.  libnb200.so
 * exports
 *   (1) the 64-bit entry points below (`nb200_*`): plain pointers and sizes,
 *       int status returns, no torch / C++ types; and
 *   (2) the exact legacy symbols (include/nb200_legacy.h) implemented on top of
 *       (1), so the reference's unmodified host objects link against it.
 *
 * Conventions
 *   - every function returns 0 (NB200_OK) or a negative NB200_E* code;
 *     nb200_last_error() returns the message for the calling thread's last
 *     failure (the host maps it to zend_throw_error).
 *   - all array arguments are DEVICE pointers to fp32 unless the name says host.
 *   - compute entry points are ASYNCHRONOUS on the context stream
 *     (nb200_stream()); entry points that return a value to the host
 *     (`*_host`) synchronise that stream before returning.
 *   - ownership: the host allocates inputs and outputs through nb200_alloc;
 *     the backend never frees host-visible arrays (same as vmalloc/vfree).
 *   - there is NO CPU fallback: without a CUDA device every entry point fails
 *     with NB200_ENODEV.
 */
#ifndef NB200_H
#define NB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NB200_OK 0
#define NB200_EINVAL (-1)  /* bad argument (shape, op id, alignment that cannot be served) */
#define NB200_ECUDA (-2)   /* CUDA runtime / driver error, see nb200_last_error() */
#define NB200_ENOMEM (-3)  /* device allocation failed ("device memory allocation failed", gpu_alloc.c:14-16) */
#define NB200_ENODEV (-4)  /* no usable sm_100 device */
#define NB200_EDOMAIN (-5) /* math domain error seen by a unary op (double_math.c:145-148 exit(1) in the reference) */

#define NB200_MAX_DIMS 8

/* Binary ops — replace cuda_{add,subtract,multiply,divide,mod,pow}_float
 * (cuda_math.h:25-29,33) and the CPU-only NDArray_Maximum/Minimum
 * (ndarray.c:852-931) and float_arctan2 (double_math.c:259).
 * Semantics are the reference's CPU ones. */
enum nb200_binary_op {
    NB200_ADD = 0, NB200_SUB = 1, NB200_MUL = 2, NB200_DIV = 3,
    NB200_MOD = 4,      /* a - floor(a/b)*b, fused multiply-subtract (arithmetics.c:787-806 AVX body) */
    NB200_POW = 5, NB200_MAXIMUM = 6, NB200_MINIMUM = 7, NB200_ARCTAN2 = 8,
    NB200_MOD_TRUNC = 9, /* fmodf: the reference's scalar tail / ref-GPU semantics (arithmetics.c:801) */
    /* comparisons -> 1.0f / 0.0f masks, all ORDERED (any NaN operand -> 0, including != : _CMP_NEQ_OQ), as the
     * _mm256_cmp_ps predicates of src/logic.c:121-660; replace cuda_float_compare_* (cuda_math.h:63,69-73) */
    NB200_CMP_EQ = 10, NB200_CMP_NE = 11, NB200_CMP_GT = 12, NB200_CMP_GE = 13, NB200_CMP_LT = 14, NB200_CMP_LE = 15,
    NB200_BIN_COUNT = 16
};

/* Unary ops — replace cuda_float_<op> (cuda_math.h:16-24,38-61,67,78-79) and
 * NDArrayMathGPU_ElementWise* (:14-15,75-76).  Semantics follow the CPU functors
 * of src/ndmath/double_math.c (line given), not the old .cu ones. */
enum nb200_unary_op {
    NB200_UN_ABS = 0,      /* :10  */ NB200_UN_SQRT = 1,     /* :19  */ NB200_UN_EXP = 2,      /* :28  */
    NB200_UN_EXP2 = 3,     /* :37  */ NB200_UN_EXPM1 = 4,    /* :46  */ NB200_UN_LOG = 5,      /* :55  */
    NB200_UN_LOG2 = 6,     /* :91  */ NB200_UN_LOG10 = 7,    /* :64  */ NB200_UN_LOG1P = 8,    /* :73  */
    NB200_UN_LOGB = 9,     /* :82  */ NB200_UN_SIN = 10,     /* :99  */ NB200_UN_COS = 11,     /* :107 */
    NB200_UN_TAN = 12,     /* :132 */ NB200_UN_ARCSIN = 13,  /* :140 */ NB200_UN_ARCCOS = 14,  /* :144 */
    NB200_UN_ARCTAN = 15,  /* :152 */ NB200_UN_SINH = 16,    /* :164 */ NB200_UN_COSH = 17,    /* :168 */
    NB200_UN_TANH = 18,    /* :172 */ NB200_UN_ARCSINH = 19, /* :176 */ NB200_UN_ARCCOSH = 20, /* :180 */
    NB200_UN_ARCTANH = 21, /* :188 */ NB200_UN_DEGREES = 22, /* :156 */ NB200_UN_RADIANS = 23, /* :160 */
    NB200_UN_RINT = 24,    /* :200 */ NB200_UN_FIX = 25,     /* :212 */ NB200_UN_TRUNC = 26,   /* :224 */
    NB200_UN_FLOOR = 27,   /* :216 */ NB200_UN_CEIL = 28,    /* :220 */ NB200_UN_SINC = 29,    /* :228 */
    NB200_UN_NEGATIVE = 30,/* :237 */ NB200_UN_POSITIVE = 31,/* :241 (== abs) */ NB200_UN_SIGN = 32, /* :246 */
    NB200_UN_RECIPROCAL = 33, /* :263 */ NB200_UN_RSQRT = 34,/* :111 fast inverse sqrt, bit-exact */
    NB200_UN_CLIP = 35,    /* :250 p0=min p1=max */
    NB200_UN_ROUND = 36,   /* :254 p0=decimals */
    NB200_UN_SQUARE = 37,  /* numpower.c:3093 Multiply(a,a) */
    NB200_UN_COUNT = 38
};

enum nb200_reduce_op { NB200_SUM = 0, NB200_PROD = 1, NB200_MIN = 2, NB200_MAX = 3 };

/* Axis-reduction order.  TREE: deterministic split of the axis (fast, HBM-bound);
 * SEQUENTIAL: ((x0 op x1) op x2)... along the axis, the exact order of the
 * reference's reduce()/_reduce() slice loop (ndarray.c:394-429) => bit-identical. */
enum nb200_reduce_order { NB200_ORDER_TREE = 0, NB200_ORDER_SEQUENTIAL = 1 };

/* nd::matmul precision.
 * AUTO (what nd::matmul passes): the fastest mode whose error bound GUARANTEES 1e-5 against cblas_sgemm for every
 *   input - FP16X3U for K >= 128 (nb200_gemm_resolve_precision; NB200_GEMM_AUTO_MODE overrides), TF32X3 below (pre-pass not
 *   worth it) and for shapes too small for the tensor path.
 * TF32X3: error-compensated 3-pass TF32 on the tcgen05 tensor pipe; guaranteed bound ~3*2^-22 per product
 *   (+ chunked round-to-nearest accumulation).  Also what the FP16X3 / FP16X3U device-side fallback runs.
 * TF32X1: single pass, fast mode (~7e-4); not a parity mode.
 * BF16X3: operands split into two bfloat16 parts each, three kind::f16 MMAs per k-step at twice the TF32 rate.
 *   Statistical accuracy only: per-product error up to 2^-16 + 2*2^-17, zero-mean - measured 1.2-2.5e-6 on random
 *   data, but coherent inputs (constant matrices) can reach ~3e-5.  Opt-in per call (or NB200_GEMM_AUTO_MODE=bf16x3).
 * FP16X3: IEEE-half parts of A scaled per row and B scaled per column by powers of two (lo parts stored x2^11; the
 *   epilogue undoes all scaling exactly).  Every non-zero element within 2^-28 of its row / column maximum keeps a
 *   2^-22 relative split error: TF32X3-class guaranteed bound, no coherent-input problem.  The split pre-pass checks
 *   that window ON THE DEVICE: a few elements outside it are taken out of the GEMM and added back by a sparse fp32
 *   repair kernel; if there are too many, the gated TF32X3 fallback enqueued with the call produces the result
 *   instead (bit-identical to a TF32X3 call).  Any alignment / leading dimension (the pre-pass repacks).
 * FP16X3U: FP16X3 with the lo parts stored UNSCALED (no 2^11 factor).  All three products then carry one scale, so a whole
 *   256-long k-chunk accumulates into ONE TMEM accumulator and the 256x256 tile of BF16X3 applies: ~8 % less tensor time than
 *   FP16X3's 256x128 tile and a 2x shorter truncating accumulation (smaller systematic bias).  Price: lo parts below 2^-14 are
 *   subnormal halves, so the split error per element is max(2^-22 |a'|, 2^-25) - still 2^-22 relative within 2^-17 of the row /
 *   column maximum, 2^-19 at the edge of the (narrower) window, 2^-20 .. 2^-21 of the maximum; per product <= 2^-18 + 2^-22 =
 *   4.1e-6 worst case, every input.  Elements outside the window: same sparse repair / gated TF32X3 fallback as FP16X3. */
enum nb200_gemm_precision { NB200_GEMM_TF32X3 = 0, NB200_GEMM_TF32X1 = 1, NB200_GEMM_BF16X3 = 2, NB200_GEMM_AUTO = 3,
                            NB200_GEMM_FP16X3 = 4, NB200_GEMM_FP16X3U = 5 };

/* ---- context / device -------------------------------------------------------- */
/* Replaces the process-global cudaSetDevice of NDArray::setDevice (numpower.c:615-635). */
int nb200_init(int device);
int nb200_shutdown(void);
int nb200_device_count(int *count);
int nb200_set_device(int device);
int nb200_get_device(int *device);
int nb200_synchronize(void);
const char *nb200_last_error(void);
/* The CUDA stream (cudaStream_t) compute is enqueued on; nb200_set_stream lets a
 * harness (e.g. torch.cuda.current_stream().cuda_stream) time with its own events. */
void *nb200_stream(void);
int nb200_set_stream(void *cuda_stream);
/* Number of kernels this library has launched since init (bench.py "gpu_launches"). */
int64_t nb200_launch_count(void);
/* CUDA graphs: record everything the nb200_* calls between begin and end enqueue on the context stream, replay it with one launch.
 * The launch-latency path for sequences of small (L2-resident) operations; no allocation and no `*_host` entry point inside a
 * capture.  The reference synchronises after every kernel (cuda_math.cu wrappers: cudaDeviceSynchronize). */
int nb200_graph_begin(void);
int nb200_graph_end(void **graph_exec);
int nb200_graph_launch(void *graph_exec);
int nb200_graph_destroy(void *graph_exec);
/* Diagnostics: when `dev_slots` (>= 16 device uint64 words; "first" slots preset to ~0, "last" slots to 0) is set, the kernels
 * of the matmul pipeline stamp %globaltimer nanoseconds into it (slot map: csrc/common.cuh); NULL switches it off.
 * scripts/gemm_timeline.py turns the stamps into profiles/r2_gemm_timeline.json. */
int nb200_trace_enable(unsigned long long *dev_slots);
/* Returns and clears the sticky math-domain flag set by arccos/arccosh/arctanh
 * (synchronises). 0 = none. */
int nb200_poll_domain_error(int *flag);

/* ---- memory: replaces vmalloc/vfree/vmemcpyd2d/vmemcpyh2d/NDArray_VFLOAT/vmemcheck
 * (gpu_alloc.h:8-15) with 64-bit sizes. -------------------------------- */
int nb200_alloc(void **dev_ptr, int64_t bytes);
int nb200_free(void *dev_ptr);
int nb200_copy_h2d(void *dev_dst, const void *host_src, int64_t bytes);
int nb200_copy_d2h(void *host_dst, const void *dev_src, int64_t bytes);
int nb200_copy_d2d(void *dev_dst, const void *dev_src, int64_t bytes);
int nb200_memset_zero(void *dev_ptr, int64_t bytes);
int nb200_mem_stats(int64_t *live_allocations, int64_t *live_bytes);
/* pinned host staging (NDArray_ToGPU / ToCPU, ndarray.c:1037-1093, use pageable memcpy) */
int nb200_host_alloc(void **host_ptr, int64_t bytes);
int nb200_host_free(void *host_ptr);

/* ---- elementwise --------------------------------------------------------------- */
/* out[idx] = a[idx·a_strides] op b[idx·b_strides] over `out_shape` (C order, out
 * contiguous).  Strides are in ELEMENTS; 0 = broadcast along that dim.  Replaces
 * cuda_<op>_float + the NDArray_Broadcast materialisation (ndarray.c:1172-1294). */
int nb200_ew_binary(int op, float *out, const float *a, const float *b, int ndim,
                    const int64_t *out_shape, const int64_t *a_strides, const int64_t *b_strides);
/* scalar operand passed by value: replaces the NDArray_Fill temp (arithmetics.c:169-181). */
int nb200_ew_binary_scalar(int op, float *out, const float *a, float scalar, int scalar_is_lhs, int64_t n);
/* out = a*b + c with two roundings (mul then add, no FMA): one pass instead of the
 * two nd:: calls PHP makes for `$a * $b + $c` (numpower.c:193-229). */
int nb200_ew_mul_add(float *out, const float *a, const float *b, const float *c, int ndim,
                     const int64_t *out_shape, const int64_t *a_strides, const int64_t *b_strides,
                     const int64_t *c_strides);
/* out[i] = f(in[i]); out may BE in (same pointer: the legacy cuda_float_<op> are in-place; dispatched to kernels without
 * __restrict__ / non-coherent loads).  Partially overlapping ranges are not supported.  The same holds for flat nb200_ew_binary. */
int nb200_ew_unary(int op, float *out, const float *in, int64_t n, float p0, float p1);
int nb200_fill(float *out, float value, int64_t n);   /* cuda_fill_float, cuda_math.h:36 */

/* ---- reductions ----------------------------------------------------------------- */
/* Full reduction to a device scalar (async) / to the host (sync).  Replaces
 * cuda_sum_float/cuda_prod_float/cuda_max_float/cuda_min_float (cuda_math.h:31-32,35,66).
 * min/max follow the CPU NaN rule of NDArray_Min/Max (ndarray.c:752-772,939-959):
 * NaN at index 0 sticks, NaN elsewhere is skipped.  Deterministic (fixed 2-stage tree). */
int nb200_reduce_full(int op, float *dev_out, const float *in, int64_t n);
int nb200_reduce_full_host(int op, float *host_out, const float *in, int64_t n);
/* Input viewed as (outer, len, inner) contiguous; out (outer, inner).  Replaces the
 * reduce()/_reduce() slice loop (ndarray.c:394-429,523-578) and NDArray_MaxAxis (:781-844). */
int nb200_reduce_axis(int op, float *out, const float *in, int64_t outer, int64_t len, int64_t inner, int order);
/* argmax/argmin along the middle dim of (outer, len, inner); out (outer, inner) holds the
 * index as float32.  Semantics of float_argmax/float_argmin (calculation.c:9-59): first
 * occurrence; argmax skips NaN unless it is element 0, argmin returns the first NaN.
 * The reference has no GPU path ("GPU not supported.", calculation.c:75-78).  len < 2^32 - 1; the index is rounded to
 * float32 like the reference's (float)i (exact below 2^24). */
int nb200_argminmax(int is_max, float *out, const float *in, int64_t outer, int64_t len, int64_t inner);
int nb200_argminmax_host(int is_max, float *host_out, const float *in, int64_t n);

/* ---- matmul --------------------------------------------------------------------- */
/* Row-major C[M,N] = A[M,K]·B[K,N] (alpha=1, beta=0).  Replaces the cublasSgemm call in
 * NDArray_FMatmul (linalg.c:54-72).  tcgen05 TF32 tensor cores, fp32 accumulate in TMEM. */
int nb200_sgemm(float *C, const float *A, const float *B, int64_t M, int64_t N, int64_t K,
                int64_t lda, int64_t ldb, int64_t ldc, int precision);
/* batch of independent products; stride* in elements (0 = shared operand). */
int nb200_sgemm_batched(float *C, const float *A, const float *B, int64_t batch, int64_t M, int64_t N, int64_t K,
                        int64_t strideA, int64_t strideB, int64_t strideC, int precision);
/* nd::matmul on HOST operands (what `$a->gpu(); nd::matmul; ->cpu()` does in three steps, NDArray_ToGPU/ToCPU
 * ndarray.c:1037-1093): B is uploaded once, then row blocks of A stream in, are multiplied and stream out, so the
 * H2D and D2H copies overlap each other (full-duplex PCIe) and the compute (copy-in stream + two worker streams that
 * run split, GEMM and download of a block in stream order).  Blocking; C_host complete on return.
 * Host buffers should be pinned (nb200_host_alloc) for full PCIe speed; pageable memory works but is staged. */
int nb200_sgemm_host(float *C_host, const float *A_host, const float *B_host, int64_t M, int64_t N, int64_t K, int precision);
/* Batch of independent products with HOST operands (contiguous [batch][M][K], [batch][K][N] -> [batch][M][N]): chunks of matrices
 * upload, multiply (nb200_sgemm_batched, any precision) and download on three streams.  Blocking. */
int nb200_sgemm_batched_host(float *C_host, const float *A_host, const float *B_host, int64_t batch, int64_t M, int64_t N, int64_t K,
                             int precision);
/* scratch the 3xTF32 split needs for an (M,N,K,batch) problem; allocated lazily from the
 * context and reused (bytes reported for capacity planning). */
int nb200_sgemm_workspace_bytes(int64_t batch, int64_t M, int64_t N, int64_t K, int precision, int64_t *bytes);
/* The concrete mode NB200_GEMM_AUTO (or any other value, returned unchanged) stands for at inner dimension K
 * (default policy, NB200_GEMM_AUTO_MODE override).  Shapes / alignments a mode cannot serve still fall back to TF32X3. */
int nb200_gemm_resolve_precision(int precision, int64_t K);
/* y[rows] = A[rows,cols]·x[cols] — replaces cuda_float_multiply_matrix_vector (cuda_math.h:62). */
int nb200_gemv(float *y, const float *A, const float *x, int64_t rows, int64_t cols);
/* nd::all (NDArray_All, src/logic.c:25-58): *host_out = 1 iff no element equals 0 (NaN counts as non-zero, the rule of the
 * reference's scalar loop; its AVX2 body compares the 8-lane mask with 0x0F and so answers 0 for every array of >= 8 elements -
 * the intended semantics are implemented, pinned by tests/logic/001-ndarray-all.phpt).  Blocking; n == 0 -> 1. */
int nb200_all(int *host_out, const float *a, int64_t n);
/* nd::allclose (NDArray_AllClose / float_allclose, src/logic.c:718-771; the reference refuses device arrays): *host_out = 1 iff
 * no element has |a - b| > atol + rtol * |b| - the reference's predicate, evaluated on element i of both arrays (its loop
 * indexes element 4i + i*stride/4, i.e. out of bounds; pinned by tests/logic/002-ndarray-allclose.phpt).  Same-shape contiguous
 * operands of n elements.  Blocking; n == 0 -> 1. */
int nb200_allclose(int *host_out, const float *a, const float *b, int64_t n, float rtol, float atol);
/* out[cols,rows] = in[rows,cols]^T, out != in — replaces cuda_float_transpose (cuda_math.h:77). */
int nb200_transpose2d(float *out, const float *in, int64_t rows, int64_t cols);

/* ---- multi-GPU shards (SURVEY.md §8 e) -------------------------------------------------------
 * ONE host process drives the G GPUs of a box.  The reference's only multi-GPU affordance is NDArray::setDevice
 * (numpower.c:615-635 -> cudaSetDevice): it has no collective and no sharded operation, so everything here is new surface, not a
 * replacement.  An array is split along its first axis into G contiguous shards ("units" = rows, matrices of a batch or elements);
 * shard s is resident on devices[s] and every entry point takes one pointer per shard.  Calls are asynchronous on the per-device
 * context streams (nb200_shard_synchronize joins them) unless they return a host value.  nb200_set_device keeps working: each
 * device has its own context, allocator pool and workspace. */
enum nb200_transport {
    NB200_XFER_NCCL = 0,   /* one grouped ncclSend / ncclRecv per call (single-process communicators, ncclCommInitAll; NCCL is
                            * loaded at run time with dlopen("libnccl.so.2") on first use) */
    NB200_XFER_P2P = 1     /* cudaMemcpyPeerAsync over NVLink peer access, one stream per peer (copy engines) */
};
int nb200_shard_init(int ndev, const int *devices);   /* devices == NULL: 0 .. ndev-1 */
int nb200_shard_finalize(void);
int nb200_shard_count(int *ndev);
int nb200_shard_device(int shard, int *device);
/* balanced contiguous split of `units`: the first units % G shards own one unit more */
int nb200_shard_range(int64_t units, int shard, int64_t *first, int64_t *count);
int nb200_shard_split(int64_t units, int nshards, int shard, int64_t *first, int64_t *count);   /* same rule, pure host arithmetic */
int nb200_shard_synchronize(void);
/* (rows, row_elems) fp32 array on devices[root] <-> row shards (shard_ptrs[s] on devices[s]; shard_ptrs[root] may point into the
 * root array itself: no copy).  Ordered after / before the work on the context streams involved. */
int nb200_shard_scatter(float *const *shard_ptrs, const float *root_src, int64_t rows, int64_t row_elems, int root, int transport);
int nb200_shard_gather(float *root_dst, const float *const *shard_ptrs, int64_t rows, int64_t row_elems, int root, int transport);
/* sharded `gpu()` / `cpu()` (NDArray_ToGPU / NDArray_ToCPU, ndarray.c:1037-1093, are one blocking copy to the current device): host
 * array <-> row shards, every shard over its own device's PCIe link, all links at once, asynchronous on the context streams for
 * pinned host memory (nb200_host_alloc); nb200_shard_synchronize() completes them. */
int nb200_shard_upload(float *const *shard_ptrs, const float *host_src, int64_t rows, int64_t row_elems);
int nb200_shard_download(float *host_dst, const float *const *shard_ptrs, int64_t rows, int64_t row_elems);
/* resident shards, same-shape operands: one launch per device, no collective (arithmetics.c:160-926 per shard) */
int nb200_shard_ew_binary(int op, float *const *out, const float *const *a, const float *const *b, int64_t rows, int64_t row_elems);
int nb200_shard_ew_mul_add(float *const *out, const float *const *a, const float *const *b, const float *const *c, int64_t rows,
                           int64_t row_elems);
int nb200_shard_ew_unary(int op, float *const *out, const float *const *in, int64_t rows, int64_t row_elems, float p0, float p1);
/* full reductions over the concatenation of the shards (n_total elements split by nb200_shard_range): per-shard partials folded on
 * the host in shard order; NDArray_Min/Max NaN rule and float_argmax/argmin first-occurrence / NaN rules (calculation.c:9-59)
 * applied to the GLOBAL index space; the argmax index is exact before the final (float) conversion. */
int nb200_shard_reduce_full(int op, float *host_out, const float *const *in, int64_t n_total);
int nb200_shard_argminmax(int is_max, float *host_out, const float *const *in, int64_t n_total);
/* batched nd::matmul, batch dimension sharded.  _sharded: resident shards.  _scatter_gather: operands and result on devices[root];
 * scatter + products + gather pipelined per `chunk` matrices per shard (chunk i multiplies while chunk i+1 lands and chunk i-1
 * returns; the root multiplies its own share in place).  elapsed_ms != NULL: synchronises and reports the device time. */
int nb200_sgemm_batched_sharded(float *const *C, const float *const *A, const float *const *B, int64_t batch, int64_t M, int64_t N, int64_t K,
                                int precision);
int nb200_sgemm_batched_scatter_gather(float *C_root, const float *A_root, const float *B_root, int64_t batch, int64_t M, int64_t N, int64_t K,
                                       int precision, int root, int transport, int64_t chunk, float *elapsed_ms);

#ifdef __cplusplus
}
#endif
#endif /* NB200_H */
