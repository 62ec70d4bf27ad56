/*
 * nb200_numpower_glue.c — Level-1 host patches for NumPower (SURVEY.md §8 f, N1; INTEGRATION.md "Level 1").
 *
 * This is the file a NumPower maintainer adds to the extension (e.g. as src/nb200_glue.c).  It is written against the
 * reference's own headers (src/ndarray.h, src/initializers.h, src/types.h) and the nb200 C-ABI (include/nb200.h), and is
 * called from ONE inserted line at the top of each hot-path function (oracle/n1_patch.py shows the exact insertions):
 *
 *   NDArray_{Add,Subtract,Multiply,Divide,Mod,Pow}_Float  -> nb200_glue_binary   (scalar by value, stride-0 broadcast:
 *                                                            no NDArray_Fill temp, no NDArray_Broadcast copy)
 *   reduce()                                              -> nb200_glue_reduce   (one launch instead of one per slice)
 *   NDArray_ArgMinMaxCommon                               -> nb200_glue_argminmax (the reference throws "GPU not supported.")
 *   NDArray_Matmul                                        -> nb200_glue_matmul   (2-D and stacks of matrices -> batched kernel)
 *   NDArray_Maximum / NDArray_Minimum                     -> nb200_glue_binary   (the reference throws "not implemented for GPU")
 *   NDArray_MaxAxis                                       -> nb200_glue_max_axis (the reference throws "GPU NDArray_MaxAxis not implemented")
 *   NDArray_Dot (N-D . 1-D)                               -> nb200_glue_dot      (every leading row, not only the last matrix)
 *   NDArray_ToGPU / NDArray_ToCPU                         -> nb200_glue_to_gpu / _to_cpu (the reference calls cudaMemcpy on emalloc'd,
 *                                                            i.e. pageable, memory: 10 / 21 GB/s; nb200_copy_h2d / d2h stage large
 *                                                            copies through pinned slots with worker threads: 37-40 GB/s; and no
 *                                                            zero-filled host copy is allocated first, ndarray.c:1051)
 *
 * Unary methods need no host patch: the legacy NDArrayMathGPU_ElementWise* drivers in libnb200.so recognise the
 * cuda_float_<op> pointer the PHP method passes (numpower.c:1648-3348) and run the op out of place.
 *
 * Every function returns NULL when the call is not a device call it serves; the reference code then runs unchanged.
 * Results are complete on return (the reference host reads them with default-stream cudaMemcpy right away).
 */
#include <php.h>
#include <string.h>
#include "ndarray.h"
#include "initializers.h"
#include "types.h"
#include "ndmath/arithmetics.h"
#include <nb200.h>

static void glue_throw(const char *what) { zend_throw_error(NULL, "%s: %s", what, nb200_last_error()); }

static NDArray *gpu_result(const int64_t *shape, int ndim) {
    int *sh = emalloc(sizeof(int) * (ndim > 0 ? ndim : 1));
    for (int i = 0; i < ndim; i++) sh[i] = (int) shape[i];
    return NDArray_Empty(sh, ndim, NDARRAY_TYPE_FLOAT32, NDARRAY_DEVICE_GPU);   /* vmalloc -> nb200_alloc */
}

/* NumPy-style broadcast of the two shapes; element strides with 0 on broadcast dims. */
static int glue_broadcast(NDArray *a, NDArray *b, int *ndim, int64_t *shape, int64_t *sa, int64_t *sb) {
    int n = NDArray_NDIM(a) > NDArray_NDIM(b) ? NDArray_NDIM(a) : NDArray_NDIM(b);
    int64_t stra = 1, strb = 1;
    if (n > NB200_MAX_DIMS) return 0;
    for (int i = n - 1; i >= 0; i--) {
        int ia = i - (n - NDArray_NDIM(a)), ib = i - (n - NDArray_NDIM(b));
        int64_t da = ia >= 0 ? NDArray_SHAPE(a)[ia] : 1, db = ib >= 0 ? NDArray_SHAPE(b)[ib] : 1;
        if (da != db && da != 1 && db != 1) return 0;
        shape[i] = da == 1 ? db : da;
        sa[i] = da == 1 ? 0 : stra;
        sb[i] = db == 1 ? 0 : strb;
        stra *= da;
        strb *= db;
    }
    *ndim = n;
    return 1;
}

NDArray *nb200_glue_binary(int op, NDArray *a, NDArray *b) {
    const int a_scalar = NDArray_NDIM(a) == 0, b_scalar = NDArray_NDIM(b) == 0;
    if (a_scalar && b_scalar) return NULL;
    NDArray *arr = a_scalar ? b : a;
    if (NDArray_DEVICE(arr) != NDARRAY_DEVICE_GPU) return NULL;
    if (a_scalar || b_scalar) {                       /* arithmetics.c:169-181 materialises the scalar; here: by value */
        NDArray *sc = a_scalar ? a : b;
        float s = NDArray_DEVICE(sc) == NDARRAY_DEVICE_GPU ? NDArray_GetFloatScalar(sc) : NDArray_FDATA(sc)[0];
        int64_t shape[NB200_MAX_DIMS];
        for (int i = 0; i < NDArray_NDIM(arr); i++) shape[i] = NDArray_SHAPE(arr)[i];
        NDArray *r = gpu_result(shape, NDArray_NDIM(arr));
        if (nb200_ew_binary_scalar(op, NDArray_FDATA(r), NDArray_FDATA(arr), s, a_scalar, NDArray_NUMELEMENTS(arr)) != NB200_OK ||
            nb200_synchronize() != NB200_OK) { glue_throw("nb200_ew_binary_scalar"); NDArray_FREE(r); return NULL; }
        return r;
    }
    if (NDArray_DEVICE(a) != NDARRAY_DEVICE_GPU || NDArray_DEVICE(b) != NDARRAY_DEVICE_GPU) return NULL;   /* reference reports the mismatch */
    int ndim;
    int64_t shape[NB200_MAX_DIMS], sa[NB200_MAX_DIMS], sb[NB200_MAX_DIMS];
    if (!glue_broadcast(a, b, &ndim, shape, sa, sb)) return NULL;       /* reference raises "Can't broadcast arrays." */
    NDArray *r = gpu_result(shape, ndim);
    if (nb200_ew_binary(op, NDArray_FDATA(r), NDArray_FDATA(a), NDArray_FDATA(b), ndim, shape, sa, sb) != NB200_OK ||
        nb200_synchronize() != NB200_OK) { glue_throw("nb200_ew_binary"); NDArray_FREE(r); return NULL; }
    return r;
}

/* reduce(array, axis, operation) ndarray.c:523-578: operation is NDArray_Add_Float (sum) or NDArray_Multiply_Float (prod) */
NDArray *nb200_glue_reduce(NDArray *array, int *axis, NDArray *(*operation)(NDArray *, NDArray *)) {
    if (NDArray_DEVICE(array) != NDARRAY_DEVICE_GPU) return NULL;
    int op = operation == NDArray_Add_Float ? NB200_SUM : (operation == NDArray_Multiply_Float ? NB200_PROD : -1);
    int ax = axis ? *axis : 0;
    if (op < 0 || ax < 0 || ax >= NDArray_NDIM(array)) return NULL;     /* reference raises its own out-of-bounds error */
    int64_t outer = 1, inner = 1, oshape[NB200_MAX_DIMS];
    int j = 0;
    for (int i = 0; i < NDArray_NDIM(array); i++) {
        if (i < ax) outer *= NDArray_SHAPE(array)[i];
        if (i > ax) inner *= NDArray_SHAPE(array)[i];
        if (i != ax) oshape[j++] = NDArray_SHAPE(array)[i];
    }
    NDArray *r = gpu_result(oshape, NDArray_NDIM(array) - 1);
    if (nb200_reduce_axis(op, NDArray_FDATA(r), NDArray_FDATA(array), outer, NDArray_SHAPE(array)[ax], inner, NB200_ORDER_TREE) != NB200_OK ||
        nb200_synchronize() != NB200_OK) { glue_throw("nb200_reduce_axis"); NDArray_FREE(r); return NULL; }
    return r;
}

/* NDArray_ArgMinMaxCommon calculation.c:73-194 (axis 128 = NDARRAY_MAX_DIMS flattens) */
NDArray *nb200_glue_argminmax(NDArray *op, int axis, bool keepdims, bool is_argmax) {
    if (NDArray_DEVICE(op) != NDARRAY_DEVICE_GPU) return NULL;
    int64_t outer = 1, m, inner = 1, oshape[NB200_MAX_DIMS];
    int ondim = 0;
    if (axis == NDARRAY_MAX_DIMS || NDArray_NDIM(op) == 0) {
        m = NDArray_NUMELEMENTS(op);
        ondim = keepdims ? NDArray_NDIM(op) : 0;
        for (int i = 0; i < ondim; i++) oshape[i] = 1;
    } else {
        if (axis < 0) axis += NDArray_NDIM(op);
        if (axis < 0 || axis >= NDArray_NDIM(op)) { zend_throw_error(NULL, "Invalid axis parameter"); return NULL; }
        m = NDArray_SHAPE(op)[axis];
        for (int i = 0; i < NDArray_NDIM(op); i++) {
            if (i < axis) outer *= NDArray_SHAPE(op)[i];
            if (i > axis) inner *= NDArray_SHAPE(op)[i];
            if (i != axis) oshape[ondim++] = NDArray_SHAPE(op)[i];
            else if (keepdims) oshape[ondim++] = 1;
        }
    }
    if (m == 0) { zend_throw_error(NULL, "attempt to get %s of an empty sequence", is_argmax ? "argmax" : "argmin"); return NULL; }
    NDArray *r = gpu_result(oshape, ondim);
    if (nb200_argminmax(is_argmax, NDArray_FDATA(r), NDArray_FDATA(op), outer, m, inner) != NB200_OK || nb200_synchronize() != NB200_OK) {
        glue_throw("nb200_argminmax"); NDArray_FREE(r); return NULL;
    }
    return r;
}

/* NDArray_Matmul linalg.c:216-245: 2-D, and stacks with equal leading dims as one batched launch */
NDArray *nb200_glue_matmul(NDArray *a, NDArray *b) {
    if (NDArray_DEVICE(a) != NDARRAY_DEVICE_GPU || NDArray_DEVICE(b) != NDARRAY_DEVICE_GPU) return NULL;
    const int nd = NDArray_NDIM(a);
    if (nd != NDArray_NDIM(b) || nd < 2 || nd > NB200_MAX_DIMS) return NULL;
    if (NDArray_SHAPE(a)[nd - 1] != NDArray_SHAPE(b)[nd - 2]) return NULL;           /* reference raises the shape mismatch */
    int64_t batch = 1, oshape[NB200_MAX_DIMS];
    for (int i = 0; i < nd - 2; i++) {
        if (NDArray_SHAPE(a)[i] != NDArray_SHAPE(b)[i]) return NULL;
        batch *= NDArray_SHAPE(a)[i];
        oshape[i] = NDArray_SHAPE(a)[i];
    }
    const int64_t M = NDArray_SHAPE(a)[nd - 2], K = NDArray_SHAPE(a)[nd - 1], N = NDArray_SHAPE(b)[nd - 1];
    oshape[nd - 2] = M;
    oshape[nd - 1] = N;
    NDArray *r = gpu_result(oshape, nd);
    int rc = batch == 1 ? nb200_sgemm(NDArray_FDATA(r), NDArray_FDATA(a), NDArray_FDATA(b), M, N, K, K, N, N, NB200_GEMM_AUTO)
                        : nb200_sgemm_batched(NDArray_FDATA(r), NDArray_FDATA(a), NDArray_FDATA(b), batch, M, N, K, M * K, K * N, M * N,
                                              NB200_GEMM_AUTO);
    if (rc != NB200_OK || nb200_synchronize() != NB200_OK) { glue_throw("nb200_sgemm"); NDArray_FREE(r); return NULL; }
    return r;
}

/* NDArray_MaxAxis ndarray.c:781-844: max over one axis, shape without that axis.  The reference's CPU loop only indexes 2-D inputs
 * correctly; on the device any ndim is served (outer x axis x inner decomposition).  NaN rule of the CPU loop (`current > best`). */
NDArray *nb200_glue_max_axis(NDArray *target, int axis) {
    if (NDArray_DEVICE(target) != NDARRAY_DEVICE_GPU) return NULL;
    if (axis < 0 || axis >= NDArray_NDIM(target)) { zend_throw_error(NULL, "Invalid axis.\n"); return NULL; }
    int64_t outer = 1, inner = 1, oshape[NB200_MAX_DIMS];
    int j = 0;
    if (NDArray_NDIM(target) > NB200_MAX_DIMS) return NULL;
    for (int i = 0; i < NDArray_NDIM(target); i++) {
        if (i < axis) outer *= NDArray_SHAPE(target)[i];
        if (i > axis) inner *= NDArray_SHAPE(target)[i];
        if (i != axis) oshape[j++] = NDArray_SHAPE(target)[i];
    }
    NDArray *r = gpu_result(oshape, NDArray_NDIM(target) - 1);
    if (nb200_reduce_axis(NB200_MAX, NDArray_FDATA(r), NDArray_FDATA(target), outer, NDArray_SHAPE(target)[axis], inner, NB200_ORDER_TREE) != NB200_OK ||
        nb200_synchronize() != NB200_OK) { glue_throw("nb200_reduce_axis"); NDArray_FREE(r); return NULL; }
    return r;
}

/* NDArray_Dot linalg.c:354-393, the N-D . 1-D case: y = A x over ALL leading rows (the reference's GPU branch passes
 * shape[ndim-2] rows to cuda_float_multiply_matrix_vector, i.e. only the first matrix of a stack is computed). */
NDArray *nb200_glue_dot(NDArray *nda, NDArray *ndb) {
    if (NDArray_DEVICE(nda) != NDARRAY_DEVICE_GPU || NDArray_DEVICE(ndb) != NDARRAY_DEVICE_GPU) return NULL;
    if (NDArray_NDIM(nda) < 2 || NDArray_NDIM(ndb) != 1 || NDArray_NDIM(nda) > NB200_MAX_DIMS) return NULL;   /* other cases: reference dispatch */
    const int nd = NDArray_NDIM(nda);
    const int64_t cols = NDArray_SHAPE(nda)[nd - 1];
    if (cols != NDArray_SHAPE(ndb)[0]) return NULL;
    int64_t rows = 1, oshape[NB200_MAX_DIMS];
    for (int i = 0; i < nd - 1; i++) { rows *= NDArray_SHAPE(nda)[i]; oshape[i] = NDArray_SHAPE(nda)[i]; }
    NDArray *r = gpu_result(oshape, nd - 1);
    if (nb200_gemv(NDArray_FDATA(r), NDArray_FDATA(nda), NDArray_FDATA(ndb), rows, cols) != NB200_OK || nb200_synchronize() != NB200_OK) {
        glue_throw("nb200_gemv"); NDArray_FREE(r); return NULL;
    }
    return r;
}

/* NDArray_ToGPU ndarray.c:1037-1068 / NDArray_ToCPU :1075-1093 for host <-> device moves (device -> device and host -> host copies
 * stay with the reference's NDArray_Copy).  nb200_copy_h2d / d2h are blocking, like the cudaMemcpy + cudaDeviceSynchronize they replace. */
NDArray *nb200_glue_to_gpu(NDArray *target) {
    if (NDArray_DEVICE(target) != NDARRAY_DEVICE_CPU || NDArray_NDIM(target) > NB200_MAX_DIMS) return NULL;
    int64_t shape[NB200_MAX_DIMS];
    for (int i = 0; i < NDArray_NDIM(target); i++) shape[i] = NDArray_SHAPE(target)[i];
    NDArray *r = gpu_result(shape, NDArray_NDIM(target));
    if (r == NULL) return NULL;
    if (nb200_copy_h2d(NDArray_FDATA(r), NDArray_FDATA(target), (int64_t) NDArray_NUMELEMENTS(target) * (int64_t) sizeof(float)) != NB200_OK) {
        glue_throw("nb200_copy_h2d"); NDArray_FREE(r); return NULL;
    }
    return r;
}

NDArray *nb200_glue_to_cpu(NDArray *target) {
    if (NDArray_DEVICE(target) != NDARRAY_DEVICE_GPU) return NULL;
    int *sh = emalloc(sizeof(int) * (NDArray_NDIM(target) > 0 ? NDArray_NDIM(target) : 1));
    memcpy(sh, NDArray_SHAPE(target), sizeof(int) * NDArray_NDIM(target));
    NDArray *r = NDArray_Empty(sh, NDArray_NDIM(target), NDARRAY_TYPE_FLOAT32, NDARRAY_DEVICE_CPU);
    if (r == NULL) return NULL;
    if (nb200_copy_d2h(NDArray_FDATA(r), NDArray_FDATA(target), (int64_t) NDArray_NUMELEMENTS(target) * (int64_t) sizeof(float)) != NB200_OK) {
        glue_throw("nb200_copy_d2h"); NDArray_FREE(r); return NULL;
    }
    return r;
}
