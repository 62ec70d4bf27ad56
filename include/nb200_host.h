/*
 * nb200_host.h — host-side mirror of NumPower's NDArray operator interface for the hot path,
 * written against the nb200 C-ABI (include/nb200.h).  PHP is not available in the build image,
 * so this C++ layer stands where the extension's L1/L2 host code stands: same
 * function names (prefixed NB_), argument meaning and error behaviour as the reference's
 *   src/ndmath/arithmetics.h:6-17, src/ndmath/linalg.h:6-24, src/ndmath/calculation.h:9,
 *   src/ndarray.h:105-143, src/initializers.h
 * but every operation runs on the B200 through libnb200 (no CPU compute path).
 * Errors: functions returning a pointer return NULL and set NB_last_error() — the message
 * strings are the ones the reference passes to zend_throw_error.
 */
#ifndef NB200_HOST_H
#define NB200_HOST_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define NB_DEVICE_CPU 0   /* NDARRAY_DEVICE_CPU, src/ndarray.h:34 */
#define NB_DEVICE_GPU 1   /* NDARRAY_DEVICE_GPU, src/ndarray.h:35 */
#define NB_MAX_DIMS_AXIS 128  /* NDARRAY_MAX_DIMS: "no axis" sentinel of argmax/argmin (numpower.c:2573-2595) */

/* Mirror of struct NDArray (src/ndarray.h:61-74) with 64-bit extents. fp32 only. */
typedef struct NB_NDArray {
    float *data;            /* host pointer (device == CPU) or device pointer (device == GPU) */
    int ndim;
    int64_t shape[8];
    int64_t numel;
    int device;
    int refcount;
    struct NB_NDArray *base; /* non-NULL: view, data not owned (iterators.c:94-111) */
} NB_NDArray;

const char *NB_last_error(void);

/* construction / residency — initializers.c:255-286,379-448 ; ndarray.c:1037-1093 */
NB_NDArray *NB_NDArray_FromHost(const float *data, int ndim, const int64_t *shape);  /* CPU array, copies */
NB_NDArray *NB_NDArray_Empty(int ndim, const int64_t *shape, int device);
NB_NDArray *NB_NDArray_ToGPU(NB_NDArray *a);   /* $a->gpu() */
NB_NDArray *NB_NDArray_ToCPU(NB_NDArray *a);   /* $a->cpu() */
NB_NDArray *NB_NDArray_Slice0(NB_NDArray *a, int64_t index);  /* $a[i]: axis-0 view (NDArrayIterator_GET) */
NB_NDArray *NB_NDArray_Reshape(NB_NDArray *a, int ndim, const int64_t *shape); /* view, manipulation.c:138-162 */
void NB_NDArray_FREE(NB_NDArray *a);           /* ndarray.c:587-632 */
int NB_NDArray_CopyToHost(NB_NDArray *a, float *dst);  /* toArray() */

/* binary arithmetic — arithmetics.c:160,293,439,566,700,825 ; ndarray.c:852,895 ; double_math.c:259 */
NB_NDArray *NB_NDArray_Add_Float(NB_NDArray *a, NB_NDArray *b);
NB_NDArray *NB_NDArray_Subtract_Float(NB_NDArray *a, NB_NDArray *b);
NB_NDArray *NB_NDArray_Multiply_Float(NB_NDArray *a, NB_NDArray *b);
NB_NDArray *NB_NDArray_Divide_Float(NB_NDArray *a, NB_NDArray *b);
NB_NDArray *NB_NDArray_Mod_Float(NB_NDArray *a, NB_NDArray *b);
NB_NDArray *NB_NDArray_Pow_Float(NB_NDArray *a, NB_NDArray *b);
NB_NDArray *NB_NDArray_Maximum(NB_NDArray *a, NB_NDArray *b);
NB_NDArray *NB_NDArray_Minimum(NB_NDArray *a, NB_NDArray *b);
NB_NDArray *NB_NDArray_Arctan2(NB_NDArray *a, NB_NDArray *b);
NB_NDArray *NB_NDArray_Binary(int nb200_binary_op, NB_NDArray *a, NB_NDArray *b);
/* fused $a * $b + $c (one kernel; bit-identical to the two calls PHP makes) */
NB_NDArray *NB_NDArray_MulAdd(NB_NDArray *a, NB_NDArray *b, NB_NDArray *c);

/* unary maps — NDArray_Map/Map1F/Map2F (ndarray.c:682-744) with the double_math.c functor ids of nb200.h */
NB_NDArray *NB_NDArray_Map(NB_NDArray *a, int nb200_unary_op, float p0, float p1);

/* reductions — arithmetics.c:36-71 ; ndarray.c:752-772,939-959,523-578 */
int NB_NDArray_Sum_Float(NB_NDArray *a, float *out);
int NB_NDArray_Float_Prod(NB_NDArray *a, float *out);
int NB_NDArray_Min(NB_NDArray *a, float *out);
int NB_NDArray_Max(NB_NDArray *a, float *out);
NB_NDArray *NB_reduce(NB_NDArray *a, int axis, int nb200_reduce_op, int nb200_reduce_order);
/* calculation.c:73-194; axis == NB_MAX_DIMS_AXIS flattens */
NB_NDArray *NB_NDArray_ArgMinMaxCommon(NB_NDArray *a, int axis, int keepdims, int is_argmax);

/* linalg.c:216-245 (2-D; 3-D stacks map to the batched kernel: SURVEY N1), :354-393, :310-345 */
NB_NDArray *NB_NDArray_Matmul(NB_NDArray *a, NB_NDArray *b, int nb200_gemm_precision);
NB_NDArray *NB_NDArray_Dot(NB_NDArray *a, NB_NDArray *b);

#ifdef __cplusplus
}
#endif
#endif
