/*
 * nb200_legacy.h — the reference's existing device-backend symbols, re-exported by
 * libnb200.so so that NumPower's UNMODIFIED host objects (numpower.c, src/ndarray.c,
 * src/ndmath/{arithmetics,linalg}.c, src/initializers.c, src/logic.c ... compiled with
 * -DHAVE_CUBLAS) link against the B200 backend instead of cuda_math.o + gpu_alloc.o.
 *
 * Every prototype below is the reference's own (file:line in /root/reference); the
 * implementations (numpower_b200/csrc/legacy.cu) are thin adapters onto include/nb200.h.
 * Contract kept from the reference: `void` returns, blocking (results visible on return),
 * errors reported through zend_throw_error (resolved from the host at load time; weak).
 * `nblocks` parameters are element counts in the reference and are ignored here.
 */
#ifndef NB200_LEGACY_H
#define NB200_LEGACY_H
#ifdef __cplusplus
extern "C" {
#endif

/* ---- src/gpu_alloc.h:8-15 -------------------------------------------------------- */
void vmalloc(void **target, unsigned int size);                 /* :8  (size < 4 GiB: SURVEY F3) */
void vfree(void *target);                                       /* :9  */
void vmemcheck(void);                                           /* :10 */
void vmemcpyd2d(char *src, char *dst, unsigned int size);       /* :11 (argument order is src, dst) */
void vmemcpyh2d(char *src, char *dst, unsigned int size);       /* :12 */
float NDArray_VFLOAT(char *target);                             /* :14 */
float NDArray_VFLOATF_I(float *target, int index);              /* :15 */

/* ---- src/ndmath/cuda/cuda_math.h:25-36,66 : binary ops, reductions, fill ------------- */
void cuda_add_float(int nblocks, float *a, float *b, float *rtn, int nelements);       /* :25 */
void cuda_subtract_float(int nblocks, float *a, float *b, float *rtn, int nelements);  /* :26 */
void cuda_divide_float(int nblocks, float *a, float *b, float *rtn, int nelements);    /* :27 */
void cuda_multiply_float(int nblocks, float *a, float *b, float *rtn, int nelements);  /* :28 */
void cuda_mod_float(int nblocks, float *a, float *b, float *rtn, int nelements);       /* :29 */
void cuda_pow_float(int nblocks, float *a, float *b, float *rtn, int nelements);       /* :33 */
float cuda_max_float(float *a, int nelements);                                         /* :31 */
float cuda_min_float(float *a, int nelements);                                         /* :32 */
int cuda_equal_float(int nblocks, float *a, float *b, int nelements);                  /* :34 */
void cuda_sum_float(int nblocks, float *a, float *rtn /* host in/out */, int nelements);  /* :35 */
void cuda_prod_float(int nblocks, float *a, float *rtn /* host in/out */, int nelements); /* :66 */
void cuda_fill_float(float *a, float value, int n);                                    /* :36 */

/* ---- cuda_math.h:16-24,38-61,67,78-79 : in-place unaries ------------------------------ */
void cuda_float_abs(int nblocks, float *d_array);        void cuda_float_expm1(int nblocks, float *d_array);
void cuda_float_exp(int nblocks, float *d_array);        void cuda_float_sqrt(int nblocks, float *d_array);
void cuda_float_log(int nblocks, float *d_array);        void cuda_float_logb(int nblocks, float *d_array);
void cuda_float_log2(int nblocks, float *d_array);       void cuda_float_log1p(int nblocks, float *d_array);
void cuda_float_log10(int nblocks, float *d_array);      void cuda_float_sin(int nblocks, float *d_array);
void cuda_float_cos(int nblocks, float *d_array);        void cuda_float_tan(int nblocks, float *d_array);
void cuda_float_arcsin(int nblocks, float *d_array);     void cuda_float_arccos(int nblocks, float *d_array);
void cuda_float_arctan(int nblocks, float *d_array);     void cuda_float_degrees(int nblocks, float *d_array);
void cuda_float_radians(int nblocks, float *d_array);    void cuda_float_sinh(int nblocks, float *d_array);
void cuda_float_cosh(int nblocks, float *d_array);       void cuda_float_tanh(int nblocks, float *d_array);
void cuda_float_arcsinh(int nblocks, float *d_array);    void cuda_float_arccosh(int nblocks, float *d_array);
void cuda_float_arctanh(int nblocks, float *d_array);    void cuda_float_rint(int nblocks, float *d_array);
void cuda_float_fix(int nblocks, float *d_array);        void cuda_float_ceil(int nblocks, float *d_array);
void cuda_float_floor(int nblocks, float *d_array);      void cuda_float_sinc(int nblocks, float *d_array);
void cuda_float_trunc(int nblocks, float *d_array);      void cuda_float_negate(int nblocks, float *d_array);
void cuda_float_sign(int nblocks, float *d_array);       void cuda_float_positive(int nblocks, float *d_array);
void cuda_float_reciprocal(int nblocks, float *d_array);
void cuda_float_clip(int nblocks, float *d_array, float minVal, float maxVal);          /* :61 */
void cuda_float_round(int nblocks, float *d_array, float decimals);                     /* :67 */
void cuda_float_arctan2(int nblocks, float *d_array, float *y_array);                   /* :44 */

/* ---- cuda_math.h:62,77 and :63,69-73 --------------------------------------------------- */
void cuda_float_multiply_matrix_vector(int nblocks, float *a_array, float *b_array, float *result, int rows, int cols);
void cuda_float_transpose(int tiledim, int blockrows, const float *d_in, float *d_out, int width, int height);
void cuda_float_compare_equal(int nblocks, float *a, float *b, float *result, int n);
void cuda_float_compare_not_equal(int nblocks, float *a, float *b, float *result, int n);
void cuda_float_compare_greater(int nblocks, float *a, float *b, float *result, int n);
void cuda_float_compare_greater_equal(int nblocks, float *a, float *b, float *result, int n);
void cuda_float_compare_less(int nblocks, float *a, float *b, float *result, int n);
void cuda_float_compare_less_equal(int nblocks, float *a, float *b, float *result, int n);

/* ---- cuda_math.h:14-15,75-76 : NDArray-level unary drivers (copy + in-place op).  They need the
 * host's struct NDArray (src/ndarray.h:52-74) and NDArray_Copy (src/initializers.h:31). ------- */
struct NDArray;
typedef void (*ElementWiseFloatGPUOperation)(int, float *);
typedef void (*ElementWiseFloatGPUOperation2F)(int, float *, float, float);
typedef void (*ElementWiseFloatGPUOperation1F)(int, float *, float);
typedef void (*ElementWiseFloatGPUOperation1N)(int, float *, float *);
struct NDArray *NDArrayMathGPU_ElementWise(struct NDArray *ndarray, ElementWiseFloatGPUOperation op);
struct NDArray *NDArrayMathGPU_ElementWise1F(struct NDArray *ndarray, ElementWiseFloatGPUOperation1F op, float val1);
struct NDArray *NDArrayMathGPU_ElementWise2F(struct NDArray *ndarray, ElementWiseFloatGPUOperation2F op, float val1, float val2);
struct NDArray *NDArrayMathGPU_ElementWise1N(struct NDArray *ndarray, ElementWiseFloatGPUOperation1N op, struct NDArray *val1);

/* ---- out-of-scope exports (dense factorizations, conv, LU, median): they must
 * LINK; calling one raises "not implemented in the B200 backend" through zend_throw_error. --- */
int cuda_svd_float(float *d_A, float *d_U, float *d_V, float *d_S, int m, int n);
int cuda_det_float(float *a, float *result, int n);
void cuda_matrix_float_inverse(float *matrix, int n);
void cuda_float_lu(float *matrix, float *L, float *U, float *P, int size);
void cuda_lstsq_float(float *A, int m, int n, float *B, int k, float *X);
void cuda_calculate_outer_product(int m, int n, float *a_array, float *b_array, float *r_array);

#ifdef __cplusplus
}
#endif
#endif
