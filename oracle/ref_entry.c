/* TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * Thin raw-pointer entry points onto the reference's OWN CPU object code
 * (compiled from /root/reference by oracle/build_ref.sh into
 * oracle/_ref/libnumpower_ref.so).  Each ref_* function wraps caller buffers
 * in the reference's `struct NDArray` (src/ndarray.h:61-74) and calls the
 * reference function named in its comment.  No arithmetic happens here.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.
 */
#include <php.h>
#include <time.h>
#include "ndarray.h"
#include "initializers.h"
#include "iterators.h"
#include "types.h"
#include "manipulation.h"
#include "ndmath/arithmetics.h"
#include "ndmath/double_math.h"
#include "ndmath/calculation.h"
#include "ndmath/linalg.h"
#include "logic.h"

/* ---- Zend runtime pieces the reference objects link against ------------- */
static char g_err[1024];
static int g_err_set = 0;

void zend_throw_error(zend_class_entry *ce, const char *format, ...) {
    (void) ce;
    va_list ap;
    va_start(ap, format);
    vsnprintf(g_err, sizeof(g_err), format, ap);
    va_end(ap);
    g_err_set = 1;
}
void zend_error(int type, const char *format, ...) {
    (void) type;
    va_list ap;
    va_start(ap, format);
    vsnprintf(g_err, sizeof(g_err), format, ap);
    va_end(ap);
    g_err_set = 1;
}
void php_error_docref(const char *docref, int type, const char *format, ...) {
    (void) docref; (void) type;
    va_list ap;
    va_start(ap, format);
    vsnprintf(g_err, sizeof(g_err), format, ap);
    va_end(ap);
    g_err_set = 1;
}
/* debug.c is not compiled (printing is out of scope); ndarray.c references these. */
char *print_matrix(double *b, int nd, int *sh, int *st, int n, int dev) {
    (void) b; (void) nd; (void) sh; (void) st; (void) n; (void) dev; return strdup("");
}
char *print_matrix_float(float *b, int nd, int *sh, int *st, int n, int dev) {
    (void) b; (void) nd; (void) sh; (void) st; (void) n; (void) dev; return strdup("");
}
#ifndef REF_ENTRY_GPU
/* linalg.c:204 references vfree outside #ifdef HAVE_CUBLAS */
void vfree(void *p) { (void) p; }
#endif

const char *ref_last_error(void) { return g_err_set ? g_err : ""; }
void ref_clear_error(void) { g_err_set = 0; g_err[0] = 0; }

/* ---- wrapping -----------------------------------------------------------
 * REF_ENTRY_GPU (drop-in build, oracle/build_dropin.sh): operands with ndim > 0 are moved to the
 * device with the reference's own NDArray_ToGPU (ndarray.c:1037-1068), the reference function then
 * takes its `NDArray_DEVICE(a) == NDARRAY_DEVICE_GPU` branch — i.e. calls the legacy cuda_* /
 * vmalloc symbols that libnb200.so now provides — and the result comes back with NDArray_ToCPU. */
static NDArray *wrap_cpu(const float *data, const int *shape, int ndim) {
    int *sh = (int *) malloc(sizeof(int) * (ndim > 0 ? ndim : 1));
    for (int i = 0; i < ndim; i++) sh[i] = shape[i];
    if (ndim == 0) sh[0] = 1;
    NDArray *a = Create_NDArray(sh, ndim, NDARRAY_TYPE_FLOAT32, NDARRAY_DEVICE_CPU);
    a->data = (char *) data;
    return a;
}
static void unwrap_cpu(NDArray *a) {
    if (a == NULL) return;
    a->data = NULL;            /* caller-owned buffer: NDArray_FREE must not efree it */
    NDArray_FREE(a);
}
#ifdef REF_ENTRY_GPU
static NDArray *wrap(const float *data, const int *shape, int ndim) {
    NDArray *c = wrap_cpu(data, shape, ndim);
    if (ndim == 0) return c;   /* PHP scalars stay 0-dim CPU arrays (ZVAL_TO_NDARRAY, numpower.c:89-117) */
    NDArray *g = NDArray_ToGPU(c);
    unwrap_cpu(c);
    return g;
}
static void unwrap(NDArray *a) {
    if (a == NULL) return;
    if (NDArray_DEVICE(a) == NDARRAY_DEVICE_GPU) NDArray_FREE(a); else unwrap_cpu(a);
}
static long take(NDArray *r, float *out, long cap) {
    if (r == NULL) return -1;
    NDArray *h = r;
    if (NDArray_DEVICE(r) == NDARRAY_DEVICE_GPU) { h = NDArray_ToCPU(r); NDArray_FREE(r); }
    long n = NDArray_NUMELEMENTS(h);
    if (n > cap) n = cap;
    memcpy(out, NDArray_FDATA(h), (size_t) n * sizeof(float));
    long total = NDArray_NUMELEMENTS(h);
    NDArray_FREE(h);
    return total;
}
#else
#define wrap wrap_cpu
#define unwrap unwrap_cpu
static long take(NDArray *r, float *out, long cap) {
    if (r == NULL) return -1;
    long n = NDArray_NUMELEMENTS(r);
    if (n > cap) n = cap;
    memcpy(out, NDArray_FDATA(r), (size_t) n * sizeof(float));
    long total = NDArray_NUMELEMENTS(r);
    NDArray_FREE(r);
    return total;
}
#endif
static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double) ts.tv_sec + 1e-9 * (double) ts.tv_nsec;
}

/* ---- binary elementwise --------------------------------------------------
 * op: 0 add (arithmetics.c:160) 1 sub (:439) 2 mul (:293) 3 div (:566)
 *     4 mod (:700) 5 pow (:825) 6 maximum (ndarray.c:852) 7 minimum (:895)
 *     8 arctan2 (ndarray.c:715 NDArray_Map1ND + double_math.c:259)
 * Returns the number of result elements (-1 on error); if `seconds` != NULL
 * it receives the wall time of the reference call alone. */
long ref_binary(int op, const float *a, const int *ashape, int andim,
                const float *b, const int *bshape, int bndim,
                float *out, long out_cap, double *seconds) {
    NDArray *na = wrap(a, ashape, andim), *nb = wrap(b, bshape, bndim), *r = NULL;
    double t0 = now_s();
    switch (op) {
        case 0: r = NDArray_Add_Float(na, nb); break;
        case 1: r = NDArray_Subtract_Float(na, nb); break;
        case 2: r = NDArray_Multiply_Float(na, nb); break;
        case 3: r = NDArray_Divide_Float(na, nb); break;
        case 4: r = NDArray_Mod_Float(na, nb); break;
        case 5: r = NDArray_Pow_Float(na, nb); break;
        case 6: r = NDArray_Maximum(na, nb); break;
        case 7: r = NDArray_Minimum(na, nb); break;
        case 8: r = NDArray_Map1ND(na, float_arctan2, nb); break;
        /* comparisons -> 0/1 masks: src/logic.c:68 (greater), :172 (less), :272 (less_equal), :378 (greater_equal),
         * :479 (equal), :580 (not_equal); ids = NB200_CMP_* of include/nb200.h */
        case 10: r = NDArray_Equal(na, nb); break;
        case 11: r = NDArray_NotEqual(na, nb); break;
        case 12: r = NDArray_Greater(na, nb); break;
        case 13: r = NDArray_GreaterEqual(na, nb); break;
        case 14: r = NDArray_Less(na, nb); break;
        case 15: r = NDArray_LessEqual(na, nb); break;
        default: break;
    }
    double t1 = now_s();
    if (seconds) *seconds = t1 - t0;
    long n = take(r, out, out_cap);
    unwrap(na); unwrap(nb);
    return n;
}

/* a*b+c exactly as unchanged PHP evaluates it: ZEND_MUL then ZEND_ADD
 * (numpower.c:193-229 -> arithmetics.c:293, :160), intermediate freed. */
long ref_mul_add(const float *a, const int *ashape, int andim,
                 const float *b, const int *bshape, int bndim,
                 const float *c, const int *cshape, int cndim,
                 float *out, long out_cap, double *seconds) {
    NDArray *na = wrap(a, ashape, andim), *nb = wrap(b, bshape, bndim), *nc = wrap(c, cshape, cndim);
    double t0 = now_s();
    NDArray *m = NDArray_Multiply_Float(na, nb);
    NDArray *r = m ? NDArray_Add_Float(m, nc) : NULL;
    if (m) NDArray_FREE(m);
    double t1 = now_s();
    if (seconds) *seconds = t1 - t0;
    long n = take(r, out, out_cap);
    unwrap(na); unwrap(nb); unwrap(nc);
    return n;
}

/* ---- unary elementwise: NDArray_Map / Map1F / Map2F (ndarray.c:682-744) ---
 * op ids follow include/nb200.h NB200_UN_* (table in SURVEY.md 8 a-3). */
typedef float (*un_fn)(float);
static un_fn un_table(int op) {
    switch (op) {
        case 0: return float_abs;      case 1: return float_sqrt;    case 2: return float_exp;
        case 3: return float_exp2;     case 4: return float_expm1;   case 5: return float_log;
        case 6: return float_log2;     case 7: return float_log10;   case 8: return float_log1p;
        case 9: return float_logb;     case 10: return float_sin;    case 11: return float_cos;
        case 12: return float_tan;     case 13: return float_arcsin; case 14: return float_arccos;
        case 15: return float_arctan;  case 16: return float_sinh;   case 17: return float_cosh;
        case 18: return float_tanh;    case 19: return float_arcsinh; case 20: return float_arccosh;
        case 21: return float_arctanh; case 22: return float_degrees; case 23: return float_radians;
        case 24: return float_rint;    case 25: return float_fix;    case 26: return float_trunc;
        case 27: return float_floor;   case 28: return float_ceil;   case 29: return float_sinc;
        case 30: return float_negate;  case 31: return float_positive; case 32: return float_sign;
        case 33: return float_reciprocal; case 34: return float_rsqrt;
        default: return NULL;
    }
}
#ifdef REF_ENTRY_GPU
/* GPU dispatch exactly as the PHP methods do it (numpower.c:1648...3348: one
 * `NDArrayMathGPU_ElementWise(nda, cuda_float_<op>)` per method).  Ops whose PHP method has no GPU
 * branch in the reference (exp2 numpower.c:3153, rsqrt :1788-1791 wrong branch) return NULL here. */
#include "ndmath/cuda/cuda_math.h"
typedef void (*gpu_fn)(int, float *);
static gpu_fn gpu_table(int op) {
    switch (op) {
        case 0: return cuda_float_abs;     case 1: return cuda_float_sqrt;    case 2: return cuda_float_exp;
        case 4: return cuda_float_expm1;   case 5: return cuda_float_log;     case 6: return cuda_float_log2;
        case 7: return cuda_float_log10;   case 8: return cuda_float_log1p;   case 9: return cuda_float_logb;
        case 10: return cuda_float_sin;    case 11: return cuda_float_cos;    case 12: return cuda_float_tan;
        case 13: return cuda_float_arcsin; case 14: return cuda_float_arccos; case 15: return cuda_float_arctan;
        case 16: return cuda_float_sinh;   case 17: return cuda_float_cosh;   case 18: return cuda_float_tanh;
        case 19: return cuda_float_arcsinh; case 20: return cuda_float_arccosh; case 21: return cuda_float_arctanh;
        case 22: return cuda_float_degrees; case 23: return cuda_float_radians; case 24: return cuda_float_rint;
        case 25: return cuda_float_fix;    case 26: return cuda_float_trunc;  case 27: return cuda_float_floor;
        case 28: return cuda_float_ceil;   case 29: return cuda_float_sinc;   case 30: return cuda_float_negate;
        case 31: return cuda_float_positive; case 32: return cuda_float_sign; case 33: return cuda_float_reciprocal;
        default: return NULL;
    }
}
#endif
long ref_unary(int op, const float *in, long n, float *out, float p0, float p1, double *seconds) {
    int shape[1] = {(int) n};
    NDArray *na = wrap(in, shape, 1), *r = NULL;
    double t0 = now_s();
#ifdef REF_ENTRY_GPU
    if (op == 35) r = NDArrayMathGPU_ElementWise2F(na, cuda_float_clip, p0, p1);
    else if (op == 36) r = NDArrayMathGPU_ElementWise1F(na, cuda_float_round, p0);
    else if (op == 37) r = NDArray_Multiply_Float(na, na);
    else { gpu_fn f = gpu_table(op); if (f) r = NDArrayMathGPU_ElementWise(na, f); }
    if (0)
#endif
    if (op == 35) r = NDArray_Map2F(na, float_clip, p0, p1);        /* clip(min,max) */
    else if (op == 36) r = NDArray_Map1F(na, float_round, p0);      /* round(decimals) */
    else if (op == 37) r = NDArray_Multiply_Float(na, na);          /* square: numpower.c:3093 */
    else { un_fn f = un_table(op); if (f) r = NDArray_Map(na, f); }
    double t1 = now_s();
    if (seconds) *seconds = t1 - t0;
    long m = take(r, out, n);
    unwrap(na);
    return m;
}

/* ---- full reductions ------------------------------------------------------
 * op: 0 sum (arithmetics.c:58) 1 prod (:36) 2 min (ndarray.c:752) 3 max (:939) */
float ref_reduce_full(int op, const float *in, long n, double *seconds) {
    int shape[1] = {(int) n};
    NDArray *na = wrap(in, shape, 1);
    float v = 0.f;
    double t0 = now_s();
    switch (op) {
        case 0: v = NDArray_Sum_Float(na); break;
        case 1: v = NDArray_Float_Prod(na); break;
        case 2: v = NDArray_Min(na); break;
        case 3: v = NDArray_Max(na); break;
        default: break;
    }
    double t1 = now_s();
    if (seconds) *seconds = t1 - t0;
    unwrap(na);
    return v;
}

/* ---- axis reductions: reduce() (ndarray.c:523-578) with Add / Multiply;
 * op 3 = NDArray_MaxAxis (ndarray.c:781-844, 2-D only). */
long ref_reduce_axis(int op, const float *in, const int *shape, int ndim, int axis,
                     float *out, long out_cap, double *seconds) {
    NDArray *na = wrap(in, shape, ndim), *r = NULL;
    int ax = axis;
    double t0 = now_s();
    if (op == 0) r = reduce(na, &ax, NDArray_Add_Float);
    else if (op == 1) r = reduce(na, &ax, NDArray_Multiply_Float);
    else if (op == 3) r = NDArray_MaxAxis(na, axis);
    double t1 = now_s();
    if (seconds) *seconds = t1 - t0;
    long n = take(r, out, out_cap);
    unwrap(na);
    return n;
}

/* ---- argmax / argmin: NDArray_ArgMinMaxCommon (calculation.c:73-194);
 * axis 128 (= NDARRAY_MAX_DIMS) flattens, as PHP_METHOD(NDArray, argmax) does
 * when no axis is given (numpower.c:2573-2595). */
long ref_argminmax(int is_max, const float *in, const int *shape, int ndim, int axis, int keepdims,
                   float *out, long out_cap, double *seconds) {
    NDArray *na = wrap(in, shape, ndim);
    double t0 = now_s();
    NDArray *r = NDArray_ArgMinMaxCommon(na, axis, keepdims != 0, is_max != 0);
    double t1 = now_s();
    if (seconds) *seconds = t1 - t0;
    long n = take(r, out, out_cap);
    unwrap(na);
    return n;
}

/* ---- matmul / dot / inner ------------------------------------------------
 * NDArray_Matmul (linalg.c:216-245) -> NDArray_FMatmul (:44-82) -> cblas_sgemm. */
long ref_matmul(const float *a, const float *b, int M, int K, int N, float *out, double *seconds) {
    int as[2] = {M, K}, bs[2] = {K, N};
    NDArray *na = wrap(a, as, 2), *nb = wrap(b, bs, 2);
    double t0 = now_s();
    NDArray *r = NDArray_Matmul(na, nb);
    double t1 = now_s();
    if (seconds) *seconds = t1 - t0;
    long n = take(r, out, (long) M * N);
    unwrap(na); unwrap(nb);
    return n;
}
/* NDArray_Matmul on N-D operands: the reference rejects ndim > 2 ("Stack of matrices not allowed", linalg.c:240-243);
 * the Level-1 patched host (oracle/build_dropin_n1.sh) serves stacks with one batched launch. */
long ref_matmul_nd(const float *a, const int *ashape, int andim, const float *b, const int *bshape, int bndim,
                   float *out, long out_cap, double *seconds) {
    NDArray *na = wrap(a, ashape, andim), *nb = wrap(b, bshape, bndim);
    double t0 = now_s();
    NDArray *r = NDArray_Matmul(na, nb);
    double t1 = now_s();
    if (seconds) *seconds = t1 - t0;
    long n = take(r, out, out_cap);
    unwrap(na); unwrap(nb);
    return n;
}
/* NDArray_Dot (linalg.c:354-393): general dispatcher (1-D.1-D, 2-D.2-D, N-D.1-D). */
long ref_dot(const float *a, const int *ashape, int andim, const float *b, const int *bshape, int bndim,
             float *out, long out_cap, double *seconds) {
    NDArray *na = wrap(a, ashape, andim), *nb = wrap(b, bshape, bndim);
    double t0 = now_s();
    NDArray *r = NDArray_Dot(na, nb);
    double t1 = now_s();
    if (seconds) *seconds = t1 - t0;
    long n = take(r, out, out_cap);
    unwrap(na); unwrap(nb);
    return n;
}

/* ---- linalg compositions next to the path: NDArray_Outer (linalg.c:724-751: cblas_sger on zeros / cuda_calculate_outer_product)
 * and NDArray_Norm(type 1) = NDArray_L1Norm (linalg.c:423-447: Transpose, Abs, one NDArray_Sum_Float per column, max). */
long ref_outer(const float *a, int m, const float *b, int n, float *out) {
    int as[1] = {m}, bs[1] = {n};
    NDArray *na = wrap(a, as, 1), *nb = wrap(b, bs, 1);
    NDArray *r = NDArray_Outer(na, nb);
    long k = take(r, out, (long) m * n);
    unwrap(na); unwrap(nb);
    return k;
}
long ref_norm1(const float *a, const int *shape, int ndim, float *out) {
    NDArray *na = wrap(a, shape, ndim);
    NDArray *r = NDArray_Norm(na, 1);
    long k = take(r, out, 1);
    unwrap(na);
    return k;
}

/* ---- logic: NDArray_All (logic.c:25-58) / NDArray_AllClose (logic.c:748-771) -------------------------
 * Only meaningful on the shapes the reference's own tests use (n < 8 for all; allclose reads out of bounds for larger
 * inputs, see oracle/port.c): these entries exist to replay tests/logic/00{1,2}-*.phpt against the real object code. */
#include "logic.h"
int ref_all(const float *a, const int *shape, int ndim) {
    NDArray *na = wrap_cpu(a, shape, ndim);
    int r = (int) NDArray_All(na);
    unwrap_cpu(na);
    return r;
}
int ref_allclose(const float *a, const float *b, const int *shape, int ndim, float rtol, float atol) {
    NDArray *na = wrap_cpu(a, shape, ndim), *nb = wrap_cpu(b, shape, ndim);
    int r = NDArray_AllClose(na, nb, rtol, atol);
    unwrap_cpu(na); unwrap_cpu(nb);
    return r;
}
