/* TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * CPU restatement ("port") of the reference's src/ndmath hot path in plain C.
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference).  Parity status: PINNED — tests/test_oracle.py checks this
 * port (a) against every phpt golden vector the reference holds for the path
 * (tests/golden/phpt_vectors.json, transcribed from tests/math/*.phpt and
 * tests/linalg/001-ndarray-matmul.phpt) and (b) against the reference's own
 * object code (oracle/_ref/libnumpower_ref.so, built by oracle/build_ref.sh)
 * on seeded random inputs.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * load this library; the product path never does.
 *
 * Third-party arithmetic: NDArray_FMatmul calls cblas_sgemm from OpenBLAS
 * (unpinned by the reference; config.m4:67-87 accepts any cblas; CI uses Ubuntu
 * 20.04 libopenblas-dev ~0.3.8).  Its published contract is the BLAS one:
 * C := alpha*A*B + beta*C evaluated in fp32; summation order unspecified.
 * port_matmul restates that contract with a plain fp32 k-ordered sum, and
 * port_matmul_f64 gives the fp64 truth used to report both sides' error.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <stdlib.h>

/* ---- binary elementwise on equal-length flat arrays ------------------------
 * The reference materialises the broadcast operand first (ndarray.c:1172-1294)
 * and then runs a flat loop; callers of the port broadcast with numpy.
 * `nbody` = number of leading elements handled by the 8-wide AVX2 body in the
 * reference ((n/8)*8 when HAVE_AVX2; pass 0 to model the scalar tail only). */
static inline long body_len(long n) { return n >= 8 ? (n / 8) * 8 : 0; }

void port_binary(int op, const float *a, const float *b, float *out, long n) {
    long nb = body_len(n), i;
    switch (op) {
    case 0: /* NDArray_Add_Float arithmetics.c:247-261 */
        for (i = 0; i < n; i++) out[i] = a[i] + b[i];
        break;
    case 1: /* NDArray_Subtract_Float arithmetics.c:526-545 */
        for (i = 0; i < n; i++) out[i] = a[i] - b[i];
        break;
    case 2: /* NDArray_Multiply_Float arithmetics.c:395-418 + fix_negative_zero :280-284:
               AVX body turns every zero product into -0.0, scalar tail turns -0.0 into +0.0 */
        for (i = 0; i < nb; i++) { float r = a[i] * b[i]; out[i] = (r == 0.0f) ? -0.0f : r; }
        for (; i < n; i++) { float r = a[i] * b[i]; out[i] = (r == 0.0f) ? 0.0f : r; }
        break;
    case 3: /* NDArray_Divide_Float arithmetics.c:662-681 */
        for (i = 0; i < n; i++) out[i] = a[i] / b[i];
        break;
    case 4: /* NDArray_Mod_Float arithmetics.c:787-806: AVX body a - floor(a/b)*b, which
               GCC contracts to one fused multiply-subtract under the reference's
               -march flags (vfnmadd; see oracle/build_ref.sh); scalar tail = fmodf */
        for (i = 0; i < nb; i++) out[i] = fmaf(-floorf(a[i] / b[i]), b[i], a[i]);
        for (; i < n; i++) out[i] = fmodf(a[i], b[i]);
        break;
    case 5: /* NDArray_Pow_Float arithmetics.c:912-914 */
        for (i = 0; i < n; i++) out[i] = powf(a[i], b[i]);
        break;
    case 6: /* NDArray_Maximum ndarray.c:880-882 */
        for (i = 0; i < n; i++) out[i] = fmaxf(a[i], b[i]);
        break;
    case 7: /* NDArray_Minimum ndarray.c:923-925 */
        for (i = 0; i < n; i++) out[i] = fminf(a[i], b[i]);
        break;
    case 8: /* NDArray_Map1ND + float_arctan2 ndarray.c:715-727, double_math.c:259-261 */
        for (i = 0; i < n; i++) out[i] = atan2f(a[i], b[i]);
        break;
    /* comparisons -> 1.0f / 0.0f, all ORDERED predicates (_mm256_cmp_ps ..._OQ / _OS and the scalar tails of
       src/logic.c:121-660): a NaN operand gives 0 for every one of them, including not_equal (_CMP_NEQ_OQ). */
    case 10: for (i = 0; i < n; i++) out[i] = a[i] == b[i] ? 1.0f : 0.0f; break;
    case 11: for (i = 0; i < n; i++) out[i] = (a[i] < b[i] || a[i] > b[i]) ? 1.0f : 0.0f; break;
    case 12: for (i = 0; i < n; i++) out[i] = a[i] > b[i] ? 1.0f : 0.0f; break;
    case 13: for (i = 0; i < n; i++) out[i] = a[i] >= b[i] ? 1.0f : 0.0f; break;
    case 14: for (i = 0; i < n; i++) out[i] = a[i] < b[i] ? 1.0f : 0.0f; break;
    case 15: for (i = 0; i < n; i++) out[i] = a[i] <= b[i] ? 1.0f : 0.0f; break;
    default: break;
    }
}

/* a*b+c as two reference calls (numpower.c:193-229): two roundings, no FMA. */
void port_mul_add(const float *a, const float *b, const float *c, float *out, long n) {
    for (long i = 0; i < n; i++) {
        volatile float m = a[i] * b[i];
        out[i] = m + c[i];
    }
}

/* ---- unary functors: src/ndmath/double_math.c (line per case) --------------- */
static float q_rsqrt(float val) { /* double_math.c:111-126 (fast inverse sqrt, one Newton step).
    The reference's expression y*(1.5f - (x2*y*y)) is FMA-contracted by GCC under its
    -march flags; this port keeps the same source expression and is compiled with the
    same flags (see __graft_entry__.build) so the contraction matches. */
    const float threehalfs = 1.5F;
    float x2 = val * 0.5F, y = val;
    uint32_t i; memcpy(&i, &y, 4);
    i = 0x5f3759df - (i >> 1);
    memcpy(&y, &i, 4);
    y = y * (threehalfs - (x2 * y * y));
    return y;
}
static float f_rint(float val) { /* double_math.c:200-210 */
    float rounded = rintf(val);
    int floorInt = (int) floorf(val);
    if (rounded - (float) floorInt == 0.5f && ((int) rounded % 2 != 0)) rounded -= 1.0f;
    return rounded;
}
static float f_sinc(float val) { /* double_math.c:228-235 */
    float pi = 3.1415927f;
    if (val == 0.0) val = 1.0e-20f;
    val = pi * val;
    return sinf(val) / val;
}

float port_unary_scalar(int op, float x, float p0, float p1) {
    switch (op) {
    case 0: return fabsf(x);              /* :10 */
    case 1: return sqrtf(x);              /* :19 */
    case 2: return expf(x);               /* :28 */
    case 3: return exp2f(x);              /* :37 */
    case 4: return expm1f(x);             /* :46 */
    case 5: return logf(x);               /* :55 */
    case 6: return log2f(x);              /* :91 */
    case 7: return log10f(x);             /* :64 */
    case 8: return log1pf(x);             /* :73 */
    case 9: return logbf(x);              /* :82 */
    case 10: return sinf(x);              /* :99 */
    case 11: return cosf(x);              /* :107 */
    case 12: return tanf(x);              /* :132 */
    case 13: return asinf(x);             /* :140 */
    case 14: return acosf(x);             /* :144 (exit(1) outside [-1,1]; inputs kept in-domain) */
    case 15: return atanf(x);             /* :152 */
    case 16: return sinhf(x);             /* :164 */
    case 17: return coshf(x);             /* :168 */
    case 18: return tanhf(x);             /* :172 */
    case 19: return asinhf(x);            /* :176 */
    case 20: return acoshf(x);            /* :180 */
    case 21: return atanhf(x);            /* :188 */
    case 22: return (float) (x * (180.0 / 3.1415926535));   /* :156 */
    case 23: return (float) (x * (3.1415926535 / 180.0));   /* :160 */
    case 24: return f_rint(x);            /* :200 */
    case 25: return truncf(x);            /* :212 fix */
    case 26: return truncf(x);            /* :224 trunc */
    case 27: return floorf(x);            /* :216 */
    case 28: return ceilf(x);             /* :220 */
    case 29: return f_sinc(x);            /* :228 */
    case 30: return -x;                   /* :237 */
    case 31: return x < 0 ? -x : x;       /* :241 positive == abs */
    case 32: return (float) ((x > 0.0f) - (x < 0.0f));      /* :246 */
    case 33: return 1 / x;                /* :263 */
    case 34: return q_rsqrt(x);           /* :111 */
    case 35: return fminf(p1, fmaxf(x, p0));                /* :250 clip(min=p0,max=p1) */
    case 36: { float f = powf(10, p0); return roundf(x * f) / f; } /* :254 round(decimals=p0) */
    case 37: return x * x;                /* numpower.c:3093 square = Multiply(a,a) */
    default: return NAN;
    }
}
/* NDArray_Map / Map1F / Map2F drivers, ndarray.c:682-744 */
void port_unary(int op, const float *in, float *out, long n, float p0, float p1) {
    for (long i = 0; i < n; i++) out[i] = port_unary_scalar(op, in[i], p0, p1);
}

/* ---- full reductions -------------------------------------------------------- */
float port_reduce_full(int op, const float *a, long n) {
    long i;
    switch (op) {
    case 0: { float v = 0; for (i = 0; i < n; i++) v += a[i]; return v; }   /* NDArray_Sum_Float arithmetics.c:58-71 */
    case 1: { float v = 1; for (i = 0; i < n; i++) v *= a[i]; return v; }   /* NDArray_Float_Prod arithmetics.c:36-49 */
    case 2: { float m = a[0]; for (i = 1; i < n; i++) if (a[i] < m) m = a[i]; return m; } /* NDArray_Min ndarray.c:764-769 */
    case 3: { float m = a[0]; for (i = 1; i < n; i++) if (a[i] > m) m = a[i]; return m; } /* NDArray_Max ndarray.c:951-956 */
    default: return NAN;
    }
}

/* ---- axis reductions: reduce()/_reduce()/apply_reduce(), ndarray.c:358-429,523-578.
 * Input viewed as (outer, len, inner); out[o,i] = ((x0 op x1) op x2) ... strictly
 * sequential along the axis starting from the first slice.  op 0 add, 1 mul,
 * 2 min, 3 max (2/3: NDArray_MaxAxis semantics ndarray.c:781-844: start from the
 * first element, replace on strict > / <). */
void port_reduce_axis(int op, const float *in, float *out, long outer, long len, long inner) {
    for (long o = 0; o < outer; o++)
        for (long i = 0; i < inner; i++) {
            const float *p = in + o * len * inner + i;
            float v = p[0];
            for (long k = 1; k < len; k++) {
                float x = p[k * inner];
                switch (op) {
                case 0: v = v + x; break;
                case 1: v = v * x; break;
                case 2: if (x < v) v = x; break;
                case 3: if (x > v) v = x; break;
                }
            }
            out[o * inner + i] = v;
        }
}

/* ---- argmax / argmin: float_argmax / float_argmin, calculation.c:9-59, driven by
 * NDArray_ArgMinMaxCommon :73-194 (axis moved last, then rows of length m). Here the
 * input is viewed as (outer, m, inner) so no transpose is needed; result is the
 * index as float32 (calculation.c:25, :52). */
void port_argminmax(int is_max, const float *in, float *out, long outer, long m, long inner) {
    for (long o = 0; o < outer; o++)
        for (long j = 0; j < inner; j++) {
            const float *p = in + o * m * inner + j;
            float mp = p[0], idx = 0;
            if (!isnan(mp)) {
                for (long i = 1; i < m; i++) {
                    float x = p[i * inner];
                    int take = is_max ? (x > mp) : !(mp <= x);
                    if (take) {
                        mp = x; idx = (float) (int) i;
                        if (isnan(mp)) break;
                    }
                }
            }
            out[o * inner + j] = idx;
        }
}

/* ---- matmul: NDArray_FMatmul linalg.c:75-79 -> cblas_sgemm(RowMajor,N,N,M,N,K,1,A,K,B,N,0,C,N) */
void port_matmul(const float *A, const float *B, float *C, long M, long K, long N) {
    #pragma omp parallel for schedule(static)
    for (long i = 0; i < M; i++) {
        float *c = C + i * N;
        for (long j = 0; j < N; j++) c[j] = 0.f;
        for (long k = 0; k < K; k++) {
            float a = A[i * K + k];
            const float *b = B + k * N;
            for (long j = 0; j < N; j++) c[j] += a * b[j];
        }
    }
}
void port_matmul_f64(const float *A, const float *B, double *C, long M, long K, long N) {
    #pragma omp parallel for schedule(static)
    for (long i = 0; i < M; i++) {
        double *c = C + i * N;
        for (long j = 0; j < N; j++) c[j] = 0.0;
        for (long k = 0; k < K; k++) {
            double a = A[i * K + k];
            const float *b = B + k * N;
            for (long j = 0; j < N; j++) c[j] += a * (double) b[j];
        }
    }
}
/* nd::all — NDArray_All, src/logic.c:25-58.  Intended semantics (the scalar loop, :43-47 / :51-56): 0 as soon as an element equals
 * 0.0, else 1; NaN is non-zero.  The AVX2 body (:29-40) compares an 8-lane movemask with 0x0F and therefore returns 0 for every
 * array of >= 8 elements: a bug the port does not restate.  Pinned by tests/logic/001-ndarray-all.phpt (n < 8: scalar loop). */
int port_all(const float *a, long n) {
    for (long i = 0; i < n; i++)
        if (a[i] == 0.0f) return 0;
    return 1;
}
/* nd::allclose — float_allclose, src/logic.c:718-738: false as soon as |a - b| > atol + rtol * |b| (a NaN difference compares
 * false and passes).  The reference's loop reads element 4i + i * strides[0] / 4 of both arrays (:727-728), i.e. runs out of
 * bounds for every i > 0; the port applies the predicate to element i.  Pinned by tests/logic/002-ndarray-allclose.phpt. */
int port_allclose(const float *a, const float *b, long n, float rtol, float atol) {
    for (long i = 0; i < n; i++) {
        float diff = fabsf(a[i] - b[i]);
        float tolerance = atol + rtol * fabsf(b[i]);
        if (diff > tolerance) return 0;
    }
    return 1;
}
/* N-D . 1-D: NDArray_Dot linalg.c:378-386 -> cblas_sgemv(RowMajor, NoTrans, rows, cols, 1, A, cols, x, 1, 0, y, 1) */
void port_gemv(const float *A, const float *x, float *y, long rows, long cols) {
    for (long i = 0; i < rows; i++) {
        float s = 0.f;
        for (long k = 0; k < cols; k++) s += A[i * cols + k] * x[k];
        y[i] = s;
    }
}
