// Elementwise kernels of the NDArray hot path (sm_100a): broadcast binary arithmetic,
// the fused a*b+c chain, and the 38 math unaries.  All are HBM-bound streaming kernels:
// 128-bit coalesced loads/stores, several independent 16-byte requests in flight per
// thread, grid sized in multiples of the SM count.  No shared memory: every input
// element is consumed exactly once (a broadcast operand is kept in registers instead).
//
// Algorithmic bytes per output element (fp32): binary full = 12 B, a*b+c fused full = 16 B,
// row/col-broadcast operands contribute ~0, unary = 8 B.
//
// Semantics follow the reference's CPU path (file:line in /root/reference):
//   add/sub/mul/div  src/ndmath/arithmetics.c:247-261, 395-418, 526-545, 662-681 (IEEE, bit-exact)
//   mod              :787-806  a - floor(a/b)*b with the multiply-subtract fused (what GCC emits for
//                              the AVX2 body under the reference's -march flags); NB200_MOD_TRUNC = fmodf tail
//   pow              :912-914  powf
//   maximum/minimum  src/ndarray.c:880-882, 923-925  fmaxf / fminf
//   arctan2          src/ndmath/double_math.c:259-261
//   unaries          src/ndmath/double_math.c (line per op in include/nb200.h)
#include "common.cuh"
#include <math_constants.h>
#include <cstdlib>

namespace nb200 {

// ------------------------------------------------------------------ functors
template <int OP>
struct BinOp {
    __device__ __forceinline__ float operator()(float a, float b, float) const {
        if constexpr (OP == NB200_ADD) return __fadd_rn(a, b);
        else if constexpr (OP == NB200_SUB) return __fsub_rn(a, b);
        else if constexpr (OP == NB200_MUL) return __fmul_rn(a, b);
        else if constexpr (OP == NB200_DIV) return __fdiv_rn(a, b);
        else if constexpr (OP == NB200_MOD) return __fmaf_rn(-floorf(__fdiv_rn(a, b)), b, a);
        else if constexpr (OP == NB200_POW) return powf(a, b);
        else if constexpr (OP == NB200_MAXIMUM) return fmaxf(a, b);
        else if constexpr (OP == NB200_MINIMUM) return fminf(a, b);
        else if constexpr (OP == NB200_ARCTAN2) return atan2f(a, b);
        else if constexpr (OP == NB200_CMP_EQ) return a == b ? 1.0f : 0.0f;
        else if constexpr (OP == NB200_CMP_NE) return (a < b || a > b) ? 1.0f : 0.0f;  // _CMP_NEQ_OQ: ordered
        else if constexpr (OP == NB200_CMP_GT) return a > b ? 1.0f : 0.0f;
        else if constexpr (OP == NB200_CMP_GE) return a >= b ? 1.0f : 0.0f;
        else if constexpr (OP == NB200_CMP_LT) return a < b ? 1.0f : 0.0f;
        else if constexpr (OP == NB200_CMP_LE) return a <= b ? 1.0f : 0.0f;
        else return fmodf(a, b);
    }
};
// scalar operand by value; LHS=true computes (s op a)
template <int OP, bool LHS>
struct BinScalarOp {
    float s;
    __device__ __forceinline__ float operator()(float a, float, float) const {
        return LHS ? BinOp<OP>()(s, a, 0.f) : BinOp<OP>()(a, s, 0.f);
    }
};
// a*b+c: two roundings, never contracted (fused kernel == two nd:: calls bit for bit)
struct MulAddOp {
    __device__ __forceinline__ float operator()(float a, float b, float c) const {
        return __fadd_rn(__fmul_rn(a, b), c);
    }
};

template <int OP>
struct UnOp {
    float p0, p1;
    int *domain_flag;
    __device__ __forceinline__ void domain(bool bad) const {
        if (bad) *domain_flag = 1;  // benign race: every writer stores 1
    }
    __device__ __forceinline__ float operator()(float x, float, float) const {
        if constexpr (OP == NB200_UN_ABS) return fabsf(x);
        else if constexpr (OP == NB200_UN_SQRT) return sqrtf(x);
        else if constexpr (OP == NB200_UN_EXP) return expf(x);
        else if constexpr (OP == NB200_UN_EXP2) return exp2f(x);
        else if constexpr (OP == NB200_UN_EXPM1) return expm1f(x);
        else if constexpr (OP == NB200_UN_LOG) return logf(x);
        else if constexpr (OP == NB200_UN_LOG2) return log2f(x);
        else if constexpr (OP == NB200_UN_LOG10) return log10f(x);
        else if constexpr (OP == NB200_UN_LOG1P) return log1pf(x);
        else if constexpr (OP == NB200_UN_LOGB) return logbf(x);
        else if constexpr (OP == NB200_UN_SIN) return sinf(x);
        else if constexpr (OP == NB200_UN_COS) return cosf(x);
        else if constexpr (OP == NB200_UN_TAN) return tanf(x);
        else if constexpr (OP == NB200_UN_ARCSIN) return asinf(x);
        else if constexpr (OP == NB200_UN_ARCCOS) {  // reference: exit(1) outside [-1,1]; here NaN + flag
            domain(x < -1.0f || x > 1.0f);
            return acosf(x);
        } else if constexpr (OP == NB200_UN_ARCTAN) return atanf(x);
        else if constexpr (OP == NB200_UN_SINH) return sinhf(x);
        else if constexpr (OP == NB200_UN_COSH) return coshf(x);
        else if constexpr (OP == NB200_UN_TANH) return tanhf(x);
        else if constexpr (OP == NB200_UN_ARCSINH) return asinhf(x);
        else if constexpr (OP == NB200_UN_ARCCOSH) {
            domain(x < 1.0f);
            return acoshf(x);
        } else if constexpr (OP == NB200_UN_ARCTANH) {
            domain(fabsf(x) >= 1.0f);
            return atanhf(x);
        } else if constexpr (OP == NB200_UN_DEGREES)  // double multiply by the reference's 11-digit pi, then round
            return (float)((double)x * (180.0 / 3.1415926535));
        else if constexpr (OP == NB200_UN_RADIANS) return (float)((double)x * (3.1415926535 / 180.0));
        else if constexpr (OP == NB200_UN_RINT) {
            float r = rintf(x);
            int fl = (int)floorf(x);
            if (r - (float)fl == 0.5f && ((int)r % 2 != 0)) r -= 1.0f;  // the reference's (no-op) fix-up, kept verbatim
            return r;
        } else if constexpr (OP == NB200_UN_FIX || OP == NB200_UN_TRUNC) return truncf(x);
        else if constexpr (OP == NB200_UN_FLOOR) return floorf(x);
        else if constexpr (OP == NB200_UN_CEIL) return ceilf(x);
        else if constexpr (OP == NB200_UN_SINC) {
            const float pi = 3.1415927f;
            if (x == 0.0f) x = 1.0e-20f;
            x = __fmul_rn(pi, x);
            return __fdiv_rn(sinf(x), x);
        } else if constexpr (OP == NB200_UN_NEGATIVE) return -x;
        else if constexpr (OP == NB200_UN_POSITIVE) return x < 0 ? -x : x;
        else if constexpr (OP == NB200_UN_SIGN) return (float)((x > 0.0f) - (x < 0.0f));
        else if constexpr (OP == NB200_UN_RECIPROCAL) return __fdiv_rn(1.0f, x);
        else if constexpr (OP == NB200_UN_RSQRT) {
            // Quake inverse sqrt, one Newton step, with the contraction GCC applies to
            // y*(1.5f - (x2*y*y)): t = x2*y ; u = fma(-t, y, 1.5f) ; y*u
            float x2 = __fmul_rn(x, 0.5f);
            unsigned int i = __float_as_uint(x);
            i = 0x5f3759dfu - (i >> 1);
            float y = __uint_as_float(i);
            float t = __fmul_rn(x2, y);
            float u = __fmaf_rn(-t, y, 1.5f);
            return __fmul_rn(y, u);
        } else if constexpr (OP == NB200_UN_CLIP) return fminf(p1, fmaxf(x, p0));
        else if constexpr (OP == NB200_UN_ROUND)  // p1 = powf(10, decimals) computed on the host (glibc, as the reference)
            return __fdiv_rn(roundf(__fmul_rn(x, p1)), p1);
        else return __fmul_rn(x, x);  // square
    }
};

// ------------------------------------------------------------------ kernels
constexpr int EW_THREADS = 256;
constexpr int EW_UNROLL = 4;

template <class T>
__device__ __forceinline__ float4 apply4(const T &f, const float4 &a, const float4 &b, const float4 &c) {
    float4 r;
    r.x = f(a.x, b.x, c.x);
    r.y = f(a.y, b.y, c.y);
    r.z = f(a.z, b.z, c.z);
    r.w = f(a.w, b.w, c.w);
    return r;
}

// Flat contiguous operands, all 16-byte aligned.  n4 float4 groups + (n & 3) tail elements.
// NC = true: non-coherent loads (ld.global.nc) - only legal when `out` aliases no input, which the __restrict__ kernels below
// promise.  The in-place kernels (out == an input: the legacy cuda_float_<op> wrappers, `$a += $b`) use plain coherent loads and no
// __restrict__; each thread still reads its own elements before it writes them, so aliasing the SAME range is well defined.
template <int NIN, class F, bool NC>
__device__ __forceinline__ void flat_vec_body(float *out, const float *a, const float *b, const float *c, int64_t n, F f) {
    const int64_t n4 = n >> 2;
    const float4 *a4 = reinterpret_cast<const float4 *>(a);
    const float4 *b4 = reinterpret_cast<const float4 *>(b);
    const float4 *c4 = reinterpret_cast<const float4 *>(c);
    float4 *o4 = reinterpret_cast<float4 *>(out);
    const int64_t tile = (int64_t)EW_THREADS * EW_UNROLL;
    for (int64_t base = (int64_t)blockIdx.x * tile; base < n4; base += (int64_t)gridDim.x * tile) {
        float4 va[EW_UNROLL], vb[EW_UNROLL], vc[EW_UNROLL];
#pragma unroll
        for (int u = 0; u < EW_UNROLL; u++) {
            int64_t i = base + (int64_t)u * EW_THREADS + threadIdx.x;
            if (i < n4) {
                if (NIN > 0) va[u] = NC ? ld_ew(a4 + i) : a4[i];
                else va[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (NIN > 1) vb[u] = NC ? ld_ew(b4 + i) : b4[i];
                if (NIN > 2) vc[u] = NC ? ld_ew(c4 + i) : c4[i];
            }
        }
#pragma unroll
        for (int u = 0; u < EW_UNROLL; u++) {
            int64_t i = base + (int64_t)u * EW_THREADS + threadIdx.x;
            if (i < n4) st_ew(o4 + i, apply4(f, va[u], NIN > 1 ? vb[u] : va[u], NIN > 2 ? vc[u] : va[u]));
        }
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        int64_t i = (n4 << 2) + threadIdx.x;
        out[i] = f(NIN > 0 ? a[i] : 0.f, NIN > 1 ? b[i] : 0.f, NIN > 2 ? c[i] : 0.f);
    }
}
template <int NIN, class F>
__global__ void __launch_bounds__(EW_THREADS) ew_flat_vec(float *__restrict__ out, const float *__restrict__ a,
                                                          const float *__restrict__ b, const float *__restrict__ c,
                                                          int64_t n, F f) {
    flat_vec_body<NIN, F, true>(out, a, b, c, n, f);
}
template <int NIN, class F>
__global__ void __launch_bounds__(EW_THREADS) ew_flat_vec_inplace(float *out, const float *a, const float *b, const float *c, int64_t n, F f) {
    flat_vec_body<NIN, F, false>(out, a, b, c, n, f);
}

// Flat, no alignment assumption (views such as $a[i] are only 4-byte aligned).
template <int NIN, class F>
__global__ void __launch_bounds__(EW_THREADS) ew_flat_scalar(float *__restrict__ out, const float *__restrict__ a,
                                                             const float *__restrict__ b, const float *__restrict__ c,
                                                             int64_t n, F f) {
    for (int64_t i = (int64_t)blockIdx.x * EW_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * EW_THREADS)
        out[i] = f(NIN > 0 ? a[i] : 0.f, NIN > 1 ? b[i] : 0.f, NIN > 2 ? c[i] : 0.f);
}
template <int NIN, class F>
__global__ void __launch_bounds__(EW_THREADS) ew_flat_scalar_inplace(float *out, const float *a, const float *b, const float *c, int64_t n, F f) {
    for (int64_t i = (int64_t)blockIdx.x * EW_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * EW_THREADS)
        out[i] = f(NIN > 0 ? a[i] : 0.f, NIN > 1 ? b[i] : 0.f, NIN > 2 ? c[i] : 0.f);
}

// 2-D broadcast: out is (R, C) contiguous.  Each operand is described by a row stride
// (elements, 0 = same row for every r) and a column mode (1 = contiguous along C,
// 0 = one value per row).  Thread (tx, ty): tx walks column groups, ty interleaves rows.
// SMASK (compile time) marks the operands that STREAM from HBM (one fresh 128-bit load per
// row); the others are broadcast operands that cost no bandwidth: a row vector (row stride 0)
// is loaded ONCE per thread and kept in registers, a column vector is one broadcast scalar
// load per row.  Only streaming operands occupy the U-deep register pipeline (U = 4 rows, 62 registers,
// 4 resident CTAs per SM).
struct Operand2D {
    const float *p;
    int64_t rs;
    int cm;
};
__host__ __device__ constexpr int popc3(int m) { return (m & 1) + ((m >> 1) & 1) + ((m >> 2) & 1); }

template <int NIN, class F, int VEC, int SMASK, int U>
__global__ void __launch_bounds__(EW_THREADS, U == 8 ? 2 : 4) ew_bcast2d(float *__restrict__ out, Operand2D A, Operand2D B, Operand2D Cc,
                                                         int64_t R, int64_t Ccols, int bx, F f) {
    constexpr int NS = popc3(SMASK);
    const int tx = threadIdx.x % bx, ty = threadIdx.x / bx, by = EW_THREADS / bx;
    const int64_t CG = Ccols / VEC;  // column groups
    const Operand2D ops[3] = {A, B, Cc};
    for (int64_t cg = (int64_t)blockIdx.x * bx + tx; cg < CG; cg += (int64_t)gridDim.x * bx) {
        const int64_t col = cg * VEC;
        float4 hoist[3];
#pragma unroll
        for (int k = 0; k < NIN; k++) {
            hoist[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (!((SMASK >> k) & 1) && ops[k].cm == 1) {   // row vector
                if (VEC == 4) hoist[k] = *reinterpret_cast<const float4 *>(ops[k].p + col);
                else hoist[k].x = ops[k].p[col];
            }
        }
        for (int64_t r0 = (int64_t)blockIdx.y * by * U + ty; r0 < R; r0 += (int64_t)gridDim.y * by * U) {
            float4 v[NS > 0 ? NS : 1][U];
            float sc[3][U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int64_t r = r0 + (int64_t)u * by;
                if (r < R) {
#pragma unroll
                    for (int k = 0; k < NIN; k++) {
                        const Operand2D &o = ops[k];
                        if ((SMASK >> k) & 1) {
                            constexpr int dummy = 0;
                            (void)dummy;
                            const int slot = popc3(SMASK & ((1 << k) - 1));
                            if (VEC == 4) v[slot][u] = ld_ew(reinterpret_cast<const float4 *>(o.p + r * o.rs + col));
                            else v[slot][u].x = ld_ew(o.p + r * o.rs + col);
                        } else if (o.cm == 0) {
                            sc[k][u] = __ldg(o.p + r * o.rs);
                        }
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int64_t r = r0 + (int64_t)u * by;
                if (r < R) {
                    float4 arg[3];
#pragma unroll
                    for (int k = 0; k < 3; k++) {
                        if (k >= NIN) arg[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                        else if ((SMASK >> k) & 1) arg[k] = v[popc3(SMASK & ((1 << k) - 1))][u];
                        else if (ops[k].cm == 1) arg[k] = hoist[k];
                        else arg[k] = make_float4(sc[k][u], sc[k][u], sc[k][u], sc[k][u]);
                    }
                    if (VEC == 4) st_ew(reinterpret_cast<float4 *>(out + r * Ccols + col), apply4(f, arg[0], arg[1], arg[2]));
                    else out[r * Ccols + col] = f(arg[0].x, arg[1].x, arg[2].x);
                }
            }
        }
    }
}

// General N-D strided fallback (<= NB200_MAX_DIMS dims, element strides, 0 = broadcast).
struct NdDesc {
    int ndim;
    int64_t shape[NB200_MAX_DIMS];
    int64_t sa[NB200_MAX_DIMS], sb[NB200_MAX_DIMS], sc[NB200_MAX_DIMS];
};
template <int NIN, class F>
__global__ void __launch_bounds__(EW_THREADS) ew_nd(float *__restrict__ out, const float *__restrict__ a,
                                                    const float *__restrict__ b, const float *__restrict__ c, int64_t n,
                                                    NdDesc d, F f) {
    for (int64_t i = (int64_t)blockIdx.x * EW_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * EW_THREADS) {
        int64_t rem = i, oa = 0, ob = 0, oc = 0;
        for (int k = d.ndim - 1; k >= 0; k--) {
            int64_t q = rem / d.shape[k], idx = rem - q * d.shape[k];
            rem = q;
            oa += idx * d.sa[k];
            if (NIN > 1) ob += idx * d.sb[k];
            if (NIN > 2) oc += idx * d.sc[k];
        }
        out[i] = f(a[oa], NIN > 1 ? b[ob] : 0.f, NIN > 2 ? c[oc] : 0.f);
    }
}

// ------------------------------------------------------------------ host-side planning
static inline int grid_for(int64_t work_items, int64_t per_block) {
    int64_t blocks = (work_items + per_block - 1) / per_block;
    // one tile per CTA (non-persistent): measured faster than a capped grid-stride grid for streaming kernels
    // (profiles/r1_copy_probe.log); the grid-stride loop only matters beyond 2^31-1 tiles
    int64_t cap = 0x7FFFFFFF;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

template <int NIN, class F>
static int launch_flat(float *out, const float *a, const float *b, const float *c, int64_t n, F f) {
    if (n == 0) return NB200_OK;
    cudaStream_t s = ctx().stream;
    bool vec = aligned16(out) && (NIN < 1 || aligned16(a)) && (NIN < 2 || aligned16(b)) && (NIN < 3 || aligned16(c));
    // out == an input (in place): the kernels without __restrict__ / non-coherent loads
    const bool inplace = (NIN > 0 && out == a) || (NIN > 1 && out == b) || (NIN > 2 && out == c);
    if (vec && n >= 4) {
        int grid = grid_for(n >> 2, (int64_t)EW_THREADS * EW_UNROLL);
        if (inplace) ew_flat_vec_inplace<NIN, F><<<grid, EW_THREADS, 0, s>>>(out, a, b, c, n, f);
        else ew_flat_vec<NIN, F><<<grid, EW_THREADS, 0, s>>>(out, a, b, c, n, f);
    } else {
        int grid = grid_for(n, EW_THREADS);
        if (inplace) ew_flat_scalar_inplace<NIN, F><<<grid, EW_THREADS, 0, s>>>(out, a, b, c, n, f);
        else ew_flat_scalar<NIN, F><<<grid, EW_THREADS, 0, s>>>(out, a, b, c, n, f);
    }
    NB_LAUNCH_CHECK();
    return NB200_OK;
}

// Collapse (shape, strides...) by merging adjacent dims that are jointly contiguous for the
// output and consistently strided (or consistently broadcast) for every input.
struct Collapsed {
    int ndim;
    int64_t shape[NB200_MAX_DIMS];
    int64_t st[3][NB200_MAX_DIMS];
};
static Collapsed collapse(int nin, int ndim, const int64_t *shape, const int64_t *const *strides) {
    Collapsed c;
    c.ndim = 0;
    for (int d = 0; d < ndim; d++) {
        if (shape[d] == 1) continue;  // size-1 dims carry no information
        bool merged = false;
        if (c.ndim > 0) {
            int p = c.ndim - 1;
            bool ok = true;
            for (int k = 0; k < nin; k++) {
                // previous dim stride must equal this dim's stride * extent (both may be 0)
                if (c.st[k][p] != strides[k][d] * shape[d]) ok = false;
            }
            if (ok) {
                c.shape[p] *= shape[d];
                for (int k = 0; k < nin; k++) c.st[k][p] = strides[k][d];
                merged = true;
            }
        }
        if (!merged) {
            c.shape[c.ndim] = shape[d];
            for (int k = 0; k < nin; k++) c.st[k][c.ndim] = strides[k][d];
            c.ndim++;
        }
    }
    if (c.ndim == 0) {
        c.ndim = 1;
        c.shape[0] = 1;
        for (int k = 0; k < nin; k++) c.st[k][0] = 1;
    }
    return c;
}

template <int NIN, class F>
static int launch_strided(float *out, const float *const *in, int ndim, const int64_t *shape,
                          const int64_t *const *strides, F f) {
    if (ndim < 0 || ndim > NB200_MAX_DIMS) return set_error(NB200_EINVAL, "ndim %d out of range [0,%d]", ndim, NB200_MAX_DIMS);
    int64_t n = 1;
    for (int d = 0; d < ndim; d++) {
        if (shape[d] < 0) return set_error(NB200_EINVAL, "negative extent");
        n *= shape[d];
    }
    if (n == 0) return NB200_OK;
    Collapsed c = collapse(NIN, ndim, shape, strides);
    cudaStream_t s = ctx().stream;
    const float *a = in[0], *b = NIN > 1 ? in[1] : in[0], *cc = NIN > 2 ? in[2] : in[0];
    // (1) everything contiguous -> flat
    if (c.ndim == 1) {
        bool flat = true;
        for (int k = 0; k < NIN; k++) flat = flat && (c.st[k][0] == 1 || c.shape[0] == 1);
        if (flat) return launch_flat<NIN, F>(out, a, b, cc, n, f);
    }
    // (2) 2-D broadcast patterns: (R, C) with per-operand col stride in {0,1}
    if (c.ndim <= 2) {
        int64_t R = c.ndim == 2 ? c.shape[0] : 1, C = c.ndim == 2 ? c.shape[1] : c.shape[0];
        Operand2D ops[3];
        bool ok = true, vec = aligned16(out) && (C % 4 == 0);
        int smask = 0;
        for (int k = 0; k < NIN; k++) {
            int64_t rs = c.ndim == 2 ? c.st[k][0] : 0, cs = c.ndim == 2 ? c.st[k][1] : c.st[k][0];
            if (cs != 0 && cs != 1) ok = false;
            ops[k].p = in[k];
            ops[k].rs = rs;
            ops[k].cm = (int)cs;
            if (cs == 1) vec = vec && aligned16(in[k]) && (rs % 4 == 0);
            if (cs == 1 && (rs != 0 || R == 1)) smask |= 1 << k;   // streams from HBM
        }
        for (int k = NIN; k < 3; k++) ops[k] = ops[0];
        if (ok) {
            const int V = vec ? 4 : 1;
            static const int u_env = getenv("NB200_BCAST_U") ? atoi(getenv("NB200_BCAST_U")) : 0;
            const int U = u_env == 8 ? 8 : 4;   // measured (scripts/bcast_probe.py): 4 rows x 4 CTAs/SM = 0.0874 ms vs 8 rows x 2 CTAs/SM = 0.0970 ms on config 3b
            int64_t CG = C / V;
            int bx = 1;
            while (bx < EW_THREADS && bx < CG) bx <<= 1;
            int by = EW_THREADS / bx;
            int64_t gx = (CG + bx - 1) / bx, gy = (R + (int64_t)by * U - 1) / ((int64_t)by * U);
            int64_t cap = (int64_t)ctx().num_sms * 16;
            if (gx > cap) gx = cap;
            if (gy > 65535) gy = 65535;
            if (gx * gy > cap * 4 && gy > 1) { gy = (cap * 4) / gx; if (gy < 1) gy = 1; }
            dim3 grid((unsigned)gx, (unsigned)gy);
#define NB_BCAST_LAUNCH(MASK)                                                                                          \
    case MASK:                                                                                                         \
        if (vec && U == 8) ew_bcast2d<NIN, F, 4, MASK, 8><<<grid, EW_THREADS, 0, s>>>(out, ops[0], ops[1], ops[2], R, C, bx, f); \
        else if (vec) ew_bcast2d<NIN, F, 4, MASK, 4><<<grid, EW_THREADS, 0, s>>>(out, ops[0], ops[1], ops[2], R, C, bx, f);   \
        else ew_bcast2d<NIN, F, 1, MASK, 4><<<grid, EW_THREADS, 0, s>>>(out, ops[0], ops[1], ops[2], R, C, bx, f);         \
        break;
            switch (smask) {
                NB_BCAST_LAUNCH(0)
                NB_BCAST_LAUNCH(1)
                NB_BCAST_LAUNCH(2)
                NB_BCAST_LAUNCH(3)
                default:
                    if constexpr (NIN == 3) {
                        switch (smask) {
                            NB_BCAST_LAUNCH(4)
                            NB_BCAST_LAUNCH(5)
                            NB_BCAST_LAUNCH(6)
                            NB_BCAST_LAUNCH(7)
                            default: break;
                        }
                    }
                    break;
            }
#undef NB_BCAST_LAUNCH
            NB_LAUNCH_CHECK();
            return NB200_OK;
        }
    }
    // (3) general N-D
    NdDesc d;
    d.ndim = c.ndim;
    for (int k = 0; k < c.ndim; k++) {
        d.shape[k] = c.shape[k];
        d.sa[k] = c.st[0][k];
        d.sb[k] = NIN > 1 ? c.st[1][k] : 0;
        d.sc[k] = NIN > 2 ? c.st[2][k] : 0;
    }
    int grid = grid_for(n, EW_THREADS);
    ew_nd<NIN, F><<<grid, EW_THREADS, 0, s>>>(out, a, b, cc, n, d, f);
    NB_LAUNCH_CHECK();
    return NB200_OK;
}

template <int OP>
static int binary_dispatch(float *out, const float *a, const float *b, int ndim, const int64_t *shape,
                           const int64_t *sa, const int64_t *sb) {
    const float *in[2] = {a, b};
    const int64_t *st[2] = {sa, sb};
    return launch_strided<2, BinOp<OP>>(out, in, ndim, shape, st, BinOp<OP>());
}

template <int OP>
static int binary_scalar_dispatch(float *out, const float *a, float s, int lhs, int64_t n) {
    if (lhs) return launch_flat<1, BinScalarOp<OP, true>>(out, a, a, a, n, BinScalarOp<OP, true>{s});
    return launch_flat<1, BinScalarOp<OP, false>>(out, a, a, a, n, BinScalarOp<OP, false>{s});
}

template <int OP>
static int unary_dispatch(float *out, const float *in, int64_t n, float p0, float p1) {
    UnOp<OP> f{p0, p1, ctx().domain_flag};
    return launch_flat<1, UnOp<OP>>(out, in, in, in, n, f);
}

struct FillOp {
    float v;
    __device__ __forceinline__ float operator()(float, float, float) const { return v; }
};

}  // namespace nb200

using namespace nb200;

#define NB_BIN_CASES(X) X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(14) X(15)
#define NB_UN_CASES(X)                                                                                      \
    X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(14) X(15) X(16) X(17) X(18) \
    X(19) X(20) X(21) X(22) X(23) X(24) X(25) X(26) X(27) X(28) X(29) X(30) X(31) X(32) X(33) X(34) X(35)  \
    X(36) X(37)

extern "C" int nb200_ew_binary(int op, float *out, const float *a, const float *b, int ndim, const int64_t *out_shape,
                               const int64_t *a_strides, const int64_t *b_strides) {
    NB_READY();
    if (!out || !a || !b || (ndim > 0 && (!out_shape || !a_strides || !b_strides)))
        return set_error(NB200_EINVAL, "nb200_ew_binary: null argument");
    switch (op) {
#define X(i) case i: return binary_dispatch<i>(out, a, b, ndim, out_shape, a_strides, b_strides);
        NB_BIN_CASES(X)
#undef X
        default: return set_error(NB200_EINVAL, "nb200_ew_binary: unknown op %d", op);
    }
}

extern "C" int nb200_ew_binary_scalar(int op, float *out, const float *a, float scalar, int scalar_is_lhs, int64_t n) {
    NB_READY();
    if (!out || !a || n < 0) return set_error(NB200_EINVAL, "nb200_ew_binary_scalar: bad argument");
    switch (op) {
#define X(i) case i: return binary_scalar_dispatch<i>(out, a, scalar, scalar_is_lhs, n);
        NB_BIN_CASES(X)
#undef X
        default: return set_error(NB200_EINVAL, "nb200_ew_binary_scalar: unknown op %d", op);
    }
}

extern "C" int nb200_ew_mul_add(float *out, const float *a, const float *b, const float *c, int ndim,
                                const int64_t *out_shape, const int64_t *a_strides, const int64_t *b_strides,
                                const int64_t *c_strides) {
    NB_READY();
    if (!out || !a || !b || !c) return set_error(NB200_EINVAL, "nb200_ew_mul_add: null argument");
    const float *in[3] = {a, b, c};
    const int64_t *st[3] = {a_strides, b_strides, c_strides};
    return launch_strided<3, MulAddOp>(out, in, ndim, out_shape, st, MulAddOp());
}

extern "C" int nb200_ew_unary(int op, float *out, const float *in, int64_t n, float p0, float p1) {
    NB_READY();
    if (!out || !in || n < 0) return set_error(NB200_EINVAL, "nb200_ew_unary: bad argument");
    if (op == NB200_UN_ROUND) p1 = powf(10.0f, p0);  // factor via the host libm, like double_math.c:255
    switch (op) {
#define X(i) case i: return unary_dispatch<i>(out, in, n, p0, p1);
        NB_UN_CASES(X)
#undef X
        default: return set_error(NB200_EINVAL, "nb200_ew_unary: unknown op %d", op);
    }
}

extern "C" int nb200_fill(float *out, float value, int64_t n) {
    NB_READY();
    if (!out || n < 0) return set_error(NB200_EINVAL, "nb200_fill: bad argument");
    return launch_flat<0, FillOp>(out, out, out, out, n, FillOp{value});
}
