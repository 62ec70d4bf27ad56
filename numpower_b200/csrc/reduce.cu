// Reductions of the NDArray hot path (sm_100a): full sum/prod/min/max, axis reductions,
// argmax/argmin.  All HBM-bound: 128-bit coalesced loads with several requests in flight
// per thread, warp-shuffle + shared-memory block combine, and a deterministic two-stage
// finish (per-block partials + "last block done" ticket) — no float atomics, so results
// are bit-reproducible run to run.
//
// Algorithmic bytes: 4 B read per input element (+4 B per output element).
//
// Reference semantics (file:line in /root/reference):
//   sum / prod     src/ndmath/arithmetics.c:58-71, 36-49   (sequential fp32; see SURVEY F1 —
//                  parity inputs are chosen so that every summation order is exact)
//   min / max      src/ndarray.c:752-772, 939-959          (NaN at index 0 sticks, later NaNs skipped)
//   axis reduce    src/ndarray.c:394-429, 523-578          (sequential along the axis from slice 0);
//                  NB200_ORDER_SEQUENTIAL reproduces that order bit for bit
//   max(axis)      src/ndarray.c:781-844
//   argmax/argmin  src/ndmath/calculation.c:9-59           (first occurrence; argmax skips NaN
//                  unless element 0 is NaN; argmin returns the first NaN; index stored as float)
#include "common.cuh"
#include <math_constants.h>

namespace nb200 {

constexpr int RED_THREADS = 512;
constexpr int RED_UNROLL = 4;
constexpr int MAX_SPLIT = 2048;

// ------------------------------------------------------------------ combine operators
template <int OP>
struct Red {
    __device__ __forceinline__ static float identity() {
        if constexpr (OP == NB200_SUM) return 0.f;
        else if constexpr (OP == NB200_PROD) return 1.f;
        else if constexpr (OP == NB200_MIN) return CUDART_INF_F;
        else return -CUDART_INF_F;
    }
    // parallel combine: min/max skip NaN (fminf/fmaxf); the index-0 rule is applied at the end
    __device__ __forceinline__ static float comb(float a, float b) {
        if constexpr (OP == NB200_SUM) return __fadd_rn(a, b);
        else if constexpr (OP == NB200_PROD) return __fmul_rn(a, b);
        else if constexpr (OP == NB200_MIN) return fminf(a, b);
        else return fmaxf(a, b);
    }
    // sequential step of the reference loops: `if (x < m) m = x` keeps a leading NaN
    __device__ __forceinline__ static float seq(float acc, float x) {
        if constexpr (OP == NB200_SUM) return __fadd_rn(acc, x);
        else if constexpr (OP == NB200_PROD) return __fmul_rn(acc, x);
        else if constexpr (OP == NB200_MIN) return (x < acc) ? x : acc;
        else return (x > acc) ? x : acc;
    }
};

template <int OP>
__device__ __forceinline__ float warp_reduce(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = Red<OP>::comb(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
template <int OP, int THREADS>
__device__ __forceinline__ float block_reduce(float v, float *smem) {
    v = warp_reduce<OP>(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) smem[warp] = v;
    __syncthreads();
    if (warp == 0) {
        v = (lane < THREADS / 32) ? smem[lane] : Red<OP>::identity();
        v = warp_reduce<OP>(v);
    }
    return v;  // valid in warp 0
}

// Accumulate a contiguous run [p, p+len) into per-thread accumulators, block-cooperatively.
// `tid`/`nthr` enumerate the cooperating threads.  Handles a 4-byte-aligned start by peeling.
template <int OP>
__device__ __forceinline__ float run_accumulate(const float *__restrict__ p, int64_t len, int64_t tid, int64_t nthr) {
    float4 acc = make_float4(Red<OP>::identity(), Red<OP>::identity(), Red<OP>::identity(), Red<OP>::identity());
    int64_t head = ((16 - (reinterpret_cast<uintptr_t>(p) & 15)) & 15) >> 2;
    if (head > len) head = len;
    if (tid < head) acc.x = Red<OP>::comb(acc.x, p[tid]);
    const float4 *p4 = reinterpret_cast<const float4 *>(p + head);
    const int64_t n4 = (len - head) >> 2;
    int64_t i = tid;
    for (; i + (RED_UNROLL - 1) * nthr < n4; i += RED_UNROLL * nthr) {
        float4 v[RED_UNROLL];
#pragma unroll
        for (int u = 0; u < RED_UNROLL; u++) v[u] = ldg_stream(p4 + i + u * nthr);
#pragma unroll
        for (int u = 0; u < RED_UNROLL; u++) {
            acc.x = Red<OP>::comb(acc.x, v[u].x);
            acc.y = Red<OP>::comb(acc.y, v[u].y);
            acc.z = Red<OP>::comb(acc.z, v[u].z);
            acc.w = Red<OP>::comb(acc.w, v[u].w);
        }
    }
    for (; i < n4; i += nthr) {
        float4 v = ldg_stream(p4 + i);
        acc.x = Red<OP>::comb(acc.x, v.x);
        acc.y = Red<OP>::comb(acc.y, v.y);
        acc.z = Red<OP>::comb(acc.z, v.z);
        acc.w = Red<OP>::comb(acc.w, v.w);
    }
    const int64_t tail0 = head + (n4 << 2);
    if (tid < len - tail0) acc.y = Red<OP>::comb(acc.y, p[tail0 + tid]);
    return Red<OP>::comb(Red<OP>::comb(acc.x, acc.y), Red<OP>::comb(acc.z, acc.w));
}

// ------------------------------------------------------------------ row reduction (inner == 1)
// rows x len, contiguous.  grid = (S, rows): block (s, r) reduces segment s of row r.  S == 1
// writes out[r] directly; S > 1 writes partials[r*S + s] and the last block of the row
// (ticket) folds the S partials in fixed order.  Full reductions are rows == 1.
template <int OP>
__global__ void __launch_bounds__(RED_THREADS) reduce_rows_kernel(float *__restrict__ out, const float *__restrict__ in,
                                                                  int64_t len, int S, float *__restrict__ partials,
                                                                  unsigned int *__restrict__ ticket) {
    __shared__ float smem[RED_THREADS / 32];
    __shared__ bool is_last;
    const int64_t r = blockIdx.y;
    const int s = blockIdx.x;
    const float *row = in + r * len;
    const int64_t seg = (((len + S - 1) / S) + 3) & ~int64_t(3);  // segment length, multiple of 4
    const int64_t l0 = (int64_t)s * seg;
    const int64_t l1 = (l0 + seg < len) ? l0 + seg : len;
    float v = Red<OP>::identity();
    if (l0 < len) v = run_accumulate<OP>(row + l0, l1 - l0, threadIdx.x, RED_THREADS);
    v = block_reduce<OP, RED_THREADS>(v, smem);
    if (S == 1) {
        if (threadIdx.x == 0) {
            if ((OP == NB200_MIN || OP == NB200_MAX) && isnan(row[0])) v = row[0];
            out[r] = v;
        }
        return;
    }
    if (threadIdx.x == 0) {
        partials[r * S + s] = v;
        __threadfence();
        unsigned int t = atomicAdd(&ticket[r], 1u);
        is_last = (t == (unsigned)S - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    float a = Red<OP>::identity();
    for (int i = threadIdx.x; i < S; i += RED_THREADS) a = Red<OP>::comb(a, __ldcg(&partials[r * S + i]));
    a = block_reduce<OP, RED_THREADS>(a, smem);
    if (threadIdx.x == 0) {
        if ((OP == NB200_MIN || OP == NB200_MAX) && isnan(row[0])) a = row[0];
        out[r] = a;
        ticket[r] = 0;  // re-arm for the next launch
    }
}

// Short rows: one warp per row (len <= 1024), 8 rows per 256-thread block.
template <int OP>
__global__ void __launch_bounds__(256) reduce_rows_warp_kernel(float *__restrict__ out, const float *__restrict__ in,
                                                               int64_t rows, int64_t len) {
    const int lane = threadIdx.x & 31;
    for (int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); r < rows; r += (int64_t)gridDim.x * 8) {
        const float *row = in + r * len;
        float v = run_accumulate<OP>(row, len, lane, 32);
        v = warp_reduce<OP>(v);
        if (lane == 0) {
            if ((OP == NB200_MIN || OP == NB200_MAX) && isnan(row[0])) v = row[0];
            out[r] = v;
        }
    }
}

// Strictly sequential row reduction (reference order): one thread per row, prefetching ahead.
template <int OP>
__global__ void __launch_bounds__(128) reduce_rows_seq_kernel(float *__restrict__ out, const float *__restrict__ in,
                                                              int64_t rows, int64_t len) {
    for (int64_t r = (int64_t)blockIdx.x * 128 + threadIdx.x; r < rows; r += (int64_t)gridDim.x * 128) {
        const float *row = in + r * len;
        float acc = row[0];
        int64_t k = 1;
        for (; k + 8 <= len; k += 8) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; u++) v[u] = row[k + u];
#pragma unroll
            for (int u = 0; u < 8; u++) acc = Red<OP>::seq(acc, v[u]);
        }
        for (; k < len; k++) acc = Red<OP>::seq(acc, row[k]);
        out[r] = acc;
    }
}

// ------------------------------------------------------------------ column reduction (inner > 1)
// in viewed as (outer, len, inner); out[o, i].  Block = BX x BY threads: tx walks `inner`
// (coalesced, VEC floats per thread), ty splits the axis inside the block; blockIdx.y
// enumerates (o, s) with s one of S segments of the axis.  Each thread accumulates its
// rows in increasing order; ty partials are folded in fixed order through shared memory.
// BY == 1 and S == 1 is exactly the reference's sequential order (SEQ = true uses Red::seq).
constexpr int COL_BX = 32;
template <int OP, int VEC, int BY, bool SEQ>
__global__ void __launch_bounds__(COL_BX *BY) reduce_cols_kernel(float *__restrict__ out, const float *__restrict__ in,
                                                                 int64_t len, int64_t inner, int S, int64_t seg,
                                                                 float *__restrict__ final_out, unsigned int *__restrict__ ticket) {
    __shared__ float smem[BY][COL_BX * VEC + 1];
    __shared__ bool is_last;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int64_t o = blockIdx.y / S;
    const int s = (int)(blockIdx.y % S);
    const int64_t col = ((int64_t)blockIdx.x * COL_BX + tx) * VEC;
    // The S segments are INTERLEAVED along the axis in chunks of BY*CU rows: block (x, s) takes chunks s, s+S, s+2S, ...
    // so at any time the whole grid sweeps one contiguous band of rows (DRAM-page friendly) instead of S far-apart bands.
    constexpr int CU = 8;                  // rows in flight per thread
    constexpr int CHUNK = BY * CU;
    (void)seg;
    float acc[VEC];
    bool started = false;
#pragma unroll
    for (int v = 0; v < VEC; v++) acc[v] = Red<OP>::identity();
    if (col < inner) {
        const float *base = in + o * len * inner + col;
        const int64_t nchunks = (len + CHUNK - 1) / CHUNK;
        for (int64_t ch = s; ch < nchunks; ch += S) {
            const int64_t l = ch * CHUNK + ty;
            if (l + (int64_t)(CU - 1) * BY < len) {
                float x[CU][VEC];
#pragma unroll
                for (int u = 0; u < CU; u++) {
                    const float *p = base + (l + (int64_t)u * BY) * inner;
                    if (VEC == 4) {
                        float4 t = ldg_stream(reinterpret_cast<const float4 *>(p));
                        x[u][0] = t.x; x[u][1 % VEC] = t.y; x[u][2 % VEC] = t.z; x[u][3 % VEC] = t.w;
                    } else {
                        x[u][0] = ldg_stream(p);
                    }
                }
#pragma unroll
                for (int u = 0; u < CU; u++)
#pragma unroll
                    for (int v = 0; v < VEC; v++) {
                        if (SEQ) acc[v] = started ? Red<OP>::seq(acc[v], x[u][v]) : x[u][v];
                        else acc[v] = Red<OP>::comb(acc[v], x[u][v]);
                        if (SEQ && v == VEC - 1) started = true;
                    }
            } else {
                for (int64_t ll = l; ll < len; ll += BY) {
                    const float *p = base + ll * inner;
#pragma unroll
                    for (int v = 0; v < VEC; v++) {
                        float xv = p[v];
                        if (SEQ) acc[v] = started ? Red<OP>::seq(acc[v], xv) : xv;
                        else acc[v] = Red<OP>::comb(acc[v], xv);
                    }
                    if (SEQ) started = true;
                }
            }
        }
    }
    if (BY > 1) {
#pragma unroll
        for (int v = 0; v < VEC; v++) smem[ty][tx * VEC + v] = acc[v];
        __syncthreads();
        if (ty == 0) {
#pragma unroll
            for (int v = 0; v < VEC; v++) {
                float a = smem[0][tx * VEC + v];
                for (int y = 1; y < BY; y++) a = Red<OP>::comb(a, smem[y][tx * VEC + v]);
                acc[v] = a;
            }
        }
    }
    if (ty == 0 && col < inner) {
        float *dst = out + (o * S + s) * inner + col;  // S == 1: the result; S > 1: partials (outer, S, inner)
#pragma unroll
        for (int v = 0; v < VEC; v++) {
            float a = acc[v];
            // index-0 NaN rule of NDArray_Min/Max: element 0 lives in segment 0; with S > 1 the
            // NaN is planted in partial 0 and picked up again by the fold pass (len == S there).
            if (!SEQ && s == 0 && (OP == NB200_MIN || OP == NB200_MAX)) {
                float first = in[o * len * inner + col + v];
                if (isnan(first)) a = first;
            }
            dst[v] = a;
        }
    }
    // S > 1 with a ticket: the last block of this (o, column tile) folds the S partials in fixed order,
    // so the split reduction is still ONE launch and deterministic.
    if (S > 1 && ticket != nullptr) {
        __threadfence();
        __syncthreads();
        if (tx == 0 && ty == 0) {
            unsigned int t = atomicAdd(&ticket[o * gridDim.x + blockIdx.x], 1u);
            is_last = (t == (unsigned)S - 1);
        }
        __syncthreads();
        if (!is_last) return;
        __threadfence();
        float a[VEC];
#pragma unroll
        for (int v = 0; v < VEC; v++) a[v] = Red<OP>::identity();
        if (col < inner) {
            for (int ss = ty; ss < S; ss += BY) {
                const float *src = out + (o * S + ss) * inner + col;
#pragma unroll
                for (int v = 0; v < VEC; v++) a[v] = Red<OP>::comb(a[v], __ldcg(src + v));
            }
        }
        __syncthreads();
#pragma unroll
        for (int v = 0; v < VEC; v++) smem[ty][tx * VEC + v] = a[v];
        __syncthreads();
        if (ty == 0 && col < inner) {
#pragma unroll
            for (int v = 0; v < VEC; v++) {
                float r = smem[0][tx * VEC + v];
                for (int y = 1; y < BY; y++) r = Red<OP>::comb(r, smem[y][tx * VEC + v]);
                if (OP == NB200_MIN || OP == NB200_MAX) {   // index-0 NaN planted in partial 0 (see above)
                    float first = __ldcg(out + (o * S) * inner + col + v);
                    if (isnan(first)) r = first;
                }
                final_out[o * inner + col + v] = r;
            }
        }
        if (tx == 0 && ty == 0) ticket[o * gridDim.x + blockIdx.x] = 0;
    }
}

// ------------------------------------------------------------------ argmax / argmin
// A candidate is packed into 64 bits so that "better" == larger integer:
//   high 32: order-preserving key of the value (argmin: inverted; NaN lowest for argmax,
//            highest for argmin), -0.0 canonicalised to +0.0 so that ties are ties;
//   low 32:  0xFFFFFFFF - index  => among equal keys the LOWEST index wins (first occurrence).
__device__ __forceinline__ unsigned int order_key(float x) {
    x = x + 0.0f;  // -0.0 -> +0.0
    unsigned int b = __float_as_uint(x);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
template <bool IS_MAX>
__device__ __forceinline__ unsigned long long pack(float v, unsigned int idx) {
    unsigned int k;
    if (isnan(v)) k = IS_MAX ? 0u : 0xFFFFFFFFu;          // argmax never prefers NaN; argmin always does
    else k = IS_MAX ? order_key(v) : ~order_key(v);
    if (!IS_MAX && !isnan(v) && k == 0xFFFFFFFFu) k = 0xFFFFFFFEu;  // keep NaN strictly best (cannot occur for finite keys)
    return ((unsigned long long)k << 32) | (unsigned long long)(0xFFFFFFFFu - idx);
}
__device__ __forceinline__ unsigned long long umax64(unsigned long long a, unsigned long long b) { return a > b ? a : b; }
__device__ __forceinline__ unsigned long long warp_max64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = umax64(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
template <int THREADS>
__device__ __forceinline__ unsigned long long block_max64(unsigned long long v, unsigned long long *smem) {
    v = warp_max64(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) smem[warp] = v;
    __syncthreads();
    if (warp == 0) {
        v = (lane < THREADS / 32) ? smem[lane] : 0ull;
        v = warp_max64(v);
    }
    return v;
}
// reference per-element rules, applied in increasing index order inside one thread
// A thread starts from (identity, ARG_NONE); the first element equal to the identity still
// claims the slot, so an all -inf (+inf) run reports its first index.
constexpr unsigned int ARG_NONE = 0xFFFFFFFFu;
template <bool IS_MAX>
__device__ __forceinline__ void arg_step(float &bv, unsigned int &bi, float x, unsigned int i) {
    if (IS_MAX) {
        if (x > bv || (bi == ARG_NONE && x == bv)) { bv = x; bi = i; }   // calculation.c:23 (NaN never taken)
    } else {
        // calculation.c:50 `!(mp <= x)` takes smaller values AND the first NaN; :53 breaks after a NaN
        if (!isnan(bv) && (!(bv <= x) || (bi == ARG_NONE && x == bv))) { bv = x; bi = i; }
    }
}

// per-thread scan of a contiguous run; indices are relative to `p` plus idx0
template <bool IS_MAX>
__device__ __forceinline__ unsigned long long arg_run(const float *__restrict__ p, int64_t len, unsigned int idx0,
                                                      int64_t tid, int64_t nthr) {
    unsigned long long best = 0ull;
    int64_t head = ((16 - (reinterpret_cast<uintptr_t>(p) & 15)) & 15) >> 2;
    if (head > len) head = len;
    if (tid < head) best = umax64(best, pack<IS_MAX>(p[tid], idx0 + (unsigned)tid));
    const float4 *p4 = reinterpret_cast<const float4 *>(p + head);
    const int64_t n4 = (len - head) >> 2;
    float bv = IS_MAX ? -CUDART_INF_F : CUDART_INF_F;
    unsigned int bi = ARG_NONE;
    int64_t i = tid;
    for (; i + (RED_UNROLL - 1) * nthr < n4; i += RED_UNROLL * nthr) {
        float4 v[RED_UNROLL];
#pragma unroll
        for (int u = 0; u < RED_UNROLL; u++) v[u] = ldg_stream(p4 + i + u * nthr);
#pragma unroll
        for (int u = 0; u < RED_UNROLL; u++) {
            unsigned int b = idx0 + (unsigned)head + (unsigned)((i + u * nthr) << 2);
            arg_step<IS_MAX>(bv, bi, v[u].x, b);
            arg_step<IS_MAX>(bv, bi, v[u].y, b + 1);
            arg_step<IS_MAX>(bv, bi, v[u].z, b + 2);
            arg_step<IS_MAX>(bv, bi, v[u].w, b + 3);
        }
    }
    for (; i < n4; i += nthr) {
        float4 v = ldg_stream(p4 + i);
        unsigned int b = idx0 + (unsigned)head + (unsigned)(i << 2);
        arg_step<IS_MAX>(bv, bi, v.x, b);
        arg_step<IS_MAX>(bv, bi, v.y, b + 1);
        arg_step<IS_MAX>(bv, bi, v.z, b + 2);
        arg_step<IS_MAX>(bv, bi, v.w, b + 3);
    }
    if (bi != ARG_NONE) best = umax64(best, pack<IS_MAX>(bv, bi));
    const int64_t tail0 = head + (n4 << 2);
    if (tid < len - tail0) best = umax64(best, pack<IS_MAX>(p[tail0 + tid], idx0 + (unsigned)(tail0 + tid)));
    return best;
}

__device__ __forceinline__ float unpack_index(unsigned long long best) {
    unsigned int idx = 0xFFFFFFFFu - (unsigned int)(best & 0xFFFFFFFFull);
    return (float)idx;  // the reference stores (float)i with int i (calculation.c:25): the same rounding below 2^31; above, where the
                        // reference's int shapes cannot go, the 32-bit index stays unsigned (was (float)(int)idx: negative results)
}

// rows x len (inner == 1).  grid = (S, rows); S > 1 uses u64 partials + ticket like reduce_rows_kernel.
template <bool IS_MAX>
__global__ void __launch_bounds__(RED_THREADS) arg_rows_kernel(float *__restrict__ out, const float *__restrict__ in,
                                                               int64_t len, int S, unsigned long long *__restrict__ partials,
                                                               unsigned int *__restrict__ ticket,
                                                               unsigned long long *__restrict__ packed_out = nullptr) {
    __shared__ unsigned long long smem[RED_THREADS / 32];
    __shared__ bool is_last;
    const int64_t r = blockIdx.y;
    const int s = blockIdx.x;
    const float *row = in + r * len;
    const int64_t seg = (((len + S - 1) / S) + 3) & ~int64_t(3);
    const int64_t l0 = (int64_t)s * seg;
    const int64_t l1 = (l0 + seg < len) ? l0 + seg : len;
    unsigned long long v = 0ull;
    if (l0 < len) v = arg_run<IS_MAX>(row + l0, l1 - l0, (unsigned)l0, threadIdx.x, RED_THREADS);
    v = block_max64<RED_THREADS>(v, smem);
    if (S == 1) {
        if (threadIdx.x == 0) {
            out[r] = isnan(row[0]) ? 0.f : unpack_index(v);  // calculation.c:14-17, :41-44
            if (packed_out) packed_out[r] = v;                // sharded callers combine the exact (key, index) words themselves
        }
        return;
    }
    if (threadIdx.x == 0) {
        partials[r * S + s] = v;
        __threadfence();
        unsigned int t = atomicAdd(&ticket[r], 1u);
        is_last = (t == (unsigned)S - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    unsigned long long a = 0ull;
    for (int i = threadIdx.x; i < S; i += RED_THREADS) a = umax64(a, __ldcg(&partials[r * S + i]));
    a = block_max64<RED_THREADS>(a, smem);
    if (threadIdx.x == 0) {
        out[r] = isnan(row[0]) ? 0.f : unpack_index(a);
        if (packed_out) packed_out[r] = a;
        ticket[r] = 0;
    }
}

template <bool IS_MAX>
__global__ void __launch_bounds__(256) arg_rows_warp_kernel(float *__restrict__ out, const float *__restrict__ in,
                                                            int64_t rows, int64_t len) {
    const int lane = threadIdx.x & 31;
    for (int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); r < rows; r += (int64_t)gridDim.x * 8) {
        const float *row = in + r * len;
        unsigned long long v = warp_max64(arg_run<IS_MAX>(row, len, 0u, lane, 32));
        if (lane == 0) out[r] = isnan(row[0]) ? 0.f : unpack_index(v);
    }
}

// (outer, m, inner) with inner > 1: thread per column, ty splits the axis, fixed-order fold.
template <bool IS_MAX, int BY>
__global__ void __launch_bounds__(COL_BX *BY) arg_cols_kernel(float *__restrict__ out, const float *__restrict__ in,
                                                              int64_t m, int64_t inner) {
    __shared__ unsigned long long smem[BY][COL_BX + 1];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int64_t o = blockIdx.y;
    const int64_t col = (int64_t)blockIdx.x * COL_BX + tx;
    unsigned long long best = 0ull;
    if (col < inner) {
        const float *base = in + o * m * inner + col;
        float bv = IS_MAX ? -CUDART_INF_F : CUDART_INF_F;
        unsigned int bi = ARG_NONE;
        int64_t l = ty;
        for (; l + 3 * BY < m; l += 4 * BY) {
            float x[4];
#pragma unroll
            for (int u = 0; u < 4; u++) x[u] = ldg_stream(base + (l + (int64_t)u * BY) * inner);
#pragma unroll
            for (int u = 0; u < 4; u++) {
                unsigned int idx = (unsigned)(l + (int64_t)u * BY);
                arg_step<IS_MAX>(bv, bi, x[u], idx);
            }
        }
        for (; l < m; l += BY) {
            float xv = base[l * inner];
            arg_step<IS_MAX>(bv, bi, xv, (unsigned)l);
        }
        if (bi != ARG_NONE) best = pack<IS_MAX>(bv, bi);
    }
    smem[ty][tx] = best;
    __syncthreads();
    if (ty == 0 && col < inner) {
        for (int y = 1; y < BY; y++) best = umax64(best, smem[y][tx]);
        float first = in[o * m * inner + col];
        out[o * inner + col] = isnan(first) ? 0.f : unpack_index(best);
    }
}

// ------------------------------------------------------------------ host-side planning
static int pick_split(int64_t rows, int64_t len) {
    // enough CTAs to cover every SM ~4x, but keep >= 16K elements per CTA
    int64_t target = (int64_t)ctx().num_sms * 4;
    if (rows >= target) return 1;
    int64_t S = (target + rows - 1) / rows;
    int64_t maxS = len / 16384;
    if (maxS < 1) maxS = 1;
    if (S > maxS) S = maxS;
    if (S > MAX_SPLIT) S = MAX_SPLIT;
    return (int)S;
}

template <int OP>
static int reduce_rows(float *out, const float *in, int64_t rows, int64_t len, int order) {
    cudaStream_t st = ctx().stream;
    if (order == NB200_ORDER_SEQUENTIAL) {
        int64_t grid = (rows + 127) / 128;
        if (grid > (int64_t)ctx().num_sms * 16) grid = (int64_t)ctx().num_sms * 16;
        reduce_rows_seq_kernel<OP><<<(unsigned)grid, 128, 0, st>>>(out, in, rows, len);
        NB_LAUNCH_CHECK();
        return NB200_OK;
    }
    if (len <= 1024 && rows >= 64) {
        int64_t grid = (rows + 7) / 8;
        if (grid > (int64_t)ctx().num_sms * 32) grid = (int64_t)ctx().num_sms * 32;
        reduce_rows_warp_kernel<OP><<<(unsigned)grid, 256, 0, st>>>(out, in, rows, len);
        NB_LAUNCH_CHECK();
        return NB200_OK;
    }
    // grid.y is limited to 65535 rows per launch
    for (int64_t r0 = 0; r0 < rows; r0 += 65535) {
        int64_t nr = rows - r0 < 65535 ? rows - r0 : 65535;
        int S = pick_split(nr, len);
        float *partials = nullptr;
        if (S > 1) {
            if (nr > 4096) S = 1;
            else {
                int rc = ensure_scratch((int64_t)nr * S * sizeof(float));
                if (rc != NB200_OK) return rc;
                partials = static_cast<float *>(ctx().scratch);
            }
        }
        dim3 grid((unsigned)S, (unsigned)nr);
        reduce_rows_kernel<OP><<<grid, RED_THREADS, 0, st>>>(out + r0, in + r0 * len, len, S, partials, ctx().ticket);
        NB_LAUNCH_CHECK();
    }
    return NB200_OK;
}

template <int OP, int VEC, int BY, bool SEQ>
static int launch_cols(float *out, const float *in, int64_t outer, int64_t len, int64_t inner, int S, int64_t seg,
                       float *final_out = nullptr, unsigned int *ticket = nullptr) {
    int64_t gx = (inner + (int64_t)COL_BX * VEC - 1) / ((int64_t)COL_BX * VEC);
    for (int64_t o0 = 0; o0 < outer * S; o0 += 65535) {
        int64_t ny = outer * S - o0 < 65535 ? outer * S - o0 : 65535;
        if (S > 1 && o0 != 0) return set_error(NB200_EINVAL, "reduce_axis: split with outer*S > 65535 unsupported");
        dim3 grid((unsigned)gx, (unsigned)ny), block(COL_BX, BY);
        // with S == 1 blockIdx.y == o - o0/S
        reduce_cols_kernel<OP, VEC, BY, SEQ><<<grid, block, 0, ctx().stream>>>(
            out + (S == 1 ? o0 * inner : 0), in + (S == 1 ? o0 * len * inner : 0), len, inner, S, seg, final_out, ticket);
        NB_LAUNCH_CHECK();
    }
    return NB200_OK;
}

template <int OP>
static int reduce_cols(float *out, const float *in, int64_t outer, int64_t len, int64_t inner, int order) {
    const bool vec = (inner % 4 == 0) && aligned16(in) && aligned16(out);
    if (order == NB200_ORDER_SEQUENTIAL) {
        if (vec && inner >= 4 * COL_BX * 64) return launch_cols<OP, 4, 1, true>(out, in, outer, len, inner, 1, len);
        return launch_cols<OP, 1, 1, true>(out, in, outer, len, inner, 1, len);
    }
    // TREE: BY = 8 threads split the axis inside a block; S segments across blocks if the grid is small
    const int VEC = vec ? 4 : 1;
    int64_t blocks = ((inner + COL_BX * VEC - 1) / (COL_BX * VEC)) * outer;
    // one balanced wave: as many CTAs as are actually co-resident (occupancy query: registers limit the 256-thread
    // CTAs to fewer than 8 per SM; sizing for 8 left a 0.3-wave tail, ncu "Waves Per SM 1.30"), so split the axis
    // into floor(slots / tiles) segments
    static int bps_vec = 0, bps_scalar = 0;
    if (!bps_vec) {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps_vec, reduce_cols_kernel<OP, 4, 8, false>, COL_BX * 8, 0);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps_scalar, reduce_cols_kernel<OP, 1, 8, false>, COL_BX * 8, 0);
        if (bps_vec < 1) bps_vec = 1;
        if (bps_scalar < 1) bps_scalar = 1;
    }
    int64_t target = (int64_t)ctx().num_sms * (vec ? bps_vec : bps_scalar);
    int S = 1;
    if (blocks * 2 <= target && len >= 64) {
        int64_t want = target / blocks, maxS = len / 32;
        S = (int)(want < maxS ? want : maxS);
        if (S < 1) S = 1;
        if (S > 256) S = 256;
        if (outer * S > 65535) S = 1;
    }
    if (S == 1) {
        if (vec) return launch_cols<OP, 4, 8, false>(out, in, outer, len, inner, 1, len);
        return launch_cols<OP, 1, 8, false>(out, in, outer, len, inner, 1, len);
    }
    {   // chunks of 64 rows are dealt round-robin to the S segments: pick the smallest S with the same makespan
        const int64_t nchunks = (len + 63) / 64;
        if (S > nchunks) S = (int)nchunks;
        const int64_t span = (nchunks + S - 1) / S;
        while (S > 1 && (nchunks + (S - 1) - 1) / (S - 1) == span) S--;
    }
    if (S <= 1) {
        if (vec) return launch_cols<OP, 4, 8, false>(out, in, outer, len, inner, 1, len);
        return launch_cols<OP, 1, 8, false>(out, in, outer, len, inner, 1, len);
    }
    int64_t seg = (len + S - 1) / S;
    int rc = ensure_scratch(outer * S * inner * (int64_t)sizeof(float));
    if (rc != NB200_OK) return rc;
    float *partials = static_cast<float *>(ctx().scratch);
    const int64_t gx = (inner + (int64_t)COL_BX * VEC - 1) / ((int64_t)COL_BX * VEC);
    if (outer * gx <= 4096) {   // one launch: last block per column tile folds the partials (ticket array has 4096 slots)
        return vec ? launch_cols<OP, 4, 8, false>(partials, in, outer, len, inner, S, seg, out, ctx().ticket)
                   : launch_cols<OP, 1, 8, false>(partials, in, outer, len, inner, S, seg, out, ctx().ticket);
    }
    rc = vec ? launch_cols<OP, 4, 8, false>(partials, in, outer, len, inner, S, seg)
             : launch_cols<OP, 1, 8, false>(partials, in, outer, len, inner, S, seg);
    if (rc != NB200_OK) return rc;
    // fold the S partials per (o, i) in fixed order; min/max index-0 NaN rule needs the raw input
    rc = vec ? launch_cols<OP, 4, 1, false>(out, partials, outer, S, inner, 1, S)
             : launch_cols<OP, 1, 1, false>(out, partials, outer, S, inner, 1, S);
    if (rc != NB200_OK) return rc;
    return NB200_OK;
}

template <int OP>
static int reduce_axis_dispatch(float *out, const float *in, int64_t outer, int64_t len, int64_t inner, int order) {
    if (inner == 1) return reduce_rows<OP>(out, in, outer, len, order);
    return reduce_cols<OP>(out, in, outer, len, inner, order);
}

template <bool IS_MAX>
static int argminmax_dispatch(float *out, const float *in, int64_t outer, int64_t len, int64_t inner) {
    cudaStream_t st = ctx().stream;
    if (len >= (int64_t)0xFFFFFFFFll) return set_error(NB200_EINVAL, "argminmax: axis length %lld >= 2^32", (long long)len);
    if (inner == 1) {
        if (len <= 1024 && outer >= 64) {
            int64_t grid = (outer + 7) / 8;
            if (grid > (int64_t)ctx().num_sms * 32) grid = (int64_t)ctx().num_sms * 32;
            arg_rows_warp_kernel<IS_MAX><<<(unsigned)grid, 256, 0, st>>>(out, in, outer, len);
            NB_LAUNCH_CHECK();
            return NB200_OK;
        }
        for (int64_t r0 = 0; r0 < outer; r0 += 65535) {
            int64_t nr = outer - r0 < 65535 ? outer - r0 : 65535;
            int S = pick_split(nr, len);
            unsigned long long *partials = nullptr;
            if (S > 1) {
                if (nr > 4096) S = 1;
                else {
                    int rc = ensure_scratch((int64_t)nr * S * sizeof(unsigned long long));
                    if (rc != NB200_OK) return rc;
                    partials = static_cast<unsigned long long *>(ctx().scratch);
                }
            }
            dim3 grid((unsigned)S, (unsigned)nr);
            arg_rows_kernel<IS_MAX><<<grid, RED_THREADS, 0, st>>>(out + r0, in + r0 * len, len, S, partials, ctx().ticket);
            NB_LAUNCH_CHECK();
        }
        return NB200_OK;
    }
    int64_t gx = (inner + COL_BX - 1) / COL_BX;
    for (int64_t o0 = 0; o0 < outer; o0 += 65535) {
        int64_t ny = outer - o0 < 65535 ? outer - o0 : 65535;
        dim3 grid((unsigned)gx, (unsigned)ny), block(COL_BX, 8);
        arg_cols_kernel<IS_MAX, 8><<<grid, block, 0, st>>>(out + o0 * inner, in + o0 * len * inner, len, inner);
        NB_LAUNCH_CHECK();
    }
    return NB200_OK;
}

// One shard of a sharded argmax / argmin (shard.cu): the packed candidate of in[0..n) - high word = ordering key (larger is better;
// NaN is 0 for argmax, 0xFFFFFFFF for argmin), low word = 0xFFFFFFFF - index of its first occurrence - WITHOUT the "leading NaN
// wins" rule, which applies to global element 0 only.  Also writes the float index next to it (unused).
int argminmax_packed(int is_max, unsigned long long *dev_out, const float *in, int64_t n) {
    if (n <= 0 || n >= (int64_t)0xFFFFFFFFll) return set_error(NB200_EINVAL, "argminmax: shard length %lld out of range", (long long)n);
    int S = pick_split(1, n);
    unsigned long long *partials = nullptr;
    if (S > 1) {
        int rc = ensure_scratch((int64_t)S * sizeof(unsigned long long));
        if (rc != NB200_OK) return rc;
        partials = static_cast<unsigned long long *>(ctx().scratch);
    }
    float *fout = ctx().dev_result + 1;
    dim3 grid((unsigned)S, 1u);
    if (is_max) arg_rows_kernel<true><<<grid, RED_THREADS, 0, ctx().stream>>>(fout, in, n, S, partials, ctx().ticket, dev_out);
    else arg_rows_kernel<false><<<grid, RED_THREADS, 0, ctx().stream>>>(fout, in, n, S, partials, ctx().ticket, dev_out);
    NB_LAUNCH_CHECK();
    return NB200_OK;
}

}  // namespace nb200

using namespace nb200;

extern "C" int nb200_reduce_axis(int op, float *out, const float *in, int64_t outer, int64_t len, int64_t inner,
                                 int order) {
    NB_READY();
    if (!out || !in || outer < 0 || len <= 0 || inner < 0)
        return set_error(NB200_EINVAL, "nb200_reduce_axis: bad argument (outer=%lld len=%lld inner=%lld)",
                         (long long)outer, (long long)len, (long long)inner);
    if (outer == 0 || inner == 0) return NB200_OK;
    switch (op) {
        case NB200_SUM: return reduce_axis_dispatch<NB200_SUM>(out, in, outer, len, inner, order);
        case NB200_PROD: return reduce_axis_dispatch<NB200_PROD>(out, in, outer, len, inner, order);
        case NB200_MIN: return reduce_axis_dispatch<NB200_MIN>(out, in, outer, len, inner, order);
        case NB200_MAX: return reduce_axis_dispatch<NB200_MAX>(out, in, outer, len, inner, order);
        default: return set_error(NB200_EINVAL, "nb200_reduce_axis: unknown op %d", op);
    }
}

extern "C" int nb200_reduce_full(int op, float *dev_out, const float *in, int64_t n) {
    if (n <= 0) return set_error(NB200_EINVAL, "nb200_reduce_full: empty input");
    return nb200_reduce_axis(op, dev_out, in, 1, n, 1, NB200_ORDER_TREE);
}

extern "C" int nb200_reduce_full_host(int op, float *host_out, const float *in, int64_t n) {
    NB_READY();
    if (!host_out) return set_error(NB200_EINVAL, "nb200_reduce_full_host: null output");
    int rc = nb200_reduce_full(op, ctx().dev_result, in, n);
    if (rc != NB200_OK) return rc;
    NB_CUDA(cudaMemcpyAsync(ctx().host_result, ctx().dev_result, sizeof(float), cudaMemcpyDeviceToHost, ctx().stream));
    NB_CUDA(cudaStreamSynchronize(ctx().stream));
    *host_out = *ctx().host_result;
    return NB200_OK;
}

extern "C" int nb200_argminmax(int is_max, float *out, const float *in, int64_t outer, int64_t len, int64_t inner) {
    NB_READY();
    if (!out || !in || outer < 0 || inner < 0) return set_error(NB200_EINVAL, "nb200_argminmax: bad argument");
    if (len <= 0) return set_error(NB200_EINVAL, "attempt to get %s of an empty sequence", is_max ? "argmax" : "argmin");
    if (outer == 0 || inner == 0) return NB200_OK;
    return is_max ? argminmax_dispatch<true>(out, in, outer, len, inner)
                  : argminmax_dispatch<false>(out, in, outer, len, inner);
}

extern "C" int nb200_argminmax_host(int is_max, float *host_out, const float *in, int64_t n) {
    NB_READY();
    if (!host_out) return set_error(NB200_EINVAL, "nb200_argminmax_host: null output");
    int rc = nb200_argminmax(is_max, ctx().dev_result, in, 1, n, 1);
    if (rc != NB200_OK) return rc;
    NB_CUDA(cudaMemcpyAsync(ctx().host_result, ctx().dev_result, sizeof(float), cudaMemcpyDeviceToHost, ctx().stream));
    NB_CUDA(cudaStreamSynchronize(ctx().stream));
    *host_out = *ctx().host_result;
    return NB200_OK;
}
