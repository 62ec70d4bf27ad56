// gemv and 2-D transpose (sm_100a): the two remaining device entry points nd::dot(N-D, 1-D)
// and argmax on a non-last axis reach in the reference.
//   nb200_gemv        replaces cuda_float_multiply_matrix_vector (src/ndmath/cuda/cuda_math.h:62;
//                     CPU oracle cblas_sgemv, src/ndmath/linalg.c:382).  HBM-bound: 4 B per matrix element.
//   nb200_transpose2d replaces cuda_float_transpose (cuda_math.h:77), whose fixed 16x16 grid is only
//                     correct up to 256x256 (cuda_math.cu:136-148).  8 B per element.
#include "common.cuh"

namespace nb200 {

// one warp per row, 128-bit loads, 4 rows in flight per warp iteration for short rows is not needed:
// rows >= #warps in practice; x is re-read through L1/L2.
__global__ void __launch_bounds__(256) gemv_kernel(float *__restrict__ y, const float *__restrict__ A,
                                                   const float *__restrict__ x, int64_t rows, int64_t cols, int vec) {
    const int lane = threadIdx.x & 31;
    for (int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); r < rows; r += (int64_t)gridDim.x * 8) {
        const float *a = A + r * cols;
        float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
        if (vec) {
            const float4 *a4 = reinterpret_cast<const float4 *>(a);
            const float4 *x4 = reinterpret_cast<const float4 *>(x);
            const int64_t n4 = cols >> 2;
            int64_t i = lane;
            for (; i + 96 < n4; i += 128) {
                float4 v0 = ldg_stream(a4 + i), v1 = ldg_stream(a4 + i + 32), v2 = ldg_stream(a4 + i + 64), v3 = ldg_stream(a4 + i + 96);
                float4 x0 = __ldg(x4 + i), x1 = __ldg(x4 + i + 32), x2 = __ldg(x4 + i + 64), x3 = __ldg(x4 + i + 96);
                acc0 += v0.x * x0.x + v0.y * x0.y + v0.z * x0.z + v0.w * x0.w;
                acc1 += v1.x * x1.x + v1.y * x1.y + v1.z * x1.z + v1.w * x1.w;
                acc2 += v2.x * x2.x + v2.y * x2.y + v2.z * x2.z + v2.w * x2.w;
                acc3 += v3.x * x3.x + v3.y * x3.y + v3.z * x3.z + v3.w * x3.w;
            }
            for (; i < n4; i += 32) {
                float4 v0 = ldg_stream(a4 + i), x0 = __ldg(x4 + i);
                acc0 += v0.x * x0.x + v0.y * x0.y + v0.z * x0.z + v0.w * x0.w;
            }
        } else {
            for (int64_t i = lane; i < cols; i += 32) acc0 += a[i] * __ldg(x + i);
        }
        float acc = (acc0 + acc1) + (acc2 + acc3);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) y[r] = acc;
    }
}

__global__ void __launch_bounds__(256) transpose_kernel(float *__restrict__ out, const float *__restrict__ in,
                                                        int64_t rows, int64_t cols) {
    __shared__ float tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    const int64_t c0 = (int64_t)blockIdx.x * 32, r0 = (int64_t)blockIdx.y * 32;
#pragma unroll
    for (int j = 0; j < 32; j += 8)
        if (r0 + ty + j < rows && c0 + tx < cols) tile[ty + j][tx] = in[(r0 + ty + j) * cols + c0 + tx];
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 32; j += 8)
        if (c0 + ty + j < cols && r0 + tx < rows) out[(c0 + ty + j) * rows + r0 + tx] = tile[tx][ty + j];
}

// ---- boolean reductions: nd::all (src/logic.c:25-58) and nd::allclose (src/logic.c:718-771) ------------------------------------
// Both are "does any element violate a predicate": every thread tests its elements (four 16-byte loads in flight per operand),
// __syncthreads_or folds the CTA, and a CTA that saw a violation stores 1 into the result word (plain store of the same value:
// no atomics, no ordering needed).  HBM-bound: 4 B (all) / 8 B (allclose) per element.
//   all:      violation = (x == 0)                      NaN is non-zero, as in the reference's scalar loop (`array[i] == 0.0`)
//   allclose: violation = |a - b| > atol + rtol * |b|   the reference's predicate (logic.c:730-733); a NaN difference is NOT a violation
template <bool CLOSE>
__global__ void __launch_bounds__(256) logic_any_kernel(int *__restrict__ violated, const float *__restrict__ a, const float *__restrict__ b,
                                                        int64_t n, float rtol, float atol, int vec) {
    auto bad = [&](float x, float y) { return CLOSE ? (fabsf(x - y) > atol + rtol * fabsf(y)) : (x == 0.0f); };
    int v = 0;
    if (vec) {
        const float4 *a4 = reinterpret_cast<const float4 *>(a), *b4 = reinterpret_cast<const float4 *>(b);
        const int64_t n4 = n >> 2, stride = (int64_t)gridDim.x * 256;
        int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
        for (; i + 3 * stride < n4; i += 4 * stride) {
            float4 x[4], y[4];
#pragma unroll
            for (int u = 0; u < 4; u++) { x[u] = ldg_stream(a4 + i + u * stride); if (CLOSE) y[u] = ldg_stream(b4 + i + u * stride); else y[u] = x[u]; }
#pragma unroll
            for (int u = 0; u < 4; u++) v |= bad(x[u].x, y[u].x) | bad(x[u].y, y[u].y) | bad(x[u].z, y[u].z) | bad(x[u].w, y[u].w);
        }
        for (; i < n4; i += stride) {
            const float4 x = ldg_stream(a4 + i), y = CLOSE ? ldg_stream(b4 + i) : x;
            v |= bad(x.x, y.x) | bad(x.y, y.y) | bad(x.z, y.z) | bad(x.w, y.w);
        }
        for (int64_t e = (n4 << 2) + (int64_t)blockIdx.x * 256 + threadIdx.x; e < n; e += stride) v |= bad(a[e], CLOSE ? b[e] : a[e]);
    } else {
        for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < n; e += (int64_t)gridDim.x * 256) v |= bad(a[e], CLOSE ? b[e] : a[e]);
    }
    if (__syncthreads_or(v) && threadIdx.x == 0) *violated = 1;
}

static int logic_any(bool close, int *host_out, const float *a, const float *b, int64_t n, float rtol, float atol) {
    int *flag = reinterpret_cast<int *>(ctx().dev_result) + 12;          // byte 48 of the 64-byte result slot
    int *hflag = reinterpret_cast<int *>(ctx().host_result) + 12;
    NB_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), ctx().stream));
    if (n > 0) {
        const int vec = aligned16(a) && (!close || aligned16(b));
        int64_t grid = (n / 4 + 1023) / 1024, cap = (int64_t)ctx().num_sms * 8;   // <= one wave of 8 CTAs per SM, >= 4 loads per thread
        if (grid > cap) grid = cap;
        if (grid < 1) grid = 1;
        if (close) logic_any_kernel<true><<<(unsigned)grid, 256, 0, ctx().stream>>>(flag, a, b, n, rtol, atol, vec);
        else logic_any_kernel<false><<<(unsigned)grid, 256, 0, ctx().stream>>>(flag, a, a, n, 0.f, 0.f, vec);
        NB_LAUNCH_CHECK();
    }
    NB_CUDA(cudaMemcpyAsync(hflag, flag, sizeof(int), cudaMemcpyDeviceToHost, ctx().stream));
    NB_CUDA(cudaStreamSynchronize(ctx().stream));
    *host_out = *hflag ? 0 : 1;
    return NB200_OK;
}

}  // namespace nb200
using namespace nb200;

extern "C" int nb200_all(int *host_out, const float *a, int64_t n) {
    NB_READY();
    if (!host_out || (!a && n > 0) || n < 0) return set_error(NB200_EINVAL, "nb200_all: bad argument");
    return logic_any(false, host_out, a, a, n, 0.f, 0.f);
}

extern "C" int nb200_allclose(int *host_out, const float *a, const float *b, int64_t n, float rtol, float atol) {
    NB_READY();
    if (!host_out || ((!a || !b) && n > 0) || n < 0) return set_error(NB200_EINVAL, "nb200_allclose: bad argument");
    return logic_any(true, host_out, a, b, n, rtol, atol);
}

extern "C" int nb200_gemv(float *y, const float *A, const float *x, int64_t rows, int64_t cols) {
    NB_READY();
    if (!y || !A || !x || rows < 0 || cols < 0) return set_error(NB200_EINVAL, "nb200_gemv: bad argument");
    if (rows == 0) return NB200_OK;
    int vec = (cols % 4 == 0) && aligned16(A) && aligned16(x);
    int64_t grid = (rows + 7) / 8, cap = (int64_t)ctx().num_sms * 32;
    if (grid > cap) grid = cap;
    gemv_kernel<<<(unsigned)grid, 256, 0, ctx().stream>>>(y, A, x, rows, cols, vec);
    NB_LAUNCH_CHECK();
    return NB200_OK;
}

extern "C" int nb200_transpose2d(float *out, const float *in, int64_t rows, int64_t cols) {
    NB_READY();
    if (!out || !in || rows < 0 || cols < 0 || out == in) return set_error(NB200_EINVAL, "nb200_transpose2d: bad argument");
    if (rows == 0 || cols == 0) return NB200_OK;
    int64_t gx = (cols + 31) / 32, gy = (rows + 31) / 32;
    if (gy > 65535) return set_error(NB200_EINVAL, "nb200_transpose2d: more than 2M rows unsupported");
    dim3 grid((unsigned)gx, (unsigned)gy);
    transpose_kernel<<<grid, 256, 0, ctx().stream>>>(out, in, rows, cols);
    NB_LAUNCH_CHECK();
    return NB200_OK;
}
