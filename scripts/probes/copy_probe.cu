// Micro-benchmark (test infrastructure): which load/store policy, unroll and grid shape gets a 1-read : 1-write streaming
// kernel closest to the HBM copy roofline on B200?  nvcc -gencode arch=compute_100a,code=sm_100a -O3 copy_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

template <int LD, int ST>
struct IO {
    static __device__ __forceinline__ float4 ld(const float4 *p) {
        float4 v;
        if (LD == 0) asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
        else if (LD == 1) v = *p;
        else asm volatile("ld.global.cs.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
        return v;
    }
    static __device__ __forceinline__ void st(float4 *p, float4 v) {
        if (ST == 0) asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
        else if (ST == 1) *p = v;
        else if (ST == 2) asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
        else asm volatile("st.global.wt.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    }
};

template <int LD, int ST, int U, int T>
__global__ void __launch_bounds__(T) k_copy(float4 *__restrict__ out, const float4 *__restrict__ in, int64_t n4) {
    const int64_t tile = (int64_t)T * U;
    for (int64_t base = (int64_t)blockIdx.x * tile; base < n4; base += (int64_t)gridDim.x * tile) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            int64_t i = base + (int64_t)u * T + threadIdx.x;
            if (i < n4) v[u] = IO<LD, ST>::ld(in + i);
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            int64_t i = base + (int64_t)u * T + threadIdx.x;
            if (i < n4) { v[u].x += 1.0f; IO<LD, ST>::st(out + i, v[u]); }
        }
    }
}

template <int LD, int ST, int U, int T>
float run(float4 *out, const float4 *in, int64_t n4, int grid_mult, float4 *flush, int64_t nflush) {
    int sms = 148;
    int64_t tiles = (n4 + (int64_t)T * U - 1) / ((int64_t)T * U);
    int64_t grid = grid_mult > 0 ? (int64_t)sms * grid_mult : tiles;
    if (grid > tiles) grid = tiles;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9, tot = 0;
    for (int it = 0; it < 8; it++) {
        cudaMemsetAsync(flush, 0, nflush);
        cudaEventRecord(e0);
        k_copy<LD, ST, U, T><<<(unsigned)grid, T>>>(out, in, n4);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (it >= 2) { tot += ms; if (ms < best) best = ms; }
    }
    return tot / 6;
}

int main() {
    const int64_t n = (int64_t)8192 * 8192, n4 = n / 4;
    float4 *a, *b, *flush;
    cudaMalloc(&a, n * 4); cudaMalloc(&b, n * 4); cudaMalloc(&flush, 256 << 20);
    cudaMemset(a, 0, n * 4);
    const double bytes = 2.0 * n * 4;
#define R(LD, ST, U, T, GM) { float ms = run<LD, ST, U, T>(b, a, n4, GM, flush, 256 << 20); \
    printf("ld=%d st=%d unroll=%d threads=%d grid=%s%d : %.4f ms  %.0f GB/s\n", LD, ST, U, T, GM > 0 ? "sms*" : "tiles/", GM, ms, bytes / ms / 1e6); }
    // cudaMemcpy D2D reference
    {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1); float tot = 0;
        for (int it = 0; it < 8; it++) { cudaMemsetAsync(flush, 0, 256 << 20); cudaEventRecord(e0); cudaMemcpyAsync(b, a, n * 4, cudaMemcpyDeviceToDevice); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (it >= 2) tot += ms; }
        printf("cudaMemcpyAsync D2D: %.4f ms %.0f GB/s\n", tot / 6, bytes / (tot / 6) / 1e6);
    }
    R(0, 0, 4, 256, 16) R(0, 0, 4, 256, 0) R(0, 0, 4, 256, 8) R(0, 0, 4, 256, 32)
    R(0, 1, 4, 256, 16) R(0, 2, 4, 256, 16) R(0, 3, 4, 256, 16)
    R(1, 1, 4, 256, 16) R(1, 1, 4, 256, 0) R(2, 0, 4, 256, 16)
    R(0, 0, 8, 256, 8) R(0, 0, 8, 256, 0) R(0, 1, 8, 256, 8) R(0, 0, 2, 256, 32) R(0, 0, 2, 256, 0)
    R(0, 0, 4, 512, 8) R(0, 0, 4, 128, 32) R(0, 0, 8, 128, 16) R(0, 0, 4, 1024, 4) R(1, 1, 4, 1024, 0)
    R(0, 0, 1, 256, 0) R(1, 1, 1, 256, 0) R(1, 1, 2, 256, 0)
    return 0;
}
