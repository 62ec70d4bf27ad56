// Shared context, error plumbing and load/store helpers for libnb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string.h>
#include <map>
#include <unordered_map>
#include <nvtx3/nvToolsExt.h>   // header-only NVTX v3: ranges cost a pointer test unless a profiler injected itself
#include "../../include/nb200.h"

#ifndef NB200_NUM_SMS_DEFAULT
#define NB200_NUM_SMS_DEFAULT 148
#endif

namespace nb200 {

struct Ctx {
    bool ready = false;
    int device = -1;
    int num_sms = NB200_NUM_SMS_DEFAULT;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    // scratch for two-stage reductions / split-axis partials (device)
    void *scratch = nullptr;
    int64_t scratch_bytes = 0;
    // GEMM workspace (3xTF32 lo parts)
    void *gemm_ws = nullptr;
    int64_t gemm_ws_bytes = 0;
    // FP16x3 control blocks at the start of gemm_ws (record counters, pre-pass barrier, column maxima), double-buffered by call
    // parity: the pre-pass of call n zeroes the block call n+1 will use, so steady-state calls need no memset (sgemm_tcgen05.cu)
    int64_t ctl_stride = -1;              // bytes per block in the current layout; -1: layout unknown (workspace used by another mode / reallocated)
    int64_t ctl_ready[2] = {0, 0};        // leading bytes of each block known to be zero
    int ctl_parity = 0;
    unsigned int *ticket = nullptr;       // last-block-done counters
    int *domain_flag = nullptr;           // sticky math-domain flag (device)
    float *host_result = nullptr;         // pinned 64-byte result slot
    float *dev_result = nullptr;          // device result slot
    int nonfinite_gen = 0;                // tag of the current x3 GEMM call (sgemm_tcgen05.cu, gemm_reset_nonfinite)
    int64_t live_allocs = 0;
    int64_t live_bytes = 0;
    int64_t launches = 0;
    int gemm_sm_reserve = 0;              // SMs the persistent GEMM grids leave free (NCCL transfer kernels running beside them, shard.cu)
    // allocation ledger + caching pool of THIS device (abi.cu): live blocks handed to the host, and freed blocks by capacity
    std::unordered_map<void *, int64_t> ledger;
    std::multimap<int64_t, void *> pool;
    int64_t pool_bytes = 0;
    unsigned long long *trace = nullptr;  // optional device buffer for %globaltimer stamps of the GEMM pipeline (nb200_trace_enable)
};

constexpr int NB200_MAX_DEVICES = 16;
Ctx &ctx();                      // context of the current device (nb200_set_device)
Ctx *ctx_of(int device);         // nullptr if out of range; ->ready tells whether the device was initialised
int set_error(int code, const char *fmt, ...);
int ensure_ready();
int ensure_scratch(int64_t bytes);
int ensure_gemm_ws(int64_t bytes);
int gemm_split_operand(const float *in, float *lo, int64_t n);
int gemm_reset_nonfinite();
int gemm_resolve_precision(int precision, int64_t K);   // NB200_GEMM_AUTO -> concrete mode
int gemm_bf16_split(const float *in, void *hi, void *lo, int64_t rows, int64_t cols);
int gemm_bf16_presplit(float *C, const void *a_hi, const void *a_lo, const void *b_hi, const void *b_lo, int64_t M, int64_t N,
                       int64_t K, int64_t ldc);
int gemm_presplit(float *C, const float *A, const float *A_lo, const float *B, const float *B_lo, int64_t M, int64_t N,
                  int64_t K, int64_t lda, int64_t ldb, int64_t ldc, int precision);

#define NB_CUDA(expr)                                                                         \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess)                                                                \
            return nb200::set_error(NB200_ECUDA, "%s failed: %s (%s:%d)", #expr,               \
                                    cudaGetErrorString(_e), __FILE__, __LINE__);               \
    } while (0)

#define NB_LAUNCH_CHECK()                                                                     \
    do {                                                                                      \
        nb200::ctx().launches++;                                                              \
        cudaError_t _e = cudaGetLastError();                                                  \
        if (_e != cudaSuccess)                                                                \
            return nb200::set_error(NB200_ECUDA, "kernel launch failed: %s (%s:%d)",           \
                                    cudaGetErrorString(_e), __FILE__, __LINE__);               \
    } while (0)

// Every C-ABI entry point opens an NVTX range named after itself (Nsight Systems / ncu --nvtx show nb200_sgemm, nb200_ew_binary, ...
// around the kernels they enqueue); the reference has no tracing hooks at all.
struct NvtxScope {
    explicit NvtxScope(const char *name) { nvtxRangePushA(name); }
    ~NvtxScope() { nvtxRangePop(); }
    NvtxScope(const NvtxScope &) = delete;
    NvtxScope &operator=(const NvtxScope &) = delete;
};

#define NB_READY()                                  \
    nb200::NvtxScope _nb_nvtx_scope(__func__);      \
    do {                                            \
        int _r = nb200::ensure_ready();             \
        if (_r != NB200_OK) return _r;              \
    } while (0)

// ---- 128-bit global access ---------------------------------------------------------------------------
// Measured on B200 (scripts/probes/copy_probe.cu -> profiles/r1_copy_probe.log, and bench.py before/after):
//  * read+write streaming kernels (elementwise): default-policy ld.global.nc / st.global with ONE tile per CTA reach
//    6.25-6.6 TB/s; the "streaming" hints (L1::no_allocate loads + st.global.cs, persistent grid) stop at 5.7 TB/s.
//  * read-only streaming kernels (reductions, argmax, gemv): the L1::no_allocate loads are the faster ones
//    (sum over 2^28: 6.13 vs 5.97 TB/s; row sums of 8192^2: 5.87 vs 4.80 TB/s).
__device__ __forceinline__ float4 ld_ew(const float4 *p) { return __ldg(p); }
__device__ __forceinline__ float ld_ew(const float *p) { return __ldg(p); }
__device__ __forceinline__ void st_ew(float4 *p, const float4 &v) { *p = v; }
__device__ __forceinline__ void st_ew(float *p, float v) { *p = v; }

__device__ __forceinline__ float4 ldg_stream(const float4 *p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ float ldg_stream(const float *p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

// ---- optional timeline of the matmul pipeline: kernels stamp %globaltimer (ns) into Ctx::trace when it is set.
// Slots: 0 pre-pass first CTA starts | 1 its phase A done | 2 its phase B1 done | 3 grid barrier passed | 4 last pre-pass CTA done |
//        5 GEMM first CTA enters | 6 it passed griddepcontrol.wait | 7 last GEMM CTA done | 8 post kernel enters | 9 post kernel done |
//        10 gated fallback enters | 11 gated fallback done.  "first" = atomicMin over the CTAs, "last" = atomicMax.
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void trace_min(unsigned long long *trace, int slot) { if (trace) atomicMin(trace + slot, global_ns()); }
__device__ __forceinline__ void trace_max(unsigned long long *trace, int slot) { if (trace) atomicMax(trace + slot, global_ns()); }

// L2 eviction priority for 128-bit accesses goes through a cache-policy operand (the plain .L2::evict_* qualifiers exist only for
// 256-bit accesses on sm_100).  evict_first: stream-once data that must not push a working set out of the 126 MB L2.
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ float4 ldg_stream_l2(const float4 *p, uint64_t pol) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ void st_l2(uint4 *p, const uint4 &v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v4.b32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_l2(uint2 *p, const uint2 &v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v2.b32 [%0], {%1,%2}, %3;" ::"l"(p), "r"(v.x), "r"(v.y), "l"(pol) : "memory");
}

static inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace nb200
