// PCIe duplex probe for the host-operand pipeline (host_pipeline.cu): which side's copy boundaries cost under duplex traffic, and
// what SM-driven transfers (loads from / stores to mapped pinned memory) reach next to the copy engines.
//   nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o scripts/probes/pcie_probe scripts/probes/pcie_probe.cu
// Prints one JSON object per line.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <algorithm>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

template <int U>
__global__ void __launch_bounds__(256) pull_kernel(float4 *__restrict__ dst, const float4 *__restrict__ src, int64_t n4) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (U - 1) * stride < n4; i += U * stride) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; u++) v[u] = __ldcs(src + i + u * stride);
#pragma unroll
        for (int u = 0; u < U; u++) dst[i + u * stride] = v[u];
    }
    for (; i < n4; i += stride) dst[i] = src[i];
}

static const size_t TOT = 64u << 20;
static float *h_in, *h_out, *d_in, *d_out, *m_in, *m_out;
static cudaStream_t s_in, s_out;

struct R { float in_ms, out_ms, all_ms; };

// in_mode / out_mode: 0 none, 1 copy engine, 2 SM kernel; chunk sizes in bytes; ctas for the kernels
static R run(int in_mode, size_t in_chunk, int in_ctas, int out_mode, size_t out_chunk, int out_ctas) {
    cudaEvent_t e0, ei, eo, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&ei)); CK(cudaEventCreate(&eo)); CK(cudaEventCreate(&e1));
    R best = {1e9f, 1e9f, 1e9f};
    for (int rep = 0; rep < 5; rep++) {
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0, s_in));
        CK(cudaStreamWaitEvent(s_out, e0, 0));
        if (in_mode) for (size_t o = 0; o < TOT; o += in_chunk) {
            if (in_mode == 1) CK(cudaMemcpyAsync((char *)d_in + o, (char *)h_in + o, in_chunk, cudaMemcpyHostToDevice, s_in));
            else pull_kernel<8><<<in_ctas, 256, 0, s_in>>>((float4 *)((char *)d_in + o), (const float4 *)((char *)m_in + o), (int64_t)(in_chunk / 16));
        }
        CK(cudaEventRecord(ei, s_in));
        if (out_mode) for (size_t o = 0; o < TOT; o += out_chunk) {
            if (out_mode == 1) CK(cudaMemcpyAsync((char *)h_out + o, (char *)d_out + o, out_chunk, cudaMemcpyDeviceToHost, s_out));
            else pull_kernel<4><<<out_ctas, 256, 0, s_out>>>((float4 *)((char *)m_out + o), (const float4 *)((char *)d_out + o), (int64_t)(out_chunk / 16));
        }
        CK(cudaEventRecord(eo, s_out));
        CK(cudaStreamWaitEvent(s_in, eo, 0));
        CK(cudaEventRecord(e1, s_in));
        CK(cudaDeviceSynchronize());
        float a, b, c;
        CK(cudaEventElapsedTime(&a, e0, ei)); CK(cudaEventElapsedTime(&b, e0, eo)); CK(cudaEventElapsedTime(&c, e0, e1));
        if (c < best.all_ms) best = {a, b, c};
    }
    return best;
}

static void report(const char *name, int in_mode, size_t in_chunk, int in_ctas, int out_mode, size_t out_chunk, int out_ctas) {
    R r = run(in_mode, in_chunk, in_ctas, out_mode, out_chunk, out_ctas);
    const double gb = TOT / 1e6;
    printf("{\"test\": \"%s\", \"in\": \"%s\", \"in_chunk_MiB\": %zu, \"in_ctas\": %d, \"out\": \"%s\", \"out_chunk_MiB\": %zu, \"out_ctas\": %d, "
           "\"in_ms\": %.3f, \"out_ms\": %.3f, \"all_ms\": %.3f, \"in_GBps\": %.1f, \"out_GBps\": %.1f}\n",
           name, in_mode == 0 ? "-" : in_mode == 1 ? "ce" : "sm", in_chunk >> 20, in_ctas, out_mode == 0 ? "-" : out_mode == 1 ? "ce" : "sm", out_chunk >> 20, out_ctas,
           r.in_ms, r.out_ms, r.all_ms, in_mode ? gb / r.in_ms : 0.0, out_mode ? gb / r.out_ms : 0.0);
    fflush(stdout);
}


// ---- emulation of nb200_sgemm_host's stream structure: which ingredient costs the H2D leg its duplex rate?
__global__ void spin_kernel(unsigned long long ns) {
    unsigned long long t0, t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    do { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); } while (t - t0 < ns);
}
static void emulate(const char *name, size_t chunk, bool rec_events, int gate, unsigned long long kernel_ns, int spin_ctas, bool head_start) {
    static cudaStream_t s_c = nullptr;
    static cudaEvent_t ev_in[64], ev_done[64];
    if (!s_c) {
        CK(cudaStreamCreateWithFlags(&s_c, cudaStreamNonBlocking));
        for (int i = 0; i < 64; i++) { CK(cudaEventCreateWithFlags(&ev_in[i], cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&ev_done[i], cudaEventDisableTiming)); }
    }
    cudaEvent_t e0, ei, eo, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&ei)); CK(cudaEventCreate(&eo)); CK(cudaEventCreate(&e1));
    const int n = (int)(TOT / chunk);
    float bi = 1e9f, bo = 1e9f, ba = 1e9f;
    for (int rep = 0; rep < 5; rep++) {
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0, s_in));
        CK(cudaStreamWaitEvent(s_out, e0, 0));
        CK(cudaStreamWaitEvent(s_c, e0, 0));
        // (head_start: a 64 MiB upload first, like B, so that the out leg starts against a busy in leg)
        if (head_start) CK(cudaMemcpyAsync(d_out, h_in, TOT, cudaMemcpyHostToDevice, s_in));
        for (int i = 0; i < n; i++) {
            CK(cudaMemcpyAsync((char *)d_in + i * chunk, (char *)h_in + i * chunk, chunk, cudaMemcpyHostToDevice, s_in));
            if (rec_events || gate) CK(cudaEventRecord(ev_in[i], s_in));
        }
        CK(cudaEventRecord(ei, s_in));
        for (int i = 0; i < n; i++) {
            if (gate == 1) CK(cudaStreamWaitEvent(s_out, ev_in[i], 0));
            if (gate == 2) {
                CK(cudaStreamWaitEvent(s_c, ev_in[i], 0));
                spin_kernel<<<spin_ctas, 192, 0, s_c>>>(kernel_ns);
                CK(cudaEventRecord(ev_done[i], s_c));
                CK(cudaStreamWaitEvent(s_out, ev_done[i], 0));
            }
            CK(cudaMemcpyAsync((char *)h_out + i * chunk, (char *)d_in + i * chunk, chunk, cudaMemcpyDeviceToHost, s_out));
        }
        CK(cudaEventRecord(eo, s_out));
        CK(cudaStreamWaitEvent(s_in, eo, 0));
        CK(cudaEventRecord(e1, s_in));
        CK(cudaDeviceSynchronize());
        float a, b, c;
        CK(cudaEventElapsedTime(&a, e0, ei)); CK(cudaEventElapsedTime(&b, e0, eo)); CK(cudaEventElapsedTime(&c, e0, e1));
        if (c < ba) { bi = a; bo = b; ba = c; }
    }
    printf("{\"test\": \"emulate_%s\", \"chunk_MiB\": %zu, \"rec_events\": %d, \"gate\": %d, \"kernel_us\": %.0f, \"spin_ctas\": %d, \"head_start\": %d, "
           "\"in_done_ms\": %.3f, \"out_done_ms\": %.3f, \"all_ms\": %.3f}\n", name, chunk >> 20, (int)rec_events, gate, kernel_ns / 1e3, spin_ctas, (int)head_start, bi, bo, ba);
    fflush(stdout);
}

// ---- structural alternatives for the same work (64 MiB head upload, then n chunks: upload, kernel, download)
//  variant 2: H2D chunks on s_in (+ event each); K worker streams round-robin: wait ev_in[i], kernel, D2H in-stream
//  variant 8: K worker streams round-robin, each chunk entirely in-stream: H2D, kernel, D2H (no cross-stream event but the head's)
//  variant 3: like 2, but the HOST waits for ev_in[i] (cudaEventSynchronize) and only then enqueues kernel + D2H: no device-side waits
static void emulate2(int variant, int K, size_t chunk, unsigned long long kernel_ns, int spin_ctas) {
    static cudaStream_t w[8];
    static cudaEvent_t ev_in[64], ev_head, ev_w[8];
    static bool init = false;
    if (!init) {
        for (int i = 0; i < 8; i++) { CK(cudaStreamCreateWithFlags(&w[i], cudaStreamNonBlocking)); CK(cudaEventCreateWithFlags(&ev_w[i], cudaEventDisableTiming)); }
        for (int i = 0; i < 64; i++) CK(cudaEventCreateWithFlags(&ev_in[i], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&ev_head, cudaEventDisableTiming));
        init = true;
    }
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int n = (int)(TOT / chunk);
    float best = 1e9f;
    for (int rep = 0; rep < 5; rep++) {
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0, s_in));
        CK(cudaMemcpyAsync(d_out, h_in, TOT, cudaMemcpyHostToDevice, s_in));   // head ("B")
        CK(cudaEventRecord(ev_head, s_in));
        for (int k = 0; k < K; k++) CK(cudaStreamWaitEvent(w[k], ev_head, 0));
        if (variant == 2 || variant == 3) {
            for (int i = 0; i < n; i++) {
                CK(cudaMemcpyAsync((char *)d_in + i * chunk, (char *)h_in + i * chunk, chunk, cudaMemcpyHostToDevice, s_in));
                CK(cudaEventRecord(ev_in[i], s_in));
            }
        }
        for (int i = 0; i < n; i++) {
            cudaStream_t x = w[i % K];
            if (variant == 2) CK(cudaStreamWaitEvent(x, ev_in[i], 0));
            if (variant == 3) CK(cudaEventSynchronize(ev_in[i]));
            if (variant == 8) CK(cudaMemcpyAsync((char *)d_in + i * chunk, (char *)h_in + i * chunk, chunk, cudaMemcpyHostToDevice, x));
            spin_kernel<<<spin_ctas, 192, 0, x>>>(kernel_ns);
            CK(cudaMemcpyAsync((char *)h_out + i * chunk, (char *)d_in + i * chunk, chunk, cudaMemcpyDeviceToHost, x));
        }
        for (int k = 0; k < K; k++) { CK(cudaEventRecord(ev_w[k], w[k])); CK(cudaStreamWaitEvent(s_in, ev_w[k], 0)); }
        CK(cudaEventRecord(e1, s_in));
        CK(cudaDeviceSynchronize());
        float c;
        CK(cudaEventElapsedTime(&c, e0, e1));
        if (c < best) best = c;
    }
    printf("{\"test\": \"emulate2\", \"variant\": %d, \"worker_streams\": %d, \"chunk_MiB\": %zu, \"kernel_us\": %.0f, \"spin_ctas\": %d, \"all_ms\": %.3f}\n",
           variant, K, chunk >> 20, kernel_ns / 1e3, spin_ctas, best);
    fflush(stdout);
}

int main() {
    CK(cudaSetDevice(0));
    CK(cudaHostAlloc(&h_in, TOT, cudaHostAllocMapped));
    CK(cudaHostAlloc(&h_out, TOT, cudaHostAllocMapped));
    for (size_t i = 0; i < TOT / 4; i++) h_in[i] = (float)(i & 1023);
    CK(cudaMalloc(&d_in, TOT)); CK(cudaMalloc(&d_out, TOT));
    CK(cudaMemset(d_out, 1, TOT));
    CK(cudaHostGetDevicePointer((void **)&m_in, h_in, 0));
    CK(cudaHostGetDevicePointer((void **)&m_out, h_out, 0));
    CK(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking));
    const size_t M = 1u << 20;
    // one direction alone
    report("in_alone", 1, 64 * M, 0, 0, 0, 0);
    report("out_alone", 0, 0, 0, 1, 64 * M, 0);
    for (int ctas : {4, 8, 16, 32, 64, 148}) report("in_alone_sm", 2, 64 * M, ctas, 0, 0, 0);
    for (int ctas : {4, 8, 16, 32, 64}) report("out_alone_sm", 0, 0, 0, 2, 64 * M, ctas);
    // duplex, copy engines, asymmetric chunking
    report("duplex_ce", 1, 64 * M, 0, 1, 64 * M, 0);
    for (size_t c : {4 * M, 8 * M, 16 * M}) {
        report("duplex_ce_in_chunked", 1, c, 0, 1, 64 * M, 0);
        report("duplex_ce_out_chunked", 1, 64 * M, 0, 1, c, 0);
        report("duplex_ce_both_chunked", 1, c, 0, 1, c, 0);
    }
    // duplex with SM-driven legs
    for (int ctas : {16, 32, 64}) {
        report("duplex_sm_in_ce_out", 2, 64 * M, ctas, 1, 64 * M, 0);
        report("duplex_sm_in_ce_out_4MiB", 2, 64 * M, ctas, 1, 4 * M, 0);
        report("duplex_sm_in_chunked_ce_out_4MiB", 2, 4 * M, ctas, 1, 4 * M, 0);
    }
    for (int ctas : {8, 16, 32}) {
        report("duplex_ce_in_sm_out", 1, 64 * M, 0, 2, 64 * M, ctas);
        report("duplex_ce_in_4MiB_sm_out_4MiB", 1, 4 * M, 0, 2, 4 * M, ctas);
        report("duplex_ce_in_8MiB_sm_out_8MiB", 1, 8 * M, 0, 2, 8 * M, ctas);
    }
    report("duplex_sm_both", 2, 64 * M, 32, 2, 64 * M, 16);
    report("duplex_sm_both_4MiB", 2, 4 * M, 32, 2, 4 * M, 16);
    for (size_t c : {4 * M, 8 * M, 16 * M}) {
        emulate("plain", c, false, 0, 0, 0, false);
        emulate("events", c, true, 0, 0, 0, false);
        emulate("gated", c, true, 1, 0, 0, false);
        emulate("gated_kernel", c, true, 2, 70000, 148, false);
        emulate("gated_kernel_small", c, true, 2, 5000, 1, false);
        emulate("head_plain", c, false, 0, 0, 0, true);
        emulate("head_gated", c, true, 1, 0, 0, true);
        emulate("head_gated_kernel", c, true, 2, 70000, 148, true);
    }
    for (size_t c : {4 * M, 8 * M}) {
        for (int K : {2, 3}) emulate2(2, K, c, 70000, 148);
        for (int K : {2, 3, 4}) emulate2(8, K, c, 70000, 148);
        for (int K : {2}) emulate2(3, K, c, 70000, 148);
    }
    emulate2(8, 3, 2 * M, 35000, 148);
    emulate2(8, 4, 2 * M, 35000, 148);
    // verify the SM pull moved the data
    CK(cudaMemset(d_in, 0, TOT));
    pull_kernel<8><<<32, 256, 0, s_in>>>((float4 *)d_in, (const float4 *)m_in, (int64_t)(TOT / 16));
    CK(cudaMemcpyAsync(h_out, d_in, TOT, cudaMemcpyDeviceToHost, s_in));
    CK(cudaStreamSynchronize(s_in));
    size_t bad = 0;
    for (size_t i = 0; i < TOT / 4; i++) bad += h_out[i] != h_in[i];
    printf("{\"test\": \"verify_sm_pull\", \"mismatches\": %zu}\n", bad);
    return 0;
}
