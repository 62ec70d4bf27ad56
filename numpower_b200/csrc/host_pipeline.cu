// Host-operand entry points: the step either side of the hot path (SURVEY.md §8 f, N3).  The reference moves whole
// arrays with a pageable, blocking cudaMemcpy on the default stream (NDArray_ToGPU / NDArray_ToCPU,
// src/ndarray.c:1037-1093) and only then computes.  nb200_sgemm_host pipelines instead:
//   copy-in stream : B (once), then A row blocks            H2D
//   compute stream : lo-split + tcgen05 GEMM per row block  (waits on the block's H2D event)
//   copy-out stream: C row blocks                           D2H (overlaps the next blocks' H2D: PCIe is full duplex)
#include "common.cuh"
#include <cstdlib>

namespace nb200 {
namespace {
struct Pipe {
    cudaStream_t s_in = nullptr, s_out = nullptr;
    float *dA = nullptr, *dAlo = nullptr, *dB = nullptr, *dBlo = nullptr, *dC = nullptr;
    int64_t capA = 0, capB = 0, capC = 0;
    int device = -1;
    static constexpr int MAXB = 64;
    cudaEvent_t ev_in[MAXB], ev_done[MAXB], ev_b = nullptr;
} g_pipes[NB200_MAX_DEVICES];   // one per device: streams, events and staging buffers cannot follow nb200_set_device

int grow(float **p, int64_t *cap, int64_t elems) {
    if (elems <= *cap) return NB200_OK;
    if (*p) NB_CUDA(cudaFree(*p));
    *p = nullptr;
    *cap = 0;
    if (cudaMalloc(p, (size_t)elems * 4) != cudaSuccess) {
        cudaGetLastError();
        return set_error(NB200_ENOMEM, "device memory allocation failed (host pipeline staging, %lld bytes)", (long long)elems * 4);
    }
    *cap = elems;
    return NB200_OK;
}
}  // namespace

void host_pipeline_release(int device) {
    if (device < 0 || device >= NB200_MAX_DEVICES) return;
    Pipe &P = g_pipes[device];
    if (P.device < 0) return;
    if (P.s_in) cudaStreamDestroy(P.s_in);
    if (P.s_out) cudaStreamDestroy(P.s_out);
    for (int i = 0; i < Pipe::MAXB; i++) { cudaEventDestroy(P.ev_in[i]); cudaEventDestroy(P.ev_done[i]); }
    if (P.ev_b) cudaEventDestroy(P.ev_b);
    if (P.dA) cudaFree(P.dA);
    if (P.dB) cudaFree(P.dB);
    if (P.dC) cudaFree(P.dC);
    P = Pipe();
}
}  // namespace nb200

using namespace nb200;

extern "C" int nb200_sgemm_host(float *C_host, const float *A_host, const float *B_host, int64_t M, int64_t N, int64_t K,
                                int precision) {
    NB_READY();
    if (!C_host || !A_host || !B_host || M < 0 || N < 0 || K < 0) return set_error(NB200_EINVAL, "nb200_sgemm_host: bad argument");
    if (precision < NB200_GEMM_TF32X3 || precision > NB200_GEMM_FP16X3) return set_error(NB200_EINVAL, "unknown precision %d", precision);
    precision = gemm_resolve_precision(precision, K);
    if (precision == NB200_GEMM_FP16X3) precision = NB200_GEMM_TF32X3;   // this path is PCIe-bound; it keeps the two-kernel TF32x3 pipeline
    if (M == 0 || N == 0) return NB200_OK;
    Ctx &c = ctx();
    Pipe &P = g_pipes[c.device];
    if (P.device != c.device) {
        NB_CUDA(cudaStreamCreateWithFlags(&P.s_in, cudaStreamNonBlocking));
        NB_CUDA(cudaStreamCreateWithFlags(&P.s_out, cudaStreamNonBlocking));
        for (int i = 0; i < Pipe::MAXB; i++) {
            NB_CUDA(cudaEventCreateWithFlags(&P.ev_in[i], cudaEventDisableTiming));
            NB_CUDA(cudaEventCreateWithFlags(&P.ev_done[i], cudaEventDisableTiming));
        }
        NB_CUDA(cudaEventCreateWithFlags(&P.ev_b, cudaEventDisableTiming));
        P.device = c.device;
    }
    // shapes the tensor path cannot serve (tiny / K,N not multiples of 4): plain three-step path
    const bool pipelined = (K % 4 == 0) && (N % 4 == 0) && K >= 32 && N >= 32 && M >= 256 && M * N * K >= (int64_t)1 << 24;
    // row block: ~8 blocks, multiple of 256 rows (one CTA-pair tile), at most MAXB blocks
    int64_t rb = M;
    if (pipelined) {
        rb = ((M / 8 + 255) / 256) * 256;
        if (rb < 256) rb = 256;
        while ((M + rb - 1) / rb > Pipe::MAXB) rb += 256;
    }
    const int64_t nblk = (M + rb - 1) / rb;
    int rc;
    if ((rc = grow(&P.dA, &P.capA, M * K)) != NB200_OK) return rc;
    if ((rc = grow(&P.dB, &P.capB, K * N)) != NB200_OK) return rc;
    if ((rc = grow(&P.dC, &P.capC, M * N)) != NB200_OK) return rc;
    if (!pipelined) {
        NB_CUDA(cudaMemcpyAsync(P.dA, A_host, (size_t)M * K * 4, cudaMemcpyHostToDevice, c.stream));
        NB_CUDA(cudaMemcpyAsync(P.dB, B_host, (size_t)K * N * 4, cudaMemcpyHostToDevice, c.stream));
        if ((rc = nb200_sgemm(P.dC, P.dA, P.dB, M, N, K, K, N, N, precision)) != NB200_OK) return rc;
        NB_CUDA(cudaMemcpyAsync(C_host, P.dC, (size_t)M * N * 4, cudaMemcpyDeviceToHost, c.stream));
        NB_CUDA(cudaStreamSynchronize(c.stream));
        return NB200_OK;
    }
    // lo parts come from the pre-pass (B once, A per row block) unless NB200_GEMM_INKERNEL=1 selects the in-kernel split
    const bool x3 = precision == NB200_GEMM_TF32X3 && getenv("NB200_GEMM_INKERNEL") == nullptr;
    const bool b3 = precision == NB200_GEMM_BF16X3;
    const int64_t Kp = (K + 7) & ~int64_t(7), Np = (N + 7) & ~int64_t(7);   // packed bf16 leading dimensions
    uint16_t *ah = nullptr, *al = nullptr, *bh = nullptr, *bl = nullptr;
    if (b3) {
        if ((rc = ensure_gemm_ws((M * Kp + K * Np) * 4 + 1024)) != NB200_OK) return rc;
        ah = static_cast<uint16_t *>(c.gemm_ws); al = ah + M * Kp; bh = al + M * Kp; bl = bh + K * Np;
    }
    if (x3) {
        if ((rc = ensure_gemm_ws((M * K + K * N) * 4 + 256)) != NB200_OK) return rc;
        P.dAlo = static_cast<float *>(c.gemm_ws);
        P.dBlo = P.dAlo + M * K;
    }
    // everything enqueued below must come after whatever the caller already has on the compute stream
    NB_CUDA(cudaEventRecord(P.ev_b, c.stream));
    NB_CUDA(cudaStreamWaitEvent(P.s_in, P.ev_b, 0));
    NB_CUDA(cudaStreamWaitEvent(P.s_out, P.ev_b, 0));
    NB_CUDA(cudaMemcpyAsync(P.dB, B_host, (size_t)K * N * 4, cudaMemcpyHostToDevice, P.s_in));
    NB_CUDA(cudaEventRecord(P.ev_b, P.s_in));
    for (int64_t i = 0; i < nblk; i++) {
        const int64_t r0 = i * rb, rows = (r0 + rb <= M) ? rb : M - r0;
        NB_CUDA(cudaMemcpyAsync(P.dA + r0 * K, A_host + r0 * K, (size_t)rows * K * 4, cudaMemcpyHostToDevice, P.s_in));
        NB_CUDA(cudaEventRecord(P.ev_in[i], P.s_in));
    }
    NB_CUDA(cudaStreamWaitEvent(c.stream, P.ev_b, 0));
    if ((x3 || b3) && (rc = gemm_reset_nonfinite()) != NB200_OK) return rc;
    if (x3 && (rc = gemm_split_operand(P.dB, P.dBlo, K * N)) != NB200_OK) return rc;
    if (b3 && (rc = gemm_bf16_split(P.dB, bh, bl, K, N)) != NB200_OK) return rc;
    for (int64_t i = 0; i < nblk; i++) {
        const int64_t r0 = i * rb, rows = (r0 + rb <= M) ? rb : M - r0;
        NB_CUDA(cudaStreamWaitEvent(c.stream, P.ev_in[i], 0));
        if (x3 && (rc = gemm_split_operand(P.dA + r0 * K, P.dAlo + r0 * K, rows * K)) != NB200_OK) return rc;
        if (b3) {
            if ((rc = gemm_bf16_split(P.dA + r0 * K, ah + r0 * Kp, al + r0 * Kp, rows, K)) != NB200_OK) return rc;
            if ((rc = gemm_bf16_presplit(P.dC + r0 * N, ah + r0 * Kp, al + r0 * Kp, bh, bl, rows, N, K, N)) != NB200_OK) return rc;
        } else if ((rc = gemm_presplit(P.dC + r0 * N, P.dA + r0 * K, x3 ? P.dAlo + r0 * K : nullptr, P.dB, x3 ? P.dBlo : nullptr, rows, N, K,
                                K, N, N, precision)) != NB200_OK)
            return rc;
        NB_CUDA(cudaEventRecord(P.ev_done[i], c.stream));
        NB_CUDA(cudaStreamWaitEvent(P.s_out, P.ev_done[i], 0));
        NB_CUDA(cudaMemcpyAsync(C_host + r0 * N, P.dC + r0 * N, (size_t)rows * N * 4, cudaMemcpyDeviceToHost, P.s_out));
    }
    // rejoin: the compute stream (the one callers time / order on) completes only after the last D2H
    NB_CUDA(cudaEventRecord(P.ev_b, P.s_out));
    NB_CUDA(cudaStreamWaitEvent(c.stream, P.ev_b, 0));
    NB_CUDA(cudaStreamSynchronize(c.stream));
    return NB200_OK;
}
