// Host-operand entry points: the step either side of the hot path (SURVEY.md §8 f, N3).  The reference moves whole
// arrays with a pageable, blocking cudaMemcpy on the default stream (NDArray_ToGPU / NDArray_ToCPU,
// src/ndarray.c:1037-1093) and only then computes.  nb200_sgemm_host pipelines instead:
//   copy-in stream     : B (once), then A row blocks                                   H2D
//   two worker streams : per row block, in stream order: wait for the block's H2D event, lo-split, tcgen05 GEMM, D2H of its C rows
//                        (the D2H overlaps the next blocks' H2D: PCIe is full duplex; why not a separate copy-out stream: see below)
// nb200_sgemm_batched_host keeps copy-in / compute / copy-out streams (upload-bound: a chunk's download is half its upload).
#include "common.cuh"
#include <cstdlib>

namespace nb200 {
namespace {
struct Pipe {
    static constexpr int MAXW = 4;
    cudaStream_t s_in = nullptr, s_out = nullptr, s_w[MAXW] = {};   // s_w: worker streams (split + GEMM + D2H of a row block, in stream order)
    float *dA = nullptr, *dAlo = nullptr, *dB = nullptr, *dBlo = nullptr, *dC = nullptr;
    int64_t capA = 0, capB = 0, capC = 0;
    int device = -1;
    static constexpr int MAXB = 64;
    cudaEvent_t ev_in[MAXB], ev_done[MAXB], ev_b = nullptr, ev_bs = nullptr, ev_w[MAXW] = {};
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

int pipe_open(Pipe &P, int device) {
    if (P.device == device) return NB200_OK;
    NB_CUDA(cudaStreamCreateWithFlags(&P.s_in, cudaStreamNonBlocking));
    NB_CUDA(cudaStreamCreateWithFlags(&P.s_out, cudaStreamNonBlocking));
    for (int i = 0; i < Pipe::MAXW; i++) {
        NB_CUDA(cudaStreamCreateWithFlags(&P.s_w[i], cudaStreamNonBlocking));
        NB_CUDA(cudaEventCreateWithFlags(&P.ev_w[i], cudaEventDisableTiming));
    }
    for (int i = 0; i < Pipe::MAXB; i++) {
        NB_CUDA(cudaEventCreateWithFlags(&P.ev_in[i], cudaEventDisableTiming));
        NB_CUDA(cudaEventCreateWithFlags(&P.ev_done[i], cudaEventDisableTiming));
    }
    NB_CUDA(cudaEventCreateWithFlags(&P.ev_b, cudaEventDisableTiming));
    NB_CUDA(cudaEventCreateWithFlags(&P.ev_bs, cudaEventDisableTiming));
    P.device = device;
    return NB200_OK;
}

// D2H by the SMs: 16-byte stores into the device alias of a pinned (mapped) host buffer.  A few CTAs are enough to fill the link
// and they sit beside the persistent GEMM grid (256 threads, no shared memory).
__global__ void __launch_bounds__(256) store_to_host_kernel(float4 *__restrict__ dst, const float4 *__restrict__ src, int64_t n4, float *dst1, const float *src1, int64_t tail) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) dst[i] = src[i];
    if (blockIdx.x == 0 && threadIdx.x < tail) dst1[threadIdx.x] = src1[threadIdx.x];
}
}  // namespace

void host_pipeline_release(int device) {
    if (device < 0 || device >= NB200_MAX_DEVICES) return;
    Pipe &P = g_pipes[device];
    if (P.device < 0) return;
    if (P.s_in) cudaStreamDestroy(P.s_in);
    if (P.s_out) cudaStreamDestroy(P.s_out);
    for (int i = 0; i < Pipe::MAXW; i++) { if (P.s_w[i]) cudaStreamDestroy(P.s_w[i]); if (P.ev_w[i]) cudaEventDestroy(P.ev_w[i]); }
    if (P.ev_bs) cudaEventDestroy(P.ev_bs);
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
    if (precision < NB200_GEMM_TF32X3 || precision > NB200_GEMM_FP16X3U) return set_error(NB200_EINVAL, "unknown precision %d", precision);
    precision = gemm_resolve_precision(precision, K);
    if (precision == NB200_GEMM_FP16X3 || precision == NB200_GEMM_FP16X3U) precision = NB200_GEMM_TF32X3;   // this path is PCIe-bound; it keeps the two-kernel TF32x3 pipeline
    if (M == 0 || N == 0) return NB200_OK;
    Ctx &c = ctx();
    Pipe &P = g_pipes[c.device];
    { const int rco = pipe_open(P, c.device); if (rco != NB200_OK) return rco; }
    // shapes the tensor path cannot serve (tiny / K,N not multiples of 4): plain three-step path
    const bool pipelined = (K % 4 == 0) && (N % 4 == 0) && K >= 32 && N >= 32 && M >= 256 && M * N * K >= (int64_t)1 << 24;
    // row blocks: ~NB200_HOST_BLOCKS (default 16) of them, multiples of 256 rows (one CTA-pair tile), at most MAXB.  A short tail
    // (fewer than 128 rows, which the tensor path might refuse: tensor_path_ok wants rows * N * K >= 64^3) joins the previous block.
    static const int64_t want_blocks = getenv("NB200_HOST_BLOCKS") ? atoll(getenv("NB200_HOST_BLOCKS")) : 16;
    int64_t rb = M;
    if (pipelined) {
        const int64_t wb = want_blocks < 1 ? 1 : want_blocks;
        rb = ((M / wb + 255) / 256) * 256;
        if (rb < 256) rb = 256;
        while ((M + rb - 1) / rb > Pipe::MAXB) rb += 256;
    }
    int64_t nblk = (M + rb - 1) / rb;
    if (nblk > 1 && M - (nblk - 1) * rb < 128) nblk--;   // the last block then has rb + (M mod rb) rows
    auto blk_rows = [&](int64_t i) { return i == nblk - 1 ? M - i * rb : rb; };
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
    // NB200_HOST_TRACE=1: per-block timeline on stderr (timing events; diagnostics only)
    static const bool trace = getenv("NB200_HOST_TRACE") != nullptr;
    cudaEvent_t tr0 = nullptr, trB = nullptr, trIn[Pipe::MAXB], trDone[Pipe::MAXB], trOut[Pipe::MAXB];
    if (trace) {
        cudaEventCreate(&tr0); cudaEventCreate(&trB);
        for (int64_t i = 0; i < nblk; i++) { cudaEventCreate(&trIn[i]); cudaEventCreate(&trDone[i]); cudaEventCreate(&trOut[i]); }
        cudaEventRecord(tr0, c.stream);
    }
    // Stream structure (measured with scripts/probes/pcie_probe.cu: profiles/r2_pcie_probe.jsonl, profiles/r2_summary.md section 5d): with the obvious three streams
    // (copy-in, compute, copy-out, an event pair per block) the upload of A drops to ~36 GB/s as soon as C blocks go out, although
    // the same copies WITHOUT a kernel in the dependency chain keep 48-49 GB/s each way.  Giving every row block's split + GEMM +
    // D2H to one of two worker streams IN STREAM ORDER (the only cross-stream edge left per block is copy-in -> worker) keeps the
    // link at its duplex rate with 4 MiB blocks.  NB200_HOST_WORKERS=0 selects the three-stream structure (A/B experiments).
    static const int workers_env = getenv("NB200_HOST_WORKERS") ? atoi(getenv("NB200_HOST_WORKERS")) : 2;
    const int W = workers_env < 0 ? 0 : workers_env > Pipe::MAXW ? Pipe::MAXW : workers_env;
    // NB200_HOST_D2H_KERNEL=<CTAs>: C blocks leave through SM stores into the mapped host buffer instead of the copy engine (experiment)
    int d2h_ctas = getenv("NB200_HOST_D2H_KERNEL") ? atoi(getenv("NB200_HOST_D2H_KERNEL")) : 0;
    float *C_map = nullptr;
    if (d2h_ctas > 0 && (cudaHostGetDevicePointer(reinterpret_cast<void **>(&C_map), C_host, 0) != cudaSuccess || !C_map)) {
        cudaGetLastError();
        d2h_ctas = 0;   // not pinned / not mapped: copy engine
    }
    // everything enqueued below must come after whatever the caller already has on the compute stream
    cudaStream_t const s_main = c.stream;
    NB_CUDA(cudaEventRecord(P.ev_b, s_main));
    NB_CUDA(cudaStreamWaitEvent(P.s_in, P.ev_b, 0));
    NB_CUDA(cudaStreamWaitEvent(P.s_out, P.ev_b, 0));
    NB_CUDA(cudaMemcpyAsync(P.dB, B_host, (size_t)K * N * 4, cudaMemcpyHostToDevice, P.s_in));
    NB_CUDA(cudaEventRecord(P.ev_b, P.s_in));
    if (trace) cudaEventRecord(trB, P.s_in);
    for (int64_t i = 0; i < nblk; i++) {
        const int64_t r0 = i * rb, rows = blk_rows(i);
        NB_CUDA(cudaMemcpyAsync(P.dA + r0 * K, A_host + r0 * K, (size_t)rows * K * 4, cudaMemcpyHostToDevice, P.s_in));
        NB_CUDA(cudaEventRecord(P.ev_in[i], P.s_in));
        if (trace) cudaEventRecord(trIn[i], P.s_in);
    }
    NB_CUDA(cudaStreamWaitEvent(s_main, P.ev_b, 0));
    if ((x3 || b3) && (rc = gemm_reset_nonfinite()) != NB200_OK) return rc;
    if (x3 && (rc = gemm_split_operand(P.dB, P.dBlo, K * N)) != NB200_OK) return rc;
    if (b3 && (rc = gemm_bf16_split(P.dB, bh, bl, K, N)) != NB200_OK) return rc;
    if (W > 0) {
        NB_CUDA(cudaEventRecord(P.ev_bs, s_main));   // B and its lo part are ready
        for (int w = 0; w < W; w++) NB_CUDA(cudaStreamWaitEvent(P.s_w[w], P.ev_bs, 0));
    }
    for (int64_t i = 0; i < nblk && rc == NB200_OK; i++) {
        const int64_t r0 = i * rb, rows = blk_rows(i);
        cudaStream_t sx = W > 0 ? P.s_w[i % W] : s_main;
        c.stream = sx;   // the GEMM launchers enqueue on the context's stream
        if (cudaStreamWaitEvent(sx, P.ev_in[i], 0) != cudaSuccess) { rc = set_error(NB200_ECUDA, "nb200_sgemm_host: cudaStreamWaitEvent failed"); break; }
        if (x3) rc = gemm_split_operand(P.dA + r0 * K, P.dAlo + r0 * K, rows * K);
        if (rc == NB200_OK && b3) {
            rc = gemm_bf16_split(P.dA + r0 * K, ah + r0 * Kp, al + r0 * Kp, rows, K);
            if (rc == NB200_OK) rc = gemm_bf16_presplit(P.dC + r0 * N, ah + r0 * Kp, al + r0 * Kp, bh, bl, rows, N, K, N);
        } else if (rc == NB200_OK) {
            rc = gemm_presplit(P.dC + r0 * N, P.dA + r0 * K, x3 ? P.dAlo + r0 * K : nullptr, P.dB, x3 ? P.dBlo : nullptr, rows, N, K, K, N, N, precision);
        }
        if (rc != NB200_OK) break;
        if (trace) cudaEventRecord(trDone[i], sx);
        cudaStream_t so = sx;
        if (W == 0) {
            cudaEventRecord(P.ev_done[i], sx);
            cudaStreamWaitEvent(P.s_out, P.ev_done[i], 0);
            so = P.s_out;
        }
        if (d2h_ctas > 0) {
            const int64_t n = rows * N, n4 = n / 4;   // (N % 4 == 0 on this path; pinned allocations are at least 16-byte aligned)
            store_to_host_kernel<<<d2h_ctas, 256, 0, so>>>(reinterpret_cast<float4 *>(C_map + r0 * N), reinterpret_cast<const float4 *>(P.dC + r0 * N), n4,
                                                           C_map + r0 * N + n4 * 4, P.dC + r0 * N + n4 * 4, n - n4 * 4);
            ctx().launches++;
        } else if (cudaMemcpyAsync(C_host + r0 * N, P.dC + r0 * N, (size_t)rows * N * 4, cudaMemcpyDeviceToHost, so) != cudaSuccess) {
            rc = set_error(NB200_ECUDA, "nb200_sgemm_host: D2H enqueue failed");
        }
        if (trace) cudaEventRecord(trOut[i], so);
    }
    c.stream = s_main;
    // rejoin: the compute stream (the one callers time / order on) completes only after the last D2H - also on the error path,
    // where copies are still in flight
    if (W > 0) {
        for (int w = 0; w < W; w++) { cudaEventRecord(P.ev_w[w], P.s_w[w]); cudaStreamWaitEvent(s_main, P.ev_w[w], 0); }
    } else {
        cudaEventRecord(P.ev_b, P.s_out);
        cudaStreamWaitEvent(s_main, P.ev_b, 0);
    }
    cudaEventRecord(P.ev_b, P.s_in);
    cudaStreamWaitEvent(s_main, P.ev_b, 0);
    NB_CUDA(cudaStreamSynchronize(s_main));
    if (rc != NB200_OK) return rc;
    if (trace) {
        float t;
        cudaEventElapsedTime(&t, tr0, trB);
        fprintf(stderr, "[nb200_sgemm_host] B in %.3f ms;", t);
        for (int64_t i = 0; i < nblk; i++) {
            float a, d, o;
            cudaEventElapsedTime(&a, tr0, trIn[i]); cudaEventElapsedTime(&d, tr0, trDone[i]); cudaEventElapsedTime(&o, tr0, trOut[i]);
            fprintf(stderr, " blk%lld in %.3f done %.3f out %.3f;", (long long)i, a, d, o);
            cudaEventDestroy(trIn[i]); cudaEventDestroy(trDone[i]); cudaEventDestroy(trOut[i]);
        }
        fprintf(stderr, "\n");
        cudaEventDestroy(tr0); cudaEventDestroy(trB);
    }
    return NB200_OK;
}

// Batch of independent products with HOST operands (config #5's per-GPU share fed from host memory): chunks of matrices move
// through the same three streams - chunk i+1 uploads (A block, B block) while chunk i is multiplied (nb200_sgemm_batched, any
// precision) and chunk i-1 downloads.  Replaces `$a->gpu(); $b->gpu(); nd::matmul per matrix; ->cpu()` (NDArray_ToGPU / ToCPU,
// ndarray.c:1037-1093: blocking pageable copies, then compute).  Blocking; C_host complete on return.
extern "C" int nb200_sgemm_batched_host(float *C_host, const float *A_host, const float *B_host, int64_t batch, int64_t M, int64_t N, int64_t K,
                                        int precision) {
    NB_READY();
    if (!C_host || !A_host || !B_host || batch < 0 || M < 0 || N < 0 || K < 0) return set_error(NB200_EINVAL, "nb200_sgemm_batched_host: bad argument");
    if (precision < NB200_GEMM_TF32X3 || precision > NB200_GEMM_FP16X3U) return set_error(NB200_EINVAL, "unknown precision %d", precision);
    if (batch == 0 || M == 0 || N == 0) return NB200_OK;
    Ctx &c = ctx();
    Pipe &P = g_pipes[c.device];
    { const int rco = pipe_open(P, c.device); if (rco != NB200_OK) return rco; }
    int rc;
    if ((rc = grow(&P.dA, &P.capA, batch * M * K)) != NB200_OK) return rc;
    if ((rc = grow(&P.dB, &P.capB, batch * K * N)) != NB200_OK) return rc;
    if ((rc = grow(&P.dC, &P.capC, batch * M * N)) != NB200_OK) return rc;
    // ~16 chunks (at least one matrix each, at most MAXB chunks): small enough that the first product starts early and the last
    // download is short (measured: 16 x 2048^2 in 11.6 ms with 16 or 4 chunks, the H2D floor being 9.7 ms)
    static const int64_t want_chunks = getenv("NB200_HOST_BLOCKS") ? atoll(getenv("NB200_HOST_BLOCKS")) : 16;
    int64_t per = (batch + (want_chunks < 1 ? 1 : want_chunks) - 1) / (want_chunks < 1 ? 1 : want_chunks);
    if (per < 1) per = 1;
    while ((batch + per - 1) / per > Pipe::MAXB) per++;
    const int64_t nchunk = (batch + per - 1) / per;
    static const bool trace = getenv("NB200_HOST_TRACE") != nullptr;   // per-chunk timeline on stderr (diagnostics only)
    cudaEvent_t tr0 = nullptr, trIn[Pipe::MAXB], trDone[Pipe::MAXB], trOut[Pipe::MAXB];
    if (trace) {
        cudaEventCreate(&tr0);
        for (int64_t i = 0; i < nchunk; i++) { cudaEventCreate(&trIn[i]); cudaEventCreate(&trDone[i]); cudaEventCreate(&trOut[i]); }
        cudaEventRecord(tr0, c.stream);
    }
    NB_CUDA(cudaEventRecord(P.ev_b, c.stream));
    NB_CUDA(cudaStreamWaitEvent(P.s_in, P.ev_b, 0));
    NB_CUDA(cudaStreamWaitEvent(P.s_out, P.ev_b, 0));
    for (int64_t i = 0; i < nchunk; i++) {
        const int64_t b0 = i * per, nb = b0 + per <= batch ? per : batch - b0;
        NB_CUDA(cudaMemcpyAsync(P.dA + b0 * M * K, A_host + b0 * M * K, (size_t)(nb * M * K) * 4, cudaMemcpyHostToDevice, P.s_in));
        NB_CUDA(cudaMemcpyAsync(P.dB + b0 * K * N, B_host + b0 * K * N, (size_t)(nb * K * N) * 4, cudaMemcpyHostToDevice, P.s_in));
        NB_CUDA(cudaEventRecord(P.ev_in[i], P.s_in));
        if (trace) cudaEventRecord(trIn[i], P.s_in);
    }
    for (int64_t i = 0; i < nchunk; i++) {
        const int64_t b0 = i * per, nb = b0 + per <= batch ? per : batch - b0;
        NB_CUDA(cudaStreamWaitEvent(c.stream, P.ev_in[i], 0));
        if ((rc = nb200_sgemm_batched(P.dC + b0 * M * N, P.dA + b0 * M * K, P.dB + b0 * K * N, nb, M, N, K, M * K, K * N, M * N, precision)) != NB200_OK)
            return rc;
        NB_CUDA(cudaEventRecord(P.ev_done[i], c.stream));
        if (trace) cudaEventRecord(trDone[i], c.stream);
        // (a chunk's download in stream order behind its product, as in nb200_sgemm_host, measures the same here: 11.34 vs 11.36 ms
        // for 16 x 2048^2 - a chunk's upload is twice its download, the link is upload-bound either way)
        NB_CUDA(cudaStreamWaitEvent(P.s_out, P.ev_done[i], 0));
        NB_CUDA(cudaMemcpyAsync(C_host + b0 * M * N, P.dC + b0 * M * N, (size_t)(nb * M * N) * 4, cudaMemcpyDeviceToHost, P.s_out));
        if (trace) cudaEventRecord(trOut[i], P.s_out);
    }
    NB_CUDA(cudaEventRecord(P.ev_b, P.s_out));
    NB_CUDA(cudaStreamWaitEvent(c.stream, P.ev_b, 0));
    NB_CUDA(cudaStreamSynchronize(c.stream));
    if (trace) {
        fprintf(stderr, "[nb200_sgemm_batched_host]");
        for (int64_t i = 0; i < nchunk; i++) {
            float a, d, o;
            cudaEventElapsedTime(&a, tr0, trIn[i]); cudaEventElapsedTime(&d, tr0, trDone[i]); cudaEventElapsedTime(&o, tr0, trOut[i]);
            fprintf(stderr, " chunk%lld in %.3f done %.3f out %.3f;", (long long)i, a, d, o);
            cudaEventDestroy(trIn[i]); cudaEventDestroy(trDone[i]); cudaEventDestroy(trOut[i]);
        }
        fprintf(stderr, "\n");
        cudaEventDestroy(tr0);
    }
    return NB200_OK;
}
