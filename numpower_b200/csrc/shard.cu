// Multi-GPU shards of the NDArray hot path (SURVEY.md §8 e): ONE host process drives the G GPUs of a box, the way a PHP
// process would after NDArray::setDevice (numpower.c:615-635 is the reference's only multi-GPU affordance: it has no collective
// and no sharded op at all).  An array is split along its first axis into G contiguous row shards, shard g resident on device g.
//   * resident shards: elementwise ops / batched matmul are independent per unit -> one launch per device on that device's context
//     stream, no data-path collective; full reductions combine G partials on the host in fixed shard order (deterministic) with
//     the reference's NaN / tie rules applied to the GLOBAL index space.
//   * an array that lives on one device: nb200_shard_scatter / nb200_shard_gather move the row shards over NVLink, either as ONE
//     grouped ncclSend/ncclRecv (single-process communicators from ncclCommInitAll) or as cudaMemcpyPeerAsync copies (one stream
//     per peer, copy engines) - the A/B the survey asks for.  nb200_sgemm_batched_scatter_gather pipelines the three steps per
//     chunk of matrices: while chunk i is multiplied, chunk i+1 lands and the product of chunk i-1 returns (root link full duplex).
// NCCL is resolved at run time (dlopen "libnccl.so.2": the process-wide copy, e.g. the one torch has loaded), so libnb200.so has no
// link-time dependency on it and single-GPU users never touch it.
#include "common.cuh"
#include <dlfcn.h>
#include <nccl.h>
#include <vector>

namespace nb200 {
namespace {

struct NcclApi {
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int *) = nullptr;
    void *handle = nullptr;
};

constexpr int MAXG = NB200_MAX_DEVICES;
struct Shards {
    int n = 0;
    int dev[MAXG];
    cudaStream_t xin[MAXG], xout[MAXG];      // transfer streams of each device (towards it / away from it)
    cudaEvent_t ev_a[MAXG], ev_b[MAXG];      // ordering between the context streams and the transfer streams
    ncclComm_t comm[MAXG];
    bool nccl_ready = false;
    NcclApi api;
} g;

#define NB_NCCL(expr)                                                                                        \
    do {                                                                                                     \
        ncclResult_t _r = (expr);                                                                            \
        if (_r != ncclSuccess)                                                                               \
            return set_error(NB200_ECUDA, "%s failed: %s (%s:%d)", #expr, g.api.GetErrorString ? g.api.GetErrorString(_r) : "?", \
                             __FILE__, __LINE__);                                                            \
    } while (0)

int load_nccl() {
    if (g.api.handle) return NB200_OK;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return set_error(NB200_ECUDA, "NCCL transport requested but libnccl.so.2 cannot be loaded (%s)", dlerror());
    NcclApi &a = g.api;
    a.CommInitAll = reinterpret_cast<decltype(a.CommInitAll)>(dlsym(h, "ncclCommInitAll"));
    a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
    a.GroupStart = reinterpret_cast<decltype(a.GroupStart)>(dlsym(h, "ncclGroupStart"));
    a.GroupEnd = reinterpret_cast<decltype(a.GroupEnd)>(dlsym(h, "ncclGroupEnd"));
    a.Send = reinterpret_cast<decltype(a.Send)>(dlsym(h, "ncclSend"));
    a.Recv = reinterpret_cast<decltype(a.Recv)>(dlsym(h, "ncclRecv"));
    a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
    a.GetVersion = reinterpret_cast<decltype(a.GetVersion)>(dlsym(h, "ncclGetVersion"));
    if (!a.CommInitAll || !a.CommDestroy || !a.GroupStart || !a.GroupEnd || !a.Send || !a.Recv || !a.GetErrorString)
        return set_error(NB200_ECUDA, "libnccl.so.2 lacks a required symbol");
    a.handle = h;
    return NB200_OK;
}

int ensure_nccl() {
    if (g.nccl_ready) return NB200_OK;
    int rc = load_nccl();
    if (rc != NB200_OK) return rc;
    NB_NCCL(g.api.CommInitAll(g.comm, g.n, g.dev));
    g.nccl_ready = true;
    return NB200_OK;
}

// RAII: the entry points below hop between device contexts and leave the caller's current device as they found it
struct DeviceScope {
    int saved;
    DeviceScope() : saved(ctx().device) {}
    ~DeviceScope() { if (saved >= 0) nb200_set_device(saved); }
};

inline void shard_range(int64_t units, int G, int s, int64_t *lo, int64_t *cnt) {
    const int64_t base = units / G, rem = units % G;   // the first `rem` shards get one extra unit (same rule as sharding.py)
    *lo = s * base + (s < rem ? s : rem);
    *cnt = base + (s < rem ? 1 : 0);
}

int check_init(const char *who) {
    if (g.n < 1) return set_error(NB200_EINVAL, "%s: call nb200_shard_init first", who);
    return NB200_OK;
}

}  // namespace
}  // namespace nb200

using namespace nb200;

extern "C" int nb200_shard_init(int ndev, const int *devices) {
    if (ndev < 1 || ndev > MAXG) return set_error(NB200_EINVAL, "nb200_shard_init: ndev %d out of range [1,%d]", ndev, MAXG);
    if (g.n) nb200_shard_finalize();
    int saved = -1;
    if (ctx().ready) saved = ctx().device;
    for (int s = 0; s < ndev; s++) {
        const int d = devices ? devices[s] : s;
        for (int q = 0; q < s; q++)
            if (g.dev[q] == d) return set_error(NB200_EINVAL, "nb200_shard_init: device %d listed twice", d);
        g.dev[s] = d;
        int rc = nb200_set_device(d);
        if (rc != NB200_OK) return rc;
        NB_CUDA(cudaStreamCreateWithFlags(&g.xin[s], cudaStreamNonBlocking));
        NB_CUDA(cudaStreamCreateWithFlags(&g.xout[s], cudaStreamNonBlocking));
        NB_CUDA(cudaEventCreateWithFlags(&g.ev_a[s], cudaEventDisableTiming));
        NB_CUDA(cudaEventCreateWithFlags(&g.ev_b[s], cudaEventDisableTiming));
    }
    g.n = ndev;
    // peer access for the cudaMemcpyPeerAsync transport (NVLink / NVSwitch: every pair)
    for (int s = 0; s < ndev; s++) {
        NB_CUDA(cudaSetDevice(g.dev[s]));
        for (int q = 0; q < ndev; q++) {
            if (q == s) continue;
            int can = 0;
            NB_CUDA(cudaDeviceCanAccessPeer(&can, g.dev[s], g.dev[q]));
            if (!can) continue;
            cudaError_t e = cudaDeviceEnablePeerAccess(g.dev[q], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) NB_CUDA(e);
            cudaGetLastError();
        }
    }
    return nb200_set_device(saved >= 0 ? saved : g.dev[0]);
}

extern "C" int nb200_shard_finalize(void) {
    if (!g.n) return NB200_OK;
    for (int s = 0; s < g.n; s++) {
        cudaSetDevice(g.dev[s]);
        cudaStreamSynchronize(g.xin[s]);
        cudaStreamSynchronize(g.xout[s]);
        if (g.nccl_ready) g.api.CommDestroy(g.comm[s]);
        cudaStreamDestroy(g.xin[s]);
        cudaStreamDestroy(g.xout[s]);
        cudaEventDestroy(g.ev_a[s]);
        cudaEventDestroy(g.ev_b[s]);
    }
    g.nccl_ready = false;
    g.n = 0;
    if (ctx().ready) cudaSetDevice(ctx().device);
    return NB200_OK;
}

extern "C" int nb200_shard_count(int *ndev) {
    if (!ndev) return set_error(NB200_EINVAL, "null argument");
    *ndev = g.n;
    return NB200_OK;
}

extern "C" int nb200_shard_device(int shard, int *device) {
    if (!device || shard < 0 || shard >= g.n) return set_error(NB200_EINVAL, "nb200_shard_device: bad argument");
    *device = g.dev[shard];
    return NB200_OK;
}

extern "C" int nb200_shard_split(int64_t units, int nshards, int shard, int64_t *first, int64_t *count) {
    if (units < 0 || nshards < 1 || shard < 0 || shard >= nshards || !first || !count) return set_error(NB200_EINVAL, "nb200_shard_split: bad argument");
    shard_range(units, nshards, shard, first, count);
    return NB200_OK;
}

extern "C" int nb200_shard_range(int64_t units, int shard, int64_t *first, int64_t *count) {
    int rc = check_init("nb200_shard_range");
    if (rc != NB200_OK) return rc;
    if (units < 0 || shard < 0 || shard >= g.n || !first || !count) return set_error(NB200_EINVAL, "nb200_shard_range: bad argument");
    shard_range(units, g.n, shard, first, count);
    return NB200_OK;
}

extern "C" int nb200_shard_synchronize(void) {
    int rc = check_init("nb200_shard_synchronize");
    if (rc != NB200_OK) return rc;
    DeviceScope scope;
    for (int s = 0; s < g.n; s++) {
        NB_CUDA(cudaSetDevice(g.dev[s]));
        NB_CUDA(cudaStreamSynchronize(g.xin[s]));
        NB_CUDA(cudaStreamSynchronize(g.xout[s]));
        NB_CUDA(cudaStreamSynchronize(ctx_of(g.dev[s])->stream));
    }
    return NB200_OK;
}

// ---- scatter / gather ------------------------------------------------------------------------------------------------------------
// Stream order: the transfers start after everything already enqueued on the root's context stream (scatter) / on every shard's
// context stream (gather), and each destination's context stream continues only after its data has landed.
extern "C" int nb200_shard_scatter(float *const *shard_ptrs, const float *root_src, int64_t rows, int64_t row_elems, int root, int transport) {
    int rc = check_init("nb200_shard_scatter");
    if (rc != NB200_OK) return rc;
    if (!shard_ptrs || !root_src || rows < 0 || row_elems < 0 || root < 0 || root >= g.n) return set_error(NB200_EINVAL, "nb200_shard_scatter: bad argument");
    if (transport == NB200_XFER_NCCL && (rc = ensure_nccl()) != NB200_OK) return rc;
    DeviceScope scope;
    Ctx *rc_ctx = ctx_of(g.dev[root]);
    NB_CUDA(cudaSetDevice(g.dev[root]));
    NB_CUDA(cudaEventRecord(g.ev_a[root], rc_ctx->stream));
    NB_CUDA(cudaStreamWaitEvent(g.xout[root], g.ev_a[root], 0));
    for (int s = 0; s < g.n; s++) {   // every destination's inbound stream starts after the root's data is ready
        if (s == root) continue;
        NB_CUDA(cudaSetDevice(g.dev[s]));
        NB_CUDA(cudaStreamWaitEvent(g.xin[s], g.ev_a[root], 0));
    }
    if (transport == NB200_XFER_NCCL) NB_NCCL(g.api.GroupStart());
    for (int s = 0; s < g.n; s++) {
        int64_t lo, cnt;
        shard_range(rows, g.n, s, &lo, &cnt);
        const int64_t bytes = cnt * row_elems * 4;
        const float *src = root_src + lo * row_elems;
        if (bytes == 0 || !shard_ptrs[s]) continue;
        if (s == root) {
            if (shard_ptrs[s] != src) {
                NB_CUDA(cudaSetDevice(g.dev[root]));
                NB_CUDA(cudaMemcpyAsync(shard_ptrs[s], src, (size_t)bytes, cudaMemcpyDeviceToDevice, rc_ctx->stream));
            }
        } else if (transport == NB200_XFER_P2P) {
            NB_CUDA(cudaSetDevice(g.dev[s]));
            NB_CUDA(cudaMemcpyPeerAsync(shard_ptrs[s], g.dev[s], src, g.dev[root], (size_t)bytes, g.xin[s]));
        } else {
            NB_NCCL(g.api.Send(src, (size_t)bytes, ncclChar, s, g.comm[root], g.xout[root]));
            NB_NCCL(g.api.Recv(shard_ptrs[s], (size_t)bytes, ncclChar, root, g.comm[s], g.xin[s]));
        }
    }
    if (transport == NB200_XFER_NCCL) NB_NCCL(g.api.GroupEnd());
    for (int s = 0; s < g.n; s++) {
        if (s == root) continue;
        NB_CUDA(cudaSetDevice(g.dev[s]));
        NB_CUDA(cudaEventRecord(g.ev_b[s], g.xin[s]));
        NB_CUDA(cudaStreamWaitEvent(ctx_of(g.dev[s])->stream, g.ev_b[s], 0));
    }
    // the root's context stream must not overwrite root_src before the sends have read it
    NB_CUDA(cudaSetDevice(g.dev[root]));
    if (transport == NB200_XFER_NCCL) {
        NB_CUDA(cudaEventRecord(g.ev_b[root], g.xout[root]));
        NB_CUDA(cudaStreamWaitEvent(rc_ctx->stream, g.ev_b[root], 0));
    } else {
        for (int s = 0; s < g.n; s++)
            if (s != root) NB_CUDA(cudaStreamWaitEvent(rc_ctx->stream, g.ev_b[s], 0));
    }
    return NB200_OK;
}

// ---- sharded residency: `$a->gpu()` / `->cpu()` of an array that lives as row shards on every GPU --------------------------------
// The reference moves an array with ONE blocking cudaMemcpy to / from the current device (NDArray_ToGPU / NDArray_ToCPU,
// src/ndarray.c:1037-1093).  Here every shard travels over its OWN device's PCIe link, all links at once: G copies are enqueued
// on the G context streams (asynchronous for pinned host memory, nb200_host_alloc), so the aggregate host<->device rate is up to G
// times one link's.  Ordered with the other work on each context stream; nb200_shard_synchronize() (or any blocking call) makes
// the host buffer reusable / the downloaded data visible.
extern "C" int nb200_shard_upload(float *const *shard_ptrs, const float *host_src, int64_t rows, int64_t row_elems) {
    int rc = check_init("nb200_shard_upload");
    if (rc != NB200_OK) return rc;
    if (!shard_ptrs || !host_src || rows < 0 || row_elems < 0) return set_error(NB200_EINVAL, "nb200_shard_upload: bad argument");
    DeviceScope scope;
    for (int s = 0; s < g.n; s++) {
        int64_t lo, cnt;
        shard_range(rows, g.n, s, &lo, &cnt);
        const int64_t bytes = cnt * row_elems * 4;
        if (bytes == 0) continue;
        if (!shard_ptrs[s]) return set_error(NB200_EINVAL, "nb200_shard_upload: shard %d has %lld rows but no buffer", s, (long long)cnt);
        NB_CUDA(cudaSetDevice(g.dev[s]));
        NB_CUDA(cudaMemcpyAsync(shard_ptrs[s], host_src + lo * row_elems, (size_t)bytes, cudaMemcpyHostToDevice, ctx_of(g.dev[s])->stream));
    }
    return NB200_OK;
}

extern "C" int nb200_shard_download(float *host_dst, const float *const *shard_ptrs, int64_t rows, int64_t row_elems) {
    int rc = check_init("nb200_shard_download");
    if (rc != NB200_OK) return rc;
    if (!shard_ptrs || !host_dst || rows < 0 || row_elems < 0) return set_error(NB200_EINVAL, "nb200_shard_download: bad argument");
    DeviceScope scope;
    for (int s = 0; s < g.n; s++) {
        int64_t lo, cnt;
        shard_range(rows, g.n, s, &lo, &cnt);
        const int64_t bytes = cnt * row_elems * 4;
        if (bytes == 0) continue;
        if (!shard_ptrs[s]) return set_error(NB200_EINVAL, "nb200_shard_download: shard %d has %lld rows but no buffer", s, (long long)cnt);
        NB_CUDA(cudaSetDevice(g.dev[s]));
        NB_CUDA(cudaMemcpyAsync(host_dst + lo * row_elems, shard_ptrs[s], (size_t)bytes, cudaMemcpyDeviceToHost, ctx_of(g.dev[s])->stream));
    }
    return NB200_OK;
}

extern "C" int nb200_shard_gather(float *root_dst, const float *const *shard_ptrs, int64_t rows, int64_t row_elems, int root, int transport) {
    int rc = check_init("nb200_shard_gather");
    if (rc != NB200_OK) return rc;
    if (!shard_ptrs || !root_dst || rows < 0 || row_elems < 0 || root < 0 || root >= g.n) return set_error(NB200_EINVAL, "nb200_shard_gather: bad argument");
    if (transport == NB200_XFER_NCCL && (rc = ensure_nccl()) != NB200_OK) return rc;
    DeviceScope scope;
    Ctx *rc_ctx = ctx_of(g.dev[root]);
    // every source's outbound stream starts after that shard's compute; the root's inbound stream after the root's own work
    for (int s = 0; s < g.n; s++) {
        NB_CUDA(cudaSetDevice(g.dev[s]));
        NB_CUDA(cudaEventRecord(g.ev_a[s], ctx_of(g.dev[s])->stream));
        NB_CUDA(cudaStreamWaitEvent(s == root ? g.xin[root] : g.xout[s], g.ev_a[s], 0));
    }
    // (peer copies run on each SOURCE's outbound stream, so the blocks of different shards travel concurrently; they must not start
    //  before the root has finished with root_dst)
    if (transport == NB200_XFER_P2P)
        for (int s = 0; s < g.n; s++) {
            if (s == root) continue;
            NB_CUDA(cudaSetDevice(g.dev[s]));
            NB_CUDA(cudaStreamWaitEvent(g.xout[s], g.ev_a[root], 0));
        }
    if (transport == NB200_XFER_NCCL) NB_NCCL(g.api.GroupStart());
    for (int s = 0; s < g.n; s++) {
        int64_t lo, cnt;
        shard_range(rows, g.n, s, &lo, &cnt);
        const int64_t bytes = cnt * row_elems * 4;
        float *dst = root_dst + lo * row_elems;
        if (bytes == 0 || !shard_ptrs[s]) continue;
        if (s == root) {
            if (shard_ptrs[s] != dst) {
                NB_CUDA(cudaSetDevice(g.dev[root]));
                NB_CUDA(cudaMemcpyAsync(dst, shard_ptrs[s], (size_t)bytes, cudaMemcpyDeviceToDevice, rc_ctx->stream));
            }
        } else if (transport == NB200_XFER_P2P) {
            NB_CUDA(cudaSetDevice(g.dev[s]));
            NB_CUDA(cudaMemcpyPeerAsync(dst, g.dev[root], shard_ptrs[s], g.dev[s], (size_t)bytes, g.xout[s]));
        } else {
            NB_NCCL(g.api.Send(shard_ptrs[s], (size_t)bytes, ncclChar, root, g.comm[s], g.xout[s]));
            NB_NCCL(g.api.Recv(dst, (size_t)bytes, ncclChar, s, g.comm[root], g.xin[root]));
        }
    }
    if (transport == NB200_XFER_NCCL) NB_NCCL(g.api.GroupEnd());
    // the root's context stream continues after every block has arrived; a source shard must not be overwritten by later work on
    // its own context stream before it has been sent
    if (transport == NB200_XFER_NCCL) {
        NB_CUDA(cudaSetDevice(g.dev[root]));
        NB_CUDA(cudaEventRecord(g.ev_b[root], g.xin[root]));
        NB_CUDA(cudaStreamWaitEvent(rc_ctx->stream, g.ev_b[root], 0));
    }
    for (int s = 0; s < g.n; s++) {
        if (s == root) continue;
        NB_CUDA(cudaSetDevice(g.dev[s]));
        NB_CUDA(cudaEventRecord(g.ev_b[s], g.xout[s]));
        NB_CUDA(cudaStreamWaitEvent(ctx_of(g.dev[s])->stream, g.ev_b[s], 0));
        if (transport == NB200_XFER_P2P) {
            NB_CUDA(cudaSetDevice(g.dev[root]));
            NB_CUDA(cudaStreamWaitEvent(rc_ctx->stream, g.ev_b[s], 0));
        }
    }
    return NB200_OK;
}

// ---- resident shards: elementwise -----------------------------------------------------------------------------------------------
extern "C" int nb200_shard_ew_binary(int op, float *const *out, const float *const *a, const float *const *b, int64_t rows, int64_t row_elems) {
    int rc = check_init("nb200_shard_ew_binary");
    if (rc != NB200_OK) return rc;
    if (!out || !a || !b || rows < 0 || row_elems < 0) return set_error(NB200_EINVAL, "nb200_shard_ew_binary: bad argument");
    DeviceScope scope;
    for (int s = 0; s < g.n; s++) {
        int64_t lo, cnt;
        shard_range(rows, g.n, s, &lo, &cnt);
        if (cnt == 0) continue;
        if ((rc = nb200_set_device(g.dev[s])) != NB200_OK) return rc;
        const int64_t n = cnt * row_elems, one = 1;
        if ((rc = nb200_ew_binary(op, out[s], a[s], b[s], 1, &n, &one, &one)) != NB200_OK) return rc;
    }
    return NB200_OK;
}

extern "C" int nb200_shard_ew_mul_add(float *const *out, const float *const *a, const float *const *b, const float *const *c, int64_t rows,
                                      int64_t row_elems) {
    int rc = check_init("nb200_shard_ew_mul_add");
    if (rc != NB200_OK) return rc;
    if (!out || !a || !b || !c || rows < 0 || row_elems < 0) return set_error(NB200_EINVAL, "nb200_shard_ew_mul_add: bad argument");
    DeviceScope scope;
    for (int s = 0; s < g.n; s++) {
        int64_t lo, cnt;
        shard_range(rows, g.n, s, &lo, &cnt);
        if (cnt == 0) continue;
        if ((rc = nb200_set_device(g.dev[s])) != NB200_OK) return rc;
        const int64_t n = cnt * row_elems, one = 1;
        if ((rc = nb200_ew_mul_add(out[s], a[s], b[s], c[s], 1, &n, &one, &one, &one)) != NB200_OK) return rc;
    }
    return NB200_OK;
}

extern "C" int nb200_shard_ew_unary(int op, float *const *out, const float *const *in, int64_t rows, int64_t row_elems, float p0, float p1) {
    int rc = check_init("nb200_shard_ew_unary");
    if (rc != NB200_OK) return rc;
    if (!out || !in || rows < 0 || row_elems < 0) return set_error(NB200_EINVAL, "nb200_shard_ew_unary: bad argument");
    DeviceScope scope;
    for (int s = 0; s < g.n; s++) {
        int64_t lo, cnt;
        shard_range(rows, g.n, s, &lo, &cnt);
        if (cnt == 0) continue;
        if ((rc = nb200_set_device(g.dev[s])) != NB200_OK) return rc;
        if ((rc = nb200_ew_unary(op, out[s], in[s], cnt * row_elems, p0, p1)) != NB200_OK) return rc;
    }
    return NB200_OK;
}

// ---- resident shards: full reductions -------------------------------------------------------------------------------------------
// Per-shard partials (launched on all devices first, then collected) folded on the host in shard order with fp32 arithmetic.
// min / max follow NDArray_Min/Max (ndarray.c:752-772, 939-959) over the GLOBAL array: a NaN sticks only at global index 0 and is
// skipped elsewhere.  The per-shard kernel applies that rule to ITS first element, so a NaN partial of a shard other than 0 means
// "this shard starts with NaN": that shard is reduced again without its leading NaNs.
extern "C" int nb200_shard_reduce_full(int op, float *host_out, const float *const *in, int64_t n_total) {
    int rc = check_init("nb200_shard_reduce_full");
    if (rc != NB200_OK) return rc;
    if (!host_out || !in || n_total <= 0) return set_error(NB200_EINVAL, "nb200_shard_reduce_full: bad argument");
    if (op < NB200_SUM || op > NB200_MAX) return set_error(NB200_EINVAL, "nb200_shard_reduce_full: unknown op %d", op);
    DeviceScope scope;
    int64_t lo[MAXG], cnt[MAXG];
    for (int s = 0; s < g.n; s++) {
        shard_range(n_total, g.n, s, &lo[s], &cnt[s]);
        if (cnt[s] == 0) continue;
        if ((rc = nb200_set_device(g.dev[s])) != NB200_OK) return rc;
        if ((rc = nb200_reduce_full(op, ctx().dev_result, in[s], cnt[s])) != NB200_OK) return rc;
        NB_CUDA(cudaMemcpyAsync(ctx().host_result, ctx().dev_result, sizeof(float), cudaMemcpyDeviceToHost, ctx().stream));
    }
    float acc = 0.f;
    bool have = false;
    for (int s = 0; s < g.n; s++) {
        if (cnt[s] == 0) continue;
        if ((rc = nb200_set_device(g.dev[s])) != NB200_OK) return rc;
        NB_CUDA(cudaStreamSynchronize(ctx().stream));
        float p = *ctx().host_result;
        bool skip = false;
        if ((op == NB200_MIN || op == NB200_MAX) && have && p != p) {
            // shard s > 0 starts with NaN: drop its leading NaNs (the reference skips them) and reduce the rest
            int64_t off = 0;
            while (p != p) {
                off++;
                if (off >= cnt[s]) { skip = true; break; }
                if ((rc = nb200_reduce_full_host(op, &p, in[s] + off, cnt[s] - off)) != NB200_OK) return rc;
            }
        }
        if (skip) continue;
        if (!have) { acc = p; have = true; continue; }
        switch (op) {
            case NB200_SUM: acc = acc + p; break;
            case NB200_PROD: acc = acc * p; break;
            case NB200_MIN: if (p < acc) acc = p; break;     // acc may be NaN (global element 0): comparisons are false, it sticks
            default: if (p > acc) acc = p; break;
        }
    }
    *host_out = acc;
    return NB200_OK;
}

namespace nb200 {
int argminmax_packed(int is_max, unsigned long long *dev_out, const float *in, int64_t n);   // reduce.cu
}

// argmax / argmin over the global index space (calculation.c:9-59): first occurrence; argmax skips NaN unless it is global element
// 0, argmin returns the first NaN.  Each shard returns its packed (ordering key, index) candidate - the same 64-bit word the kernel
// reduces internally, so the index stays exact beyond 2^24; the host takes the best key, lowest shard on ties (= lowest global
// index because shards are contiguous), and only then rounds to float like the reference's `(float)i`.
extern "C" int nb200_shard_argminmax(int is_max, float *host_out, const float *const *in, int64_t n_total) {
    int rc = check_init("nb200_shard_argminmax");
    if (rc != NB200_OK) return rc;
    if (!host_out || !in) return set_error(NB200_EINVAL, "nb200_shard_argminmax: bad argument");
    if (n_total <= 0) return set_error(NB200_EINVAL, "attempt to get %s of an empty sequence", is_max ? "argmax" : "argmin");
    DeviceScope scope;
    int64_t lo[MAXG], cnt[MAXG];
    float first_elem = 0.f;
    for (int s = 0; s < g.n; s++) {
        shard_range(n_total, g.n, s, &lo[s], &cnt[s]);
        if (cnt[s] == 0) continue;
        if ((rc = nb200_set_device(g.dev[s])) != NB200_OK) return rc;
        unsigned long long *slot = reinterpret_cast<unsigned long long *>(ctx().dev_result) + 2;   // bytes 16..23 of the 64-byte result slot
        if ((rc = argminmax_packed(is_max, slot, in[s], cnt[s])) != NB200_OK) return rc;
        NB_CUDA(cudaMemcpyAsync(reinterpret_cast<unsigned long long *>(ctx().host_result) + 2, slot, 8, cudaMemcpyDeviceToHost, ctx().stream));
        if (lo[s] == 0) NB_CUDA(cudaMemcpyAsync(ctx().host_result, in[s], 4, cudaMemcpyDeviceToHost, ctx().stream));
    }
    unsigned int best_key = 0;
    int64_t best_idx = -1;
    for (int s = 0; s < g.n; s++) {
        if (cnt[s] == 0) continue;
        if ((rc = nb200_set_device(g.dev[s])) != NB200_OK) return rc;
        NB_CUDA(cudaStreamSynchronize(ctx().stream));
        const unsigned long long packed = *(reinterpret_cast<unsigned long long *>(ctx().host_result) + 2);
        if (lo[s] == 0) first_elem = *ctx().host_result;
        const unsigned int key = (unsigned int)(packed >> 32), idx = 0xFFFFFFFFu - (unsigned int)(packed & 0xFFFFFFFFull);
        if (best_idx < 0 || key > best_key) { best_key = key; best_idx = lo[s] + (int64_t)idx; }
    }
    if (first_elem != first_elem) best_idx = 0;   // calculation.c:14-17, :41-44: a leading NaN wins outright
    *host_out = (float)best_idx;                  // the reference stores (float)i with int i: same rounding below 2^31, no wrap above
    return NB200_OK;
}

// ---- batched matmul -------------------------------------------------------------------------------------------------------------
// resident shards: shard s holds batch matrices [lo_s, lo_s + cnt_s) of A, B and C
extern "C" int nb200_sgemm_batched_sharded(float *const *C, const float *const *A, const float *const *B, int64_t batch, int64_t M, int64_t N,
                                           int64_t K, int precision) {
    int rc = check_init("nb200_sgemm_batched_sharded");
    if (rc != NB200_OK) return rc;
    if (!C || !A || !B || batch < 0 || M < 0 || N < 0 || K < 0) return set_error(NB200_EINVAL, "nb200_sgemm_batched_sharded: bad argument");
    DeviceScope scope;
    for (int s = 0; s < g.n; s++) {
        int64_t lo, cnt;
        shard_range(batch, g.n, s, &lo, &cnt);
        if (cnt == 0) continue;
        if ((rc = nb200_set_device(g.dev[s])) != NB200_OK) return rc;
        if ((rc = nb200_sgemm_batched(C[s], A[s], B[s], cnt, M, N, K, M * K, K * N, M * N, precision)) != NB200_OK) return rc;
    }
    return NB200_OK;
}

// Operands and result on ONE device (`root`): scatter + compute + gather, pipelined per chunk of `chunk` matrices per shard.
// Step t of the schedule enqueues, in this order:  transfer-in of chunk t (A and B blocks to every peer), the products of chunk t
// on every device (they wait for their inputs through events), transfer-out of chunk t (C blocks back to the root).  P2P: inbound
// and outbound copies use different streams per device (copy engines), so chunk t+1 lands and chunk t-1 returns while chunk t is
// multiplied.  NCCL: one group per step carries both directions (chunk t out, chunk t-2 back) and the GEMMs leave 16 SMs to the
// transfer kernels.  The root multiplies its own share in place, without any copy.
// elapsed_ms (optional): device time from the first transfer to the arrival of the last result block on the root (synchronises).
extern "C" int nb200_sgemm_batched_scatter_gather(float *C_root, const float *A_root, const float *B_root, int64_t batch, int64_t M, int64_t N,
                                                  int64_t K, int precision, int root, int transport, int64_t chunk, float *elapsed_ms) {
    int rc = check_init("nb200_sgemm_batched_scatter_gather");
    if (rc != NB200_OK) return rc;
    if (!C_root || !A_root || !B_root || batch < 0 || M <= 0 || N <= 0 || K <= 0 || root < 0 || root >= g.n)
        return set_error(NB200_EINVAL, "nb200_sgemm_batched_scatter_gather: bad argument");
    if (transport == NB200_XFER_NCCL && (rc = ensure_nccl()) != NB200_OK) return rc;
    DeviceScope scope;
    const int G = g.n;
    int64_t lo[MAXG], cnt[MAXG], max_cnt = 0;
    for (int s = 0; s < G; s++) { shard_range(batch, G, s, &lo[s], &cnt[s]); if (s != root && cnt[s] > max_cnt) max_cnt = cnt[s]; }
    if (chunk <= 0) chunk = 8;
    const int64_t steps = (max_cnt + chunk - 1) / chunk;
    // staging on the peers: whole shards of A, B and C (config #5: 3 x 2 GiB per device)
    float *sa[MAXG] = {}, *sb[MAXG] = {}, *sc[MAXG] = {};
    std::vector<cudaEvent_t> ev_in((size_t)G * (steps + 1)), ev_done((size_t)G * (steps + 1));
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    auto cleanup = [&]() {
        for (int s = 0; s < G; s++) {
            nb200_set_device(g.dev[s]);
            if (sa[s]) nb200_free(sa[s]);
            if (sb[s]) nb200_free(sb[s]);
            if (sc[s]) nb200_free(sc[s]);
            for (int64_t t = 0; t <= steps; t++) {
                if (ev_in[s * (steps + 1) + t]) cudaEventDestroy(ev_in[s * (steps + 1) + t]);
                if (ev_done[s * (steps + 1) + t]) cudaEventDestroy(ev_done[s * (steps + 1) + t]);
            }
        }
        if (t0) cudaEventDestroy(t0);
        if (t1) cudaEventDestroy(t1);
        for (int s = 0; s < G; s++) ctx_of(g.dev[s])->gemm_sm_reserve = 0;
    };
    for (auto &e : ev_in) e = nullptr;
    for (auto &e : ev_done) e = nullptr;
#define SG_TRY(expr) do { int _rc = (expr); if (_rc != NB200_OK) { cleanup(); return _rc; } } while (0)
#define SG_CUDA(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { cleanup(); return set_error(NB200_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); } } while (0)
#define SG_NCCL(expr) do { ncclResult_t _r = (expr); if (_r != ncclSuccess) { cleanup(); return set_error(NB200_ECUDA, "%s failed: %s (%s:%d)", #expr, g.api.GetErrorString(_r), __FILE__, __LINE__); } } while (0)
    for (int s = 0; s < G; s++) {
        SG_TRY(nb200_set_device(g.dev[s]));
        if (s != root && cnt[s] > 0) {
            SG_TRY(nb200_alloc(reinterpret_cast<void **>(&sa[s]), cnt[s] * M * K * 4));
            SG_TRY(nb200_alloc(reinterpret_cast<void **>(&sb[s]), cnt[s] * K * N * 4));
            SG_TRY(nb200_alloc(reinterpret_cast<void **>(&sc[s]), cnt[s] * M * N * 4));
        }
        for (int64_t t = 0; t <= steps; t++) {
            SG_CUDA(cudaEventCreateWithFlags(&ev_in[s * (steps + 1) + t], cudaEventDisableTiming));
            SG_CUDA(cudaEventCreateWithFlags(&ev_done[s * (steps + 1) + t], cudaEventDisableTiming));
        }
    }
    // NCCL's send / receive kernels need SMs of their own: a persistent GEMM grid that owns every SM would serialise them behind
    // the products (measured at N = 2: 25.7 ms against 8.0 ms with the copy-engine transport).  The GEMMs of this call leave 16 SMs free.
    const int reserve = transport == NB200_XFER_NCCL ? 16 : 0;
    for (int s = 0; s < G; s++) ctx_of(g.dev[s])->gemm_sm_reserve = reserve;
    Ctx *rctx = ctx_of(g.dev[root]);
    SG_CUDA(cudaSetDevice(g.dev[root]));
    SG_CUDA(cudaEventCreate(&t0));
    SG_CUDA(cudaEventCreate(&t1));
    // everything starts after what the caller already has on the root's context stream
    SG_CUDA(cudaEventRecord(t0, rctx->stream));
    SG_CUDA(cudaStreamWaitEvent(g.xout[root], t0, 0));
    SG_CUDA(cudaStreamWaitEvent(g.xin[root], t0, 0));
    for (int s = 0; s < G; s++) {
        if (s == root) continue;
        SG_CUDA(cudaSetDevice(g.dev[s]));
        SG_CUDA(cudaStreamWaitEvent(g.xin[s], t0, 0));
        SG_CUDA(cudaStreamWaitEvent(g.xout[s], t0, 0));
    }
    // the root's own share: one batched call on its context stream, concurrent with the transfers
    if (cnt[root] > 0) {
        SG_TRY(nb200_set_device(g.dev[root]));
        SG_TRY(nb200_sgemm_batched(C_root + lo[root] * M * N, A_root + lo[root] * M * K, B_root + lo[root] * K * N, cnt[root], M, N, K, M * K, K * N,
                                   M * N, precision));
    }
    if (transport == NB200_XFER_NCCL) {
        // NCCL orders the operations of one communicator even across streams, so both directions of a step go into ONE group (one
        // kernel per device: the root sends the A / B blocks of chunk t and receives the C blocks of chunk t-2 at the same time), all
        // on the devices' inbound transfer streams.  Chunk t-2 was multiplied while chunk t-1 travelled.
        for (int64_t t = 0; t < steps + 2; t++) {
            // the C blocks of chunk t-2 leave only after their products
            if (t >= 2)
                for (int s = 0; s < G; s++) {
                    if (s == root || cnt[s] - (t - 2) * chunk <= 0) continue;
                    SG_CUDA(cudaSetDevice(g.dev[s]));
                    SG_CUDA(cudaStreamWaitEvent(g.xin[s], ev_done[s * (steps + 1) + (t - 2)], 0));
                }
            SG_NCCL(g.api.GroupStart());
            for (int s = 0; s < G; s++) {
                if (s == root) continue;
                if (t < steps) {
                    const int64_t c0 = t * chunk, cn = cnt[s] - c0 < chunk ? cnt[s] - c0 : chunk;
                    if (cn > 0) {
                        SG_NCCL(g.api.Send(A_root + (lo[s] + c0) * M * K, (size_t)(cn * M * K * 4), ncclChar, s, g.comm[root], g.xin[root]));
                        SG_NCCL(g.api.Send(B_root + (lo[s] + c0) * K * N, (size_t)(cn * K * N * 4), ncclChar, s, g.comm[root], g.xin[root]));
                        SG_NCCL(g.api.Recv(sa[s] + c0 * M * K, (size_t)(cn * M * K * 4), ncclChar, root, g.comm[s], g.xin[s]));
                        SG_NCCL(g.api.Recv(sb[s] + c0 * K * N, (size_t)(cn * K * N * 4), ncclChar, root, g.comm[s], g.xin[s]));
                    }
                }
                if (t >= 2) {
                    const int64_t c0 = (t - 2) * chunk, cn = cnt[s] - c0 < chunk ? cnt[s] - c0 : chunk;
                    if (cn > 0) {
                        SG_NCCL(g.api.Send(sc[s] + c0 * M * N, (size_t)(cn * M * N * 4), ncclChar, root, g.comm[s], g.xin[s]));
                        SG_NCCL(g.api.Recv(C_root + (lo[s] + c0) * M * N, (size_t)(cn * M * N * 4), ncclChar, s, g.comm[root], g.xin[root]));
                    }
                }
            }
            SG_NCCL(g.api.GroupEnd());
            // products of chunk t on the peers, after this step's transfer
            if (t < steps)
                for (int s = 0; s < G; s++) {
                    if (s == root) continue;
                    const int64_t c0 = t * chunk, cn = cnt[s] - c0 < chunk ? cnt[s] - c0 : chunk;
                    if (cn <= 0) continue;
                    SG_TRY(nb200_set_device(g.dev[s]));
                    Ctx &cs = ctx();
                    SG_CUDA(cudaEventRecord(ev_in[s * (steps + 1) + t], g.xin[s]));
                    SG_CUDA(cudaStreamWaitEvent(cs.stream, ev_in[s * (steps + 1) + t], 0));
                    SG_TRY(nb200_sgemm_batched(sc[s] + c0 * M * N, sa[s] + c0 * M * K, sb[s] + c0 * K * N, cn, M, N, K, M * K, K * N, M * N, precision));
                    SG_CUDA(cudaEventRecord(ev_done[s * (steps + 1) + t], cs.stream));
                }
        }
    } else {
        for (int64_t t = 0; t < steps; t++) {
            // (1) transfer-in of chunk t
            if (transport == NB200_XFER_NCCL) SG_NCCL(g.api.GroupStart());
            for (int s = 0; s < G; s++) {
                if (s == root) continue;
                const int64_t c0 = t * chunk, cn = cnt[s] - c0 < chunk ? cnt[s] - c0 : chunk;
                if (cn <= 0) continue;
                const float *srcA = A_root + (lo[s] + c0) * M * K, *srcB = B_root + (lo[s] + c0) * K * N;
                if (transport == NB200_XFER_P2P) {
                    SG_CUDA(cudaSetDevice(g.dev[s]));
                    SG_CUDA(cudaMemcpyPeerAsync(sa[s] + c0 * M * K, g.dev[s], srcA, g.dev[root], (size_t)(cn * M * K * 4), g.xin[s]));
                    SG_CUDA(cudaMemcpyPeerAsync(sb[s] + c0 * K * N, g.dev[s], srcB, g.dev[root], (size_t)(cn * K * N * 4), g.xin[s]));
                } else {
                    SG_NCCL(g.api.Send(srcA, (size_t)(cn * M * K * 4), ncclChar, s, g.comm[root], g.xout[root]));
                    SG_NCCL(g.api.Send(srcB, (size_t)(cn * K * N * 4), ncclChar, s, g.comm[root], g.xout[root]));
                    SG_NCCL(g.api.Recv(sa[s] + c0 * M * K, (size_t)(cn * M * K * 4), ncclChar, root, g.comm[s], g.xin[s]));
                    SG_NCCL(g.api.Recv(sb[s] + c0 * K * N, (size_t)(cn * K * N * 4), ncclChar, root, g.comm[s], g.xin[s]));
                }
            }
            if (transport == NB200_XFER_NCCL) SG_NCCL(g.api.GroupEnd());
            // (2) products of chunk t on the peers (after their inputs), (3) transfer-out of chunk t (after the products)
            for (int s = 0; s < G; s++) {
                if (s == root) continue;
                const int64_t c0 = t * chunk, cn = cnt[s] - c0 < chunk ? cnt[s] - c0 : chunk;
                if (cn <= 0) continue;
                SG_TRY(nb200_set_device(g.dev[s]));
                Ctx &cs = ctx();
                SG_CUDA(cudaEventRecord(ev_in[s * (steps + 1) + t], g.xin[s]));
                SG_CUDA(cudaStreamWaitEvent(cs.stream, ev_in[s * (steps + 1) + t], 0));
                SG_TRY(nb200_sgemm_batched(sc[s] + c0 * M * N, sa[s] + c0 * M * K, sb[s] + c0 * K * N, cn, M, N, K, M * K, K * N, M * N, precision));
                SG_CUDA(cudaEventRecord(ev_done[s * (steps + 1) + t], cs.stream));
            }
            if (transport == NB200_XFER_NCCL) SG_NCCL(g.api.GroupStart());
            for (int s = 0; s < G; s++) {
                if (s == root) continue;
                const int64_t c0 = t * chunk, cn = cnt[s] - c0 < chunk ? cnt[s] - c0 : chunk;
                if (cn <= 0) continue;
                float *dst = C_root + (lo[s] + c0) * M * N;
                if (transport == NB200_XFER_P2P) {
                    SG_CUDA(cudaSetDevice(g.dev[s]));
                    SG_CUDA(cudaStreamWaitEvent(g.xout[s], ev_done[s * (steps + 1) + t], 0));
                    SG_CUDA(cudaMemcpyPeerAsync(dst, g.dev[root], sc[s] + c0 * M * N, g.dev[s], (size_t)(cn * M * N * 4), g.xout[s]));
                } else {
                    SG_CUDA(cudaSetDevice(g.dev[s]));
                    SG_CUDA(cudaStreamWaitEvent(g.xout[s], ev_done[s * (steps + 1) + t], 0));
                    SG_NCCL(g.api.Send(sc[s] + c0 * M * N, (size_t)(cn * M * N * 4), ncclChar, root, g.comm[s], g.xout[s]));
                    SG_NCCL(g.api.Recv(dst, (size_t)(cn * M * N * 4), ncclChar, s, g.comm[root], g.xin[root]));
                }
            }
            if (transport == NB200_XFER_NCCL) SG_NCCL(g.api.GroupEnd());
        }
}
    // join: the root's context stream continues after every result block has arrived (and after its own share)
    SG_CUDA(cudaSetDevice(g.dev[root]));
    if (transport == NB200_XFER_NCCL) {
        SG_CUDA(cudaEventRecord(ev_done[root * (steps + 1) + steps], g.xin[root]));
        SG_CUDA(cudaStreamWaitEvent(rctx->stream, ev_done[root * (steps + 1) + steps], 0));
    }
    for (int s = 0; s < G; s++) {
        if (s == root) continue;
        SG_CUDA(cudaSetDevice(g.dev[s]));
        SG_CUDA(cudaEventRecord(ev_done[s * (steps + 1) + steps], transport == NB200_XFER_NCCL ? g.xin[s] : g.xout[s]));
        SG_CUDA(cudaStreamWaitEvent(ctx_of(g.dev[s])->stream, ev_done[s * (steps + 1) + steps], 0));   // staging reuse stays ordered
        SG_CUDA(cudaSetDevice(g.dev[root]));
        SG_CUDA(cudaStreamWaitEvent(rctx->stream, ev_done[s * (steps + 1) + steps], 0));
    }
    SG_CUDA(cudaSetDevice(g.dev[root]));
    SG_CUDA(cudaEventRecord(t1, rctx->stream));
    if (elapsed_ms) {
        SG_CUDA(cudaEventSynchronize(t1));
        SG_CUDA(cudaEventElapsedTime(elapsed_ms, t0, t1));
    }
    // the staging blocks go back to their pools; later allocations reuse them only on the context streams, which are ordered after
    // the transfers above.  Events are destroyed lazily by the runtime once they have completed.
    cleanup();
    return NB200_OK;
#undef SG_TRY
#undef SG_CUDA
#undef SG_NCCL
}
