// Context, error and memory entry points of the C-ABI (include/nb200.h).
// Replaces src/gpu_alloc.c (vmalloc/vfree/vmemcpy*/NDArray_VFLOAT/vmemcheck, :11-54 in
// /root/reference) with 64-bit sizes, status returns and an allocation ledger, and the
// process-global cudaSetDevice of NDArray::setDevice (numpower.c:615-635).
#include "common.cuh"
#include <unordered_map>
#include <map>
#include <cstdlib>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

namespace nb200 {

// One context per device.  The reference keeps a process-global "current device" (NDArray::setDevice -> cudaSetDevice,
// numpower.c:615-635); nb200_set_device switches the current context WITHOUT tearing the others down, so a single host process can
// drive several GPUs (include/nb200.h "multi-GPU shards").  Scratch, workspace, streams, the allocation ledger and the caching pool
// all belong to their device; a block freed while another device is current goes back to its owner's pool.
static Ctx g_ctxs[NB200_MAX_DEVICES];
static int g_current = 0;
// Caching allocator behind nb200_alloc / vmalloc (SURVEY.md §8 f, N3).  The reference pays a cudaMalloc and a
// device-synchronising cudaFree for every result array (gpu_alloc.c:11-34); here freed blocks go to a size-keyed pool
// and are reused by later requests of a similar size.  Reuse is safe without synchronisation because every kernel of
// this library on a device runs on that device's context stream (a block handed out again is only touched by later work on that
// stream).  NB200_NO_CACHE=1 restores plain cudaMalloc/cudaFree.
static int g_cache_enabled = -1;
static bool cache_enabled() {
    if (g_cache_enabled < 0) g_cache_enabled = getenv("NB200_NO_CACHE") ? 0 : 1;
    return g_cache_enabled == 1;
}
static int64_t round_capacity(int64_t bytes) {
    if (bytes < 512) return 512;
    if (bytes < ((int64_t)1 << 20)) return (bytes + 511) & ~int64_t(511);
    return (bytes + ((int64_t)2 << 20) - 1) & ~(((int64_t)2 << 20) - 1);   // 2 MiB granules
}
static void pool_release_all(Ctx &c) {
    for (auto &kv : c.pool) cudaFree(kv.second);
    c.pool.clear();
    c.pool_bytes = 0;
}
static thread_local char g_err[512] = "";

Ctx &ctx() { return g_ctxs[g_current]; }
Ctx *ctx_of(int device) { return (device >= 0 && device < NB200_MAX_DEVICES) ? &g_ctxs[device] : nullptr; }

int set_error(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

static int init_device(int device) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        return set_error(NB200_ENODEV, "no CUDA device available (%s); libnb200 has no CPU fallback",
                         e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    }
    if (device < 0 || device >= count || device >= NB200_MAX_DEVICES)
        return set_error(NB200_EINVAL, "device %d out of range [0,%d)", device, count < NB200_MAX_DEVICES ? count : NB200_MAX_DEVICES);
    cudaDeviceProp prop;
    NB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return set_error(NB200_ENODEV, "device %d is sm_%d%d; libnb200 is built for sm_100a only", device, prop.major,
                         prop.minor);
    NB_CUDA(cudaSetDevice(device));
    Ctx &c = g_ctxs[device];
    c.device = device;
    c.num_sms = prop.multiProcessorCount;
    NB_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
    c.own_stream = true;
    NB_CUDA(cudaMalloc(&c.ticket, 4096 * sizeof(unsigned int)));
    NB_CUDA(cudaMemset(c.ticket, 0, 4096 * sizeof(unsigned int)));
    NB_CUDA(cudaMalloc(&c.domain_flag, sizeof(int)));
    NB_CUDA(cudaMemset(c.domain_flag, 0, sizeof(int)));
    NB_CUDA(cudaMalloc(&c.dev_result, 64));
    NB_CUDA(cudaMallocHost(&c.host_result, 64));
    c.scratch = nullptr;
    c.scratch_bytes = 0;
    c.gemm_ws = nullptr;
    c.gemm_ws_bytes = 0;
    c.ready = true;
    return NB200_OK;
}

int ensure_ready() {
    if (ctx().ready) return NB200_OK;
    const int rc = init_device(g_current);
    return rc;
}

int ensure_scratch(int64_t bytes) {
    Ctx &c = ctx();
    if (bytes <= c.scratch_bytes) return NB200_OK;
    if (c.scratch) {
        NB_CUDA(cudaStreamSynchronize(c.stream));
        NB_CUDA(cudaFree(c.scratch));
        c.scratch = nullptr;
        c.scratch_bytes = 0;
    }
    int64_t want = bytes < (int64_t)(1 << 20) ? (int64_t)(1 << 20) : bytes;
    if (cudaMalloc(&c.scratch, (size_t)want) != cudaSuccess) {
        cudaGetLastError();
        return set_error(NB200_ENOMEM, "device memory allocation failed (scratch %lld bytes)", (long long)want);
    }
    c.scratch_bytes = want;
    return NB200_OK;
}

int ensure_gemm_ws(int64_t bytes) {
    Ctx &c = ctx();
    c.ctl_stride = -1;   // whoever asks for the workspace may overwrite the FP16x3 control blocks; gemm_fp16x3 re-validates its own layout
    if (bytes <= c.gemm_ws_bytes) return NB200_OK;
    if (c.gemm_ws) {
        NB_CUDA(cudaStreamSynchronize(c.stream));
        NB_CUDA(cudaFree(c.gemm_ws));
        c.gemm_ws = nullptr;
        c.gemm_ws_bytes = 0;
    }
    if (cudaMalloc(&c.gemm_ws, (size_t)bytes) != cudaSuccess) {
        cudaGetLastError();
        return set_error(NB200_ENOMEM, "device memory allocation failed (gemm workspace %lld bytes)", (long long)bytes);
    }
    c.gemm_ws_bytes = bytes;
    return NB200_OK;
}

void host_pipeline_release(int device);   // host_pipeline.cu: staging buffers, streams and events of that device
void staging_release(int device);         // below: pinned staging slots of the pageable-copy path

static void shutdown_device(int device) {
    Ctx &c = g_ctxs[device];
    if (!c.ready) return;
    cudaSetDevice(device);
    cudaStreamSynchronize(c.stream);
    host_pipeline_release(device);
    staging_release(device);
    if (c.own_stream && c.stream) cudaStreamDestroy(c.stream);
    if (c.scratch) cudaFree(c.scratch);
    if (c.gemm_ws) cudaFree(c.gemm_ws);
    if (c.ticket) cudaFree(c.ticket);
    if (c.domain_flag) cudaFree(c.domain_flag);
    if (c.dev_result) cudaFree(c.dev_result);
    if (c.host_result) cudaFreeHost(c.host_result);
    pool_release_all(c);
    // blocks the host still holds stay allocated (the host frees them, or leaks them, exactly as with vmalloc); the ledger is
    // dropped so that a later nb200_free of such a block releases it with cudaFree instead of pooling it in a new context
    c.ledger.clear();
    const int64_t launches = c.launches;
    c = Ctx();
    c.launches = launches;
}

}  // namespace nb200

using namespace nb200;

extern "C" int nb200_init(int device) { return nb200_set_device(device); }

extern "C" int nb200_set_device(int device) {
    if (device < 0 || device >= NB200_MAX_DEVICES) return set_error(NB200_EINVAL, "device %d out of range", device);
    if (!g_ctxs[device].ready) {
        const int rc = init_device(device);
        if (rc != NB200_OK) return rc;
    } else {
        NB_CUDA(cudaSetDevice(device));
    }
    g_current = device;
    return NB200_OK;
}

extern "C" int nb200_shutdown(void) {
    for (int d = 0; d < NB200_MAX_DEVICES; d++) shutdown_device(d);
    g_current = 0;
    return NB200_OK;
}

extern "C" int nb200_device_count(int *count) {
    if (!count) return set_error(NB200_EINVAL, "null argument");
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess) {
        cudaGetLastError();
        *count = 0;
        return set_error(NB200_ENODEV, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    }
    return NB200_OK;
}

extern "C" int nb200_get_device(int *device) {
    NB_READY();
    if (!device) return set_error(NB200_EINVAL, "null argument");
    *device = ctx().device;
    return NB200_OK;
}

extern "C" int nb200_synchronize(void) {
    NB_READY();
    NB_CUDA(cudaStreamSynchronize(ctx().stream));
    return NB200_OK;
}

extern "C" const char *nb200_last_error(void) { return g_err; }

extern "C" void *nb200_stream(void) { return ctx().ready ? (void *)ctx().stream : nullptr; }

extern "C" int nb200_set_stream(void *cuda_stream) {
    NB_READY();
    Ctx &c = ctx();
    NB_CUDA(cudaStreamSynchronize(c.stream));
    if (c.own_stream && c.stream) cudaStreamDestroy(c.stream);
    c.stream = static_cast<cudaStream_t>(cuda_stream);
    c.own_stream = false;
    return NB200_OK;
}

extern "C" int64_t nb200_launch_count(void) {
    int64_t n = 0;
    for (int d = 0; d < NB200_MAX_DEVICES; d++) n += g_ctxs[d].launches;
    return n;
}

// ---- CUDA graphs: launch-latency path for sequences of small operations -------------------------------------------------------
// A 1024^2 nd::add moves 12 MiB that live in L2: the kernel takes ~2.5 us, a stream launch costs about as much again and the host
// call path more.  Capturing a sequence of nb200_* calls once and replaying it removes both per-op costs.  Everything enqueued on the
// context stream between begin and end is recorded (no allocation, no host-returning entry point inside a capture).
extern "C" int nb200_graph_begin(void) {
    NB_READY();
    NB_CUDA(cudaStreamBeginCapture(ctx().stream, cudaStreamCaptureModeThreadLocal));
    return NB200_OK;
}

extern "C" int nb200_graph_end(void **graph_exec) {
    NB_READY();
    if (!graph_exec) return set_error(NB200_EINVAL, "null argument");
    cudaGraph_t graph = nullptr;
    NB_CUDA(cudaStreamEndCapture(ctx().stream, &graph));
    cudaGraphExec_t exec = nullptr;
    cudaError_t e = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) return set_error(NB200_ECUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(e));
    *graph_exec = exec;
    return NB200_OK;
}

extern "C" int nb200_graph_launch(void *graph_exec) {
    NB_READY();
    if (!graph_exec) return set_error(NB200_EINVAL, "null argument");
    NB_CUDA(cudaGraphLaunch(static_cast<cudaGraphExec_t>(graph_exec), ctx().stream));
    return NB200_OK;
}

extern "C" int nb200_graph_destroy(void *graph_exec) {
    if (graph_exec) NB_CUDA(cudaGraphExecDestroy(static_cast<cudaGraphExec_t>(graph_exec)));
    return NB200_OK;
}

extern "C" int nb200_trace_enable(unsigned long long *dev_slots) {
    NB_READY();
    ctx().trace = dev_slots;
    return NB200_OK;
}

extern "C" int nb200_poll_domain_error(int *flag) {
    NB_READY();
    if (!flag) return set_error(NB200_EINVAL, "null argument");
    Ctx &c = ctx();
    int h = 0;
    NB_CUDA(cudaMemcpyAsync(&h, c.domain_flag, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
    NB_CUDA(cudaStreamSynchronize(c.stream));
    if (h) NB_CUDA(cudaMemsetAsync(c.domain_flag, 0, sizeof(int), c.stream));
    *flag = h;
    return NB200_OK;
}

// ---- memory ------------------------------------------------------------------------
extern "C" int nb200_alloc(void **dev_ptr, int64_t bytes) {
    NB_READY();
    if (!dev_ptr || bytes < 0) return set_error(NB200_EINVAL, "nb200_alloc: bad argument");
    *dev_ptr = nullptr;
    Ctx &c = ctx();
    const int64_t cap = round_capacity(bytes);
    if (cache_enabled()) {
        auto it = c.pool.lower_bound(cap);
        if (it != c.pool.end() && it->first <= cap + cap / 4) {   // accept up to 25 % slack
            *dev_ptr = it->second;
            c.pool_bytes -= it->first;
            c.ledger[*dev_ptr] = it->first;
            c.live_allocs++;
            c.live_bytes += it->first;
            c.pool.erase(it);
            return NB200_OK;
        }
    }
    if (cudaMalloc(dev_ptr, (size_t)cap) != cudaSuccess) {
        cudaGetLastError();
        pool_release_all(c);   // give cached blocks back to the driver and retry once
        if (cudaMalloc(dev_ptr, (size_t)cap) != cudaSuccess) {
            cudaGetLastError();
            *dev_ptr = nullptr;
            return set_error(NB200_ENOMEM, "device memory allocation failed");  // gpu_alloc.c:15 message
        }
    }
    c.live_allocs++;
    c.live_bytes += cap;
    c.ledger[*dev_ptr] = cap;
    return NB200_OK;
}

extern "C" int nb200_free(void *dev_ptr) {
    NB_READY();
    if (!dev_ptr) return NB200_OK;
    // the block belongs to the device it was allocated on, which need not be the current one
    Ctx *owner = nullptr;
    if (ctx().ledger.count(dev_ptr)) owner = &ctx();
    for (int d = 0; d < NB200_MAX_DEVICES && !owner; d++)
        if (g_ctxs[d].ready && g_ctxs[d].ledger.count(dev_ptr)) owner = &g_ctxs[d];
    if (!owner) {
        // not in any ledger: a block that outlived its context (nb200_shutdown / re-init) is released, anything else is an error
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, dev_ptr) == cudaSuccess && at.type == cudaMemoryTypeDevice) {
            NB_CUDA(cudaFree(dev_ptr));
            return NB200_OK;
        }
        cudaGetLastError();
        return set_error(NB200_EINVAL, "nb200_free: pointer %p was not allocated by nb200_alloc", dev_ptr);
    }
    Ctx &c = *owner;
    const int64_t cap = c.ledger[dev_ptr];
    c.ledger.erase(dev_ptr);
    c.live_allocs--;
    c.live_bytes -= cap;
    // keep at most 32 GiB cached per device; larger pools are trimmed by releasing the biggest blocks first
    if (cache_enabled() && cap <= ((int64_t)8 << 30)) {
        c.pool.emplace(cap, dev_ptr);
        c.pool_bytes += cap;
        while (c.pool_bytes > ((int64_t)32 << 30) && !c.pool.empty()) {
            auto big = std::prev(c.pool.end());
            NB_CUDA(cudaFree(big->second));
            c.pool_bytes -= big->first;
            c.pool.erase(big);
        }
        return NB200_OK;
    }
    NB_CUDA(cudaFree(dev_ptr));   // (synchronises the device: no kernel can still use the block)
    return NB200_OK;
}

// ---- `$a->gpu()` / `->cpu()` from PAGEABLE host memory (SURVEY.md section 8 f, N3) ---------------------------------------------
// PHP's arrays are emalloc'd, i.e. pageable.  A plain cudaMemcpy from pageable memory goes through the driver's single-threaded
// staging copy: measured on the B200 boxes 10.3 GB/s host-to-device and 21.2 GB/s device-to-host against 55.5 / 57.1 GB/s for
// pinned memory (scripts/pageable_probe.py).  Large pageable copies are therefore staged HERE: chunks move through a ring of pinned
// slots, a small pool of worker threads does the pageable <-> pinned memcpy of a chunk in parallel slices, and the DMA of chunk i
// runs while chunk i+1 is being staged.  Pinned / registered / managed pointers and small copies take the direct path.
// Measured (256 MiB, 16-core host): host-to-device 10.3 -> 37-40 GB/s, device-to-host 21.2 -> 36-38 GB/s; 2 / 4 / 8 / 16 threads:
// 23 / 39 / 38 / 38 GB/s in, 16 / 28 / 38 / 36 out (4 MiB chunks).
// NB200_STAGE_THREADS (default 8, 0 = always the direct path), NB200_STAGE_CHUNK_MB (default 4).
namespace nb200 {
class CopyPool {
  public:
    explicit CopyPool(int workers) {
        for (int i = 0; i < workers; i++) threads_.emplace_back([this, i] { run(i); });
    }
    ~CopyPool() {
        { std::lock_guard<std::mutex> lk(m_); stop_ = true; gen_++; }
        cv_.notify_all();
        for (auto &t : threads_) t.join();
    }
    // memcpy(dst, src, n) cut into workers + 1 slices (the caller copies the first one)
    void copy(char *dst, const char *src, size_t n) {
        const size_t parts = threads_.size() + 1;
        if (parts == 1 || n < ((size_t)1 << 20)) { memcpy(dst, src, n); return; }
        const size_t per = ((n / parts) + 4095) & ~size_t(4095);
        { std::lock_guard<std::mutex> lk(m_); dst_ = dst; src_ = src; n_ = n; per_ = per; pending_ = (int)threads_.size(); gen_++; }
        cv_.notify_all();
        memcpy(dst, src, per < n ? per : n);
        std::unique_lock<std::mutex> lk(m_);
        done_.wait(lk, [this] { return pending_ == 0; });
    }
  private:
    void run(int idx) {
        uint64_t seen = 0;
        for (;;) {
            char *dst; const char *src; size_t n, per;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return gen_ != seen; });
                seen = gen_;
                if (stop_) return;
                dst = dst_; src = src_; n = n_; per = per_;
            }
            const size_t off = (size_t)(idx + 1) * per;
            if (off < n) memcpy(dst + off, src + off, n - off < per ? n - off : per);
            { std::lock_guard<std::mutex> lk(m_); if (--pending_ == 0) done_.notify_one(); }
        }
    }
    std::vector<std::thread> threads_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    char *dst_ = nullptr; const char *src_ = nullptr;
    size_t n_ = 0, per_ = 0;
    int pending_ = 0;
    uint64_t gen_ = 0;
    bool stop_ = false;
};
struct Staging {
    static constexpr int SLOTS = 3;
    char *slot[SLOTS] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev[SLOTS] = {nullptr, nullptr, nullptr};
    size_t chunk = 0;
    bool ok = false;
};
static Staging g_stage[NB200_MAX_DEVICES];
static CopyPool *g_copy_pool = nullptr;
static int stage_threads() {
    static const int n = getenv("NB200_STAGE_THREADS") ? atoi(getenv("NB200_STAGE_THREADS")) : 8;
    return n < 0 ? 0 : (n > 32 ? 32 : n);
}
// host pointer the DMA engines cannot read directly (plain malloc / emalloc memory)?
static bool is_pageable(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return true; }
    return a.type == cudaMemoryTypeUnregistered;
}
static int staging_get(Staging **out) {
    const int dev = ctx().device;
    if (dev < 0 || dev >= NB200_MAX_DEVICES) return set_error(NB200_EINVAL, "device index out of range");
    Staging &S = g_stage[dev];
    if (!S.ok) {
        static const size_t mb = getenv("NB200_STAGE_CHUNK_MB") ? (size_t)atoll(getenv("NB200_STAGE_CHUNK_MB")) : 4;
        S.chunk = (mb < 1 ? 1 : mb) << 20;
        for (int i = 0; i < Staging::SLOTS; i++) {
            if (cudaMallocHost(reinterpret_cast<void **>(&S.slot[i]), S.chunk) != cudaSuccess) {
                cudaGetLastError();
                return set_error(NB200_ENOMEM, "pinned staging allocation failed");
            }
            NB_CUDA(cudaEventCreateWithFlags(&S.ev[i], cudaEventDisableTiming));
        }
        S.ok = true;
    }
    if (!g_copy_pool) g_copy_pool = new CopyPool(stage_threads() - 1);
    *out = &S;
    return NB200_OK;
}
void staging_release(int device) {   // nb200_shutdown
    if (device < 0 || device >= NB200_MAX_DEVICES) return;
    Staging &S = g_stage[device];
    for (int i = 0; i < Staging::SLOTS; i++) {
        if (S.slot[i]) cudaFreeHost(S.slot[i]);
        if (S.ev[i]) cudaEventDestroy(S.ev[i]);
    }
    S = Staging();
}
static int staged_copy(char *dev, char *host, size_t bytes, bool to_device) {
    Staging *S = nullptr;
    int rc = staging_get(&S);
    if (rc != NB200_OK) return rc;
    cudaStream_t st = ctx().stream;
    const size_t C = S->chunk;
    const size_t nchunks = (bytes + C - 1) / C;
    if (to_device) {
        for (size_t i = 0; i < nchunks; i++) {
            const int sl = (int)(i % Staging::SLOTS);
            const size_t off = i * C, n = bytes - off < C ? bytes - off : C;
            if (i >= (size_t)Staging::SLOTS) NB_CUDA(cudaEventSynchronize(S->ev[sl]));   // the DMA that last read this slot is done
            g_copy_pool->copy(S->slot[sl], host + off, n);
            NB_CUDA(cudaMemcpyAsync(dev + off, S->slot[sl], n, cudaMemcpyHostToDevice, st));
            NB_CUDA(cudaEventRecord(S->ev[sl], st));
        }
        NB_CUDA(cudaStreamSynchronize(st));
    } else {
        // DMA of chunk i + 1 and i + 2 in flight while chunk i is copied out of its slot
        size_t issued = 0;
        auto issue = [&](size_t i) -> int {
            const int sl = (int)(i % Staging::SLOTS);
            const size_t off = i * C, n = bytes - off < C ? bytes - off : C;
            NB_CUDA(cudaMemcpyAsync(S->slot[sl], dev + off, n, cudaMemcpyDeviceToHost, st));
            NB_CUDA(cudaEventRecord(S->ev[sl], st));
            return NB200_OK;
        };
        for (; issued < nchunks && issued < (size_t)Staging::SLOTS - 1; issued++)
            if ((rc = issue(issued)) != NB200_OK) return rc;
        for (size_t i = 0; i < nchunks; i++) {
            if (issued < nchunks) { if ((rc = issue(issued)) != NB200_OK) return rc; issued++; }
            const int sl = (int)(i % Staging::SLOTS);
            const size_t off = i * C, n = bytes - off < C ? bytes - off : C;
            NB_CUDA(cudaEventSynchronize(S->ev[sl]));
            g_copy_pool->copy(host + off, S->slot[sl], n);
        }
    }
    return NB200_OK;
}
static bool use_staging(const void *host, int64_t bytes) {
    return stage_threads() > 0 && bytes >= ((int64_t)16 << 20) && is_pageable(host);
}
}  // namespace nb200

extern "C" int nb200_copy_h2d(void *dev_dst, const void *host_src, int64_t bytes) {
    NB_READY();
    if (bytes < 0 || (bytes > 0 && (!dev_dst || !host_src))) return set_error(NB200_EINVAL, "nb200_copy_h2d: bad argument");
    if (use_staging(host_src, bytes))
        return staged_copy(static_cast<char *>(dev_dst), const_cast<char *>(static_cast<const char *>(host_src)), (size_t)bytes, true);
    NB_CUDA(cudaMemcpyAsync(dev_dst, host_src, (size_t)bytes, cudaMemcpyHostToDevice, ctx().stream));
    NB_CUDA(cudaStreamSynchronize(ctx().stream));
    return NB200_OK;
}

extern "C" int nb200_copy_d2h(void *host_dst, const void *dev_src, int64_t bytes) {
    NB_READY();
    if (bytes < 0 || (bytes > 0 && (!host_dst || !dev_src))) return set_error(NB200_EINVAL, "nb200_copy_d2h: bad argument");
    if (use_staging(host_dst, bytes))
        return staged_copy(const_cast<char *>(static_cast<const char *>(dev_src)), static_cast<char *>(host_dst), (size_t)bytes, false);
    NB_CUDA(cudaMemcpyAsync(host_dst, dev_src, (size_t)bytes, cudaMemcpyDeviceToHost, ctx().stream));
    NB_CUDA(cudaStreamSynchronize(ctx().stream));
    return NB200_OK;
}

extern "C" int nb200_copy_d2d(void *dev_dst, const void *dev_src, int64_t bytes) {
    NB_READY();
    if (bytes < 0 || (bytes > 0 && (!dev_dst || !dev_src))) return set_error(NB200_EINVAL, "nb200_copy_d2d: bad argument");
    NB_CUDA(cudaMemcpyAsync(dev_dst, dev_src, (size_t)bytes, cudaMemcpyDeviceToDevice, ctx().stream));
    return NB200_OK;
}

extern "C" int nb200_memset_zero(void *dev_ptr, int64_t bytes) {
    NB_READY();
    if (bytes < 0 || (bytes > 0 && !dev_ptr)) return set_error(NB200_EINVAL, "nb200_memset_zero: bad argument");
    NB_CUDA(cudaMemsetAsync(dev_ptr, 0, (size_t)bytes, ctx().stream));
    return NB200_OK;
}

extern "C" int nb200_mem_stats(int64_t *live_allocations, int64_t *live_bytes) {
    if (live_allocations) *live_allocations = ctx().live_allocs;
    if (live_bytes) *live_bytes = ctx().live_bytes;
    return NB200_OK;
}

extern "C" int nb200_host_alloc(void **host_ptr, int64_t bytes) {
    NB_READY();
    if (!host_ptr || bytes < 0) return set_error(NB200_EINVAL, "nb200_host_alloc: bad argument");
    if (cudaMallocHost(host_ptr, bytes == 0 ? 16 : (size_t)bytes) != cudaSuccess) {
        cudaGetLastError();
        return set_error(NB200_ENOMEM, "pinned host allocation failed");
    }
    return NB200_OK;
}

extern "C" int nb200_host_free(void *host_ptr) {
    NB_READY();
    if (host_ptr) NB_CUDA(cudaFreeHost(host_ptr));
    return NB200_OK;
}
