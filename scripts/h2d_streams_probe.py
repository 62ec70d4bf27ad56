"""Does splitting one pinned H2D copy over several streams (copy engines) raise the PCIe rate?  gpurun_out/h2d_streams.json"""
import json
import os
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
n = 64 << 20
pin = torch.empty(n, dtype=torch.float32).pin_memory(); pin.fill_(1.0)
dev = torch.empty(n, dtype=torch.float32, device="cuda")
res = {}
for k in (1, 2, 3, 4, 8):
    streams = [torch.cuda.Stream() for _ in range(k)]
    part = n // k
    def go():
        for i, s in enumerate(streams):
            with torch.cuda.stream(s):
                dev[i * part:(i + 1) * part].copy_(pin[i * part:(i + 1) * part], non_blocking=True)
    for direction in ("h2d",):
        for _ in range(2):
            go()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in streams:
            s.wait_event(e0)
        for _ in range(5):
            go()
        for s in streams:
            torch.cuda.current_stream().wait_stream(s)
        e1.record()
        torch.cuda.synchronize()
        res[f"h2d_{k}_streams_GBps"] = n * 4 * 5 / e0.elapsed_time(e1) / 1e6
# chunk size sensitivity on one stream
for chunk_mb in (1, 4, 16):
    c = (chunk_mb << 20) // 4
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(0, n, c):
        dev[i:i + c].copy_(pin[i:i + c], non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    res[f"h2d_1_stream_{chunk_mb}MiB_chunks_GBps"] = n * 4 / e0.elapsed_time(e1) / 1e6
# write-combined host memory via cudaHostAlloc flag
import ctypes
rt = ctypes.CDLL("libcudart.so.12") if os.path.exists("/usr/local/cuda/lib64/libcudart.so.12") else None
try:
    rt = ctypes.CDLL("/usr/local/cuda/lib64/libcudart.so.12")
    p = ctypes.c_void_p()
    assert rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(n * 4), ctypes.c_uint(4)) == 0   # cudaHostAllocWriteCombined
    ctypes.memset(p, 1, n * 4)
    s = torch.cuda.current_stream().cuda_stream
    for _ in range(2):
        rt.cudaMemcpyAsync(ctypes.c_void_p(dev.data_ptr()), p, ctypes.c_size_t(n * 4), 1, ctypes.c_void_p(s))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        rt.cudaMemcpyAsync(ctypes.c_void_p(dev.data_ptr()), p, ctypes.c_size_t(n * 4), 1, ctypes.c_void_p(s))
    e1.record()
    torch.cuda.synchronize()
    res["h2d_write_combined_GBps"] = n * 4 * 5 / e0.elapsed_time(e1) / 1e6
except Exception as e:
    res["wc_error"] = str(e)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "h2d_streams.json"), "w"), indent=1)
print(json.dumps(res, indent=1))
