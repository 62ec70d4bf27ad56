"""How much of a buffer that was just read is still in L2 when it is read again?  nb200_reduce_full (sum, streaming loads) over
buffers of 8..128 MiB, back to back (warm) vs with an L2 flush between the launches (cold).  Effective GB/s of each."""
import ctypes as C
import json
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import numpower_b200 as nb

lib = nb.lib()
assert lib.nb200_init(0) == 0
assert lib.nb200_set_stream(C.c_void_p(torch.cuda.current_stream().cuda_stream)) == 0
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
res = torch.empty(16, device="cuda")
for mib in (8, 16, 32, 48, 64, 96, 128, 256):
    n = (mib << 20) // 4
    x = torch.rand(n, device="cuda")
    fn = lambda: lib.nb200_reduce_full(0, res.data_ptr(), x.data_ptr(), n)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        fn()
    e1.record()
    torch.cuda.synchronize()
    warm = e0.elapsed_time(e1) / 20
    cold = 0.0
    for _ in range(10):
        flush.zero_()
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        cold += e0.elapsed_time(e1) / 10
    print(json.dumps({"MiB": mib, "warm_us": round(warm * 1e3, 2), "cold_us": round(cold * 1e3, 2), "warm_GBps": round(n * 4 / warm / 1e6), "cold_GBps": round(n * 4 / cold / 1e6)}), flush=True)
