"""e2e probe: nb200_sgemm_host 4096^2 from pinned host buffers, with NB200_HOST_TRACE timeline, and raw PCIe numbers
(H2D alone, D2H alone, both directions at once)."""
import ctypes as C, os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import numpower_b200 as nb
lib = nb.lib()
assert lib.nb200_init(0) == 0
n = 4096
ha, hb, hc = (torch.rand(n, n).pin_memory() for _ in range(3))
def ev(): return torch.cuda.Event(enable_timing=True)
for i in range(4):
    assert lib.nb200_sgemm_host(hc.data_ptr(), ha.data_ptr(), hb.data_ptr(), n, n, n, 3) == 0
import time
t0 = time.perf_counter()
for i in range(10):
    lib.nb200_sgemm_host(hc.data_ptr(), ha.data_ptr(), hb.data_ptr(), n, n, n, 3)
print("sgemm_host ms/call (wall):", (time.perf_counter() - t0) / 10 * 1e3, file=sys.stderr)
pin = torch.empty(64 << 20, dtype=torch.float32).pin_memory(); pin2 = torch.empty(64 << 20, dtype=torch.float32).pin_memory()
d1 = torch.empty(64 << 20, dtype=torch.float32, device="cuda"); d2 = torch.empty(64 << 20, dtype=torch.float32, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
def h2d():
    with torch.cuda.stream(s1): d1.copy_(pin, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): pin2.copy_(d2, non_blocking=True)
def both(): h2d(); d2h()
gb = pin.numel() * 4 / 1e9
print(json.dumps({"h2d_GBps": gb / timed(h2d), "d2h_GBps": gb / timed(d2h), "duplex_each_GBps": gb / timed(both)}), file=sys.stderr)
# batched host pipeline: 16 x 2048^2
nb_, m = 16, 2048
hA, hB, hC = (torch.rand(nb_, m, m).pin_memory() for _ in range(3))
for i in range(3):
    assert lib.nb200_sgemm_batched_host(hC.data_ptr(), hA.data_ptr(), hB.data_ptr(), nb_, m, m, m, 3) == 0
t0 = time.perf_counter()
for i in range(5):
    lib.nb200_sgemm_batched_host(hC.data_ptr(), hA.data_ptr(), hB.data_ptr(), nb_, m, m, m, 3)
print("sgemm_batched_host ms/call (wall):", (time.perf_counter() - t0) / 5 * 1e3, file=sys.stderr)
