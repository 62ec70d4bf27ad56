"""Launches each hot kernel a few times at its BASELINE size so that `ncu -k regex:...` can capture it.
Used only under ncu (gpurun); numbers printed under a profiler are never bench values."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpower_b200 as nb

lib = nb.lib()
assert lib.nb200_init(0) == 0
lib.nb200_set_stream(C.c_void_p(torch.cuda.current_stream().cuda_stream))
g = torch.Generator(device="cuda").manual_seed(0)
which = sys.argv[1:] or ["gemm", "ew", "reduce"]
reps = 2
if "gemm" in which:
    n = 4096
    a = torch.rand(n, n, device="cuda", generator=g); b = torch.rand(n, n, device="cuda", generator=g); c = torch.empty(n, n, device="cuda")
    for _ in range(reps):
        assert lib.nb200_sgemm(c.data_ptr(), a.data_ptr(), b.data_ptr(), n, n, n, n, n, n, 0) == 0
    for _ in range(reps):
        assert lib.nb200_sgemm(c.data_ptr(), a.data_ptr(), b.data_ptr(), n, n, n, n, n, n, 1) == 0
    del a, b, c
if "gemm_bf16" in which:
    n = 4096
    a = torch.rand(n, n, device="cuda", generator=g); b = torch.rand(n, n, device="cuda", generator=g); c = torch.empty(n, n, device="cuda")
    for _ in range(reps + 1):
        assert lib.nb200_sgemm(c.data_ptr(), a.data_ptr(), b.data_ptr(), n, n, n, n, n, n, 2) == 0
    del a, b, c
if "gemm_fp16" in which:
    n = 4096
    a = torch.rand(n, n, device="cuda", generator=g); b = torch.rand(n, n, device="cuda", generator=g); c = torch.empty(n, n, device="cuda")
    for _ in range(reps + 1):
        assert lib.nb200_sgemm(c.data_ptr(), a.data_ptr(), b.data_ptr(), n, n, n, n, n, n, 4) == 0
    del a, b, c
if "gemm_fp16u" in which:
    n = 4096
    a = torch.rand(n, n, device="cuda", generator=g); b = torch.rand(n, n, device="cuda", generator=g); c = torch.empty(n, n, device="cuda")
    for _ in range(reps + 1):
        assert lib.nb200_sgemm(c.data_ptr(), a.data_ptr(), b.data_ptr(), n, n, n, n, n, n, 5) == 0
    del a, b, c
if "gemm_auto" in which:   # what nd::matmul runs (NB200_GEMM_AUTO)
    n = 4096
    a = torch.rand(n, n, device="cuda", generator=g); b = torch.rand(n, n, device="cuda", generator=g); c = torch.empty(n, n, device="cuda")
    for _ in range(reps + 1):
        assert lib.nb200_sgemm(c.data_ptr(), a.data_ptr(), b.data_ptr(), n, n, n, n, n, n, 3) == 0
    del a, b, c
if "ew" in which:
    m = 8192
    x = torch.rand(m, m, device="cuda", generator=g); y = torch.rand(m, m, device="cuda", generator=g); z = torch.rand(m, m, device="cuda", generator=g)
    out = torch.empty(m, m, device="cuda")
    shp = (C.c_int64 * 2)(m, m); full = (C.c_int64 * 2)(m, 1); rowv = (C.c_int64 * 2)(0, 1); colv = (C.c_int64 * 2)(1, 0)
    for _ in range(reps):
        assert lib.nb200_ew_mul_add(out.data_ptr(), x.data_ptr(), y.data_ptr(), z.data_ptr(), 2, shp, full, full, full) == 0
        assert lib.nb200_ew_mul_add(out.data_ptr(), x.data_ptr(), y.data_ptr(), z.data_ptr(), 2, shp, full, rowv, colv) == 0
        assert lib.nb200_ew_binary(0, out.data_ptr(), x.data_ptr(), y.data_ptr(), 2, shp, full, full) == 0
        assert lib.nb200_ew_unary(2, out.data_ptr(), x.data_ptr(), m * m, 0.0, 0.0) == 0
    ax = torch.empty(m, device="cuda")
    for _ in range(reps):
        assert lib.nb200_reduce_axis(0, ax.data_ptr(), x.data_ptr(), 1, m, m, 0) == 0
        assert lib.nb200_reduce_axis(0, ax.data_ptr(), x.data_ptr(), m, m, 1, 0) == 0
    del x, y, z, out
if "reduce" in which:
    big = torch.rand(1 << 28, device="cuda", generator=g)
    res = torch.empty(16, device="cuda")
    for _ in range(reps):
        assert lib.nb200_reduce_full(0, res.data_ptr(), big.data_ptr(), 1 << 28) == 0
        assert lib.nb200_argminmax(1, res.data_ptr(), big.data_ptr(), 1, 1 << 28, 1) == 0
torch.cuda.synchronize()
print("done")
