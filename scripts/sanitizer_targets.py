"""Small-shape pass over every kernel family for compute-sanitizer (memcheck / racecheck / synccheck).
Shapes are ragged on purpose (non-multiples of the vector width / tile sizes, unaligned views)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpower_b200 as nb

nd, A = nb.nd, nb.NDArray.array
r = np.random.default_rng(0)
a = A(r.random((37, 53), dtype=np.float32)).gpu()
b = A(r.random((37, 53), dtype=np.float32)).gpu()
row = A(r.random(53, dtype=np.float32)).gpu()
col = A(r.random((37, 1), dtype=np.float32)).gpu()
for op in ("add", "sub", "mul", "div", "mod", "pow", "maximum", "minimum", "arctan2", "equal", "less"):
    nd.binary(op, a, b).toArray(); nd.binary(op, a, row).toArray(); nd.binary(op, a, col).toArray(); nd.binary(op, a, 2.0).toArray()
(a[1] + a[2]).toArray()                      # 4-byte aligned views
nd.mul_add(a, row, col).toArray(); nd.mul_add(a, b, b).toArray()
x4 = A(r.random((3, 1, 5, 7), dtype=np.float32)).gpu(); y4 = A(r.random((4, 1, 7), dtype=np.float32)).gpu()
nd.add(x4, y4).toArray()                     # generic N-D path
for op in ("exp", "log", "sin", "sqrt", "rsqrt", "sinc", "sign", "rint", "degrees"):
    nd.unary(op, a).toArray()
nd.clip(a, 0.2, 0.8).toArray(); nd.round(a, 2).toArray()
big = A(r.random(100003, dtype=np.float32)).gpu()
for op in ("sum", "prod", "min", "max"):
    nd.reduce(op, big); nd.reduce(op, a, 0).toArray(); nd.reduce(op, a, 1).toArray()
t3 = A(r.random((5, 70, 9), dtype=np.float32)).gpu()
for ax in (0, 1, 2):
    nd.sum(t3, ax).toArray(); nd.sum(t3, ax, order=nb.ORDER_SEQUENTIAL).toArray(); nd.argmax(t3, ax).toArray(); nd.argmin(t3, ax).toArray()
tall = A(r.random((3000, 5), dtype=np.float32)).gpu(); nd.sum(tall, 0).toArray()
wide = A(r.random((3, 70001), dtype=np.float32)).gpu(); nd.sum(wide, 1).toArray(); nd.argmax(wide, 1).toArray()
nd.argmax(big); nd.argmin(big)
m1, m2 = A(r.random((300, 136), dtype=np.float32)).gpu(), A(r.random((136, 200), dtype=np.float32)).gpu()
nd.matmul(m1, m2).toArray(); nd.matmul(m1, m2, nb.TF32X1).toArray()            # tcgen05 path, ragged tiles
nd.matmul(A(r.random((3, 130, 64), dtype=np.float32)).gpu(), A(r.random((3, 64, 260), dtype=np.float32)).gpu()).toArray()
nd.matmul(A(r.random((5, 7), dtype=np.float32)).gpu(), A(r.random((7, 3), dtype=np.float32)).gpu()).toArray()   # SIMT path
if os.environ.get("SANITIZE_16BIT", "1") == "1":
    # 16-bit operand modes: odd K / N (repacking split), merged 256-column tile incl. the split tail, scaled half mode
    m3, m4 = A(r.random((300, 137), dtype=np.float32)).gpu(), A(r.random((137, 201), dtype=np.float32)).gpu()
    for prec in (nb.BF16X3, nb.FP16X3, nb.FP16X3U):
        nd.matmul(m1, m2, prec).toArray(); nd.matmul(m3, m4, prec).toArray()
        nd.matmul(A(r.random((512, 160), dtype=np.float32)).gpu(), A(r.random((160, 512), dtype=np.float32)).gpu(), prec).toArray()
        nd.matmul(A(r.random((3, 260, 72), dtype=np.float32)).gpu(), A(r.random((3, 72, 264), dtype=np.float32)).gpu(), prec).toArray()
    # out-of-window elements: sparse repair (records spread over the grid) and the gated TF32x3 fallback, both 16-bit scaled modes
    m5 = r.random((300, 256), dtype=np.float32) + 0.25
    m6 = r.random((256, 264), dtype=np.float32) + 0.25
    m5[17, 5] = 2.0 ** -40; m6[30, 21] = 2.0 ** -40
    m7 = m5.copy(); m7[:40, :128] *= np.float32(2.0 ** -40)
    for prec in (nb.FP16X3, nb.FP16X3U):
        nd.matmul(A(m5).gpu(), A(m6).gpu(), prec).toArray(); nd.matmul(A(m7).gpu(), A(m6).gpu(), prec).toArray()
nd.dot(m1, A(r.random(136, dtype=np.float32)).gpu()).toArray()
# boolean reductions, transpose, in-place elementwise (out == in), CUDA-graph capture / replay
nd.all(big); nd.allclose(big, big); nd.all(a[1]); nd.allclose(a[1], a[2])
import ctypes as C
lib = nb.lib()
t_in, t_out = A(r.random((33, 65), dtype=np.float32)).gpu(), A(np.zeros((65, 33), np.float32)).gpu()
assert lib.nb200_transpose2d(t_out.data_ptr, t_in.data_ptr, 33, 65) == 0
assert lib.nb200_ew_unary(1, big.data_ptr, big.data_ptr, big.size, 0.0, 0.0) == 0
shp, st = (C.c_int64 * 1)(big.size), (C.c_int64 * 1)(1)
assert lib.nb200_ew_binary(0, big.data_ptr, big.data_ptr, big.data_ptr, 1, shp, st, st) == 0
assert lib.nb200_graph_begin() == 0
assert lib.nb200_ew_binary(2, big.data_ptr, big.data_ptr, big.data_ptr, 1, shp, st, st) == 0
g = C.c_void_p()
assert lib.nb200_graph_end(C.byref(g)) == 0
assert lib.nb200_graph_launch(g) == 0 and lib.nb200_synchronize() == 0 and lib.nb200_graph_destroy(g) == 0
# host-operand pipelines (worker streams, ragged last block), outer product / L1 norm compositions
hm, hk, hn = 1025, 136, 264
ha, hb = r.random((hm, hk), dtype=np.float32), r.random((hk, hn), dtype=np.float32)
hc = np.empty((hm, hn), np.float32)
for prec in (0, 2, 3):
    assert lib.nb200_sgemm_host(hc.ctypes.data, ha.ctypes.data, hb.ctypes.data, hm, hn, hk, prec) == 0
hA, hB = r.random((5, 130, 72), dtype=np.float32), r.random((5, 72, 264), dtype=np.float32)
hC = np.empty((5, 130, 264), np.float32)
assert lib.nb200_sgemm_batched_host(hC.ctypes.data, hA.ctypes.data, hB.ctypes.data, 5, 130, 264, 72, 3) == 0
nd.outer(row, A(r.random(41, dtype=np.float32)).gpu()).toArray(); nd.norm(a, 1)
print("sanitizer targets done;", nb.lib().nb200_launch_count(), "launches")
