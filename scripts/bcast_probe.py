"""Times the row/column-broadcast fused chain (config 3b) for NB200_BCAST_U = 4 and 8 in one process pair."""
import ctypes as C, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch, numpower_b200 as nb
    lib = nb.lib(); assert lib.nb200_init(0) == 0
    lib.nb200_set_stream(C.c_void_p(torch.cuda.current_stream().cuda_stream))
    m = 8192
    x = torch.rand(m, m, device="cuda"); y = torch.rand(m, m, device="cuda"); z = torch.rand(m, m, device="cuda"); out = torch.empty(m, m, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    shp = (C.c_int64 * 2)(m, m); full = (C.c_int64 * 2)(m, 1); rowv = (C.c_int64 * 2)(0, 1); colv = (C.c_int64 * 2)(1, 0)
    def t(fn):
        for _ in range(3): fn()
        tot = 0
        for _ in range(10):
            flush.zero_(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize(); tot += e0.elapsed_time(e1)
        return tot / 10
    a = t(lambda: lib.nb200_ew_mul_add(out.data_ptr(), x.data_ptr(), y.data_ptr(), z.data_ptr(), 2, shp, full, rowv, colv))
    b = t(lambda: lib.nb200_ew_binary(0, out.data_ptr(), x.data_ptr(), y.data_ptr(), 2, shp, full, rowv))
    c = t(lambda: lib.nb200_ew_binary(2, out.data_ptr(), x.data_ptr(), z.data_ptr(), 2, shp, full, colv))
    print(f"U={os.environ.get('NB200_BCAST_U')} mul_add row+col {a:.4f} ms  add row {b:.4f} ms  mul col {c:.4f} ms  (ideal @6486 GB/s: {2*m*m*4/6486e6:.4f} ms)")
else:
    for rep in range(2):
        for u in ("8", "4"):
            p = subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, NB200_BCAST_U=u), capture_output=True, text=True)
            print(p.stdout.strip() or p.stderr[-500:])
