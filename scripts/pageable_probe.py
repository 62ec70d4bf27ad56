"""H2D / D2H rates of nb200_copy_h2d / nb200_copy_d2h from pageable vs pinned host memory (what `$a->gpu()` / `->cpu()` pay)."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpower_b200 as nb

lib = nb.lib()
assert lib.nb200_init(0) == 0
n = 64 << 20
dev = torch.empty(n, dtype=torch.float32, device="cuda")
res = {}
for kind in ("pageable", "pinned"):
    host = torch.ones(n, dtype=torch.float32)
    if kind == "pinned":
        host = host.pin_memory()
    for name, fn in (("h2d", lambda: lib.nb200_copy_h2d(dev.data_ptr(), host.data_ptr(), n * 4)),
                     ("d2h", lambda: lib.nb200_copy_d2h(host.data_ptr(), dev.data_ptr(), n * 4))):
        fn()
        best = 1e9
        for _ in range(5):
            t0 = time.perf_counter()
            assert fn() == 0
            best = min(best, time.perf_counter() - t0)
        res[f"{kind}_{name}_GBps"] = n * 4 / best / 1e9
print(json.dumps(res))
