"""GPU parity tests of the multi-GPU shard entry points (include/nb200.h "multi-GPU shards", csrc/shard.cu): one host process,
G devices, through the C-ABI.  They run in `pytest -m gpu` whenever >= 2 GPUs are visible (skipped on a 1-GPU box; the scatter /
gather round trip with one shard still runs there).  Checker: the oracle on the CONCATENATED array, i.e. the reference's semantics
over the global index space."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle
from helpers import rel_err

ORACLE = oracle.ref if oracle.ref.available else oracle.port
pytestmark = pytest.mark.gpu
NCCL, P2P = 0, 1


@pytest.fixture(scope="module")
def grp():
    import torch
    import numpower_b200 as nb
    from numpower_b200.sharding import ShardGroup
    n = torch.cuda.device_count()
    g = ShardGroup(list(range(min(n, 8))))
    yield g, nb, torch
    g.close()
    assert nb.lib().nb200_set_device(0) == 0


def _need2(g):
    if g.n < 2:
        pytest.skip("needs >= 2 GPUs")


def _put(torch, g, arr, units):
    """numpy (units, ...) -> one torch tensor per shard on its device"""
    return [torch.from_numpy(np.ascontiguousarray(arr[lo:hi])).to(f"cuda:{g.devices[s]}") for s, (lo, hi) in enumerate(g.split(units))]


@pytest.mark.parametrize("transport", [NCCL, P2P])
def test_scatter_gather_round_trip_is_bit_exact(grp, transport):
    g, nb, torch = grp
    r = np.random.default_rng(5)
    for rows, cols in ((1000, 37), (7, 4096), (g.n, 1), (3, 5)):       # ragged: fewer rows than shards leaves empty shards
        a = r.random((rows, cols), dtype=np.float32)
        for root in sorted({0, g.n - 1}):
            src = torch.from_numpy(a).to(f"cuda:{g.devices[root]}")
            shards = g.empty_shards(rows, (cols,))
            g.scatter(shards, src, root=root, transport=transport)
            g.synchronize()
            for s, (lo, hi) in enumerate(g.split(rows)):
                np.testing.assert_array_equal(shards[s].cpu().numpy(), a[lo:hi])
            back = torch.zeros_like(src)
            g.gather(back, shards, root=root, transport=transport)
            g.synchronize()
            np.testing.assert_array_equal(back.cpu().numpy(), a)


def test_sharded_upload_download_round_trip_and_use(grp):
    """nb200_shard_upload / nb200_shard_download: the sharded `gpu()` / `cpu()` - host array <-> row shards over every device's own
    PCIe link (pinned and pageable host memory), bit-exact, and the uploaded shards feed a sharded op directly (stream order)."""
    g, nb, torch = grp
    lib = nb.lib()
    r = np.random.default_rng(15)
    for rows, cols, pinned in ((1000, 37, True), (7, 4096, False), (max(g.n - 1, 1), 3, True), (2053, 1024, True)):
        a = r.random((rows, cols), dtype=np.float32)
        host = torch.from_numpy(a.copy())
        if pinned:
            host = host.pin_memory()
        shards = g.empty_shards(rows, (cols,))
        assert lib.nb200_shard_upload(g.ptrs(shards), host.data_ptr(), rows, cols) == 0, lib.nb200_last_error()
        out = g.empty_shards(rows, (cols,))
        assert lib.nb200_shard_ew_binary(2, g.ptrs(out), g.ptrs(shards), g.ptrs(shards), rows, cols) == 0, lib.nb200_last_error()   # a * a
        back = torch.empty(rows, cols)
        if pinned:
            back = back.pin_memory()
        assert lib.nb200_shard_download(back.data_ptr(), g.ptrs(out), rows, cols) == 0, lib.nb200_last_error()
        g.synchronize()
        for s, (lo, hi) in enumerate(g.split(rows)):
            np.testing.assert_array_equal(shards[s].cpu().numpy(), a[lo:hi])
        np.testing.assert_array_equal(back.numpy(), ORACLE.binary("mul", a, a))


def test_sharded_elementwise_matches_oracle(grp):
    g, nb, torch = grp
    lib = nb.lib()
    r = np.random.default_rng(6)
    rows, cols = 1031, 520
    a, b, c = (r.random((rows, cols), dtype=np.float32) - 0.5 for _ in range(3))
    A, B, Cc = _put(torch, g, a, rows), _put(torch, g, b, rows), _put(torch, g, c, rows)
    out = g.empty_shards(rows, (cols,))
    assert lib.nb200_shard_ew_binary(0, g.ptrs(out), g.ptrs(A), g.ptrs(B), rows, cols) == 0, lib.nb200_last_error()
    g.synchronize()
    got = np.concatenate([o.cpu().numpy() for o in out])
    np.testing.assert_array_equal(got, ORACLE.binary("add", a, b))
    assert lib.nb200_shard_ew_mul_add(g.ptrs(out), g.ptrs(A), g.ptrs(B), g.ptrs(Cc), rows, cols) == 0, lib.nb200_last_error()
    g.synchronize()
    got = np.concatenate([o.cpu().numpy() for o in out])
    np.testing.assert_array_equal(got, ORACLE.mul_add(a, b, c))
    assert lib.nb200_shard_ew_unary(2, g.ptrs(out), g.ptrs(A), rows, cols, 0.0, 0.0) == 0      # exp
    g.synchronize()
    got = np.concatenate([o.cpu().numpy() for o in out])
    assert rel_err(got, ORACLE.unary("exp", a)).max() <= 1e-5


def test_sharded_full_reductions_global_semantics(grp):
    g, nb, torch = grp
    r = np.random.default_rng(7)
    n = 100_003
    x = (r.integers(-64, 65, size=n).astype(np.float32) / 64)            # exactly summable: any order gives the same bits
    ops = {"sum": 0, "prod": 1, "min": 2, "max": 3}
    def check(arr):
        sh = _put(torch, g, arr, len(arr))
        for name, op in ops.items():
            if name == "prod":
                continue
            got, exp = g.reduce_full(op, sh, len(arr)), float(ORACLE.reduce_full(name, arr))
            assert got == exp or (got != got and exp != exp), (name, got, exp)
        for is_max in (True, False):
            got, exp = g.argminmax(is_max, sh, len(arr)), float(ORACLE.argminmax(is_max, arr))
            assert got == exp, ("argmax" if is_max else "argmin", got, exp)
    check(x)
    # ties across shard boundaries: the lowest global index wins
    y = np.zeros(n, np.float32)
    y[[10, n // 2, n - 3]] = 3.0
    y[[11, n // 2 + 1]] = -3.0
    check(y)
    # NaN rules over the GLOBAL index space: NaN at a shard's first element (not global 0) is skipped by min/max/argmax and taken by argmin
    for s in range(1, g.n):
        z = x.copy()
        lo, _ = g.split(n)[s]
        z[lo] = np.nan
        z[lo + 1] = -100.0
        z[lo + 2] = 100.0
        check(z)
        z[lo:lo + 5] = np.nan                                              # a run of leading NaNs in the shard
        check(z)
    z = x.copy()
    z[0] = np.nan                                                          # global element 0: sticks / wins
    check(z)
    # prod on a small exactly-representable set
    p = np.full(40, 1.0, np.float32)
    p[[3, 17, 33]] = [2.0, -0.5, 4.0]
    assert g.reduce_full(1, _put(torch, g, p, 40), 40) == float(ORACLE.reduce_full("prod", p))


def test_argmax_index_beyond_2pow24_is_exact_before_the_float_cast(grp):
    """A maximum planted at an odd global index > 2^24 inside the last shard: the combine works on the exact integer index and only
    then rounds like the reference's (float)i."""
    g, nb, torch = grp
    n = (1 << 25) + 11
    planted = (1 << 25) - 5 if g.n > 1 else (1 << 24) + 7
    shards = [torch.zeros(hi - lo, device=f"cuda:{g.devices[s]}") for s, (lo, hi) in enumerate(g.split(n))]
    for s, (lo, hi) in enumerate(g.split(n)):
        if lo <= planted < hi:
            shards[s][planted - lo] = 5.0
    assert g.argminmax(True, shards, n) == float(np.float32(planted))
    assert g.argminmax(False, shards, n) == 0.0


@pytest.mark.parametrize("transport", [NCCL, P2P])
def test_batched_matmul_scatter_compute_gather_pipeline(grp, transport):
    g, nb, torch = grp
    lib = nb.lib()
    r = np.random.default_rng(9)
    batch, M, N, K = 2 * g.n + 3, 256, 264, 200
    a, b = r.random((batch, M, K), dtype=np.float32), r.random((batch, K, N), dtype=np.float32)
    exp = np.stack([ORACLE.matmul(a[i], b[i]) for i in range(batch)])
    for root in sorted({0, g.n - 1}):
        dev = f"cuda:{g.devices[root]}"
        A, B = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)
        for chunk in (1, 2, 64):
            Cc = torch.full((batch, M, N), float("nan"), device=dev)
            ms = C.c_float()
            assert lib.nb200_sgemm_batched_scatter_gather(Cc.data_ptr(), A.data_ptr(), B.data_ptr(), batch, M, N, K, nb.GEMM_AUTO, root, transport,
                                                          chunk, C.byref(ms)) == 0, lib.nb200_last_error()
            assert rel_err(Cc.cpu().numpy(), exp).max() <= 1e-5, (root, chunk)
            assert ms.value > 0
        # asynchronous form (no elapsed_ms): ordered with the root's context stream, joined by nb200_shard_synchronize
        Cc = torch.full((batch, M, N), float("nan"), device=dev)
        assert lib.nb200_sgemm_batched_scatter_gather(Cc.data_ptr(), A.data_ptr(), B.data_ptr(), batch, M, N, K, nb.GEMM_AUTO, root, transport, 2, None) == 0
        g.synchronize()
        assert rel_err(Cc.cpu().numpy(), exp).max() <= 1e-5


def test_resident_sharded_batched_matmul_and_device_switching(grp):
    g, nb, torch = grp
    lib = nb.lib()
    r = np.random.default_rng(10)
    batch, M, N, K = g.n + 1, 256, 128, 160
    a, b = r.random((batch, M, K), dtype=np.float32), r.random((batch, K, N), dtype=np.float32)
    A, B = _put(torch, g, a, batch), _put(torch, g, b, batch)
    Cs = g.empty_shards(batch, (M, N))
    assert lib.nb200_sgemm_batched_sharded(g.ptrs(Cs), g.ptrs(A), g.ptrs(B), batch, M, N, K, nb.GEMM_AUTO) == 0, lib.nb200_last_error()
    g.synchronize()
    got = np.concatenate([c.cpu().numpy() for c in Cs])
    exp = np.stack([ORACLE.matmul(a[i], b[i]) for i in range(batch)])
    assert rel_err(got, exp).max() <= 1e-5
    # per-device state.  A block allocated on the last device and freed while device 0 is current goes back to ITS pool;
    # a >48 KB-shared-memory GEMM launch works on every device after switching back and forth.
    last = g.devices[-1]
    assert lib.nb200_set_device(last) == 0
    p = C.c_void_p()
    assert lib.nb200_alloc(C.byref(p), 1 << 20) == 0
    assert lib.nb200_set_device(g.devices[0]) == 0
    live, nbytes = C.c_int64(), C.c_int64()
    assert lib.nb200_mem_stats(C.byref(live), C.byref(nbytes)) == 0
    before = live.value
    assert lib.nb200_free(p) == 0
    # device 0's ledger untouched (a one-GPU box has a single device: the block WAS device 0's)
    assert lib.nb200_mem_stats(C.byref(live), C.byref(nbytes)) == 0 and live.value == before - (1 if last == g.devices[0] else 0)
    assert lib.nb200_set_device(last) == 0
    q = C.c_void_p()
    assert lib.nb200_alloc(C.byref(q), 1 << 20) == 0
    if last != g.devices[0]:
        assert q.value == p.value                                                                   # came back from the owner's pool
    assert lib.nb200_free(q) == 0
    assert lib.nb200_set_device(g.devices[0]) == 0
