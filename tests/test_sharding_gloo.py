"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: shard partitioning, scatter/gather round trip,
deterministic combination of per-rank partials, and the sharded batched-matmul flow with the oracle port
standing in for the device kernel (test only)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_shard_range_partitions_exactly():
    from numpower_b200.sharding import shard_range, shard_sizes
    for n in (0, 1, 7, 8, 1024, 1025):
        for w in (1, 2, 3, 4, 8):
            spans = [shard_range(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(shard_sizes(n, w)) - min(shard_sizes(n, w)) <= 1


def test_c_abi_shard_split_matches_the_python_rule():
    """nb200_shard_split (pure host arithmetic, callable without a GPU) == sharding.shard_range."""
    import ctypes as C
    import numpower_b200 as nb
    from numpower_b200.sharding import shard_range
    lib = nb.lib()
    for n in (0, 1, 7, 8, 1023, 1024, 1025, 2 ** 28 + 3):
        for w in (1, 2, 3, 4, 8, 16):
            for r in range(w):
                lo, cnt = C.c_int64(), C.c_int64()
                assert lib.nb200_shard_split(n, w, r, C.byref(lo), C.byref(cnt)) == 0
                assert (lo.value, lo.value + cnt.value) == shard_range(n, w, r)
    lo, cnt = C.c_int64(), C.c_int64()
    assert lib.nb200_shard_split(10, 0, 0, C.byref(lo), C.byref(cnt)) != 0
    n = C.c_int(-1)
    assert lib.nb200_shard_count(C.byref(n)) == 0 and n.value == 0          # no group without nb200_shard_init


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    import oracle
    from numpower_b200 import sharding as sh
    rng = np.random.default_rng(0)
    batch, m, k, n = 5, 16, 24, 8
    a = rng.random((batch, m, k), dtype=np.float32)
    b = rng.random((batch, k, n), dtype=np.float32)
    A = torch.from_numpy(a) if rank == 0 else None
    Bm = torch.from_numpy(b) if rank == 0 else None
    a_sh = sh.scatter_axis0(A, a.shape)
    b_sh = sh.scatter_axis0(Bm, b.shape)
    lo, hi = sh.shard_range(batch, world, rank)
    assert a_sh.shape[0] == hi - lo and np.array_equal(a_sh.numpy(), a[lo:hi])
    c_sh = torch.stack([torch.from_numpy(oracle.port.matmul(a_sh[i].numpy(), b_sh[i].numpy())) for i in range(hi - lo)]) \
        if hi > lo else torch.empty((0, m, n))
    full = sh.gather_axis0(c_sh, batch)
    ok = True
    if rank == 0:
        exp = np.stack([oracle.port.matmul(a[i], b[i]) for i in range(batch)])
        ok = np.array_equal(full.numpy(), exp)
    # full reductions: exact-set data so the order does not matter, NaN rules for min/max
    x = (rng.integers(-64, 65, size=1000).astype(np.float32) / 64)
    xs = x[slice(*sh.shard_range(1000, world, rank))]
    tot = sh.allreduce_partials(float(oracle.port.reduce_full("sum", xs)), "sum")
    ok = ok and tot == float(oracle.port.reduce_full("sum", x))
    mx = sh.allreduce_partials(float(oracle.port.reduce_full("max", xs)), "max")
    ok = ok and mx == float(x.max())
    # argmax with a tie across ranks: the lower rank (lower global index) must win
    y = np.zeros(1000, np.float32)
    y[10] = 3.0
    y[900] = 3.0
    lo2, hi2 = sh.shard_range(1000, world, rank)
    ys = y[lo2:hi2]
    li = int(oracle.port.argminmax(True, ys))
    gi = sh.allreduce_argminmax(float(ys[li]), li, lo2, True)
    ok = ok and gi == 10.0
    # NaN at a shard boundary.  Reference rule over the GLOBAL array (oracle on the whole array): a NaN sticks only at
    # global index 0; [.., | NaN, -100] must still find -100, and a NaN at index 0 must win.
    blo, _ = sh.shard_range(1000, world, world - 1)
    z = x.copy()
    z[blo] = np.nan
    z[blo + 1] = -100.0
    for zz in (z, np.concatenate([[np.float32(np.nan)], z[1:]]).astype(np.float32)):
        zs = zz[lo2:hi2]
        for op in ("min", "max"):
            part = sh.local_minmax_partial(lambda off, cnt: float(oracle.port.reduce_full(op, zs[off:off + cnt])), len(zs), rank)
            got = sh.allreduce_partials(part, op)
            exp = float(oracle.port.reduce_full(op, zz))
            ok = ok and (got == exp or (got != got and exp != exp))
        # argmax / argmin: NaN-position-neutral local candidates (what the packed device word carries), combined with the global rules
        for is_max in (True, False):
            nanmask = np.isnan(zs)
            if is_max:
                li = int(np.nanargmax(zs)) if not nanmask.all() else 0
            else:
                li = int(np.argmax(nanmask)) if nanmask.any() else int(np.argmin(zs))
            gi = sh.allreduce_argminmax(float(zs[li]), li, lo2, is_max, first_value=float(zs[0]))
            ok = ok and gi == float(oracle.port.argminmax(is_max, zz))
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_world2_scatter_compute_gather_and_partials():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=100) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    assert sorted(res) == [(0, True), (1, True)]
