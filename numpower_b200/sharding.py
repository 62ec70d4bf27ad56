"""Multi-GPU sharding of the NDArray hot path (SURVEY.md §8 e): one process per GPU, launched by torchrun;
torch.distributed (NCCL over NVLink on the GPU box, gloo in the CPU tests) is only the plumbing.

The path shards naturally: elementwise ops, batched matmul and axis reductions over a non-reduced axis are
independent per unit, so RESIDENT shards need no data-path collective at all.  Collectives appear only
 (a) when an array that lives on one rank has to be scattered / gathered (`scatter_axis0` / `gather_axis0`,
     one batched send/recv each), and
 (b) to combine the tiny per-rank partials of a full reduction (`allreduce_*`), in fixed rank order so the
     result is deterministic.
Nothing here computes on array data; compute is the caller's libnb200 calls on its shard.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced split of `n` units: the first n % world ranks get one extra unit."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_sizes(n: int, world: int) -> List[int]:
    return [shard_range(n, world, r)[1] - shard_range(n, world, r)[0] for r in range(world)]


def scatter_axis0(full: torch.Tensor | None, shape: Sequence[int], root: int = 0, dtype=torch.float32, device=None) -> torch.Tensor:
    """Root holds `full` (shape = `shape`); every rank returns its contiguous axis-0 shard.
    One grouped send/recv (NCCL: ncclGroupStart/End around per-peer ncclSend/ncclRecv)."""
    rank, world = dist.get_rank(), dist.get_world_size()
    lo, hi = shard_range(shape[0], world, rank)
    device = device if device is not None else (full.device if full is not None else torch.device("cpu"))
    out = torch.empty((hi - lo, *shape[1:]), dtype=dtype, device=device)
    if world == 1:
        out.copy_(full)
        return out
    ops = []
    if rank == root:
        for r in range(world):
            a, b = shard_range(shape[0], world, r)
            if r == root:
                out.copy_(full[a:b])
            elif b > a:
                ops.append(dist.P2POp(dist.isend, full[a:b].contiguous(), r))
    elif hi > lo:
        ops.append(dist.P2POp(dist.irecv, out, root))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return out


def gather_axis0(shard: torch.Tensor, n: int, root: int = 0) -> torch.Tensor | None:
    """Inverse of scatter_axis0: root returns the (n, ...) array, other ranks None."""
    rank, world = dist.get_rank(), dist.get_world_size()
    if world == 1:
        return shard.clone()
    full = torch.empty((n, *shard.shape[1:]), dtype=shard.dtype, device=shard.device) if rank == root else None
    ops = []
    if rank == root:
        for r in range(world):
            a, b = shard_range(n, world, r)
            if r == root:
                full[a:b].copy_(shard)
            elif b > a:
                ops.append(dist.P2POp(dist.irecv, full[a:b], r))
    elif shard.shape[0] > 0:
        ops.append(dist.P2POp(dist.isend, shard.contiguous(), root))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return full


def _gather_small(t: torch.Tensor) -> List[torch.Tensor]:
    world = dist.get_world_size()
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t)
    return outs


def allreduce_partials(partial: float, op: str, device=None) -> float:
    """Combine per-rank partials of a full reduction in FIXED rank order (deterministic).  op in {sum, prod, min, max}.
    min / max follow NDArray_Min / NDArray_Max over the GLOBAL array (ndarray.c:752-772, 939-959: `if (a[i] < m) m = a[i]`): a NaN
    sticks only when it is global element 0 and is skipped everywhere else.  The per-shard kernel applies that rule to its own first
    element, so ranks > 0 must pass the partial of their shard WITHOUT its leading NaN run (`local_minmax_partial` does that); here a
    NaN from rank 0 sticks and a NaN from a later rank (an all-NaN shard) is skipped, exactly like the sequential loop."""
    t = torch.tensor([partial], dtype=torch.float32, device=device)
    parts = [float(x.item()) for x in _gather_small(t)]
    acc = torch.tensor(parts[0], dtype=torch.float32)
    for p in parts[1:]:
        q = torch.tensor(p, dtype=torch.float32)
        if op == "sum":
            acc = acc + q
        elif op == "prod":
            acc = acc * q
        elif op == "min":
            acc = q if bool(q < acc) else acc          # false for NaN on either side: acc NaN sticks, q NaN is skipped
        else:
            acc = q if bool(q > acc) else acc
    return float(acc)


def local_minmax_partial(reduce_fn, n: int, rank: int) -> float:
    """Per-rank partial for allreduce_partials(min|max).  reduce_fn(offset, count) -> the kernel's result over shard[offset:offset+count]
    (NaN iff that sub-range starts with NaN).  Rank 0 keeps the kernel's rule; later ranks drop their leading NaNs first."""
    p = reduce_fn(0, n)
    if rank == 0:
        return p
    off = 0
    while p != p:
        off += 1
        if off >= n:
            return float("nan")        # all NaN: skipped by the combine
        p = reduce_fn(off, n - off)
    return p


def allreduce_argminmax(value: float, local_index: int, offset: int, is_max: bool, device=None, first_value: float | None = None) -> float:
    """Per-rank candidate (value at the local arg index, local index) -> global index as float32 (`(float)i`, calculation.c:25).
    Rules of float_argmax / float_argmin (calculation.c:9-59) over the GLOBAL index space: first occurrence wins, ties across ranks go
    to the lowest rank (= lowest global index: shards are contiguous); argmax never takes a NaN unless it is global element 0; argmin
    takes the first NaN anywhere.  `value` must come from a NaN-position-neutral local scan (for argmax: the shard's arg over its
    non-NaN elements, NaN only if the shard is all NaN); `first_value` = the shard's first element (only rank 0's is used)."""
    fv = value if first_value is None else first_value
    t = torch.tensor([value, float(offset + local_index), fv], dtype=torch.float64, device=device)
    parts = [x.tolist() for x in _gather_small(t)]
    if parts[0][2] != parts[0][2]:
        return 0.0                                            # leading NaN wins outright (calculation.c:14-17, :41-44)
    best = None
    for v, gi, _ in parts:   # rank order
        nan = v != v
        if best is None:
            best = (v, gi)
        elif is_max:
            if (best[0] != best[0] and not nan) or (not nan and v > best[0]):
                best = (v, gi)
        else:
            if best[0] != best[0]:
                continue                                      # an earlier rank already holds the first NaN
            if nan or v < best[0]:
                best = (v, gi)
    return float(torch.tensor(best[1], dtype=torch.float32))


# ------------------------------------------------------------------------------------------------ single-process shard group (C-ABI)
class ShardGroup:
    """Thin face of the nb200_shard_* entry points (include/nb200.h): one host process, G GPUs, one pointer per shard.
    Tensors are torch CUDA tensors living on the group's devices; nothing here computes on array data."""

    NCCL, P2P = 0, 1

    def __init__(self, devices: Sequence[int] | None = None):
        import ctypes as C
        import numpower_b200 as nb
        self.C, self.lib = C, nb.lib()
        n = torch.cuda.device_count() if devices is None else len(devices)
        devs = list(range(n)) if devices is None else list(devices)
        arr = (C.c_int * n)(*devs)
        self._check(self.lib.nb200_shard_init(n, arr))
        self.devices, self.n = devs, n

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(self.lib.nb200_last_error().decode())

    def close(self):
        self._check(self.lib.nb200_shard_finalize())

    def split(self, units: int) -> List[Tuple[int, int]]:
        return [shard_range(units, self.n, s) for s in range(self.n)]

    def ptrs(self, tensors):
        return (self.C.c_void_p * self.n)(*[(t.data_ptr() if t is not None and t.numel() else None) for t in tensors])

    def empty_shards(self, units: int, tail_shape: Sequence[int]) -> List[torch.Tensor]:
        return [torch.empty((hi - lo, *tail_shape), dtype=torch.float32, device=torch.device("cuda", self.devices[s]))
                for s, (lo, hi) in enumerate(self.split(units))]

    def scatter(self, shards, root_tensor, root: int = 0, transport: int = 0):
        rows, row_elems = root_tensor.shape[0], root_tensor[0].numel() if root_tensor.shape[0] else 0
        self._check(self.lib.nb200_shard_scatter(self.ptrs(shards), root_tensor.data_ptr(), rows, row_elems, root, transport))

    def gather(self, root_tensor, shards, root: int = 0, transport: int = 0):
        rows, row_elems = root_tensor.shape[0], root_tensor[0].numel() if root_tensor.shape[0] else 0
        self._check(self.lib.nb200_shard_gather(root_tensor.data_ptr(), self.ptrs(shards), rows, row_elems, root, transport))

    def synchronize(self):
        self._check(self.lib.nb200_shard_synchronize())

    def reduce_full(self, op: int, shards, n_total: int) -> float:
        out = self.C.c_float()
        self._check(self.lib.nb200_shard_reduce_full(op, self.C.byref(out), self.ptrs(shards), n_total))
        return out.value

    def argminmax(self, is_max: bool, shards, n_total: int) -> float:
        out = self.C.c_float()
        self._check(self.lib.nb200_shard_argminmax(1 if is_max else 0, self.C.byref(out), self.ptrs(shards), n_total))
        return out.value
