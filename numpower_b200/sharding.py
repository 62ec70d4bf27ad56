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
    """Combine per-rank partials of a full reduction in FIXED rank order (deterministic).
    op in {sum, prod, min, max}; min/max use the skip-NaN rule of the kernels."""
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
            acc = torch.fmin(acc, q)
        else:
            acc = torch.fmax(acc, q)
    return float(acc)


def allreduce_argminmax(value: float, local_index: int, offset: int, is_max: bool, device=None) -> float:
    """Per-rank (value, local index) -> global index as float32 (calculation.c:25 semantics): best value wins,
    ties go to the LOWEST rank (= lowest global index because shards are contiguous)."""
    t = torch.tensor([value, float(offset + local_index)], dtype=torch.float64, device=device)
    parts = [x.tolist() for x in _gather_small(t)]
    best = None
    for v, gi in parts:   # rank order
        if best is None:
            best = (v, gi)
        elif (is_max and v > best[0]) or ((not is_max) and v < best[0]):
            best = (v, gi)
    return float(torch.tensor(best[1], dtype=torch.float32))
