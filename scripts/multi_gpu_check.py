"""torchrun entry (one rank per GPU, NCCL): sharded hot path end to end on real GPUs.
 1. batched nd::matmul with the batch scattered from rank 0 over NVLink (sharding.scatter_axis0), computed per shard
    through the C-ABI, gathered back (gather_axis0); rank 0 checks sampled matrices against the oracle.
 2. sharded nd::sum / nd::argmax over a 1-D array: per-rank partial through the C-ABI + fixed-order combine.
Prints one JSON line (rank 0) with the parity verdicts and the scatter/compute/gather timings.
"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import numpower_b200 as nb
    from numpower_b200 import sharding as sh
    lib = nb.lib()
    assert lib.nb200_init(local) == 0
    assert lib.nb200_set_stream(C.c_void_p(torch.cuda.current_stream().cuda_stream)) == 0
    dev = torch.device("cuda", local)
    per_rank = int(os.environ.get("NB200_CHECK_BATCH_PER_RANK", "16"))
    batch, n = per_rank * world, 2048
    out = {"world": world, "batch": batch, "n": n}
    A = Bm = None
    if rank == 0:
        g = torch.Generator(device=dev).manual_seed(10)
        A = torch.rand(batch, n, n, device=dev, generator=g)
        Bm = torch.rand(batch, n, n, device=dev, generator=g)

    def timed(fn):
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return r, float(t.item())

    # warm-up of NCCL p2p channels and the kernels
    a_sh = sh.scatter_axis0(A, (batch, n, n), device=dev)
    b_sh = sh.scatter_axis0(Bm, (batch, n, n), device=dev)
    c_sh = torch.empty_like(a_sh)
    nloc = a_sh.shape[0]

    def compute():
        assert lib.nb200_sgemm_batched(c_sh.data_ptr(), a_sh.data_ptr(), b_sh.data_ptr(), nloc, n, n, n, n * n, n * n, n * n, 0) == 0
    compute()
    sh.gather_axis0(c_sh, batch)

    (a_sh, b_sh), t_scatter = timed(lambda: (sh.scatter_axis0(A, (batch, n, n), device=dev), sh.scatter_axis0(Bm, (batch, n, n), device=dev)))
    _, t_compute = timed(compute)
    full, t_gather = timed(lambda: sh.gather_axis0(c_sh, batch))
    if rank == 0:
        import oracle
        chk = oracle.ref if oracle.ref.available else oracle.port
        worst = 0.0
        for i in sorted({0, batch // 2, batch - 1}):
            exp = chk.matmul(A[i].cpu().numpy(), Bm[i].cpu().numpy())
            got = full[i].cpu().numpy()
            worst = max(worst, float(np.abs(got - exp).max() / np.abs(exp).max()))
        moved_out = (world - 1) / world * 2 * batch * n * n * 4   # bytes leaving rank 0 in the scatter
        moved_in = (world - 1) / world * batch * n * n * 4        # bytes entering rank 0 in the gather
        out.update(matmul_max_rel_err=worst, matmul_ok=bool(worst <= 1e-5), scatter_ms=t_scatter, compute_ms=t_compute, gather_ms=t_gather,
                   scatter_GBps_root_egress=moved_out / t_scatter / 1e6, gather_GBps_root_ingress=moved_in / t_gather / 1e6,
                   compute_tflops_total=batch * 2.0 * n ** 3 / t_compute / 1e9,
                   end_to_end_tflops=batch * 2.0 * n ** 3 / (t_scatter + t_compute + t_gather) / 1e9)
    del A, Bm, a_sh, b_sh, c_sh

    # ---- sharded full reductions over a 2^26-element array (exact-set data: any order gives the same sum)
    N = 1 << 26
    x = None
    if rank == 0:
        xs = np.random.default_rng(8).choice(np.array([-1, 0, 0, 1], np.float32), size=N)
        xs[40_000_001] = 7.0
        xs[50_000_001] = 7.0   # later tie on a higher rank must lose
        x = torch.from_numpy(xs).to(dev)
    x_sh = sh.scatter_axis0(x, (N,), device=dev)
    lo, hi = sh.shard_range(N, world, rank)
    part = C.c_float()
    assert lib.nb200_reduce_full_host(0, C.byref(part), x_sh.data_ptr(), hi - lo) == 0
    total = sh.allreduce_partials(part.value, "sum", device=dev)
    idx = C.c_float()
    assert lib.nb200_argminmax_host(1, C.byref(idx), x_sh.data_ptr(), hi - lo) == 0
    li = int(idx.value)
    val = float(x_sh[li].item())
    gidx = sh.allreduce_argminmax(val, li, lo, True, device=dev)
    if rank == 0:
        out.update(sum_ok=bool(total == float(xs.astype(np.float64).sum())), argmax_ok=bool(gidx == float(np.float32(40_000_001))),
                   sharded_sum=total, sharded_argmax=gidx)
        print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
