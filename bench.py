#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 NDArray backend (contract: task prompt, section 4).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N = 1 : workload = BASELINE.json configs[1] "nd::matmul 4096x4096 fp32": one step = one nd::matmul through the C-ABI
        (nb200_sgemm, NB200_GEMM_AUTO = FP16x3: error-compensated three-product scheme on tcgen05 kind::f16 whose bound
        guarantees 1e-5).  `value` = useful TFLOP/s with operands resident in HBM; `e2e` = same call fed from pinned HOST
        buffers (H2D of A,B and D2H of C inside the timed region).  Every other BASELINE config (a*b+c 8192^2 chain,
        sum/argmax over 2^28, axis sums, nd::add 1024^2, the batched-matmul per-GPU share) is measured in the same run and
        reported under roofline.per_config / cpu_baseline.per_config (GB/s or TFLOP/s, fraction of the measured roofline, and
        the reference CPU path on the FULL config, BASELINE.md §4 protocol).
N > 1 : launched by torchrun, one rank per GPU: BASELINE.json configs[4] batched matmul, the batch dimension sharded with no
        data-path collective (128 matrices of 2048^2 per rank; N = 8 is exactly 1024 x (2048x2048)).  Weak scaling; value =
        total useful TFLOP/s over all ranks, max-over-ranks time.  The line also carries per-rank times / clocks, the same
        workload on ONE GPU of the same box in the same run (`scaling_base`), the sharded HBM-bound configs, the host-operand
        pipeline (`e2e`) and the single-process NVLink scatter + compute + gather variant (`scatter_gather`, NCCL vs P2P).
--impl reference : the reference's own CPU implementation (oracle/_ref = NumPower's object code -> OpenBLAS cblas_sgemm)
        timed on the host cores for the same metric/config.
"""
from __future__ import annotations

import argparse
import csv
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MATMUL_N = 4096
SHARD_BATCH, SHARD_N = 128, 2048
GEMM_AUTO = 3                      # include/nb200.h NB200_GEMM_AUTO
# what NB200_GEMM_AUTO can resolve to (nb200_gemm_resolve_precision): name, dtype string, MMA kind of the roofline peak
MODES = {
    0: ("tf32x3", "tf32x3 (fp32 in/out, error-compensated 3-pass TF32, fp32 accumulate)", "tf32"),
    2: ("bf16x3", "bf16x3 (fp32 in/out, operands split into 2 bf16 parts, 3 MMAs, fp32 accumulate)", "bf16"),
    4: ("fp16x3", "fp16x3 (fp32 in/out, row/column-scaled operands split into 2 half parts, 3 kind::f16 MMAs, fp32 accumulate; "
                  "sparse repair / TF32x3 fallback decided on the device)", "bf16"),
    5: ("fp16x3u", "fp16x3u (fp32 in/out, row/column-scaled operands split into 2 half parts with the lo part unscaled, 3 kind::f16 MMAs into "
                   "one accumulator per chunk (merged 256x256 tile), fp32 accumulate; sparse repair / TF32x3 fallback decided on the device)", "bf16"),
}
METRIC = "nd::matmul useful TFLOP/s (fp32 in/out)"
FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}
NVLINK_GBS_PER_DIR = 900.0
# ncu --set full summaries (scripts/ncu_extract.py) the roofline traffic figures are READ from at run time (profiles/ travels with
# the repo); kernel-name fragment -> which record it feeds
NCU_FILES = ["profiles/r2_ncu_matmul_auto.csv", "profiles/r2_ncu_hbm.csv"]


def config_single():
    return {"workload": f"nd::matmul {MATMUL_N}x{MATMUL_N} fp32 (BASELINE configs[1])"}


def config_multi(world):
    return {"workload": f"batched nd::matmul {SHARD_BATCH} x ({SHARD_N}x{SHARD_N}) per GPU x {world} GPUs, batch sharded "
                        f"(BASELINE configs[4] = 1024 x (2048x2048) at 8 GPUs)"}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        d["_source"] = "measured (MEASURED_PEAKS.json)"
        return d
    d = dict(FALLBACK_PEAKS)
    d["_source"] = "fallback (B200_PROFILING.md)"
    return d


def load_ncu_traffic():
    """kernel name -> dram bytes (read + write) per launch, from the committed ncu --set full summaries."""
    out = {}
    for rel in NCU_FILES:
        p = os.path.join(ROOT, rel)
        if not os.path.exists(p):
            continue
        rows = list(csv.reader(open(p)))
        if len(rows) < 2:
            continue
        hdr = rows[0]
        def col(name):
            for i, h in enumerate(hdr):
                if h.startswith(name):
                    return i, h
            return None, None
        ri, rh = col("dram__bytes_read.sum")
        wi, wh = col("dram__bytes_write.sum")
        ti, _ = col("gpu__time_duration.sum")
        pi, _ = col("sm__pipe_tensor_cycles_active")
        if ri is None or wi is None:
            continue
        def scale(h):
            return 1e9 if "Gbyte" in h else 1e6 if "Mbyte" in h else 1e3 if "Kbyte" in h else 1.0
        for r in rows[1:]:
            try:
                rec = {"dram_bytes": float(r[ri]) * scale(rh) + float(r[wi]) * scale(wh), "source": rel}
                if ti is not None:
                    rec["ncu_us"] = float(r[ti])
                if pi is not None:
                    rec["tensor_pipe_active_pct"] = float(r[pi])
            except (ValueError, IndexError):
                continue
            out.setdefault(r[0].strip(), []).append(rec)
    return out


def ncu_lookup(traffic, *fragments, near=None):
    """Sum the DRAM bytes of the kernels whose names contain ALL fragments of one tuple each (several captures of a kernel: the one
    whose traffic is closest to `near`, else the first); (None, None, None) if one is missing."""
    total, srcs, pipe = 0.0, set(), None
    for frag in fragments:
        frag = frag if isinstance(frag, tuple) else (frag,)
        recs = next((v for k, v in traffic.items() if all(f in k for f in frag)), None)
        if not recs:
            return None, None, None
        hit = min(recs, key=lambda r: abs(r["dram_bytes"] - near)) if near else recs[0]
        total += hit["dram_bytes"]
        srcs.add(hit["source"])
        if hit.get("tensor_pipe_active_pct"):
            pipe = hit["tensor_pipe_active_pct"]
    return total, sorted(srcs), pipe


class ClockSampler:
    """nvidia-smi clocks / power / throttle reasons of ONE GPU (every rank samples its own).  nvidia-smi needs the better part of a
    second to start, so the sampler is started well before the timed region and every sample carries its own timestamp: stop()
    keeps the samples that fall inside the [t0, t1] window the caller marks around the timed region (mark_start / mark_end)."""

    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "20"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def mark_start(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.1)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)      # gone before anything else is timed (its NVML queries hold driver locks)
        except Exception:
            self.proc.kill()

        def summarise(lines):
            sm, mx, reasons, power = [], None, set(), []
            for _, ln in lines:
                f = [x.strip() for x in ln.split(",")]
                if len(f) < 10:
                    continue
                try:
                    sm.append(float(f[2]))
                    mx = float(f[3])
                    power.append(float(f[4]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[6:10]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            return {"sm_mhz": statistics.median(sm) if sm else None, "sm_min_mhz": min(sm) if sm else None, "sm_max_mhz": mx,
                    "reasons": sorted(reasons), "samples": len(sm), "power_w_max": max(power) if power else None,
                    "power_w_median": statistics.median(power) if power else None}
        if self.t0 is None:
            return summarise(self.lines)
        t1 = self.t1 if self.t1 is not None else time.time()
        # a sample printed at time t describes the ~20 ms before it
        inside = [x for x in self.lines if self.t0 <= x[0] <= t1 + 0.03]
        out = summarise(inside)
        out["window_s"] = t1 - self.t0
        if out["samples"] == 0:        # region shorter than nvidia-smi's period: fall back to the nearest samples around it
            near = sorted(self.lines, key=lambda x: abs(x[0] - (self.t0 + t1) / 2))[:3]
            out = summarise(near)
            out["window_s"] = t1 - self.t0
            out["note"] = "timed region shorter than the sampling period: nearest samples"
        return out


# ------------------------------------------------------------------------------------------------ reference arm
def _oracle_impl():
    import oracle
    use_ref = oracle.ref.available
    impl = oracle.ref if use_ref else oracle.port
    if use_ref:
        oracle.ref.lib  # load
        oracle.ref.set_blas_threads(os.cpu_count() or 1)   # torchrun exports OMP_NUM_THREADS=1
    return oracle, impl, use_ref


def run_reference(args):
    """The reference's own CPU path for the arm's config: NDArray_Matmul -> cblas_sgemm (oracle/_ref), driver's steps / warmup."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import numpy as np
    oracle, impl, use_ref = _oracle_impl()
    rng = np.random.default_rng(3)
    if args.gpus > 1:
        # same workload as our multi-GPU arm (batched 2048^2 matmuls = a loop of NDArray_Matmul calls),
        # each step a bounded sample: 4 of the 128 x N matrices
        n, per_step, config = SHARD_N, 4, config_multi(args.gpus)
        sample = f"each step = {per_step} of the {SHARD_BATCH * args.gpus} matrices (loop of NDArray_Matmul -> cblas_sgemm)"
    else:
        n, per_step, config = MATMUL_N, 1, config_single()
        sample = "each step = one full 4096^3 nd::matmul (NDArray_Matmul -> cblas_sgemm)"
    a, b = rng.random((n, n), dtype=np.float32), rng.random((n, n), dtype=np.float32)
    for _ in range(args.warmup):
        for _ in range(per_step):
            impl.matmul(a, b)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        for _ in range(per_step):
            impl.matmul(a, b)
    dt = (time.perf_counter() - t0) / args.steps
    info = oracle.ref.blas_info() if use_ref else {}
    cores = info.get("threads", os.cpu_count())
    val = per_step * 2.0 * n ** 3 / dt / 1e12
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "TFLOP/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config,
        "cpu_baseline": {"value": val, "unit": "TFLOP/s", "cores": cores, "kind": "reference" if use_ref else "port",
                         "sample": sample, "blas": info.get("config", "")},
        "e2e": {"value": val, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# ------------------------------------------------------------------------------------------------ our arm
class Bench:
    def __init__(self, device: int):
        import torch
        import numpower_b200 as nb
        self.torch, self.nb = torch, nb
        self.lib = nb.lib()
        torch.cuda.set_device(device)
        self.check(self.lib.nb200_init(device))
        # run the library on torch's current stream so torch.cuda.Event brackets its kernels
        self.check(self.lib.nb200_set_stream(C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        self.flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def check(self, rc):
        if rc != 0:
            raise RuntimeError(self.lib.nb200_last_error().decode())

    def flush_l2(self):
        self.flush_buf.zero_()

    def time_steps(self, fn, steps, warmup, flush=False):
        """ms per step, CUDA events on the launching stream, sync on both sides."""
        torch = self.torch
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if not flush:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / steps
        total = 0.0
        for _ in range(steps):
            self.flush_l2()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            total += e0.elapsed_time(e1)
        return total / steps


    def time_rotating(self, fn, n_sets, rounds, warm_rounds=1):
        """ms per launch of fn(i), i cycling over n_sets DISTINCT operand sets whose combined size is many times the L2: every launch
        reads data that cannot be cached (the other sets were streamed in between), launches go back to back, one event pair around
        all of them - the kernel's steady-state rate without a per-launch event / drain gap."""
        torch = self.torch
        for _ in range(warm_rounds * n_sets):
            fn(_ % n_sets)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for r in range(rounds * n_sets):
            fn(r % n_sets)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / (rounds * n_sets)


def best_of(fn, reps, warm=1):
    for _ in range(warm):
        fn()
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        best = min(best, time.perf_counter() - t0)
    return best


def cpu_baselines_full(host):
    """Reference CPU path on the FULL BASELINE configs (same host buffers as the GPU run, 1 warm-up + best of 5,
    20 for config 1).  Elementwise / reductions / argmax are single-threaded in the reference; cblas_sgemm uses every core."""
    oracle, impl, use_ref = _oracle_impl()
    info = oracle.ref.blas_info() if use_ref else {}
    kind = "reference" if use_ref else "port"
    threads = info.get("threads", os.cpu_count())
    out = {}
    a, b = host["mm_a"], host["mm_b"]
    t = best_of(lambda: impl.matmul(a, b), 5)
    out["matmul_4096"] = {"ms": t * 1e3, "TFLOPs": 2.0 * MATMUL_N ** 3 / t / 1e12, "cores": threads, "protocol": "best of 5, full config"}
    s, s2 = host["add_a"], host["add_b"]
    t = best_of(lambda: impl.binary("add", s, s2), 20)
    out["add_1024sq"] = {"ms": t * 1e3, "GBps": 3 * s.nbytes / t / 1e9, "cores": 1, "protocol": "best of 20, full config (config 1)"}
    x, y, z = host["chain_x"], host["chain_y"], host["chain_z"]
    t = best_of(lambda: impl.mul_add(x, y, z), 5)
    out["chain_mul_add_8192sq"] = {"ms": t * 1e3, "GBps_algorithmic_fused": 4 * x.nbytes / t / 1e9, "GBps_two_calls": 6 * x.nbytes / t / 1e9,
                                   "cores": 1, "protocol": "best of 5, full config, two nd:: calls (Multiply then Add) as PHP does"}
    big = host["big"]
    t = best_of(lambda: impl.reduce_full("sum", big), 5)
    out["sum_2pow28"] = {"ms": t * 1e3, "GBps": big.nbytes / t / 1e9, "cores": 1, "protocol": "best of 5, full config",
                         "result": float(impl.reduce_full("sum", big)), "note": "sequential fp32 accumulator saturates at 2^24 on U[0,1) data (SURVEY F1)"}
    t = best_of(lambda: impl.argminmax(True, big), 5)
    out["argmax_2pow28"] = {"ms": t * 1e3, "GBps": big.nbytes / t / 1e9, "cores": 1, "protocol": "best of 5, full config"}
    ax = host["chain_x"]
    t = best_of(lambda: impl.reduce_axis("sum", ax, 0), 3)
    out["sum_axis0_8192sq"] = {"ms": t * 1e3, "GBps": ax.nbytes / t / 1e9, "cores": 1,
                               "protocol": "best of 3, full config: reduce() = 8191 slice-wise NDArray_Add_Float calls (ndarray.c:394-429); axis 1 degenerates to one "
                                           "allocation per element (12 s measured) and is not timed"}
    ba, bb = host["bm_a"], host["bm_b"]
    t = best_of(lambda: impl.matmul(ba, bb), 8)
    out["batched_matmul_1024x2048sq"] = {"ms_per_slice": t * 1e3, "TFLOPs": 2.0 * SHARD_N ** 3 / t / 1e12, "cores": threads,
                                         "protocol": "8 of the 1024 slices timed (best), extrapolated: %.1f s for config 5" % (t * 1024)}
    return out, {"kind": kind, "threads": threads, "blas_core": info.get("core", ""), "blas": info.get("config", ""), "nproc": os.cpu_count()}


def run_single(args):
    peaks = load_peaks()
    traffic = load_ncu_traffic()
    B = Bench(0)
    torch, lib = B.torch, B.lib
    n = MATMUL_N
    # pinned host buffers of the e2e arm, allocated first (see the note on the e2e spread before the e2e loop)
    ha, hb, hc = (torch.empty(n, n, dtype=torch.float32, pin_memory=True) for _ in range(3))
    g = torch.Generator(device="cuda").manual_seed(3)
    a = torch.rand(n, n, device="cuda", generator=g)
    b = torch.rand(n, n, device="cuda", generator=g)
    c = torch.empty(n, n, device="cuda")
    flops = 2.0 * n ** 3
    host = {}

    def mm(prec):
        B.check(lib.nb200_sgemm(c.data_ptr(), a.data_ptr(), b.data_ptr(), n, n, n, n, n, n, prec))

    sampler = ClockSampler(0)
    sampler.start()
    time.sleep(0.6)                 # nvidia-smi start-up
    sampler.mark_start()
    auto_mode = int(lib.nb200_gemm_resolve_precision(GEMM_AUTO, n))        # the mode nd::matmul runs at this K
    auto_name, auto_dtype, auto_kind = MODES[auto_mode]
    for _ in range(args.warmup):
        mm(GEMM_AUTO)
    torch.cuda.synchronize()
    launches0 = lib.nb200_launch_count()
    ms = B.time_steps(lambda: mm(GEMM_AUTO), args.steps, 0)
    launches_timed = lib.nb200_launch_count() - launches0
    mode_ms = {auto_name: ms}
    for name, prec in (("tf32x3", 0), ("bf16x3", 2), ("fp16x3", 4), ("fp16x3u", 5), ("tf32x1", 1)):
        if name not in mode_ms:
            mode_ms[name] = B.time_steps(lambda: mm(prec), args.steps, args.warmup)
    # accuracy of each mode against an fp64 product of 64 sampled rows (reported, the parity tests assert it)
    rows = torch.randperm(n, device="cuda", generator=g)[:64]
    truth = a[rows].double() @ b.double()
    mode_err = {}
    for name, prec in (("tf32x3", 0), ("bf16x3", 2), ("fp16x3", 4), ("fp16x3u", 5), ("tf32x1", 1)):
        mm(prec)
        mode_err[name] = float(((c[rows].double() - truth) / truth).abs().max())
    del truth
    # The sampler stops HERE: its window covers the timed matmul loop and the other-mode loops (the headline loop alone is shorter
    # than one nvidia-smi period); the e2e loop below is PCIe-bound and is not the timed region of `value`.
    # (The e2e figure varies from PROCESS to process for the same build: 2.81 / 3.25 / 3.04 / 3.29 / 2.95 / 2.87 ms in six bench.py runs
    # on six boxes, while scripts/duplex_probe.py - the same call in a process of its own - measured 2.77-2.80 ms in every one of ~30
    # processes on the same kind of box, including right before and after a 2.95 ms bench run.  A process is fast or slow for all of
    # its calls.  Ruled out in isolation: the nvidia-smi sampler, the legacy default stream, other streams of the process, the order
    # of the measurements, NUMA placement (one node).  Open: profiles/r2_summary.md section 5e.)
    sampler.mark_end()
    clocks = sampler.stop()

    # ---- e2e: pinned host buffers, H2D(A,B) + matmul + D2H(C) per step, through the C-ABI
    ha.copy_(a, non_blocking=False); hb.copy_(b, non_blocking=False)
    host["mm_a"], host["mm_b"] = ha.numpy(), hb.numpy()
    nbytes = n * n * 4

    def e2e_step():
        # the public host-operand call: uploads B, streams row blocks of A in / C out around the tcgen05 GEMM
        B.check(lib.nb200_sgemm_host(hc.data_ptr(), ha.data_ptr(), hb.data_ptr(), n, n, n, GEMM_AUTO))

    e2e_steps = max(3, min(args.steps, 10))
    ms_e2e = B.time_steps(e2e_step, e2e_steps, 3)
    # raw PCIe ceilings for the e2e number: 256 MiB pinned copies, each direction alone and both at once
    pin = torch.empty(64 << 20, dtype=torch.float32).pin_memory()
    pin2 = torch.empty(64 << 20, dtype=torch.float32).pin_memory()
    dbuf = torch.empty(64 << 20, dtype=torch.float32, device="cuda")
    dbuf2 = torch.empty(64 << 20, dtype=torch.float32, device="cuda")
    t_h2d = B.time_steps(lambda: dbuf.copy_(pin, non_blocking=True), 5, 2)
    t_d2h = B.time_steps(lambda: pin.copy_(dbuf, non_blocking=True), 5, 2)
    side = torch.cuda.Stream()

    def duplex():
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            pin2.copy_(dbuf2, non_blocking=True)
        dbuf.copy_(pin, non_blocking=True)
        torch.cuda.current_stream().wait_stream(side)
    t_dup = B.time_steps(duplex, 5, 2)
    gbs = pin.numel() * 4 / 1e6
    pcie = {"h2d_GBps": gbs / t_h2d, "d2h_GBps": gbs / t_d2h, "duplex_each_GBps": gbs / t_dup}
    del pin, pin2, dbuf, dbuf2
    mm(GEMM_AUTO)                                                    # resident result of the same mode for comparison
    torch.cuda.synchronize()
    e2e_err = float(((hc.cuda() - c).abs() / c.abs()).max())       # host pipeline (TF32x3 kernels) vs resident AUTO: both inside 1e-5

    # ---- the other BASELINE configs: HBM-bound.  Two timings each: `ms` = launches back to back over NSETS distinct operand sets
    # (>= 1 GiB streamed between two uses of a set, 8x the 126 MB L2: nothing can be cached; one event pair around all launches),
    # `ms_isolated` = one launch between its own event pair after an explicit L2 flush (adds the ~5 us event / drain gap per launch).
    hbm = peaks["hbm_gbs"]
    per = {}
    m = 8192
    NSETS = 4
    xs = [torch.rand(m, m, device="cuda", generator=g) for _ in range(NSETS)]
    ys = [torch.rand(m, m, device="cuda", generator=g) for _ in range(NSETS)]
    zs = [torch.rand(m, m, device="cuda", generator=g) for _ in range(NSETS)]
    outs = [torch.empty(m, m, device="cuda") for _ in range(NSETS)]
    tmps = [torch.empty(m, m, device="cuda") for _ in range(NSETS)]
    x, y, z, out, tmp = xs[0], ys[0], zs[0], outs[0], tmps[0]
    shp = (C.c_int64 * 2)(m, m)
    full = (C.c_int64 * 2)(m, 1)
    rowv = (C.c_int64 * 2)(0, 1)
    colv = (C.c_int64 * 2)(1, 0)
    reps = 10

    def rec_hbm(name, ms_, bytes_, config, kernels, ms_iso=None):
        tb, src, _ = ncu_lookup(traffic, *kernels, near=bytes_) if kernels else (None, None, None)
        per[name] = {"config": config, "bound": "hbm", "ms": ms_, "achieved": bytes_ / ms_ / 1e6, "peak": hbm, "unit": "GB/s",
                     "frac": bytes_ / ms_ / 1e6 / hbm, "algorithmic_bytes": bytes_, "traffic": tb, "traffic_source": src,
                     "timing": f"back to back over {NSETS} distinct operand sets (inputs >> L2), one event pair"}
        if ms_iso is not None:
            per[name]["ms_isolated"] = ms_iso
            per[name]["frac_isolated"] = bytes_ / ms_iso / 1e6 / hbm

    def both(fn_i, rounds=5):
        """(steady-state ms over the rotating sets, isolated ms after an L2 flush)"""
        return B.time_rotating(fn_i, NSETS, rounds), B.time_steps(lambda: fn_i(0), reps, 3, flush=True)

    t, ti = both(lambda i: B.check(lib.nb200_ew_mul_add(outs[i].data_ptr(), xs[i].data_ptr(), ys[i].data_ptr(), zs[i].data_ptr(), 2, shp, full, full, full)))
    rec_hbm("chain_fused_8192sq", t, 4 * m * m * 4, "configs[2] a*b+c 8192^2, one fused call (nb200_ew_mul_add)", [("ew_flat_vec", "MulAdd")], ti)

    def unfused(i):
        B.check(lib.nb200_ew_binary(2, tmps[i].data_ptr(), xs[i].data_ptr(), ys[i].data_ptr(), 2, shp, full, full))
        B.check(lib.nb200_ew_binary(0, outs[i].data_ptr(), tmps[i].data_ptr(), zs[i].data_ptr(), 2, shp, full, full))
    t, ti = both(unfused)
    rec_hbm("chain_two_calls_8192sq", t, 6 * m * m * 4, "configs[2] a*b+c 8192^2 as the two nd:: calls unchanged PHP makes", None, ti)
    t, ti = both(lambda i: B.check(lib.nb200_ew_mul_add(outs[i].data_ptr(), xs[i].data_ptr(), ys[i].data_ptr(), zs[i].data_ptr(), 2, shp, full, rowv, colv)))
    rec_hbm("chain_fused_row_col_broadcast_8192sq", t, 2 * m * m * 4 + 2 * m * 4, "configs[2] broadcast variant: b row vector, c column vector", [("ew_bcast2d", "MulAdd")], ti)
    t, ti = both(lambda i: B.check(lib.nb200_ew_unary(2, outs[i].data_ptr(), xs[i].data_ptr(), m * m, 0.0, 0.0)))
    rec_hbm("unary_exp_8192sq", t, 2 * m * m * 4, "nd::exp 8192^2 (math unary)", None, ti)
    res = torch.empty(16, device="cuda")
    ax = torch.empty(m, device="cuda")
    # (the 8192^2 reductions read one 256 MiB array per launch: all four operand arrays of every set serve as inputs, 16 x 256 MiB)
    red_in = xs + ys + zs + outs
    for o_ in outs:
        o_.copy_(xs[0])
    t = B.time_rotating(lambda i: B.check(lib.nb200_reduce_axis(0, ax.data_ptr(), red_in[i].data_ptr(), 1, m, m, 0)), len(red_in), 2)
    ti = B.time_steps(lambda: B.check(lib.nb200_reduce_axis(0, ax.data_ptr(), x.data_ptr(), 1, m, m, 0)), reps, 3, flush=True)
    rec_hbm("sum_axis0_8192sq", t, m * m * 4, "nd::sum(axis=0) 8192^2 (axis reduction, north_star)", [("reduce_cols_kernel",)], ti)
    t = B.time_rotating(lambda i: B.check(lib.nb200_reduce_axis(0, ax.data_ptr(), red_in[i].data_ptr(), m, m, 1, 0)), len(red_in), 2)
    ti = B.time_steps(lambda: B.check(lib.nb200_reduce_axis(0, ax.data_ptr(), x.data_ptr(), m, m, 1, 0)), reps, 3, flush=True)
    rec_hbm("sum_axis1_8192sq", t, m * m * 4, "nd::sum(axis=1) 8192^2 (axis reduction, north_star)", [("reduce_rows_kernel",)], ti)
    for k_ in ("sum_axis0_8192sq", "sum_axis1_8192sq"):
        per[k_]["timing"] = f"back to back over {len(red_in)} distinct 256 MiB inputs (4 GiB >> L2), one event pair"
    host["chain_x"], host["chain_y"], host["chain_z"] = x.cpu().numpy(), y.cpu().numpy(), z.cpu().numpy()
    del y, z, out, tmp, x, xs, ys, zs, outs, tmps, red_in
    bigs = [torch.rand(1 << 28, device="cuda", generator=g) for _ in range(NSETS)]
    big = bigs[0]
    t = B.time_rotating(lambda i: B.check(lib.nb200_reduce_full(0, res.data_ptr(), bigs[i].data_ptr(), 1 << 28)), NSETS, 3)
    ti = B.time_steps(lambda: B.check(lib.nb200_reduce_full(0, res.data_ptr(), big.data_ptr(), 1 << 28)), reps, 3, flush=True)
    rec_hbm("sum_2pow28", t, 1 << 30, "configs[3] nd::sum over 2^28", [("reduce_rows_kernel",)], ti)
    gpu_sum = float(res[0].item())
    t = B.time_rotating(lambda i: B.check(lib.nb200_argminmax(1, res.data_ptr(), bigs[i].data_ptr(), 1, 1 << 28, 1)), NSETS, 3)
    B.check(lib.nb200_argminmax(1, res.data_ptr(), big.data_ptr(), 1, 1 << 28, 1))
    ti = B.time_steps(lambda: B.check(lib.nb200_argminmax(1, res.data_ptr(), big.data_ptr(), 1, 1 << 28, 1)), reps, 3, flush=True)
    rec_hbm("argmax_2pow28", t, 1 << 30, "configs[3] nd::argmax over 2^28", [("arg_rows_kernel",)], ti)
    per["sum_2pow28"]["result"] = gpu_sum
    per["sum_2pow28"]["fp64_truth"] = float(big.double().sum().item())
    # the 8192^2 axis sums are only 256 MiB (~45 us): launch + fold latency is visible, so also the 1 GiB shape
    ax2 = torch.empty(32768, device="cuda")
    t = B.time_rotating(lambda i: B.check(lib.nb200_reduce_axis(0, ax2.data_ptr(), bigs[i].data_ptr(), 1, 32768, 8192, 0)), NSETS, 3)
    ti = B.time_steps(lambda: B.check(lib.nb200_reduce_axis(0, ax2.data_ptr(), big.data_ptr(), 1, 32768, 8192, 0)), reps, 3, flush=True)
    rec_hbm("sum_axis0_32768x8192", t, 1 << 30, "nd::sum(axis=0) 32768x8192 (1 GiB)", None, ti)
    t = B.time_rotating(lambda i: B.check(lib.nb200_reduce_axis(0, ax2.data_ptr(), bigs[i].data_ptr(), 32768, 8192, 1, 0)), NSETS, 3)
    ti = B.time_steps(lambda: B.check(lib.nb200_reduce_axis(0, ax2.data_ptr(), big.data_ptr(), 32768, 8192, 1, 0)), reps, 3, flush=True)
    rec_hbm("sum_axis1_32768x8192", t, 1 << 30, "nd::sum(axis=1) 32768x8192 (1 GiB)", None, ti)
    host["big"] = big.cpu().numpy()
    del big, bigs
    # configs[0]: nd::add 1024x1024 (launch-latency bound on a GPU: 12 MiB of traffic, L2-resident)
    s = torch.rand(1024, 1024, device="cuda"); s2 = torch.rand(1024, 1024, device="cuda"); so = torch.empty(1024, 1024, device="cuda")
    s1 = (C.c_int64 * 1)(1 << 20); st1 = (C.c_int64 * 1)(1)
    t = B.time_steps(lambda: B.check(lib.nb200_ew_binary(0, so.data_ptr(), s.data_ptr(), s2.data_ptr(), 1, s1, st1, st1)), 50, 5)
    per["add_1024sq_l2_warm"] = {"config": "configs[0] nd::add 1024x1024 on the GPU, back to back (L2-resident: launch-latency bound, reported, not a target)",
                                 "bound": "launch latency", "ms": t, "achieved": 3 * (1 << 22) / t / 1e6, "unit": "GB/s", "algorithmic_bytes": 3 << 22}
    # the launch-latency path: 50 nd::add calls recorded once (nb200_graph_begin / end) and replayed with one launch
    # (the legacy default stream cannot be captured: record and replay on a side stream, which is also where the events are recorded)
    gexec = C.c_void_p()
    cap_stream = torch.cuda.Stream()
    torch.cuda.synchronize()
    with torch.cuda.stream(cap_stream):
        B.check(lib.nb200_set_stream(C.c_void_p(cap_stream.cuda_stream)))
        B.check(lib.nb200_graph_begin())
        for _ in range(50):
            B.check(lib.nb200_ew_binary(0, so.data_ptr(), s.data_ptr(), s2.data_ptr(), 1, s1, st1, st1))
        B.check(lib.nb200_graph_end(C.byref(gexec)))
        tg = B.time_steps(lambda: B.check(lib.nb200_graph_launch(gexec)), 20, 3) / 50
        B.check(lib.nb200_graph_destroy(gexec))
    torch.cuda.synchronize()
    B.check(lib.nb200_set_stream(C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    per["add_1024sq_graph_replay"] = {"config": "configs[0] nd::add 1024x1024: 50 calls captured into one CUDA graph (nb200_graph_*), per-op time of a replay",
                                      "bound": "launch latency", "ms": tg, "achieved": 3 * (1 << 22) / tg / 1e6, "unit": "GB/s", "algorithmic_bytes": 3 << 22}
    host["add_a"], host["add_b"] = s.cpu().numpy(), s2.cpu().numpy()

    # configs[4]'s per-GPU share on ONE GPU (weak-scaling base for bench.py --gpus N): 128 x (2048x2048)
    ab = torch.rand(SHARD_BATCH, SHARD_N, SHARD_N, device="cuda", generator=g)
    bb_ = torch.rand(SHARD_BATCH, SHARD_N, SHARD_N, device="cuda", generator=g)
    cb = torch.empty(SHARD_BATCH, SHARD_N, SHARD_N, device="cuda")
    sn = SHARD_N
    t = B.time_steps(lambda: B.check(lib.nb200_sgemm_batched(cb.data_ptr(), ab.data_ptr(), bb_.data_ptr(), SHARD_BATCH, sn, sn, sn,
                                                             sn * sn, sn * sn, sn * sn, GEMM_AUTO)), 5, 3)
    bf16_peak = peaks["bf16_tflops"]         # measured cuBLAS bf16 rate: the kind::f16 MMA ceiling
    tf32_peak = bf16_peak / 2.0              # tcgen05 kind::tf32 runs at half the bf16 rate
    sust = peaks.get("bf16_tflops_sustained", bf16_peak) / (2.0 if auto_kind == "tf32" else 1.0)
    per["batched_matmul_128x2048sq_1gpu"] = {"config": "configs[4] per-GPU share (128 of the 1024 matrices) on one GPU, resident", "bound": "tensor", "ms": t,
                                            "achieved": SHARD_BATCH * 2.0 * sn ** 3 / t / 1e9, "peak": sust, "unit": "TFLOP/s",
                                            "frac": SHARD_BATCH * 2.0 * sn ** 3 / t / 1e9 / sust,
                                            "pipe_frac": 3 * SHARD_BATCH * 2.0 * sn ** 3 / t / 1e9 / sust,
                                            "note": "sustained (6+ ms per step): rated against bf16_tflops_sustained"}
    host["bm_a"], host["bm_b"] = ab[0].cpu().numpy(), bb_[0].cpu().numpy()
    del ab, bb_, cb
    auto_peak = tf32_peak if auto_kind == "tf32" else bf16_peak
    useful = flops / ms / 1e9
    mode_notes = {
        "tf32x3": ("NB200_GEMM_TF32X3: three kind::tf32 MMAs per k-step, guaranteed bound; what AUTO runs for K < 128 and the FP16x3 fallback", tf32_peak),
        "bf16x3": ("NB200_GEMM_BF16X3 (opt-in): two bf16 parts per operand, three kind::f16 MMAs; statistical accuracy (zero-mean split "
                   "error: fine on random data, up to ~3e-5 on coherent inputs)", bf16_peak),
        "fp16x3": ("NB200_GEMM_FP16X3: half parts of row/column-scaled operands (22-bit elements inside a 2^28 window, "
                   "guaranteed bound), one persistent pre-pass launch, PDL chain pre-pass -> GEMM -> repair/fallback", bf16_peak),
        "fp16x3u": ("NB200_GEMM_FP16X3U: FP16x3 with unscaled lo parts (split error max(2^-22 |a'|, 2^-25); per-product bound 2^-18 + 2^-22 = 4.1e-6 for "
                    "every input inside the 2^-20 window), one accumulator per 256-long chunk -> the merged 256x256 tile", bf16_peak),
    }
    modes = {}
    for name, (note, pk) in mode_notes.items():
        tm = mode_ms[name]
        modes[name] = {"ms": tm, "useful_tflops": flops / tm / 1e9, "pipe_executed_tflops": 3 * flops / tm / 1e9,
                       "pipe_frac": 3 * flops / tm / 1e9 / pk, "max_rel_err_vs_fp64": mode_err[name], "is_auto": name == auto_name, "note": note}
    modes["tf32x1"] = {"ms": mode_ms["tf32x1"], "useful_tflops": flops / mode_ms["tf32x1"] / 1e9,
                       "frac_tf32_peak": flops / mode_ms["tf32x1"] / 1e9 / tf32_peak, "max_rel_err_vs_fp64": mode_err["tf32x1"],
                       "note": "single-pass TF32 fast mode (not a parity mode)"}
    gemm_kernels = {"fp16x3": [("prep16_coop_kernel",), ("sgemm_tf32_kernel", "GemmCfg<2, 128, 3, 0, 1, 0, 1>")],
                    "fp16x3u": [("prep16_coop_kernel",), ("sgemm_tf32_kernel", "GemmCfg<2, 256, 3, 0, 1, 1, 1, 1>")],
                    "tf32x3": [("split_tf32_kernel",), ("sgemm_tf32_kernel", "GemmCfg<2, 128, 3>")],
                    "bf16x3": [("split_bf16_flat_kernel",), ("sgemm_tf32_kernel", "GemmCfg<2, 256, 3, 0, 1, 1>")]}[auto_name]
    tb, tsrc, pipe_active = ncu_lookup(traffic, *gemm_kernels)
    cpu_per, cpu_info = cpu_baselines_full(host)
    cpu_mm = cpu_per["matmul_4096"]
    for k_gpu, k_cpu, fld in (("chain_two_calls_8192sq", "chain_mul_add_8192sq", "GBps_two_calls"), ("chain_fused_8192sq", "chain_mul_add_8192sq", "GBps_algorithmic_fused"),
                              ("sum_2pow28", "sum_2pow28", "GBps"), ("argmax_2pow28", "argmax_2pow28", "GBps"), ("sum_axis0_8192sq", "sum_axis0_8192sq", "GBps"),
                              ("add_1024sq_l2_warm", "add_1024sq", "GBps")):
        per[k_gpu]["cpu_reference"] = cpu_per[k_cpu][fld]
        per[k_gpu]["speedup_vs_cpu_reference"] = per[k_gpu]["achieved"] / cpu_per[k_cpu][fld]
    per["batched_matmul_128x2048sq_1gpu"]["cpu_reference"] = cpu_per["batched_matmul_1024x2048sq"]["TFLOPs"]

    line = {
        "metric": METRIC, "value": useful, "unit": "TFLOP/s",
        "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": auto_dtype,
        "data": "synthetic",
        "config": config_single(),
        "roofline": {"bound": "tensor", "achieved": useful, "peak": auto_peak, "unit": "TFLOP/s", "frac": useful / auto_peak,
                     "traffic": tb,
                     "traffic_source": (f"dram__bytes_read.sum + dram__bytes_write.sum per launch of the pre-pass + GEMM kernels, read at run time from {tsrc} "
                                        "(ncu --set full of the same call); algorithmic minimum 3 x 64 MiB = 0.201 GB: the GEMM is tensor-bound, re-reads are L2-served")
                                       if tb else "no ncu summary found under profiles/ for this mode: null",
                     "tensor_pipe_active_pct_ncu": pipe_active,
                     "peak_source": peaks["_source"] + (": bf16_tflops / 2 (tf32 = half the bf16 MMA rate)" if auto_kind == "tf32" else ": bf16_tflops (cuBLAS bf16 8192^3, burst)"),
                     "pipe_executed_tflops": 3 * useful, "pipe_frac": 3 * useful / auto_peak,
                     "note": "achieved counts the algorithmic 2*M*N*K flops of the fp32 product; the error-compensated scheme executes 3x that "
                             "on the tensor pipe (pipe_frac), so frac is bounded by 1/3; per_config = every other BASELINE config measured in this run",
                     "workload_detail": f"nb200_sgemm, NB200_GEMM_AUTO (= {auto_name}), max rel err vs fp64 {mode_err[auto_name]:.2e} (tolerance 1e-5); "
                                        "operands 128 MiB + result 64 MiB exceed the 126 MB L2; HBM-bound configs flush L2 between timed launches; CUDA events on the launching stream",
                     "matmul_modes": modes,
                     "per_config": per},
        "cpu_baseline": {"value": cpu_mm["TFLOPs"], "unit": "TFLOP/s", "cores": cpu_mm["cores"], "kind": cpu_info["kind"],
                         "sample": "best of 5 full 4096^3 nd::matmul calls (NDArray_Matmul -> cblas_sgemm)", "ms": cpu_mm["ms"],
                         "blas_core": cpu_info["blas_core"], "blas": cpu_info["blas"], "nproc": cpu_info["nproc"],
                         "per_config": cpu_per},
        "e2e": {"value": flops / ms_e2e / 1e9, "unit": "TFLOP/s", "h2d_bytes_per_step": 2 * nbytes, "d2h_bytes_per_step": nbytes,
                "ms_per_step": ms_e2e, "steps": e2e_steps, "api": "nb200_sgemm_host (pinned host buffers, pipelined H2D/compute/D2H)",
                "max_rel_diff_vs_resident": e2e_err, "pcie_measured": pcie,
                "mode": "NB200_GEMM_AUTO (" + ("the host pipeline keeps the TF32x3 kernels: PCIe-bound, per-row-block splits" if auto_name in ("fp16x3", "fp16x3u") else auto_name) + ")",
                "pcie_bound_ms": 2 * nbytes / pcie["h2d_GBps"] / 1e6,
                "pcie_bound_duplex_ms": nbytes / pcie["h2d_GBps"] / 1e6 + nbytes / pcie["duplex_each_GBps"] / 1e6,
                "frac_of_pcie_bound": (2 * nbytes / pcie["h2d_GBps"] / 1e6) / ms_e2e,
                "frac_of_pcie_bound_duplex": (nbytes / pcie["h2d_GBps"] / 1e6 + nbytes / pcie["duplex_each_GBps"] / 1e6) / ms_e2e,
                "note": "H2D of A and B (128 MiB) is the floor; the D2H of C overlaps the second half of it, where the link runs at its measured "
                        "duplex rate (pcie_bound_duplex_ms)"},
        "gpu_launches": int(launches_timed),
        "clocks": clocks,
    }
    emit(line)
    return 0


def run_multi(args):
    import datetime
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(minutes=20))
    cpu_pg = dist.new_group(backend="gloo", timeout=datetime.timedelta(minutes=20))   # long waits park on the CPU, not in a spinning NCCL kernel
    peaks = load_peaks()
    B = Bench(local)
    lib = B.lib
    nb_, n = SHARD_BATCH, SHARD_N
    g = torch.Generator(device="cuda").manual_seed(10 + rank)   # resident shard, generated on the owning GPU
    a = torch.rand(nb_, n, n, device="cuda", generator=g)
    b = torch.rand(nb_, n, n, device="cuda", generator=g)
    c = torch.empty(nb_, n, n, device="cuda")
    flops_rank = nb_ * 2.0 * n ** 3

    def step():
        B.check(lib.nb200_sgemm_batched(c.data_ptr(), a.data_ptr(), b.data_ptr(), nb_, n, n, n, n * n, n * n, n * n, GEMM_AUTO))

    def device_barrier():
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()                 # (nvidia-smi needs ~0.5 s to start: the window is marked around the timed region below)
    for _ in range(args.warmup):
        step()
    # ---- scaling base: the same per-GPU workload on ONE GPU of this box while the others idle (rank 0; same steps)
    device_barrier()
    base_ms = None
    if rank == 0:
        base_ms = B.time_steps(step, args.steps, 1)
    dist.barrier(group=cpu_pg)
    # ---- the timed region: every rank multiplies its resident shard
    for _ in range(2):
        step()
    device_barrier()
    launches0 = lib.nb200_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.mark_start()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    device_barrier()
    sampler.mark_end()
    my_ms = e0.elapsed_time(e1) / args.steps
    clocks = sampler.stop()
    launches = lib.nb200_launch_count() - launches0
    t = torch.tensor([my_ms], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    per_rank = [None] * world
    dist.all_gather_object(per_rank, {"rank": rank, "gpu": local, "ms_per_step": my_ms, "clocks": clocks}, group=cpu_pg)

    # ---- sharded HBM-bound configs, weak scaling: every rank runs the op on its resident shard (no data-path collective; the
    #      full reductions add one 4-byte-per-rank combine, see numpower_b200/sharding.py and nb200_shard_reduce_full)
    del c
    hbm = peaks["hbm_gbs"]
    m = 8192
    x = a.view(-1)[: m * m].view(m, m)
    y = b.view(-1)[: m * m].view(m, m)
    z = a.view(-1)[m * m: 2 * m * m].view(m, m)
    out = torch.empty(m, m, device="cuda")
    shp = (C.c_int64 * 2)(m, m)
    full = (C.c_int64 * 2)(m, 1)
    big = b.view(-1)[: 1 << 28]
    res = torch.empty(16, device="cuda")
    sharded = {}
    for name, fn, bytes_ in (
        ("chain_fused_8192sq_per_gpu", lambda: B.check(lib.nb200_ew_mul_add(out.data_ptr(), x.data_ptr(), y.data_ptr(), z.data_ptr(), 2, shp, full, full, full)), 4 * m * m * 4),
        ("sum_2pow28_per_gpu", lambda: B.check(lib.nb200_reduce_full(0, res.data_ptr(), big.data_ptr(), 1 << 28)), 1 << 30),
        ("argmax_2pow28_per_gpu", lambda: B.check(lib.nb200_argminmax(1, res.data_ptr(), big.data_ptr(), 1, 1 << 28, 1)), 1 << 30),
    ):
        device_barrier()
        tt = B.time_steps(fn, 10, 3, flush=True)
        tm = torch.tensor([tt], device="cuda")
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        tmax = float(tm.item())
        sharded[name] = {"ms_max_over_ranks": tmax, "aggregate_GBps": world * bytes_ / tmax / 1e6, "per_gpu_GBps": bytes_ / tmax / 1e6,
                         "frac_hbm_per_gpu": bytes_ / tmax / 1e6 / hbm, "scaling": "weak", "collective": "none" if "chain" in name else "4-byte partial per rank"}
    del out

    # ---- e2e: each rank feeds matrices of its shard from pinned host memory through the chunked three-stream pipeline
    e2e_batch = 16
    ha = torch.empty(e2e_batch, n, n).pin_memory(); hb = torch.empty(e2e_batch, n, n).pin_memory(); hc = torch.empty(e2e_batch, n, n).pin_memory()
    ha.copy_(a[:e2e_batch].cpu()); hb.copy_(b[:e2e_batch].cpu())
    nbytes = e2e_batch * n * n * 4

    def e2e_step():
        B.check(lib.nb200_sgemm_batched_host(hc.data_ptr(), ha.data_ptr(), hb.data_ptr(), e2e_batch, n, n, n, GEMM_AUTO))

    e2e_step()
    device_barrier()
    e0.record()
    for _ in range(3):
        e2e_step()
    e1.record()
    device_barrier()
    my_e2e = e0.elapsed_time(e1) / 3
    # PCIe floor of THIS rank while every rank copies at once (host memory / root complexes are shared)
    dbuf = torch.empty(64 << 20, dtype=torch.float32, device="cuda")
    pin = ha.view(-1)[: 64 << 20]
    device_barrier()
    t_h2d = B.time_steps(lambda: dbuf.copy_(pin, non_blocking=True), 3, 1)
    device_barrier()
    h2d_gbps = pin.numel() * 4 / t_h2d / 1e6
    # ... and with the D2H of C running against it (the pipeline's steady state: both directions of the link busy)
    dbuf2 = torch.empty(64 << 20, dtype=torch.float32, device="cuda")
    pin2 = hc.view(-1)[: 64 << 20]
    side = torch.cuda.Stream()

    def duplex():
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            pin2.copy_(dbuf2, non_blocking=True)
        dbuf.copy_(pin, non_blocking=True)
        torch.cuda.current_stream().wait_stream(side)
    device_barrier()
    t_dup = B.time_steps(duplex, 3, 1)
    device_barrier()
    dup_gbps = pin.numel() * 4 / t_dup / 1e6
    # floor of one step: the D2H bytes overlap as many H2D bytes at the duplex rate, the remaining H2D bytes run at the one-way rate
    floor_duplex_ms = nbytes / dup_gbps / 1e6 + nbytes / h2d_gbps / 1e6
    del dbuf, dbuf2
    t2 = torch.tensor([my_e2e], device="cuda")
    dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    ms_e2e = float(t2.item())
    e2e_rank = [None] * world
    dist.all_gather_object(e2e_rank, {"rank": rank, "ms_per_step": my_e2e, "h2d_GBps_under_contention": h2d_gbps,
                                      "duplex_each_GBps_under_contention": dup_gbps,
                                      "pcie_bound_ms": 2 * nbytes / h2d_gbps / 1e6, "frac_of_pcie_bound": (2 * nbytes / h2d_gbps / 1e6) / my_e2e,
                                      "pcie_bound_duplex_ms": floor_duplex_ms, "frac_of_pcie_bound_duplex": floor_duplex_ms / my_e2e}, group=cpu_pg)
    del a, b, ha, hb, hc
    torch.cuda.empty_cache()

    # ---- scatter + compute + gather over NVLink: ONE process (rank 0) drives all N GPUs through nb200_shard_* while the other
    #      ranks park on a CPU barrier.  Operands and result live on GPU 0; batch = 128 x N (N = 8: BASELINE configs[4]).
    sg = None
    if rank == 0:
        try:
            sg = scatter_gather_bench(B, world)
        except Exception as ex:  # reported, never fatal for the headline
            sg = {"error": str(ex)[:300]}
    dist.barrier(group=cpu_pg)

    if rank == 0:
        auto_name, auto_dtype, auto_kind = MODES[int(lib.nb200_gemm_resolve_precision(GEMM_AUTO, n))]
        peak = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]) / (2.0 if auto_kind == "tf32" else 1.0)
        total = world * flops_rank / ms / 1e9
        per_gpu = flops_rank / ms / 1e9
        base_tflops = flops_rank / base_ms / 1e9
        slow = max(per_rank, key=lambda r: r["ms_per_step"])
        line = {
            "metric": METRIC, "value": total, "unit": "TFLOP/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": auto_dtype,
            "data": "synthetic",
            "config": config_multi(world),
            "scaling_base": {"workload": "the same 128 x (2048x2048) per-GPU share on ONE GPU of this box, other GPUs idle, same run, same steps",
                             "ms_per_step": base_ms, "value": base_tflops, "unit": "TFLOP/s",
                             "speedup_vs_base": total / base_tflops, "efficiency_vs_base": total / base_tflops / world},
            "per_rank": per_rank,
            "limiter": {"slowest_rank": slow["rank"], "slowest_ms": slow["ms_per_step"], "fastest_ms": min(r["ms_per_step"] for r in per_rank),
                        "slowest_rank_clocks": slow["clocks"],
                        "note": "no collective in the timed region: the max-over-ranks time is the slowest GPU's sustained tensor-pipe rate under its power cap "
                                "(per_rank carries every rank's time, SM clocks, power and throttle reasons)"},
            "roofline": {"bound": "tensor", "achieved": per_gpu, "peak": peak, "unit": "TFLOP/s", "frac": per_gpu / peak,
                         "traffic": None, "peak_source": peaks["_source"] + (": bf16_tflops_sustained / 2" if auto_kind == "tf32" else ": bf16_tflops_sustained"),
                         "pipe_executed_tflops": 3 * per_gpu, "pipe_frac": 3 * per_gpu / peak,
                         "note": "per-GPU figures; frac counts the algorithmic flops (bounded by 1/3), pipe_frac the executed ones",
                         "workload_detail": f"resident shards, no data-path collective; per-rank operands {2 * nb_ * n * n * 4 >> 20} MiB exceed L2; CUDA events, max over ranks",
                         "per_config": sharded},
            "e2e": {"value": world * e2e_batch * 2.0 * n ** 3 / ms_e2e / 1e9, "unit": "TFLOP/s", "h2d_bytes_per_step": 2 * nbytes,
                    "d2h_bytes_per_step": nbytes, "ms_per_step": ms_e2e, "api": "nb200_sgemm_batched_host (chunked H2D / compute / D2H on three streams per rank)",
                    "per_rank": e2e_rank,
                    "note": f"{e2e_batch} matrices per rank per step from pinned host buffers; every rank shares the host's memory system and PCIe root complexes, so the "
                            "per-rank floor is measured with all ranks copying at once (h2d_GBps_under_contention)"},
            "scatter_gather": sg,
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        emit(line)
    dist.destroy_process_group()
    return 0


def scatter_gather_bench(B, world):
    """configs[4] with operands and result on ONE GPU: nb200_sgemm_batched_scatter_gather over all `world` GPUs, NCCL vs P2P."""
    torch, lib = B.torch, B.lib
    n, batch = SHARD_N, SHARD_BATCH * world
    devs = (C.c_int * world)(*range(world))
    B.check(lib.nb200_shard_init(world, devs))
    g = torch.Generator(device="cuda").manual_seed(77)
    A = torch.rand(batch, n, n, device="cuda", generator=g)
    Bm = torch.rand(batch, n, n, device="cuda", generator=g)
    Cc = torch.empty(batch, n, n, device="cuda")
    out = {"batch": batch, "matrix": n, "root": 0, "chunk": 8, "nvlink_GBps_per_direction": NVLINK_GBS_PER_DIR}
    egress = 2.0 * (world - 1) / world * batch * n * n * 4      # A and B blocks leaving the root
    ingress = 1.0 * (world - 1) / world * batch * n * n * 4     # C blocks coming back
    idx = [0, batch // 2 - 1, batch // 2, batch - 1]
    truth = [A[i].double() @ Bm[i].double() for i in idx]
    tmp = torch.empty(batch, n, n, device="cuda")        # target of the link-only gather (keeps Cc intact)
    for name, transport in (("nccl", 0), ("p2p", 1)):
        ms = C.c_float()
        best = 1e30
        Cc.fill_(float("nan"))
        for _ in range(3):      # first call warms the channels / peer mappings
            B.check(lib.nb200_sgemm_batched_scatter_gather(Cc.data_ptr(), A.data_ptr(), Bm.data_ptr(), batch, n, n, n, GEMM_AUTO, 0, transport, 8, C.byref(ms)))
            best = min(best, ms.value)
        # parity of the pipeline result: sampled matrices (root's share, both sides of a shard boundary, the last one) vs fp64
        err = max(float(((Cc[i].double() - tr) / tr).abs().max()) for i, tr in zip(idx, truth))
        # the link alone: scatter of A (no compute) and gather of C
        shards = [torch.empty(((batch // world) + 1) * n * n, device=f"cuda:{d}") for d in range(world)]
        ptrs = (C.c_void_p * world)(*[s.data_ptr() for s in shards])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        B.check(lib.nb200_shard_scatter(ptrs, A.data_ptr(), batch, n * n, 0, transport)); B.check(lib.nb200_shard_synchronize())
        e0.record(); B.check(lib.nb200_shard_scatter(ptrs, A.data_ptr(), batch, n * n, 0, transport)); e1.record(); B.check(lib.nb200_shard_synchronize())
        torch.cuda.synchronize()
        t_sc = e0.elapsed_time(e1)
        e0.record(); B.check(lib.nb200_shard_gather(tmp.data_ptr(), ptrs, batch, n * n, 0, transport)); e1.record(); B.check(lib.nb200_shard_synchronize())
        torch.cuda.synchronize()
        t_ga = e0.elapsed_time(e1)
        del shards
        one = (world - 1) / world * batch * n * n * 4
        out[name] = {"ms": best, "max_rel_err_vs_fp64_sampled": err, "useful_tflops": batch * 2.0 * n ** 3 / best / 1e9,
                     "root_egress_GBps": egress / best / 1e6, "root_ingress_GBps": ingress / best / 1e6,
                     "egress_frac_of_nvlink": egress / best / 1e6 / NVLINK_GBS_PER_DIR,
                     "scatter_only": {"ms": t_sc, "root_egress_GBps": one / t_sc / 1e6, "frac_of_nvlink": one / t_sc / 1e6 / NVLINK_GBS_PER_DIR},
                     "gather_only": {"ms": t_ga, "root_ingress_GBps": one / t_ga / 1e6, "frac_of_nvlink": one / t_ga / 1e6 / NVLINK_GBS_PER_DIR}}
    out["note"] = ("single host process, nb200_shard_* C-ABI; the pipelined call is link-bound (SURVEY F10): its egress figure divides the A+B bytes by the whole "
                   "scatter+compute+gather time; scatter_only / gather_only time the link alone")
    del A, Bm, Cc, tmp, truth
    B.check(lib.nb200_shard_finalize())
    B.check(lib.nb200_set_device(0))
    return out


_REAL_STDOUT = None


def guard_stdout():
    """stdout carries exactly ONE line, the JSON result.  Libraries that write to file descriptor 1 (NCCL prints its
    version banner there when NCCL_DEBUG is set on the box) are diverted to stderr; emit() writes to the saved descriptor."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = sys.stderr


def emit(line: dict):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    guard_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        return run_multi(args)
    return run_single(args)


if __name__ == "__main__":
    sys.exit(main())
