"""nb200_sgemm_host 4096^2 from pinned host buffers, one child process per environment (profiles/r2_summary.md sections 5d, 5e).
Switches of the child (env): NB200_HOST_* (host_pipeline.cu), PROBE_N, PROBE_BATCHED=<n> (nb200_sgemm_batched_host, n x 2048^2),
PROBE_EXTRA_STREAMS=<n>, PROBE_NULL_STREAM=1 (enqueue on torch's current = legacy default stream, as bench.py does), PROBE_SMI=<ms>
(an nvidia-smi poller beside the calls), PROBE_BENCHLIKE=1 (150 resident products in every mode first, then CUDA-event timing of
10 calls), PROBE_AFFINITY=<numa node>.  `--pure`: chunked duplex copies on torch streams first.  Output: JSON lines on stdout."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _node_cpus():
    import glob
    nodes = {}
    for d in sorted(glob.glob("/sys/devices/system/node/node[0-9]*")):
        cpus = set()
        for part in open(d + "/cpulist").read().strip().split(","):
            if not part:
                continue
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        nodes[int(d.rsplit("node", 1)[1])] = cpus
    return nodes


def child():
    aff_note = None
    if os.environ.get("PROBE_AFFINITY"):   # bind the process to one NUMA node's cores BEFORE anything is allocated or pinned
        nodes, allowed = _node_cpus(), os.sched_getaffinity(0)
        want = nodes.get(int(os.environ["PROBE_AFFINITY"]), set()) & allowed
        aff_note = {"nodes": {k: len(v) for k, v in nodes.items()}, "allowed": len(allowed), "bound_to": len(want)}
        if want:
            os.sched_setaffinity(0, want)
    import torch
    import numpower_b200 as nb
    lib = nb.lib()
    assert lib.nb200_init(0) == 0
    n = int(os.environ.get("PROBE_N", "4096"))
    if os.environ.get("PROBE_NULL_STREAM"):   # what bench.py does: the library enqueues on torch's current stream = the legacy default stream
        import ctypes
        assert lib.nb200_set_stream(ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)) == 0
    extra = [torch.cuda.Stream() for _ in range(int(os.environ.get("PROBE_EXTRA_STREAMS", "0")))]   # other streams of the process (hardware queue aliasing?)
    for st in extra:
        with torch.cuda.stream(st):
            torch.zeros(16, device="cuda").add_(1.0)
    torch.cuda.synchronize()
    smi = None
    if os.environ.get("PROBE_SMI"):   # bench.py's clock sampler running beside the calls: do its NVML queries stall the enqueues?
        smi = subprocess.Popen(["nvidia-smi", "--query-gpu=timestamp,clocks.sm,power.draw,clocks_event_reasons.active", "--format=csv,noheader,nounits",
                                "-i", "0", "-lms", os.environ["PROBE_SMI"]], stdout=subprocess.DEVNULL)
        time.sleep(0.8)
    import atexit
    atexit.register(lambda: smi and smi.terminate())
    if os.environ.get("PROBE_BATCHED"):
        nb_, m = int(os.environ["PROBE_BATCHED"]), 2048
        hA, hB, hC = (torch.rand(nb_, m, m).pin_memory() for _ in range(3))
        for _ in range(3):
            assert lib.nb200_sgemm_batched_host(hC.data_ptr(), hA.data_ptr(), hB.data_ptr(), nb_, m, m, m, 3) == 0
        ts = []
        for _ in range(8):
            t0 = time.perf_counter()
            lib.nb200_sgemm_batched_host(hC.data_ptr(), hA.data_ptr(), hB.data_ptr(), nb_, m, m, m, 3)
            ts.append((time.perf_counter() - t0) * 1e3)
        ref = (hA[nb_ - 1].double()[:64] @ hB[nb_ - 1].double()).float()
        err = float(((hC[nb_ - 1][:64] - ref).abs() / ref.abs()).max())
        ts.sort()
        print(json.dumps({"env": {k: v for k, v in os.environ.items() if k.startswith("NB200_HOST") or k.startswith("PROBE")}, "batched": nb_, "ms_min": ts[0],
                          "ms_median": ts[len(ts) // 2], "h2d_floor_ms_at_55GBps": nb_ * m * m * 8 / 55e6, "max_rel_err": err}))
        return
    ha, hb, hc = (torch.rand(n, n).pin_memory() for _ in range(3))
    if os.environ.get("PROBE_BENCHLIKE"):   # what bench.py does before its e2e loop: a burst of resident matmuls in every mode
        a = torch.rand(n, n, device="cuda"); b = torch.rand(n, n, device="cuda"); c = torch.empty(n, n, device="cuda")
        for prec in (3, 0, 2, 4, 5, 1):
            for _ in range(25):
                assert lib.nb200_sgemm(c.data_ptr(), a.data_ptr(), b.data_ptr(), n, n, n, n, n, n, prec) == 0
        torch.cuda.synchronize()
        ha.copy_(a.cpu()); hb.copy_(b.cpu())
    for _ in range(3 if os.environ.get("PROBE_BENCHLIKE") else 4):
        assert lib.nb200_sgemm_host(hc.data_ptr(), ha.data_ptr(), hb.data_ptr(), n, n, n, 3) == 0
    if os.environ.get("PROBE_BENCHLIKE"):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            lib.nb200_sgemm_host(hc.data_ptr(), ha.data_ptr(), hb.data_ptr(), n, n, n, 3)
        e1.record(); torch.cuda.synchronize()
        print(json.dumps({"benchlike_event_ms_per_call": e0.elapsed_time(e1) / 10}))
    ts = []
    for _ in range(12):
        t0 = time.perf_counter()
        lib.nb200_sgemm_host(hc.data_ptr(), ha.data_ptr(), hb.data_ptr(), n, n, n, 3)
        ts.append((time.perf_counter() - t0) * 1e3)
    ref = (ha.double()[:64] @ hb.double()).float()
    err = float(((hc[:64] - ref).abs() / ref.abs()).max())
    ts.sort()
    print(json.dumps({"env": {k: v for k, v in os.environ.items() if k.startswith("NB200_HOST") or k.startswith("PROBE") or k == "CUDA_DEVICE_MAX_CONNECTIONS"}, "n": n, "ms_min": ts[0], "ms_median": ts[len(ts) // 2], "ms_mean": sum(ts) / len(ts), "affinity": aff_note, "ms_all": [round(t, 3) for t in ts], "max_rel_err_rows0_63": err}))


def pure_copies():
    import torch
    tot = 16 << 20   # floats = 64 MiB
    pin_in = torch.empty(tot, dtype=torch.float32).pin_memory(); pin_in.fill_(1.0)
    pin_out = torch.empty(tot, dtype=torch.float32).pin_memory()
    d_in = torch.empty(tot, dtype=torch.float32, device="cuda")
    d_out = torch.ones(tot, dtype=torch.float32, device="cuda")
    for nstreams in (1, 2):
        s_in = [torch.cuda.Stream() for _ in range(nstreams)]
        s_out = [torch.cuda.Stream() for _ in range(nstreams)]
        for chunk_mb in (2, 4, 8, 16, 64):
            c = (chunk_mb << 20) // 4
            nchunk = tot // c

            def go(gated=True, do_in=True, do_out=True):
                evs = []
                if do_in:
                    for i in range(nchunk):
                        s = s_in[i % nstreams]
                        with torch.cuda.stream(s):
                            d_in[i * c:(i + 1) * c].copy_(pin_in[i * c:(i + 1) * c], non_blocking=True)
                            e = torch.cuda.Event(); e.record(s); evs.append(e)
                if do_out:
                    for i in range(nchunk):
                        s = s_out[i % nstreams]
                        if gated and do_in:
                            s.wait_event(evs[i])
                        with torch.cuda.stream(s):
                            pin_out[i * c:(i + 1) * c].copy_(d_out[i * c:(i + 1) * c], non_blocking=True)

            res = {"part": "pure_copies", "streams_per_direction": nstreams, "chunk_MiB": chunk_mb}
            for name, kw in (("in_only", dict(do_out=False)), ("out_only", dict(do_in=False)), ("both_gated", dict()), ("both_free", dict(gated=False))):
                go(**kw); torch.cuda.synchronize()
                best = 1e9
                for _ in range(4):
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    go(**kw)
                    torch.cuda.synchronize()
                    best = min(best, time.perf_counter() - t0)
                res[name + "_ms"] = round(best * 1e3, 3)
                res[name + "_GBps_each"] = round(tot * 4 / best / 1e9, 1)
            print(json.dumps(res), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
        sys.exit(0)
    if "--pure" in sys.argv:
        pure_copies()
    bl = {"PROBE_BENCHLIKE": "1", "PROBE_NULL_STREAM": "1"}
    variants = [{}, {"NB200_HOST_WORKERS": "0", "NB200_HOST_BLOCKS": "8"}, {"NB200_HOST_BLOCKS": "8"}, {"NB200_HOST_BLOCKS": "32"}, {"NB200_HOST_WORKERS": "3"},
                {"PROBE_NULL_STREAM": "1"}, {"PROBE_SMI": "20"}, {"PROBE_EXTRA_STREAMS": "8"}, dict(bl), dict(bl), {"PROBE_N": "8192"}, {"PROBE_N": "2048"},
                {"PROBE_BATCHED": "16"}]
    for v in variants:
        env = dict(os.environ); env.update(v)
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=env, capture_output=True, text=True, timeout=300)
        sys.stdout.write(r.stdout)
        if r.returncode != 0:
            print(json.dumps({"env": v, "rc": r.returncode, "stderr": r.stderr[-600:]}))
        sys.stdout.flush()
    # one traced call of the two most interesting variants
    for v in ({}, {"NB200_HOST_WORKERS": "0", "NB200_HOST_BLOCKS": "8"}):
        env = dict(os.environ); env.update(v); env["NB200_HOST_TRACE"] = "1"
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=env, capture_output=True, text=True, timeout=300)
        lines = [ln for ln in r.stderr.splitlines() if ln.startswith("[nb200_sgemm_host]")]
        print(json.dumps({"trace_env": v, "last": lines[-1] if lines else r.stderr[-300:]}), flush=True)
