"""Which host NUMA node feeds GPU 0 fastest?  Pinned 256 MiB buffers allocated under each node's CPU affinity,
H2D / D2H timed with CUDA events.  Writes gpurun_out/numa_probe.json."""
import glob
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def cpulist(s):
    out = []
    for part in s.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        out += list(range(int(a), int(b or a) + 1))
    return out


def main():
    res = {"nodes": {}, "topo": subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout[-1500:]}
    all_cpus = sorted(os.sched_getaffinity(0))
    res["affinity_at_start"] = [all_cpus[0], all_cpus[-1], len(all_cpus)]
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(0)
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        res["nvml_cpu_affinity_words"] = [hex(int(m)) for m in mask]
        try:
            res["nvml_numa_node"] = pynvml.nvmlDeviceGetNumaNodeId(h)
        except Exception as e:  # older nvml
            res["nvml_numa_node"] = str(e)
    except Exception as e:
        res["nvml_error"] = str(e)
    torch.cuda.init()
    dev = torch.empty(64 << 20, dtype=torch.float32, device="cuda")
    for path in sorted(glob.glob("/sys/devices/system/node/node[0-9]*")):
        node = int(path.rsplit("node", 1)[1])
        cpus = [c for c in cpulist(open(path + "/cpulist").read()) if c in all_cpus]
        if not cpus:
            res["nodes"][node] = {"skipped": "no allowed cpus"}
            continue
        os.sched_setaffinity(0, cpus)
        pin = torch.empty(64 << 20, dtype=torch.float32).pin_memory()
        pin.fill_(1.0)
        out = {}
        for name, fn in (("h2d", lambda: dev.copy_(pin, non_blocking=True)), ("d2h", lambda: pin.copy_(dev, non_blocking=True))):
            for _ in range(2):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(5):
                fn()
            e1.record()
            torch.cuda.synchronize()
            out[name + "_GBps"] = pin.numel() * 4 * 5 / e0.elapsed_time(e1) / 1e6
        out["cpus"] = [cpus[0], cpus[-1], len(cpus)]
        res["nodes"][node] = out
        del pin
        os.sched_setaffinity(0, all_cpus)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "numa_probe.json"), "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
