#!/usr/bin/env bash
# usage: gpu_multi.sh <N> <tag>   (torchrun bench.py --gpus N exactly as the driver launches it, plus the reference arm)
N=$1; tag=$2
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${tag}_bench_n$N.json 2> gpurun_out/${tag}_bench_n$N.err
echo "rc=$?"; tail -5 gpurun_out/${tag}_bench_n$N.err | cut -c1-400
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_bench_n$N.json").read().strip().splitlines()[-1])
except Exception as e:
    print("UNREADABLE", e); raise SystemExit
print("N", d["n_gpus"], "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "base", {k: (round(v, 3) if isinstance(v, float) else v) for k, v in d["scaling_base"].items() if k != "workload"})
for r in d["per_rank"]: print("  rank", r["rank"], round(r["ms_per_step"], 3), {k: r["clocks"].get(k) for k in ("sm_mhz", "sm_min_mhz", "power_w_max", "power_w_median", "reasons")})
for k, v in d["roofline"]["per_config"].items(): print("  sharded", k, {a: (round(b, 3) if isinstance(b, float) else b) for a, b in v.items()})
print("  e2e", round(d["e2e"]["value"], 1), round(d["e2e"]["ms_per_step"], 2), [(r["rank"], round(r["ms_per_step"], 2), round(r["h2d_GBps_under_contention"], 1), round(r.get("duplex_each_GBps_under_contention", 0), 1), round(r["frac_of_pcie_bound"], 2), round(r.get("frac_of_pcie_bound_duplex", 0), 2)) for r in d["e2e"]["per_rank"]])
print("  scatter_gather", json.dumps(d["scatter_gather"])[:1500])
PY
