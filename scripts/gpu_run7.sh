#!/usr/bin/env bash
# full GPU validation + BF16x3 probe + ncu capture of the BF16x3 GEMM + bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
tail -6 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
PROBE_PRECISIONS=2 timeout 300 python scripts/gemm_probe.py cg2_bn128 2>&1 | tail -2
bash scripts/gpu_ncu3.sh
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -5 gpurun_out/bench_n1.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_n1.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "roofline", d["roofline"]["frac"], d["roofline"]["pipe_frac"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"], "clocks", d["clocks"])
for k, v in d["extras"].items():
    if isinstance(v, dict) and "ms" in v: print(k, {a: (round(b, 4) if isinstance(b, float) else b) for a, b in v.items() if a != "note"})
PY
