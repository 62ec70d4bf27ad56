#!/usr/bin/env bash
# round-end validation: full GPU pytest, smoke, FP16x3/BF16x3 probe, bench, memcheck over every kernel family
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
tail -15 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
PROBE_PRECISIONS=4,2 timeout 300 python scripts/gemm_probe.py cg2_bn256 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -3 gpurun_out/bench_n1.err
timeout 400 compute-sanitizer --tool memcheck --print-limit 20 python scripts/sanitizer_targets.py > gpurun_out/sanitizer_memcheck_final.log 2>&1; tail -4 gpurun_out/sanitizer_memcheck_final.log
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_n1.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "roofline", d["roofline"]["frac"], d["roofline"]["pipe_frac"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"], "clocks", d["clocks"])
for k, v in d["extras"].items():
    if isinstance(v, dict) and "ms" in v: print(k, {a: (round(b, 4) if isinstance(b, float) else b) for a, b in v.items() if a not in ("note",)})
for l in open("gpurun_out/gemm_probe.jsonl").read().strip().splitlines()[-16:]:
    d = json.loads(l)
    if "ms" in d: print(d["precision"], d["M"], d["K"], d["N"], round(d["ms"], 4), round(d["useful_tflops"], 1), "%.2e" % d["max_rel"])
    elif "error" in d or d.get("exit"): print(d)
PY
