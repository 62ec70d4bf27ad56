#!/usr/bin/env bash
# round-end validation with both AUTO policies: default build, and NB200_GEMM_AUTO_MODE=fp16x3
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; tail -8 gpurun_out/pytest_gpu.log | cut -c1-250
echo "== AUTO=fp16x3 pytest"; NB200_GEMM_AUTO_MODE=fp16x3 timeout 600 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider -k "(matmul or sgemm or config2 or config5 or dropin or dot or statistics) and not auto_is" > gpurun_out/pytest_gpu_auto_fp16.log 2>&1; tail -6 gpurun_out/pytest_gpu_auto_fp16.log | cut -c1-250
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
python scripts/gemm_probe.py child cg2_bn128 4 4096x4096x4096 8192x8192x8192 2048x2048x2048 1024x1024x1024 > gpurun_out/probe_fp16x3.jsonl 2>&1
timeout 400 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -2 gpurun_out/bench_n1.err
NB200_GEMM_AUTO_MODE=fp16x3 timeout 400 python bench.py > gpurun_out/bench_n1_fp16x3.json 2> gpurun_out/bench_n1_fp16x3.err; tail -2 gpurun_out/bench_n1_fp16x3.err
python - <<'PY'
import json
for l in open("gpurun_out/probe_fp16x3.jsonl"):
    if l.startswith("{"):
        d = json.loads(l)
        print("probe", d["precision"], d["M"], round(d.get("ms", 0), 4), round(d.get("useful_tflops", 0), 1), "%.2e" % d.get("max_rel", -1), "%.2e" % d.get("mean_signed_rel", 0), d.get("error", ""))
    else:
        print(l.strip()[:200])
for f in ("bench_n1.json", "bench_n1_fp16x3.json"):
    try:
        d = json.loads(open("gpurun_out/" + f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "UNREADABLE", e); continue
    print(f, "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 4), d["dtype"][:30], "frac", round(d["roofline"]["frac"], 3), round(d["roofline"]["pipe_frac"], 3), "e2e", round(d["e2e"]["value"], 1), "launches", d["gpu_launches"], d["clocks"]["reasons"])
    for k, v in d["extras"].items():
        if k.startswith("matmul_") or k.startswith("batched"): print("   ", k, {a: (round(b, 4) if isinstance(b, float) else b) for a, b in v.items() if a != "note"})
PY
