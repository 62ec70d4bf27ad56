#!/usr/bin/env bash
mkdir -p gpurun_out; rm -f gpurun_out/gemm_probe.jsonl
timeout 900 python scripts/gemm_probe.py cg1_bn128 cg2_bn128 cg2_bn256 > gpurun_out/gemm_probe.log 2>&1
python - <<'PY'
import json
for l in open("gpurun_out/gemm_probe.jsonl"):
    d = json.loads(l)
    if "M" in d:
        print(d["variant"], d["precision"], d["M"], d.get("rc"), "max_rel=%.2e mean=%.2e" % (d.get("max_rel", -1), d.get("mean_signed_rel", 0)),
              "ms=%.4f useful=%.1f pipe=%.1f" % (d.get("ms", 0), d.get("useful_tflops", 0), d.get("pipe_tflops", 0)))
    elif d.get("exit"):
        print(d)
PY
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
tail -12 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 3000 gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2>&1; tail -c 600 gpurun_out/bench_ref.json
