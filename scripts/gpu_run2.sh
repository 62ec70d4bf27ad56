#!/usr/bin/env bash
mkdir -p gpurun_out; rm -f gpurun_out/gemm_probe.jsonl
timeout 1500 python scripts/gemm_probe.py > gpurun_out/gemm_probe.log 2>&1
timeout 900 python -m pytest tests -m gpu -q --timeout=300 -p no:cacheprovider -x -k "matmul or dot or config2 or dropin or sequential" > gpurun_out/pytest_gemm.log 2>&1
python - <<'PY'
import json
for l in open("gpurun_out/gemm_probe.jsonl"):
    d = json.loads(l)
    if "M" in d:
        print(d["variant"], d["precision"], d["M"], d.get("rc"), "max_rel=%.2e mean=%.2e" % (d.get("max_rel", -1), d.get("mean_signed_rel", 0)),
              "ms=%.4f useful=%.1f pipe=%.1f" % (d.get("ms", 0), d.get("useful_tflops", 0), d.get("pipe_tflops", 0)),
              "trunc=%.2e rnd=%.2e" % (d.get("max_rel_vs_truncated_inputs", -1), d.get("max_rel_vs_rounded_inputs", -1)))
    elif d.get("exit"):
        print(d)
PY
tail -15 gpurun_out/pytest_gemm.log
