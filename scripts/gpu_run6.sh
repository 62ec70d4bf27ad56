#!/usr/bin/env bash
mkdir -p gpurun_out; rm -f gpurun_out/gemm_probe.jsonl
echo "== in-kernel split (default)"; timeout 600 python scripts/gemm_probe.py cg1_bn128 cg2_bn128 > gpurun_out/gemm_probe_ink.log 2>&1
mv gpurun_out/gemm_probe.jsonl gpurun_out/gemm_probe_ink.jsonl
echo "== pre-pass split"; NB200_GEMM_PRESPLIT=1 timeout 600 python scripts/gemm_probe.py cg2_bn128 > gpurun_out/gemm_probe_pre.log 2>&1
mv gpurun_out/gemm_probe.jsonl gpurun_out/gemm_probe_pre.jsonl
python - <<'PY'
import json
for f in ("ink", "pre"):
    for l in open(f"gpurun_out/gemm_probe_{f}.jsonl"):
        d = json.loads(l)
        if "M" in d and d["precision"] == "x3":
            print(f, d["variant"], d["M"], d.get("rc"), "max_rel=%.2e mean=%.2e" % (d.get("max_rel", -1), d.get("mean_signed_rel", 0)),
                  "ms=%.4f useful=%.1f pipe=%.1f" % (d.get("ms", 0), d.get("useful_tflops", 0), d.get("pipe_tflops", 0)))
        elif d.get("exit"):
            print(f, d)
PY
timeout 900 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider -x -k "matmul or dot or config2 or config5 or dropin or sgemm" > gpurun_out/pytest_gemm.log 2>&1
tail -5 gpurun_out/pytest_gemm.log
