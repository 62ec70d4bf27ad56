#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout=600 -p no:cacheprovider -k "matmul or sgemm or fp16 or config2 or config5 or dot or dropin" > gpurun_out/r2d_pytest_matmul.log 2>&1; tail -12 gpurun_out/r2d_pytest_matmul.log | cut -c1-300
echo "== probe fp16x3 (PDL on)"
timeout 300 python scripts/gemm_probe.py child auto 4 4096x4096x4096 8192x8192x8192 2048x2048x2048 1024x1024x1024 > gpurun_out/r2d_probe_fp16x3.jsonl 2>&1; cut -c1-400 gpurun_out/r2d_probe_fp16x3.jsonl | tail -4
echo "== probe fp16x3 (NB200_PDL=0)"
NB200_PDL=0 timeout 300 python scripts/gemm_probe.py child auto 4 4096x4096x4096 2048x2048x2048 > gpurun_out/r2d_probe_fp16x3_nopdl.jsonl 2>&1; cut -c1-400 gpurun_out/r2d_probe_fp16x3_nopdl.jsonl | tail -2
echo "== probe tf32x3 / bf16x3"
timeout 300 python scripts/gemm_probe.py child auto 0 4096x4096x4096 > gpurun_out/r2d_probe_tf32x3.jsonl 2>&1; cut -c1-400 gpurun_out/r2d_probe_tf32x3.jsonl | tail -1
timeout 300 python scripts/gemm_probe.py child auto 2 4096x4096x4096 > gpurun_out/r2d_probe_bf16x3.jsonl 2>&1; cut -c1-400 gpurun_out/r2d_probe_bf16x3.jsonl | tail -1
echo "== launch list (ncu, not a bench value)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2d_launches_probe.csv python scripts/gemm_probe.py child auto 4 4096x4096x4096 > gpurun_out/r2d_ncu_probe.log 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open("gpurun_out/r2d_launches_probe.csv")) if len(r) > 10]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
for r in rows[1:][-5:]:
    print(r[ki][:80], r[vi], r[ui])
PY
