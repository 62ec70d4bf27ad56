#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout=600 -p no:cacheprovider -k "matmul or sgemm or fp16 or config2 or config5 or dot or dropin" > gpurun_out/r2b_pytest_matmul.log 2>&1; tail -15 gpurun_out/r2b_pytest_matmul.log | cut -c1-300
for tile in 256 0; do
  echo "== probe fp16x3 tile=$tile"
  NB200_FP16_TILE=$tile timeout 300 python scripts/gemm_probe.py child auto 4 4096x4096x4096 8192x8192x8192 2048x2048x2048 1024x1024x1024 4097x4097x4097 > gpurun_out/r2b_probe_fp16x3_tile$tile.jsonl 2>&1
  cut -c1-400 gpurun_out/r2b_probe_fp16x3_tile$tile.jsonl | tail -6
done
echo "== launch list (ncu, not a bench value)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2b_launches_probe.csv python scripts/gemm_probe.py child auto 4 4096x4096x4096 > gpurun_out/r2b_ncu_probe.log 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open("gpurun_out/r2b_launches_probe.csv")) if len(r) > 10]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
for r in rows[1:][-14:]:
    print(r[ki][:80], r[vi], r[ui])
PY
