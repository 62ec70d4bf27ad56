#!/usr/bin/env bash
# round 2, first GPU pass: FP16x3-as-AUTO validation (both tile shapes), probe timings, launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q --timeout=600 -p no:cacheprovider -k "matmul or sgemm or fp16 or config2 or config5 or dot or dropin" > gpurun_out/r2a_pytest_matmul.log 2>&1; tail -15 gpurun_out/r2a_pytest_matmul.log | cut -c1-300
for tile in 128 256 0; do
  echo "== probe fp16x3 tile=$tile"
  NB200_FP16_TILE=$tile timeout 300 python scripts/gemm_probe.py child auto 4 4096x4096x4096 8192x8192x8192 2048x2048x2048 1024x1024x1024 4097x4097x4097 > gpurun_out/r2a_probe_fp16x3_tile$tile.jsonl 2>&1
  cut -c1-400 gpurun_out/r2a_probe_fp16x3_tile$tile.jsonl | tail -6
done
echo "== probe fp16x3 old prepass"
NB200_FP16_PREPASS=0 timeout 300 python scripts/gemm_probe.py child auto 4 4096x4096x4096 > gpurun_out/r2a_probe_fp16x3_oldprepass.jsonl 2>&1; cut -c1-400 gpurun_out/r2a_probe_fp16x3_oldprepass.jsonl | tail -2
echo "== probe bf16x3 / tf32x3"
timeout 300 python scripts/gemm_probe.py child auto 2 4096x4096x4096 > gpurun_out/r2a_probe_bf16x3.jsonl 2>&1; cut -c1-400 gpurun_out/r2a_probe_bf16x3.jsonl | tail -1
timeout 300 python scripts/gemm_probe.py child auto 0 4096x4096x4096 > gpurun_out/r2a_probe_tf32x3.jsonl 2>&1; cut -c1-400 gpurun_out/r2a_probe_tf32x3.jsonl | tail -1
echo "== launch list (ncu, not a bench value)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2a_launches_probe.csv python scripts/gemm_probe.py child auto 4 4096x4096x4096 > gpurun_out/r2a_ncu_probe.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r2a_launches_probe.csv")) if len(r) > 10]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
last = collections.OrderedDict()
for r in rows[1:][-40:]:
    print(r[ki][:70], r[vi], r[ui])
PY
timeout 400 python bench.py > gpurun_out/r2a_bench_n1.json 2> gpurun_out/r2a_bench_n1.err; tail -3 gpurun_out/r2a_bench_n1.err; cut -c1-1500 gpurun_out/r2a_bench_n1.json
