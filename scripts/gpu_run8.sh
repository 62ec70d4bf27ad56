#!/usr/bin/env bash
# full GPU validation + ncu capture of the BF16x3 GEMM (merged 256x256 tile) + bench + launch list of the bench command
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
ncu --set full --clock-control none --import-source on -k regex:"sgemm_tf32_kernel|split_bf16" -s 4 -c 2 -o gpurun_out/prof_r1e \
    python scripts/profile_targets.py gemm_bf16 > gpurun_out/ncu_r1e.log 2>&1
tail -1 gpurun_out/ncu_r1e.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -5 gpurun_out/bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_r1e.csv python bench.py --steps 3 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_n1.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "roofline", d["roofline"]["frac"], d["roofline"]["pipe_frac"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"], "clocks", d["clocks"])
for k, v in d["extras"].items():
    if isinstance(v, dict) and "ms" in v: print(k, {a: (round(b, 4) if isinstance(b, float) else b) for a, b in v.items() if a != "note"})
PY
