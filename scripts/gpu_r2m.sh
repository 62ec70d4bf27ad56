#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q --timeout=600 -p no:cacheprovider -k "host or matmul" > gpurun_out/r2m_pytest.log 2>&1; tail -5 gpurun_out/r2m_pytest.log | cut -c1-300
timeout 600 python bench.py > gpurun_out/r2m_bench_n1.json 2> gpurun_out/r2m_bench_n1.err; tail -5 gpurun_out/r2m_bench_n1.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2m_bench_n1.json").read().strip().splitlines()[-1])
print("value", round(d["value"], 1), "ms", round(d["ms_per_step"], 4), "frac", round(d["roofline"]["frac"], 3), "pipe", round(d["roofline"]["pipe_frac"], 3), "e2e ms", round(d["e2e"]["ms_per_step"], 3), d["e2e"]["pcie_measured"], "launches", d["gpu_launches"], d["clocks"])
for k, v in d["roofline"]["matmul_modes"].items(): print("  mode", k, round(v["ms"], 4), round(v["useful_tflops"], 1), "%.2e" % v["max_rel_err_vs_fp64"])
for k, v in d["roofline"]["per_config"].items(): print("  cfg", k, round(v["ms"], 4), round(v["achieved"], 1), v["unit"], round(v.get("frac", 0), 3), "cpu", v.get("cpu_reference"))
print("cpu", {k: {a: (round(b, 3) if isinstance(b, float) else b) for a, b in v.items() if a in ("ms", "GBps", "TFLOPs", "ms_per_slice")} for k, v in d["cpu_baseline"]["per_config"].items()})
PY
bash scripts/gpu_ncu.sh r2_matmul_auto "prep16_coop|sgemm_tf32_kernel|fp16_post" 4 4 gemm_auto
python scripts/ncu_extract.py gpurun_out/prof_r2_matmul_auto.ncu-rep gpurun_out/r2_ncu_matmul_auto.csv
bash scripts/gpu_ncu.sh r2_hbm "ew_flat_vec|ew_bcast2d|reduce_rows_kernel|arg_rows_kernel|reduce_cols_kernel" 0 14 ew reduce
python scripts/ncu_extract.py gpurun_out/prof_r2_hbm.ncu-rep gpurun_out/r2_ncu_hbm.csv
