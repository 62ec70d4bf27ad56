"""Prints the interesting fields of a bench.py JSON line (used by scripts/gpu_stage.sh)."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value", round(d["value"], 1), "ms", round(d["ms_per_step"], 4), "frac", round(d["roofline"]["frac"], 3), "pipe", round(d["roofline"]["pipe_frac"], 3),
      "traffic", d["roofline"].get("traffic"), "e2e ms", round(d["e2e"]["ms_per_step"], 3), d["e2e"].get("pcie_measured"), "launches", d["gpu_launches"], d["clocks"])
for k, v in d["roofline"]["matmul_modes"].items():
    print("  mode", k, round(v["ms"], 4), round(v["useful_tflops"], 1), "%.2e" % v["max_rel_err_vs_fp64"])
for k, v in d["roofline"]["per_config"].items():
    print("  cfg", k, round(v["ms"], 5), round(v["achieved"], 1), v["unit"], round(v.get("frac", 0), 3))
print("cpu", {k: {a: (round(b, 3) if isinstance(b, float) else b) for a, b in v.items() if a in ("ms", "GBps", "TFLOPs", "ms_per_slice")}
              for k, v in d["cpu_baseline"]["per_config"].items()})
