"""CPU tests of the measurement contract: the reference arm runs here and prints the agreed JSON line, the product arm fails loudly
without a GPU (no CPU fallback), and the committed round-2 record carries every key the driver and the judge read."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
             "data", "config", "e2e", "gpu_launches"}


def _run(*args, timeout=300):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=timeout, cwd=ROOT)


def test_reference_arm_runs_on_cpu_and_prints_the_contract_line():
    import oracle
    if not (oracle.ref.available or os.path.exists(os.path.join(ROOT, "oracle", "liboracle_port.so"))):
        pytest.skip("oracle not built")
    r = _run("--impl", "reference", "--steps", "2", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-800:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and BASE_KEYS <= set(line), sorted(BASE_KEYS - set(line))
    assert line["steps"] == 2 and line["warmup"] == 1                       # the driver's steps / warm-up are honoured
    assert line["metric"].startswith("nd::matmul") and line["unit"] == "TFLOP/s" and line["higher_is_better"] is True
    assert "4096" in line["config"]["workload"]
    assert line["value"] > 0 and abs(line["value"] - 2 * 4096 ** 3 / line["ms_per_step"] / 1e9) <= 1e-6 * line["value"]
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0


@pytest.mark.skipif(__import__("torch").cuda.is_available(), reason="needs a box without a GPU")
def test_product_arm_fails_loudly_without_a_gpu():
    r = _run("--steps", "1", "--warmup", "1", timeout=120)
    assert r.returncode != 0
    assert not any(ln.lstrip().startswith("{") and '"value"' in ln for ln in r.stdout.splitlines())   # no number from a fallback


def test_committed_round2_record_has_every_contract_key():
    path = os.path.join(ROOT, "profiles", "r2_bench_n1.json")
    line = json.loads(open(path).read().strip().splitlines()[-1])
    assert BASE_KEYS | {"roofline", "cpu_baseline", "clocks"} <= set(line)
    assert line["n_gpus"] == 1 and line["scaling"] == "weak" and line["vs_baseline"] is None and line["data"] == "synthetic"
    rf = line["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(rf)
    assert rf["bound"] == "tensor" and rf["unit"] == "TFLOP/s" and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9
    assert rf["achieved"] == line["value"] and rf["traffic"] and rf["traffic"] > 2.0e8        # read from the ncu summary, not null
    assert abs(line["value"] - 2 * 4096 ** 3 / line["ms_per_step"] / 1e9) <= 1e-6 * line["value"]
    # every other BASELINE config rides in the record the driver keeps
    per = rf["per_config"]
    for k in ("chain_fused_8192sq", "chain_two_calls_8192sq", "sum_2pow28", "argmax_2pow28", "sum_axis0_8192sq", "add_1024sq_l2_warm",
              "batched_matmul_128x2048sq_1gpu"):
        assert k in per and per[k]["ms"] > 0, k
    cb = line["cpu_baseline"]
    assert {"value", "unit", "cores", "kind", "sample"} <= set(cb) and cb["kind"] == "reference" and cb["cores"] >= 1
    e = line["e2e"]
    assert e["h2d_bytes_per_step"] == 2 * 4096 * 4096 * 4 and e["d2h_bytes_per_step"] == 4096 * 4096 * 4
    assert e["unit"] == line["unit"] and 0 < e["value"] < line["value"]                         # copies inside the timed region
    assert line["gpu_launches"] > 0
    ck = line["clocks"]
    assert ck["sm_mhz"] and ck["sm_max_mhz"] and not ({"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(ck["reasons"]))
