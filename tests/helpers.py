"""Shared helpers for the parity tests (test infrastructure)."""
import json
import math
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "phpt_vectors.json")


def load_golden():
    recs = json.load(open(GOLDEN))["records"]
    for r in recs:
        r["expected"] = [float("nan") if v == "NAN" else float(v) for v in r["expected"]]
    return recs


def php14(v: float) -> float:
    """What PHP prints for a float32 result at precision=14, read back as a float."""
    v = float(v)
    if math.isnan(v) or math.isinf(v):
        return v
    return float(f"{v:.14G}")


def assert_php_equal(got, expected, what=""):
    got = np.asarray(got, dtype=np.float32).reshape(-1)
    assert got.size == len(expected), f"{what}: size {got.size} != {len(expected)}"
    for i, (g, e) in enumerate(zip(got, expected)):
        if math.isnan(e):
            assert math.isnan(float(g)), f"{what}[{i}]: expected NAN got {g}"
        else:
            assert php14(g) == e, f"{what}[{i}]: printed {php14(g)!r} != expected {e!r}"


def run_golden_record(backend, r):
    """Evaluate one phpt record on a backend exposing binary/unary/reduce_full/
    reduce_axis/matmul (oracle.ref, oracle.port or the numpower_b200 host mirror)."""
    op, kw = r["op"], r["kwargs"]
    vals = [np.asarray(o["value"], dtype=np.float32) for o in r["operands"]]
    if op in ("add", "sub", "mul", "div", "mod", "pow"):
        return backend.binary(op, vals[0], vals[1])
    if op == "matmul":
        return backend.matmul(vals[0], vals[1])
    if op in ("sum", "prod", "max", "min"):
        if kw.get("axis") is None:
            return backend.reduce_full(op, vals[0])
        return backend.reduce_axis(op, vals[0], kw["axis"])
    if op == "clip":
        return backend.unary("clip", vals[0], float(kw["min"]), float(kw["max"]))
    if op == "round":
        return backend.unary("round", vals[0], float(kw["precision"]), 0.0)
    return backend.unary(op, vals[0])


def rel_err(got, ref):
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    both_nan = np.isnan(got) & np.isnan(ref)
    same_inf = np.isinf(ref) & (got == ref)
    denom = np.where(ref == 0, 1.0, np.abs(ref))
    err = np.abs(got - ref) / denom
    err = np.where(both_nan | same_inf, 0.0, err)
    err = np.where(np.isnan(err), np.inf, err)
    return err
