"""CPU tests: pin the oracle.

1. the reference's own object code (oracle/_ref) reproduces every phpt golden vector;
2. the C restatement (oracle/port.c) reproduces them too;
3. port == reference object code on seeded random inputs (bit-exact for the
   arithmetic/index ops, <= 2 ulp-ish for libm, 1e-5 for sgemm).
"""
import numpy as np
import pytest

import oracle
from helpers import assert_php_equal, load_golden, rel_err, run_golden_record

RECS = load_golden()
IDS = [f"{r['file'].split('/')[-1].split('.')[0]}#{i}" for i, r in enumerate(RECS)]

needs_ref = pytest.mark.skipif(not oracle.ref.available, reason="oracle/_ref not built")


@needs_ref
@pytest.mark.parametrize("rec", RECS, ids=IDS)
def test_reference_object_code_reproduces_phpt(rec):
    assert_php_equal(run_golden_record(oracle.ref, rec), rec["expected"], rec["file"])


@pytest.mark.parametrize("rec", RECS, ids=IDS)
def test_port_reproduces_phpt(rec):
    assert_php_equal(run_golden_record(oracle.port, rec), rec["expected"], rec["file"])


def _rng(seed):
    return np.random.default_rng(seed)


@needs_ref
@pytest.mark.parametrize("op", ["add", "sub", "mul", "div", "mod", "pow", "maximum", "minimum", "arctan2"])
@pytest.mark.parametrize("n", [1, 7, 8, 9, 1000, 4099])
def test_port_vs_ref_binary(op, n):
    r = _rng(n)
    a = (r.random(n, dtype=np.float32) * 8 - 4).astype(np.float32)
    b = (r.random(n, dtype=np.float32) * 3 + 0.25).astype(np.float32)
    if op == "pow":
        a = np.abs(a) + 0.1
    if op == "mod":
        b[::3] *= -1  # mixed signs: body is floored (fused), tail is fmodf — the port models both
    g, e = oracle.port.binary(op, a, b), oracle.ref.binary(op, a, b)
    if op in ("pow", "arctan2"):
        assert rel_err(g, e).max() <= 1e-6
    else:
        np.testing.assert_array_equal(g, e)
        if op == "mul":
            z = np.zeros(n, np.float32)
            np.testing.assert_array_equal(np.signbit(oracle.port.binary(op, a, z)), np.signbit(oracle.ref.binary(op, a, z)))


@needs_ref
@pytest.mark.parametrize("shape_a,shape_b", [((5, 7), ()), ((5, 7), (7,)), ((6, 3), (6, 1)), ((4, 4), (1, 4)), ((3,), (3,))])
def test_port_vs_ref_broadcast_subset(shape_a, shape_b):
    r = _rng(11)
    a = r.standard_normal(shape_a).astype(np.float32)
    b = r.standard_normal(shape_b).astype(np.float32) + 3
    for op in ("add", "sub", "mul", "div"):
        np.testing.assert_array_equal(oracle.port.binary(op, a, b), oracle.ref.binary(op, a, b))
        np.testing.assert_array_equal(oracle.port.binary(op, b, a), oracle.ref.binary(op, b, a))


UNARY_DOMAINS = {
    "sqrt": (0, 50), "log": (1e-3, 50), "log2": (1e-3, 50), "log10": (1e-3, 50), "log1p": (-0.9, 50),
    "logb": (1e-3, 50), "arcsin": (-1, 1), "arccos": (-1, 1), "arccosh": (1, 50), "arctanh": (-0.99, 0.99),
    "exp": (-20, 20), "exp2": (-20, 20), "expm1": (-20, 20), "sinh": (-10, 10), "cosh": (-10, 10),
    "rsqrt": (1e-3, 100), "reciprocal": (0.01, 50),
}


@needs_ref
@pytest.mark.parametrize("op", [k for k in oracle.UN_OPS])
def test_port_vs_ref_unary(op):
    lo, hi = UNARY_DOMAINS.get(op, (-30, 30))
    x = (_rng(5).random(5000, dtype=np.float32) * (hi - lo) + lo).astype(np.float32)
    x[:4] = np.clip(np.array([0.0, 0.5, 1.0, 2.5], np.float32), lo, hi)
    p0, p1 = (-1.5, 2.5) if op == "clip" else ((2.0, 0.0) if op == "round" else (0.0, 0.0))
    g, e = oracle.port.unary(op, x, p0, p1), oracle.ref.unary(op, x, p0, p1)
    np.testing.assert_array_equal(g, e)  # same libm, same flags => identical bits


@needs_ref
def test_port_vs_ref_reductions():
    r = _rng(3)
    x = r.integers(-64, 65, size=(37, 53)).astype(np.float32) / 64
    for op in ("sum", "prod", "min", "max"):
        assert oracle.port.reduce_full(op, x) == oracle.ref.reduce_full(op, x) or op == "prod"
    for axis in (0, 1):
        np.testing.assert_array_equal(oracle.port.reduce_axis("sum", x, axis), oracle.ref.reduce_axis("sum", x, axis))
        np.testing.assert_array_equal(oracle.port.reduce_axis("max", x, axis), oracle.ref.reduce_axis("max", x, axis))
    y = r.random((4, 5, 6), dtype=np.float32)
    for axis in (0, 1, 2):
        np.testing.assert_array_equal(oracle.port.reduce_axis("sum", y, axis), oracle.ref.reduce_axis("sum", y, axis))
        g, e = oracle.port.reduce_axis("prod", y, axis), oracle.ref.reduce_axis("prod", y, axis)
        np.testing.assert_array_equal(g, e)


@needs_ref
def test_sum_is_sequential_fp32_and_saturates():
    """the reference's sum is a strictly sequential fp32 accumulation."""
    x = np.ones(2**24 + 100, np.float32)
    assert oracle.ref.reduce_full("sum", x) == np.float32(2**24)
    assert oracle.port.reduce_full("sum", x) == np.float32(2**24)


@needs_ref
def test_port_vs_ref_argminmax():
    r = _rng(9)
    x = r.integers(0, 50, size=(6, 7, 5)).astype(np.float32)  # many ties -> first index wins
    for is_max in (True, False):
        np.testing.assert_array_equal(oracle.port.argminmax(is_max, x), oracle.ref.argminmax(is_max, x))
        for axis in (0, 1, 2):
            for kd in (False, True):
                np.testing.assert_array_equal(oracle.port.argminmax(is_max, x, axis, kd),
                                              oracle.ref.argminmax(is_max, x, axis, kd))
    y = x.reshape(-1).copy()
    y[17] = np.nan
    y[40] = np.nan
    # calculation.c:23 uses `*ip > mp` (NaN compares false => a NaN after index 0 is SKIPPED
    # by argmax) while :50 uses `!(mp <= *ip)` (=> the first NaN WINS argmin).
    assert oracle.ref.argminmax(False, y) == 17 == oracle.port.argminmax(False, y)
    no_nan = np.where(np.isnan(y), -np.inf, y)
    assert oracle.ref.argminmax(True, y) == np.argmax(no_nan) == oracle.port.argminmax(True, y)
    y[0] = np.nan
    assert oracle.ref.argminmax(True, y) == 0 == oracle.port.argminmax(True, y)


@needs_ref
@pytest.mark.parametrize("mkn", [(2, 2, 2), (17, 33, 9), (128, 256, 64), (300, 200, 100)])
def test_port_vs_ref_matmul(mkn):
    m, k, n = mkn
    r = _rng(m)
    a, b = r.random((m, k), dtype=np.float32), r.random((k, n), dtype=np.float32)
    g, e = oracle.port.matmul(a, b), oracle.ref.matmul(a, b)
    t = oracle.port.matmul_f64(a, b)
    assert rel_err(g, e).max() <= 1e-5
    assert rel_err(e, t).max() <= 1e-5
    x = r.random(k, dtype=np.float32)
    assert rel_err(oracle.port.gemv(a, x), oracle.ref.dot(a, x)).max() <= 1e-5


@needs_ref
@pytest.mark.parametrize("op", ["equal", "not_equal", "greater", "greater_equal", "less", "less_equal"])
def test_port_vs_ref_comparisons(op):
    r = _rng(13)
    a = r.integers(-3, 4, size=1003).astype(np.float32)
    b = r.integers(-3, 4, size=1003).astype(np.float32)
    np.testing.assert_array_equal(oracle.port.binary(op, a, b), oracle.ref.binary(op, a, b))
    if op not in ("less", "equal"):
        # NDArray_Less and NDArray_Equal loop over the UN-broadcast operands (logic.c:229-244, :534-553 use
        # nda/ndb instead of a_broad/b_broad) and read out of bounds when shapes differ: broadcast parity is
        # only defined for the other four predicates.
        np.testing.assert_array_equal(oracle.port.binary(op, a.reshape(17, 59), b[:59]), oracle.ref.binary(op, a.reshape(17, 59), b[:59]))


# ------------------------------------------------------------------ nd::all / nd::allclose (SURVEY §8 f, N2)
def _logic_vectors():
    import json
    import os
    return json.load(open(os.path.join(os.path.dirname(__file__), "golden", "logic_vectors.json")))["vectors"]


@pytest.mark.parametrize("rec", _logic_vectors(), ids=lambda r: f"{r['op']}-{r['expect']}")
def test_all_allclose_golden_vectors(rec):
    """tests/logic/001-ndarray-all.phpt and 002-ndarray-allclose.phpt: the port and (where built) the reference's object code."""
    args = [np.asarray(a, np.float32) for a in rec["args"]]
    fn = getattr(oracle.port, rec["op"])
    assert fn(*args) == rec["expect"]
    if oracle.ref.available and hasattr(oracle.ref.lib, "ref_all"):
        if len(args) == 2 and np.array_equal(args[0], args[1]):
            args[1] = args[0]      # the phpt passes the SAME object twice: the reference's out-of-bounds reads then see equal memory
        assert getattr(oracle.ref, rec["op"])(*args) == rec["expect"]


def _compare_vectors():
    import json
    import os
    return json.load(open(os.path.join(os.path.dirname(__file__), "golden", "compare_vectors.json")))["vectors"]


@pytest.mark.parametrize("rec", _compare_vectors(), ids=lambda r: f"{r['op']}-{np.asarray(r['args'][0]).ndim}d")
def test_comparison_and_transpose_golden_vectors(rec):
    """tests/logic/003..008-*.phpt: the port and the reference's object code print the same 0 / 1 masks; the transpose vectors of
    tests/manipulation/001-ndarray-transpose.phpt pin the layout the GPU test checks nb200_transpose2d against (numpy's .T)."""
    exp = np.asarray(rec["expect"], np.float32)
    if rec["op"] == "transpose":
        x = np.asarray(rec["args"][0], np.float32)
        np.testing.assert_array_equal(x.T if x.ndim == 2 else x.reshape(exp.shape), exp)
        return
    a, b = (np.asarray(v, np.float32) for v in rec["args"])
    np.testing.assert_array_equal(oracle.port.binary(rec["op"], a, b), exp)
    if oracle.ref.available:
        np.testing.assert_array_equal(oracle.ref.binary(rec["op"], a, b), exp)


def test_all_allclose_intended_semantics_where_the_reference_loops_are_broken():
    """oracle/port.c documents the two reference bugs; this pins what the port (and the kernels checked against it) do instead."""
    x = np.arange(1, 41, dtype=np.float32)
    assert oracle.port.all(x) == 1                                  # the reference's AVX2 body answers 0 for any n >= 8
    if oracle.ref.available and hasattr(oracle.ref.lib, "ref_all"):
        assert oracle.ref.all(x) == 0                               # (documented bug: mask compared with 0x0F, logic.c:33)
        assert oracle.ref.all(x[:7]) == 1                           # scalar loop: correct
    x[17] = 0.0
    assert oracle.port.all(x) == 0
    x[17] = np.nan
    assert oracle.port.all(x) == 1                                  # NaN is non-zero (scalar loop rule, numpy's rule)
    a = np.linspace(1, 2, 33, dtype=np.float32)
    b = a.copy()
    b[20] += np.float32(1e-3)
    assert oracle.port.allclose(a, a) == 1 and oracle.port.allclose(a, b) == 0
    assert oracle.port.allclose(a, b, rtol=1e-2) == 1 and oracle.port.allclose(a, b, rtol=0.0, atol=2e-3) == 1
    b[20] = np.nan
    assert oracle.port.allclose(a, b) == 1                          # `diff > tolerance` is false for NaN (logic.c:732)
    with pytest.raises(RuntimeError, match="Shape mismatch"):
        oracle.port.allclose(a, a[:5])
