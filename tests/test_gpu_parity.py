"""GPU parity tests (run on the B200 box: `pytest -m gpu`).

Every test drives the CUDA path through the C-ABI (libnb200.so via the host mirror) and checks it
against the oracle (the reference's own CPU object code when oracle/_ref is present, else the C port)
on the same seeded inputs.  Bars (BASELINE.json north_star): bit-exact for arithmetic / index work,
<= 1e-5 relative for fp32 transcendental and matmul results (tolerance written at each assert).
"""
import math

import numpy as np
import pytest

import oracle
from helpers import assert_php_equal, load_golden, php14, rel_err, run_golden_record

pytestmark = pytest.mark.gpu

ORACLE = oracle.ref if oracle.ref.available else oracle.port
RTOL = 1e-5  # north_star: "within 1e-5 relative for fp32"

EXACT_OPS = {"add", "sub", "mul", "div", "mod", "abs", "sign", "square", "ceil", "fix", "floor", "rint", "round",
             "trunc", "clip", "sqrt", "sum", "prod", "max", "min", "matmul", "degrees", "radians"}


@pytest.fixture(scope="module")
def nb():
    import numpower_b200 as nb
    nb.lib()  # must load: no fallback
    return nb


def _rng(seed):
    return np.random.default_rng(seed)


def eq_zero_sign_free(g, e):
    """Bit-exact up to the sign of zero (the reference itself does not pin it: its AVX body turns
    every zero product into -0.0 and its scalar tail into +0.0, arithmetics.c:280-284,409-416)."""
    g, e = np.asarray(g), np.asarray(e)
    assert g.shape == e.shape, (g.shape, e.shape)
    both_nan = np.isnan(g) & np.isnan(e)
    ok = (g == e) | both_nan
    assert ok.all(), f"{(~ok).sum()} mismatches, first at {np.argwhere(~ok)[0]}: {g[~ok][0]!r} vs {e[~ok][0]!r}"


# --------------------------------------------------------------------- golden vectors
RECS = load_golden()
IDS = [f"{r['file'].split('/')[-1].split('.')[0]}#{i}" for i, r in enumerate(RECS)]


@pytest.mark.parametrize("rec", RECS, ids=IDS)
def test_phpt_golden_on_gpu(nb, rec):
    got = run_golden_record(nb.GoldenBackend, rec)
    if rec["op"] in EXACT_OPS:
        assert_php_equal(got, rec["expected"], rec["file"])
    else:  # libm-backed functors: CUDA libm vs glibc differ by <= 1-2 ulp
        got = np.asarray(got, np.float32).reshape(-1)
        for g, e in zip(got, rec["expected"]):
            if math.isnan(e):
                assert math.isnan(float(g))
            else:
                assert abs(php14(g) - e) <= RTOL * abs(e) + 1e-12, (rec["file"], g, e)


# --------------------------------------------------------------------- binary elementwise
@pytest.mark.parametrize("op", ["add", "sub", "mul", "div", "mod", "pow", "maximum", "minimum", "arctan2"])
@pytest.mark.parametrize("n", [1, 7, 8, 1000, 4099, 1 << 20])
def test_binary_flat_vs_oracle(nb, op, n):
    r = _rng(n + len(op))
    a = (r.random(n, dtype=np.float32) * 8 - 4).astype(np.float32)
    b = (r.random(n, dtype=np.float32) * 3 + 0.25).astype(np.float32)
    if op == "pow":
        a = np.abs(a) + 0.1
    A, B = nb.NDArray.array(a).gpu(), nb.NDArray.array(b).gpu()
    got = nb.nd.binary(op, A, B).toArray()
    exp = ORACLE.binary(op, a, b)
    if op in ("pow", "arctan2"):
        assert rel_err(got, exp).max() <= RTOL
    elif op == "mod":
        # the reference computes the first (n//8)*8 elements with the fused floored formula and the
        # tail with fmodf (arithmetics.c:787-806); same-sign operands agree everywhere up to rounding
        nbody = (n // 8) * 8
        eq_zero_sign_free(got[:nbody], exp[:nbody])
    else:
        eq_zero_sign_free(got, exp)


def test_mod_mixed_signs_body_is_floored(nb):
    r = _rng(77)
    n = 4096
    a = (r.random(n, dtype=np.float32) * 20 - 10).astype(np.float32)
    b = (r.random(n, dtype=np.float32) * 4 - 2).astype(np.float32)
    b[np.abs(b) < 0.1] = 0.5
    got = nb.nd.mod(nb.NDArray.array(a).gpu(), nb.NDArray.array(b).gpu()).toArray()
    eq_zero_sign_free(got, ORACLE.binary("mod", a, b))  # n % 8 == 0: all body


BCAST = [((512, 384), ()), ((512, 384), (384,)), ((300, 257), (257,)), ((640, 128), (640, 1)), ((64, 64), (1, 64)),
         ((33, 5), (33, 1)), ((7, 1), (7, 9))]


@pytest.mark.parametrize("sa,sb", BCAST)
@pytest.mark.parametrize("op", ["add", "sub", "mul", "div"])
def test_binary_broadcast_vs_oracle(nb, sa, sb, op):
    r = _rng(len(sa) * 31 + len(sb))
    a = r.standard_normal(sa).astype(np.float32)
    b = (r.standard_normal(sb) + 3).astype(np.float32)
    A = nb.NDArray.array(a).gpu()
    B = nb.NDArray.array(b).gpu() if b.ndim else nb.NDArray.array(b)
    eq_zero_sign_free(nb.nd.binary(op, A, B).toArray(), ORACLE.binary(op, a, b))
    eq_zero_sign_free(nb.NDArray(nb.lib().NB_NDArray_Binary(nb.ndarray.BIN[op], B._h, A._h)).toArray(),
                      ORACLE.binary(op, b, a))


def test_general_nd_broadcast_matches_numpy_semantics(nb):
    """Stride-0 broadcasting is a superset of the reference's special cases."""
    r = _rng(5)
    a = r.standard_normal((4, 1, 6, 5)).astype(np.float32)
    b = r.standard_normal((3, 1, 5)).astype(np.float32)
    got = nb.nd.add(nb.NDArray.array(a).gpu(), nb.NDArray.array(b).gpu()).toArray()
    np.testing.assert_array_equal(got, a + b)


def test_views_unaligned_rows(nb):
    """$a[i] views are only 4-byte aligned: row 1 of an (N,3) array."""
    r = _rng(9)
    a = r.standard_normal((5, 3)).astype(np.float32)
    A = nb.NDArray.array(a).gpu()
    eq_zero_sign_free((A[1] + A[2]).toArray(), a[1] + a[2])
    eq_zero_sign_free((A + A[0]).toArray(), ORACLE.binary("add", a, a[0]))
    assert nb.nd.sum(A[1]) == pytest.approx(float(ORACLE.reduce_full("sum", a[1])), rel=1e-6)


def test_mul_add_fused_equals_two_calls_and_oracle(nb):
    r = _rng(21)
    for shp_b, shp_c in [((257, 1031), (257, 1031)), ((1031,), (257, 1)), ((257, 1), (1031,))]:
        a = r.random((257, 1031), dtype=np.float32)
        b = r.random(shp_b, dtype=np.float32)
        c = r.random(shp_c, dtype=np.float32)
        A, B, Cc = (nb.NDArray.array(x).gpu() for x in (a, b, c))
        fused = nb.nd.mul_add(A, B, Cc).toArray()
        two = (A * B + Cc).toArray()
        np.testing.assert_array_equal(fused, two)          # no FMA contraction: fused == unfused bit for bit
        eq_zero_sign_free(fused, oracle.port.mul_add(a, b, c))
        if oracle.ref.available:
            eq_zero_sign_free(fused, oracle.ref.mul_add(a, b, c))


# --------------------------------------------------------------------- unary
UNARY_DOMAINS = {
    "sqrt": (0, 50), "log": (1e-3, 50), "log2": (1e-3, 50), "log10": (1e-3, 50), "log1p": (-0.9, 50),
    "logb": (1e-3, 50), "arcsin": (-1, 1), "arccos": (-1, 1), "arccosh": (1, 50), "arctanh": (-0.99, 0.99),
    "exp": (-20, 20), "exp2": (-20, 20), "expm1": (-20, 20), "sinh": (-10, 10), "cosh": (-10, 10),
    "rsqrt": (1e-3, 100), "reciprocal": (0.01, 50),
}
BITEXACT_UNARY = {"abs", "sqrt", "degrees", "radians", "rint", "fix", "trunc", "floor", "ceil", "negative", "positive",
                  "sign", "reciprocal", "rsqrt", "clip", "round", "square", "logb"}


@pytest.mark.parametrize("op", list(oracle.UN_OPS))
def test_unary_vs_oracle(nb, op):
    lo, hi = UNARY_DOMAINS.get(op, (-30, 30))
    x = (_rng(5).random(100003, dtype=np.float32) * (hi - lo) + lo).astype(np.float32)
    x[:4] = np.clip(np.array([0.0, 0.5, 1.0, 2.5], np.float32), lo, hi)
    p0, p1 = (-1.5, 2.5) if op == "clip" else ((2.0, 0.0) if op == "round" else (0.0, 0.0))
    got = nb.nd.unary(op, nb.NDArray.array(x).gpu(), p0, p1).toArray()
    exp = ORACLE.unary(op, x, p0, p1)
    if op in BITEXACT_UNARY:
        eq_zero_sign_free(got, exp)
    elif op in ("sin", "cos", "tan", "sinc"):
        # near the zeros of sin/cos a relative bound is ill-posed; both libms are within 2 ulp of the
        # true value, so bound the error by 1e-5 of max(|result|, ulp-scale of the argument)
        err = np.abs(got.astype(np.float64) - exp)
        assert (err <= RTOL * np.maximum(np.abs(exp), 1e-2)).all()
    else:
        assert rel_err(got, exp).max() <= RTOL


def test_domain_error_is_reported_not_fatal(nb):
    """double_math.c:145-148 exit(1)s on arccos(|x|>1); the backend raises instead."""
    A = nb.NDArray.array(np.array([0.5, 2.0], np.float32)).gpu()
    with pytest.raises(nb.BackendError, match="arccos"):
        nb.nd.unary("arccos", A)
    assert np.isfinite(nb.nd.unary("arccos", nb.NDArray.array(np.array([0.5], np.float32)).gpu()).toArray()).all()


# --------------------------------------------------------------------- reductions
def _set_p(n, seed):  # values in {-1,0,1}: every partial sum is an exact integer (SURVEY §8 d, set P)
    return _rng(seed).choice(np.array([-1, 0, 0, 1], np.float32), size=n)


def _set_p2(shape, seed):  # k/64, k in [-64,64]: exact in any order
    return (_rng(seed).integers(-64, 65, size=shape).astype(np.float32) / 64).astype(np.float32)


@pytest.mark.parametrize("n", [1, 5, 1023, 4096, 100003, (1 << 22) + 5])
def test_full_reductions_exact_sets(nb, n):
    x = _set_p(n, n)
    A = nb.NDArray.array(x).gpu()
    assert nb.nd.sum(A) == float(ORACLE.reduce_full("sum", x))
    y = _set_p2(n, n + 1)
    B = nb.NDArray.array(y).gpu()
    assert nb.nd.sum(B) == float(ORACLE.reduce_full("sum", y))
    assert nb.nd.max(B) == float(ORACLE.reduce_full("max", y))
    assert nb.nd.min(B) == float(ORACLE.reduce_full("min", y))
    z = _rng(n).choice(np.array([1, 1, 1, -1, 2, 0.5], np.float32), size=min(n, 4000))
    assert nb.nd.prod(nb.NDArray.array(z).gpu()) == float(ORACLE.reduce_full("prod", z))


def test_sum_uniform_vs_fp64_truth(nb):
    """Set U: the reference's sequential fp32 sum drifts/saturates; compare with fp64."""
    x = _rng(9).random(1 << 24, dtype=np.float32)
    got = nb.nd.sum(nb.NDArray.array(x).gpu())
    truth = float(x.astype(np.float64).sum())
    assert abs(got - truth) <= RTOL * truth
    # and it is reproducible bit for bit (no float atomics)
    assert got == nb.nd.sum(nb.NDArray.array(x).gpu())


def test_min_max_nan_rules(nb):
    x = _rng(3).standard_normal(5000).astype(np.float32)
    x[100] = np.nan
    for op in ("min", "max"):
        assert nb.nd.reduce(op, nb.NDArray.array(x).gpu()) == float(ORACLE.reduce_full(op, x))  # later NaN skipped
    x[0] = np.nan
    for op in ("min", "max"):
        assert math.isnan(nb.nd.reduce(op, nb.NDArray.array(x).gpu())) and math.isnan(float(ORACLE.reduce_full(op, x)))


AXIS_SHAPES = [((37, 53), 0), ((37, 53), 1), ((4, 5, 6), 1), ((3, 1000, 8), 1), ((2048, 96), 0), ((96, 2048), 1),
               ((5, 7, 2048), 2), ((3000, 5), 0), ((1, 70000), 1), ((70000, 3), 0),
               # more slices than a grid dimension holds (outer > 65535), long middle axes with a tiny / odd inner extent
               ((70000, 40, 3), 1), ((66000, 5), 1), ((3, 70001, 2), 1), ((2, 3, 70001), 1), ((70001, 2, 4), 1)]


@pytest.mark.parametrize("shape,axis", AXIS_SHAPES)
def test_axis_reductions_exact_set_tree(nb, shape, axis):
    x = _set_p2(shape, sum(shape))
    A = nb.NDArray.array(x).gpu()
    np.testing.assert_array_equal(nb.nd.sum(A, axis).toArray(), oracle.port.reduce_axis("sum", x, axis))
    np.testing.assert_array_equal(nb.nd.max(A, axis).toArray(), oracle.port.reduce_axis("max", x, axis))
    np.testing.assert_array_equal(nb.nd.min(A, axis).toArray(), oracle.port.reduce_axis("min", x, axis))


@pytest.mark.parametrize("shape,axis", AXIS_SHAPES[:8])
def test_axis_sum_sequential_order_bit_exact_on_random_data(nb, shape, axis):
    """NB200_ORDER_SEQUENTIAL reproduces the reference's slice-loop order (ndarray.c:394-429)."""
    x = _rng(1).standard_normal(shape).astype(np.float32)
    A = nb.NDArray.array(x).gpu()
    got = nb.nd.sum(A, axis, order=nb.ORDER_SEQUENTIAL).toArray()
    np.testing.assert_array_equal(got, oracle.port.reduce_axis("sum", x, axis))
    if oracle.ref.available and x.size < 20000:
        np.testing.assert_array_equal(got, oracle.ref.reduce_axis("sum", x, axis))
    tree = nb.nd.sum(A, axis).toArray()
    assert np.abs(tree - got).max() <= 1e-5 * np.abs(x).sum(axis=axis).max()  # same values up to summation order


# --------------------------------------------------------------------- argmax / argmin
@pytest.mark.parametrize("shape,axis", AXIS_SHAPES[8:])
def test_argminmax_many_slices_and_long_axes(nb, shape, axis):
    """argmax / argmin on the shapes with more slices than a grid dimension holds and on long middle axes; ties everywhere
    (values 0..49): first occurrence, as float_argmax / float_argmin (calculation.c:9-59)."""
    x = _rng(sum(shape)).integers(0, 50, size=shape).astype(np.float32)
    A = nb.NDArray.array(x).gpu()
    np.testing.assert_array_equal(nb.nd.argmax(A, axis).toArray(), oracle.port.argminmax(True, x, axis))
    np.testing.assert_array_equal(nb.nd.argmin(A, axis).toArray(), oracle.port.argminmax(False, x, axis))


def test_argminmax_axes_and_ties(nb):
    x = _rng(9).integers(0, 50, size=(6, 70, 5)).astype(np.float32)
    A = nb.NDArray.array(x).gpu()
    for is_max, fn in ((True, nb.nd.argmax), (False, nb.nd.argmin)):
        assert fn(A) == float(ORACLE.argminmax(is_max, x))
        for axis in (0, 1, 2):
            for kd in (False, True):
                got = fn(A, axis, kd).toArray()
                np.testing.assert_array_equal(got, oracle.port.argminmax(is_max, x, axis, kd))
    y = _rng(4).integers(0, 3, size=(300, 2000)).astype(np.float32)  # long rows, heavy ties
    Y = nb.NDArray.array(y).gpu()
    np.testing.assert_array_equal(nb.nd.argmax(Y, 1).toArray(), oracle.port.argminmax(True, y, 1))
    np.testing.assert_array_equal(nb.nd.argmin(Y, 0).toArray(), oracle.port.argminmax(False, y, 0))


def test_argminmax_nan_rules(nb):
    x = _rng(2).standard_normal(100000).astype(np.float32)
    x[777] = np.nan
    x[99000] = np.nan
    A = nb.NDArray.array(x).gpu()
    assert nb.nd.argmin(A) == float(ORACLE.argminmax(False, x)) == 777.0     # first NaN wins argmin
    assert nb.nd.argmax(A) == float(ORACLE.argminmax(True, x))               # argmax skips later NaNs
    x[0] = np.nan
    A = nb.NDArray.array(x).gpu()
    assert nb.nd.argmax(A) == 0.0 == float(ORACLE.argminmax(True, x))
    assert nb.nd.argmin(A) == 0.0
    z = np.full(5000, -np.inf, np.float32)
    assert nb.nd.argmax(nb.NDArray.array(z).gpu()) == 0.0
    z = np.array([0.0, -0.0, 0.0], np.float32)
    assert nb.nd.argmin(nb.NDArray.array(z).gpu()) == 0.0 == float(ORACLE.argminmax(False, z))


def test_argmax_large_index_float_rounding(nb):
    """The index is returned as float32 (calculation.c:25): an odd index > 2^24 rounds to even."""
    n = (1 << 24) + 4099
    x = np.zeros(n, np.float32)
    idx = (1 << 24) + 1001
    x[idx] = 5.0
    x[idx + 2000] = 5.0  # later tie must lose
    got = nb.nd.argmax(nb.NDArray.array(x).gpu())
    assert got == float(np.float32(idx)) and got != float(idx)


# --------------------------------------------------------------------- matmul
def _matmul_check(nb, a, b, precision, tol):
    got = nb.nd.matmul(nb.NDArray.array(a).gpu(), nb.NDArray.array(b).gpu(), precision).toArray()
    exp = ORACLE.matmul(a, b)
    err = rel_err(got, exp).max()
    assert err <= tol, f"max rel err {err:.3e} > {tol}"
    return got


@pytest.mark.parametrize("mkn", [(2, 2, 2), (2, 2, 1), (17, 33, 9), (64, 64, 64), (100, 40, 100)])
def test_matmul_small_shapes_simt_path(nb, mkn):
    m, k, n = mkn
    r = _rng(m * 7 + n)
    _matmul_check(nb, r.random((m, k), dtype=np.float32), r.random((k, n), dtype=np.float32), nb.TF32X3, RTOL)


@pytest.mark.parametrize("mkn", [(128, 128, 256), (256, 512, 256), (384, 1024, 640), (1000, 520, 776), (1024, 1024, 1024),
                                 (130, 36, 260)])
def test_matmul_tf32x3_vs_cblas_sgemm(nb, mkn):
    """Positive inputs, per-element relative error <= 1e-5 against the reference's cblas_sgemm."""
    m, k, n = mkn
    r = _rng(m + k + n)
    _matmul_check(nb, r.random((m, k), dtype=np.float32), r.random((k, n), dtype=np.float32), nb.TF32X3, RTOL)


@pytest.mark.parametrize("mkn", [(128, 128, 256), (256, 512, 256), (384, 1024, 640), (1000, 520, 776), (1024, 1024, 1024),
                                 (130, 36, 260), (333, 77, 129), (257, 1001, 67), (2048, 2048, 512)])
def test_matmul_bf16x3_vs_cblas_sgemm(nb, mkn):
    """BF16x3 (two bf16 parts per operand, three kind::f16 MMAs): positive inputs, per-element relative error <= 1e-5
    against cblas_sgemm, including K and N that are not multiples of 8 (the pre-pass repacks with padded rows)."""
    m, k, n = mkn
    r = _rng(m + k + n)
    _matmul_check(nb, r.random((m, k), dtype=np.float32), r.random((k, n), dtype=np.float32), nb.BF16X3, RTOL)


def test_matmul_auto_is_the_guaranteed_mode(nb):
    """AUTO (nd::matmul's default) == FP16X3 bit for bit for K >= 128 and == TF32X3 below: the modes whose error bound holds
    for every input.  BF16X3 is opt-in (include/nb200.h); on a constant matrix pair its coherent split error shows, AUTO's
    does not."""
    r = _rng(77)
    for (m, k, n), mode in (((256, 512, 256), nb.lib().nb200_gemm_resolve_precision(nb.GEMM_AUTO, 512)), ((256, 96, 256), nb.TF32X3)):
        a, b = r.random((m, k), dtype=np.float32), r.random((k, n), dtype=np.float32)
        A, B = nb.NDArray.array(a).gpu(), nb.NDArray.array(b).gpu()
        np.testing.assert_array_equal(nb.nd.matmul(A, B).toArray(), nb.nd.matmul(A, B, mode).toArray())
    # coherent inputs: every product carries the same split error (values chosen next to a bf16 rounding boundary)
    a = np.full((256, 512), 1.00390613, np.float32)
    b = np.full((512, 256), 1.00390613, np.float32)
    A, B = nb.NDArray.array(a).gpu(), nb.NDArray.array(b).gpu()
    exp = ORACLE.matmul(a, b)
    assert rel_err(nb.nd.matmul(A, B).toArray(), exp).max() <= RTOL
    assert rel_err(nb.nd.matmul(A, B, nb.BF16X3).toArray(), exp).max() <= 5e-5    # documented statistical mode


def _matmul_contract_cases(nb, prec, seed, gather_tol):
    """Everything the 1e-5 contract has to survive, through one precision mode: random, coherent, row/column dynamic range, signed
    (norm-wise), inf/NaN propagation, ragged shapes, out-of-window repair / fallback, batch with a shared operand."""
    r = _rng(seed)
    # (K = 512 / 1024: one warp per row of A in the pre-pass, 2048: one CTA per row, 9000: two-pass rows; 1001: repacking split)
    for (m, k, n) in ((256, 512, 256), (512, 1024, 768), (1000, 520, 776), (300, 136, 264), (257, 1001, 267), (384, 2048, 520), (136, 9000, 264)):
        _matmul_check(nb, r.random((m, k), dtype=np.float32), r.random((k, n), dtype=np.float32), prec, RTOL)
    # coherent: one product repeated K times, values next to rounding boundaries of the 11-bit / 8-bit parts
    for val in (1.00390613, 1.0004883, 0.33333334, 1.9990234):
        a = np.full((256, 640), val, np.float32)
        b = np.full((640, 512), val * 0.7501221, np.float32)
        got = nb.nd.matmul(nb.NDArray.array(a).gpu(), nb.NDArray.array(b).gpu(), prec).toArray()
        assert rel_err(got, ORACLE.matmul(a, b)).max() <= RTOL
    # rows / columns scaled by 2^-60 .. 2^60
    a2 = (r.random((512, 160), dtype=np.float32) * np.exp2(r.integers(-60, 61, size=(512, 1))).astype(np.float32)).astype(np.float32)
    b2 = (r.random((160, 512), dtype=np.float32) * np.exp2(r.integers(-60, 61, size=(1, 512))).astype(np.float32)).astype(np.float32)
    exp2 = ORACLE.matmul(a2, b2)
    ok = np.isfinite(exp2) & (np.abs(exp2) > 1e-30)
    got2 = nb.nd.matmul(nb.NDArray.array(a2).gpu(), nb.NDArray.array(b2).gpu(), prec).toArray()
    assert rel_err(got2[ok], exp2[ok]).max() <= RTOL
    # elements spread over 2^-20 .. 1 inside every row / column, gathered one by one through a permutation matrix
    g = ((r.random((256, 512), dtype=np.float32) + 0.5) * np.exp2(r.integers(-20, 1, size=(256, 512))).astype(np.float32)).astype(np.float32)
    perm = r.permutation(512)
    pm = np.zeros((512, 512), np.float32)
    pm[perm, np.arange(512)] = 1.0
    got = nb.nd.matmul(nb.NDArray.array(g).gpu(), nb.NDArray.array(pm).gpu(), prec).toArray()
    assert rel_err(got, g[:, perm]).max() <= gather_tol
    # inf / NaN propagate like cblas_sgemm
    a3 = r.random((256, 256), dtype=np.float32)
    b3 = r.random((256, 512), dtype=np.float32)
    a3[3, 7] = np.inf; a3[100, 5] = -np.inf; b3[9, 200] = np.nan; b3[7, 50] = 0.0
    exp3 = ORACLE.matmul(a3, b3)
    got3 = nb.nd.matmul(nb.NDArray.array(a3).gpu(), nb.NDArray.array(b3).gpu(), prec).toArray()
    np.testing.assert_array_equal(np.isnan(got3), np.isnan(exp3))
    np.testing.assert_array_equal(np.isposinf(got3), np.isposinf(exp3))
    np.testing.assert_array_equal(np.isneginf(got3), np.isneginf(exp3))
    fin = np.isfinite(exp3)
    assert rel_err(got3[fin], exp3[fin]).max() <= 2e-3
    # signed inputs, norm-wise
    a4 = (r.random((640, 1024), dtype=np.float32) * 2 - 1).astype(np.float32)
    b4 = (r.random((1024, 512), dtype=np.float32) * 2 - 1).astype(np.float32)
    got4 = nb.nd.matmul(nb.NDArray.array(a4).gpu(), nb.NDArray.array(b4).gpu(), prec).toArray()
    scale = (np.abs(a4).astype(np.float64) @ np.abs(b4).astype(np.float64)).max()
    assert np.abs(got4.astype(np.float64) - ORACLE.matmul(a4, b4)).max() / scale <= RTOL
    # out-of-window elements: repaired (few) / fallback (many) - the fallback is bit for bit the TF32X3 result
    a5 = r.random((384, 256), dtype=np.float32) + 0.25
    b5 = r.random((256, 512), dtype=np.float32) + 0.25
    a5[17, 5] = a5[17].max() * np.float32(2.0 ** -40)
    b5[:, 11] = 0.0
    b5[5, 11] = 1.5
    exp5 = ORACLE.matmul(a5, b5)
    got5 = nb.nd.matmul(nb.NDArray.array(a5).gpu(), nb.NDArray.array(b5).gpu(), prec).toArray()
    assert rel_err(got5[17, 11], exp5[17, 11]) <= RTOL and rel_err(got5, exp5).max() <= RTOL
    a6 = a5.copy()
    a6[:40, :128] *= np.float32(2.0 ** -40)
    A6, B5 = nb.NDArray.array(a6).gpu(), nb.NDArray.array(b5).gpu()
    np.testing.assert_array_equal(nb.nd.matmul(A6, B5, prec).toArray(), nb.nd.matmul(A6, B5, nb.TF32X3).toArray())
    # batch with a shared B through the C-ABI
    lib = nb.lib()
    batch, M, N, K = 3, 256, 264, 200
    a7, b7 = r.random((batch, M, K), dtype=np.float32), r.random((K, N), dtype=np.float32)
    da, db, dc = _dev(nb, a7), _dev(nb, b7), _dev(nb, np.zeros((batch, M, N), np.float32))
    assert lib.nb200_sgemm_batched(dc, da, db, batch, M, N, K, M * K, 0, M * N, prec) == 0, lib.nb200_last_error()
    got7 = _fetch(nb, dc, (batch, M, N))
    for i in range(batch):
        assert rel_err(got7[i], ORACLE.matmul(a7[i], b7)).max() <= RTOL
    for p in (da, db, dc):
        lib.nb200_free(p)


@pytest.mark.parametrize("tile", ["128", "256"])
def test_matmul_auto_both_fp16_tiles_all_contract_cases(nb, tile, monkeypatch):
    """The two FP16x3 tile shapes (256x128 with the separate cross accumulator; merged 256x256 whose per-k-block accumulator
    folds the 2^11-scaled cross products with tcgen05.mma's scale-input-d).  NB200_FP16_TILE is read per call."""
    monkeypatch.setenv("NB200_FP16_TILE", tile)
    _matmul_contract_cases(nb, nb.FP16X3, 1000 + int(tile), 2.0 ** -19)


def test_matmul_fp16x3u_all_contract_cases(nb):
    """FP16x3U (half hi parts + UNSCALED half lo parts, one accumulator per chunk, merged 256x256 tile): the same contract cases.
    A gathered element carries the split error of ONE element: <= 2^-19 at the edge of the window (include/nb200.h)."""
    _matmul_contract_cases(nb, nb.FP16X3U, 1500, 2.0 ** -18.9)


def test_matmul_auto_contract_cases(nb):
    """Whatever NB200_GEMM_AUTO resolves to (nb200_gemm_resolve_precision) passes the contract cases too."""
    _matmul_contract_cases(nb, nb.GEMM_AUTO, 1700, 2.0 ** -18.9)


@pytest.mark.parametrize("mkn", [(128, 128, 256), (384, 1024, 640), (1000, 520, 776), (333, 77, 129), (257, 1001, 67), (2048, 2048, 512),
                                 (4096, 256, 4096)])
def test_matmul_fp16x3u_vs_cblas_sgemm(nb, mkn):
    m, k, n = mkn
    r = _rng(m + k + n + 2)
    _matmul_check(nb, r.random((m, k), dtype=np.float32), r.random((k, n), dtype=np.float32), nb.FP16X3U, RTOL)


def test_matmul_unaligned_operands_stay_on_the_tensor_path(nb):
    """Leading dimensions that are not multiples of 4 and 4-byte aligned bases (row views): the FP16x3 pre-pass repacks the
    operands, so AUTO and TF32X3 calls neither fail nor drop to the fp32 SIMT kernel (one tcgen05 GEMM launch is counted by
    its runtime: a 1030^3 SIMT product takes > 1 ms, the tensor path a few tens of microseconds); results within 1e-5."""
    import time
    lib = nb.lib()
    r = _rng(4097)
    M = N = K = 1030                      # 1030 % 4 == 2
    a, b = r.random((M + 1, K), dtype=np.float32), r.random((K, N), dtype=np.float32)
    da, db, dc = _dev(nb, a), _dev(nb, b), _dev(nb, np.zeros((M, N), np.float32))
    a_view = da.value + 4 * K             # row 1 of A: 8-byte aligned base, ld % 4 == 2
    exp = ORACLE.matmul(np.ascontiguousarray(a[1:]), b)
    for prec in (nb.GEMM_AUTO, nb.TF32X3):
        assert lib.nb200_sgemm(dc, a_view, db, M, N, K, K, N, N, prec) == 0, lib.nb200_last_error()
        assert rel_err(_fetch(nb, dc, (M, N)), exp).max() <= RTOL
    lib.nb200_synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        assert lib.nb200_sgemm(dc, a_view, db, M, N, K, K, N, N, nb.GEMM_AUTO) == 0
    lib.nb200_synchronize()
    per_call = (time.perf_counter() - t0) / 10
    assert per_call < 1e-3, f"{per_call * 1e3:.2f} ms per 1030^3 call: not on the tensor path"
    for p in (da, db, dc):
        lib.nb200_free(p)


def test_fp16x3_shared_operand_fallback_in_a_late_chunk(nb, monkeypatch):
    """batch processed in several workspace chunks, B shared (stride 0), and only a LATE chunk of A holds too many
    out-of-window elements.  The fallback of that chunk needs the TF32 lo parts of the shared B although chunk 0 (eligible)
    never wrote them: the post kernel splits both operands of every chunk that falls back."""
    lib = nb.lib()
    r = _rng(34)
    batch, M, N, K = 5, 256, 128, 512
    a = r.random((batch, M, K), dtype=np.float32) + 0.25
    b = r.random((K, N), dtype=np.float32) + 0.25
    a[3, :40, :128] *= np.float32(2.0 ** -40)          # 5120 out-of-window elements in batch 3 only (> the 4096-record cap)
    da, db = _dev(nb, a), _dev(nb, b)
    dc = _dev(nb, np.zeros((batch, M, N), np.float32))
    monkeypatch.setenv("NB200_GEMM_WS_BUDGET_MB", "1")
    assert lib.nb200_sgemm_batched(dc, da, db, batch, M, N, K, M * K, 0, M * N, nb.FP16X3) == 0, lib.nb200_last_error()
    got = _fetch(nb, dc, (batch, M, N))
    for i in range(batch):
        assert rel_err(got[i], ORACLE.matmul(a[i], b)).max() <= RTOL, i
    for p in (da, db, dc):
        lib.nb200_free(p)


@pytest.mark.parametrize("mkn", [(128, 128, 256), (384, 1024, 640), (1000, 520, 776), (333, 77, 129), (257, 1001, 67), (2048, 2048, 512)])
def test_matmul_fp16x3_vs_cblas_sgemm(nb, mkn):
    """FP16x3 (half parts of row-scaled A / column-scaled B): positive inputs, per-element relative error <= 1e-5."""
    m, k, n = mkn
    r = _rng(m + k + n + 1)
    _matmul_check(nb, r.random((m, k), dtype=np.float32), r.random((k, n), dtype=np.float32), nb.FP16X3, RTOL)


def test_matmul_fp16x3_coherent_inputs_dynamic_range_and_specials(nb):
    # the constant pair that costs BF16x3 1.5e-5
    a = np.full((256, 512), 1.00390613, np.float32)
    b = np.full((512, 256), 1.00390613, np.float32)
    got = nb.nd.matmul(nb.NDArray.array(a).gpu(), nb.NDArray.array(b).gpu(), nb.FP16X3).toArray()
    assert rel_err(got, ORACLE.matmul(a, b)).max() <= RTOL
    # rows / columns scaled by 2^-60 .. 2^60
    r = _rng(42)
    a2 = (r.random((256, 160), dtype=np.float32) * np.exp2(r.integers(-60, 61, size=(256, 1))).astype(np.float32)).astype(np.float32)
    b2 = (r.random((160, 128), dtype=np.float32) * np.exp2(r.integers(-60, 61, size=(1, 128))).astype(np.float32)).astype(np.float32)
    exp2 = ORACLE.matmul(a2, b2)
    ok = np.isfinite(exp2) & (np.abs(exp2) > 1e-30)
    got2 = nb.nd.matmul(nb.NDArray.array(a2).gpu(), nb.NDArray.array(b2).gpu(), nb.FP16X3).toArray()
    assert rel_err(got2[ok], exp2[ok]).max() <= RTOL
    # inf / NaN propagate like cblas_sgemm
    a3 = r.random((256, 128), dtype=np.float32)
    b3 = r.random((128, 256), dtype=np.float32)
    a3[3, 7] = np.inf; a3[100, 5] = -np.inf; b3[9, 200] = np.nan; b3[7, 50] = 0.0
    exp3 = ORACLE.matmul(a3, b3)
    got3 = nb.nd.matmul(nb.NDArray.array(a3).gpu(), nb.NDArray.array(b3).gpu(), nb.FP16X3).toArray()
    np.testing.assert_array_equal(np.isnan(got3), np.isnan(exp3))
    np.testing.assert_array_equal(np.isposinf(got3), np.isposinf(exp3))
    np.testing.assert_array_equal(np.isneginf(got3), np.isneginf(exp3))
    fin = np.isfinite(exp3)
    assert rel_err(got3[fin], exp3[fin]).max() <= 2e-3
    # signed inputs, norm-wise; batched with a shared B through the C-ABI
    a4 = (r.random((640, 1024), dtype=np.float32) * 2 - 1).astype(np.float32)
    b4 = (r.random((1024, 512), dtype=np.float32) * 2 - 1).astype(np.float32)
    got4 = nb.nd.matmul(nb.NDArray.array(a4).gpu(), nb.NDArray.array(b4).gpu(), nb.FP16X3).toArray()
    scale = (np.abs(a4).astype(np.float64) @ np.abs(b4).astype(np.float64)).max()
    assert np.abs(got4.astype(np.float64) - ORACLE.matmul(a4, b4)).max() / scale <= RTOL
    lib = nb.lib()
    batch, M, N, K = 5, 256, 136, 200
    a5, b5 = r.random((batch, M, K), dtype=np.float32), r.random((K, N), dtype=np.float32)
    da, db, dc = _dev(nb, a5), _dev(nb, b5), _dev(nb, np.zeros((batch, M, N), np.float32))
    assert lib.nb200_sgemm_batched(dc, da, db, batch, M, N, K, M * K, 0, M * N, 4) == 0, lib.nb200_last_error()
    got5 = _fetch(nb, dc, (batch, M, N))
    for i in range(batch):
        assert rel_err(got5[i], ORACLE.matmul(a5[i], b5)).max() <= RTOL
    for p in (da, db, dc):
        lib.nb200_free(p)


def test_matmul_fp16x3_device_side_fallback_repair_and_gather(nb):
    """FP16x3 keeps 22 bits of every element within 2^-28 of its row (A) / column (B) maximum.  A few elements outside that
    window are repaired by a sparse rank-1 update after the GEMM; if there are too many (or a subnormal storm) the split
    pre-pass marks the call on the device and the gated TF32x3 fallback produces the result — bit-identical to TF32X3."""
    r = _rng(55)
    a = r.random((384, 256), dtype=np.float32) + 0.25
    b = r.random((256, 320), dtype=np.float32) + 0.25
    A, B = nb.NDArray.array(a).gpu(), nb.NDArray.array(b).gpu()
    strict = nb.nd.matmul(A, B, nb.TF32X3).toArray()
    fast = nb.nd.matmul(A, B, nb.FP16X3).toArray()
    assert rel_err(fast, ORACLE.matmul(a, b)).max() <= RTOL
    assert not np.array_equal(fast, strict)                  # eligible data really took the FP16 path
    # (1) repair: one element of A 2^-40 below its row maximum, made observable by a column of B that selects only it;
    #     one element of B 2^-40 below its column maximum, observable because the large element's partner column of A is 0
    a1, b1 = a.copy(), b.copy()
    a1[17, 5] = a1[17].max() * np.float32(2.0 ** -40)
    b1[:, 11] = 0.0
    b1[5, 11] = 1.5                                           # C[17, 11] = a1[17, 5] * 1.5
    a1[:, 9] = 0.0
    b1[:, 21] = 0.0
    b1[9, 21] = 1.0                                           # multiplies the zero column of A: contributes nothing
    b1[30, 21] = np.float32(2.0 ** -40)                      # C[:, 21] = a1[:, 30] * 2^-40
    exp1 = ORACLE.matmul(a1, b1)
    got1 = nb.nd.matmul(nb.NDArray.array(a1).gpu(), nb.NDArray.array(b1).gpu(), nb.FP16X3).toArray()
    assert rel_err(got1[17, 11], exp1[17, 11]) <= RTOL and rel_err(got1[:, 21], exp1[:, 21]).max() <= RTOL
    ok = np.abs(exp1) > 0
    assert rel_err(got1[ok], exp1[ok]).max() <= RTOL
    assert not np.array_equal(got1, nb.nd.matmul(nb.NDArray.array(a1).gpu(), nb.NDArray.array(b1).gpu(), nb.TF32X3).toArray())
    # (2) too many out-of-window elements: the fallback takes over, bit for bit the TF32X3 result
    a_bad = a.copy()
    a_bad[:40, :128] *= np.float32(2.0 ** -40)               # 5120 elements far below their rows' maxima
    Ab = nb.NDArray.array(a_bad).gpu()
    np.testing.assert_array_equal(nb.nd.matmul(Ab, B, nb.FP16X3).toArray(), nb.nd.matmul(Ab, B, nb.TF32X3).toArray())
    # the next (eligible) call is not affected by the previous call's flag
    np.testing.assert_array_equal(nb.nd.matmul(A, B, nb.FP16X3).toArray(), fast)
    # (3) batch with a shared B that carries a repaired element: the record applies to every matrix of the batch
    lib = nb.lib()
    batch = 3
    a3 = r.random((batch, 256, 256), dtype=np.float32) + 0.25
    a3[:, :, 9] = 0.0
    da, db, dc = _dev(nb, a3), _dev(nb, b1), _dev(nb, np.zeros((batch, 256, 320), np.float32))
    assert lib.nb200_sgemm_batched(dc, da, db, batch, 256, 320, 256, 256 * 256, 0, 256 * 320, 4) == 0, lib.nb200_last_error()
    got3 = _fetch(nb, dc, (batch, 256, 320))
    for i in range(batch):
        e3 = ORACLE.matmul(a3[i], b1)
        assert rel_err(got3[i][:, 21], e3[:, 21]).max() <= RTOL
    for p in (da, db, dc):
        lib.nb200_free(p)
    # (4) gather through a permutation matrix: every output IS one input element, also the ones 2^-20 below their row maximum
    g = (r.random((256, 256), dtype=np.float32) + 0.5) * np.exp2(r.integers(-20, 1, size=(256, 256))).astype(np.float32)
    perm = r.permutation(256)
    pm = np.zeros((256, 256), np.float32)
    pm[perm, np.arange(256)] = 1.0
    got = nb.nd.matmul(nb.NDArray.array(g.astype(np.float32)).gpu(), nb.NDArray.array(pm).gpu(), nb.FP16X3).toArray()
    assert rel_err(got, g.astype(np.float32)[:, perm]).max() <= 2.0 ** -20


@pytest.mark.parametrize("prec", ["TF32X3", "BF16X3"])
def test_matmul_signed_inputs_normwise_per_mode(nb, prec):
    r = _rng(13)
    a = (r.random((640, 1024), dtype=np.float32) * 2 - 1).astype(np.float32)
    b = (r.random((1024, 512), dtype=np.float32) * 2 - 1).astype(np.float32)
    got = nb.nd.matmul(nb.NDArray.array(a).gpu(), nb.NDArray.array(b).gpu(), getattr(nb, prec)).toArray()
    exp = ORACLE.matmul(a, b)
    scale = (np.abs(a).astype(np.float64) @ np.abs(b).astype(np.float64)).max()
    assert np.abs(got.astype(np.float64) - exp).max() / scale <= RTOL


def test_matmul_signed_inputs_normwise(nb):
    """Signed inputs: per-element relative error is ill-posed under cancellation; use the norm-wise
    metric of SURVEY §8 d: max|G-R| / max_ij (|A|.|B|)_ij <= 1e-5."""
    r = _rng(12)
    a = (r.random((512, 768), dtype=np.float32) * 2 - 1).astype(np.float32)
    b = (r.random((768, 384), dtype=np.float32) * 2 - 1).astype(np.float32)
    got = nb.nd.matmul(nb.NDArray.array(a).gpu(), nb.NDArray.array(b).gpu()).toArray()
    exp = ORACLE.matmul(a, b)
    scale = (np.abs(a).astype(np.float64) @ np.abs(b).astype(np.float64)).max()
    assert np.abs(got.astype(np.float64) - exp).max() / scale <= RTOL


def test_matmul_tf32x1_fast_mode_is_tf32_accurate(nb):
    r = _rng(6)
    a, b = r.random((512, 1024), dtype=np.float32), r.random((1024, 512), dtype=np.float32)
    _matmul_check(nb, a, b, nb.TF32X1, 2e-3)  # single-pass TF32: ~2^-11 per product


def test_matmul_batched_stack_equals_loop(nb):
    """The reference rejects stacks (linalg.c:240-243); the oracle is a loop of 2-D matmuls."""
    r = _rng(8)
    a, b = r.random((5, 256, 128), dtype=np.float32), r.random((5, 128, 384), dtype=np.float32)
    got = nb.nd.matmul(nb.NDArray.array(a).gpu(), nb.NDArray.array(b).gpu()).toArray()
    for i in range(5):
        assert rel_err(got[i], ORACLE.matmul(a[i], b[i])).max() <= RTOL


def test_matmul_error_behaviour(nb):
    A = nb.NDArray.array(np.ones((4, 5), np.float32)).gpu()
    B = nb.NDArray.array(np.ones((4, 5), np.float32)).gpu()
    with pytest.raises(nb.BackendError, match="Shape mismatch for matmul"):
        nb.nd.matmul(A, B)
    with pytest.raises(nb.BackendError, match="Device mismatch"):
        nb.nd.matmul(A, nb.NDArray.array(np.ones((5, 4), np.float32)))


def test_dot_variants(nb):
    r = _rng(10)
    a, x = r.random((300, 1000), dtype=np.float32), r.random(1000, dtype=np.float32)
    got = nb.nd.dot(nb.NDArray.array(a).gpu(), nb.NDArray.array(x).gpu()).toArray()
    assert rel_err(got, ORACLE.dot(a, x) if oracle.ref.available else oracle.port.gemv(a, x)).max() <= RTOL
    # more rows than a grid dimension holds, odd row length (no 16-byte alignment past row 0), 4-row x long vectors; dyadic: exact
    for rows, cols in ((70001, 37), (4, 100003), (1, 8)):
        a2, x2 = _set_p2((rows, cols), rows), _set_p2(cols, cols)
        got2 = nb.nd.dot(nb.NDArray.array(a2).gpu(), nb.NDArray.array(x2).gpu()).toArray()
        np.testing.assert_array_equal(got2, (a2.astype(np.float64) @ x2.astype(np.float64)).astype(np.float32))
    v = _set_p2(5000, 3)
    w = _set_p2(5000, 4)
    got = float(nb.nd.dot(nb.NDArray.array(v).gpu(), nb.NDArray.array(w).gpu()).toArray())
    assert got == pytest.approx(float((v.astype(np.float64) * w).sum()), rel=1e-6)


# --------------------------------------------------------------------- BASELINE sizes: size-independent properties
def test_config3_chain_full_size_properties(nb):
    """8192x8192 a*b+c: fused == unfused bit for bit; row/col broadcast == materialised; spot rows vs oracle."""
    n = 8192
    r = _rng(5)
    a = r.random((n, n), dtype=np.float32)
    brow = r.random(n, dtype=np.float32)
    ccol = r.random((n, 1), dtype=np.float32)
    A, B, Cc = nb.NDArray.array(a).gpu(), nb.NDArray.array(brow).gpu(), nb.NDArray.array(ccol).gpu()
    fused = nb.nd.mul_add(A, B, Cc)
    two = A * B + Cc
    d = (fused - two)
    assert nb.nd.max(d) == 0.0 and nb.nd.min(d) == 0.0
    rows = [0, 1, 4095, 8191]
    host = fused.toArray()
    for i in rows:
        eq_zero_sign_free(host[i], oracle.port.mul_add(a[i], brow, ccol[i]))
    # linearity-style checksum: sum over an exact-set variant is order independent
    p = _set_p2((n, n), 1)
    P = nb.NDArray.array(p).gpu()
    s_full = nb.nd.sum(P)
    s_rows = nb.nd.sum(nb.nd.sum(P, 1))
    s_cols = nb.nd.sum(nb.nd.sum(P, 0))
    assert s_full == s_rows == s_cols == float(p.astype(np.float64).sum())
    # ... and the axis sums themselves against the oracle (exact set: every summation order gives the same bits)
    np.testing.assert_array_equal(nb.nd.sum(P, 0).toArray(), oracle.port.reduce_axis("sum", p, 0))
    np.testing.assert_array_equal(nb.nd.sum(P, 1).toArray(), oracle.port.reduce_axis("sum", p, 1))
    assert s_full == float(ORACLE.reduce_full("sum", p))


def test_config4_sum_argmax_2pow28(nb):
    n = 1 << 28
    x = _set_p(n, 8)
    A = nb.NDArray.array(x).gpu()
    assert nb.nd.sum(A) == float(x.astype(np.float64).sum())   # exact set: any order gives the same integer
    assert nb.nd.sum(A) == float(ORACLE.reduce_full("sum", x))  # the reference's own sequential fp32 loop on the full 2^28 input
    x[123456789] = 7.0
    x[200000001] = 7.0
    A = nb.NDArray.array(x).gpu()
    assert nb.nd.argmax(A) == float(np.float32(123456789))
    assert nb.nd.argmin(A) == float(np.argmin(x))
    assert nb.nd.argmax(A) == float(ORACLE.argminmax(True, x)) and nb.nd.argmin(A) == float(ORACLE.argminmax(False, x))
    assert nb.nd.max(A) == float(ORACLE.reduce_full("max", x)) and nb.nd.min(A) == float(ORACLE.reduce_full("min", x))


def test_config2_matmul_4096_checksum(nb):
    """4096^3: compare 64 sampled rows against the reference's cblas_sgemm, and the full result through
    the checksum identity  (1^T A) B == 1^T (A B)  in fp64."""
    n = 4096
    r = _rng(3)
    a, b = r.random((n, n), dtype=np.float32), r.random((n, n), dtype=np.float32)
    got = nb.nd.matmul(nb.NDArray.array(a).gpu(), nb.NDArray.array(b).gpu()).toArray()
    rows = r.choice(n, 64, replace=False)
    exp = ORACLE.matmul(a[rows], b)
    assert rel_err(got[rows], exp).max() <= RTOL
    lhs = a.astype(np.float64).sum(0) @ b.astype(np.float64)
    rhs = got.astype(np.float64).sum(0)
    assert np.abs(lhs - rhs).max() / np.abs(lhs).max() <= RTOL   # systematic part of the error (measured ~1.4e-6)


# --------------------------------------------------------------------- C-ABI level GEMM edge cases
def _dev(nb, arr):
    import ctypes as C
    lib = nb.lib()
    p = C.c_void_p()
    assert lib.nb200_alloc(C.byref(p), arr.nbytes) == 0
    assert lib.nb200_copy_h2d(p, arr.ctypes.data, arr.nbytes) == 0
    return p


def _fetch(nb, p, shape):
    out = np.empty(shape, np.float32)
    assert nb.lib().nb200_copy_d2h(out.ctypes.data, p, out.nbytes) == 0
    return out


def test_sgemm_leading_dimensions_and_ragged_edges(nb):
    """lda/ldb/ldc larger than the logical extents (sub-matrix views) and M, N, K not multiples of the tile."""
    lib = nb.lib()
    r = _rng(31)
    M, N, K, lda, ldb, ldc = 300, 200, 136, 160, 256, 208
    abuf, bbuf = r.random((M, lda), dtype=np.float32), r.random((K, ldb), dtype=np.float32)
    cbuf = np.full((M, ldc), -1.0, np.float32)
    da, db, dc = _dev(nb, abuf), _dev(nb, bbuf), _dev(nb, cbuf)
    for prec, tol in ((0, RTOL), (1, 2e-3), (2, RTOL), (3, RTOL)):
        assert lib.nb200_sgemm(dc, da, db, M, N, K, lda, ldb, ldc, prec) == 0, lib.nb200_last_error()
        got = _fetch(nb, dc, (M, ldc))
        exp = ORACLE.matmul(np.ascontiguousarray(abuf[:, :K]), np.ascontiguousarray(bbuf[:, :N]))
        assert rel_err(got[:, :N], exp).max() <= tol
        assert (got[:, N:] == -1.0).all()   # padding columns untouched
    for p in (da, db, dc):
        lib.nb200_free(p)


def test_sgemm_batched_shared_operand_and_many_batches(nb):
    """strideB = 0 (one B for the whole batch) and a batch large enough to exercise several tiles per matrix."""
    lib = nb.lib()
    r = _rng(32)
    batch, M, N, K = 9, 256, 128, 96
    a, b = r.random((batch, M, K), dtype=np.float32), r.random((K, N), dtype=np.float32)
    da, db = _dev(nb, a), _dev(nb, b)
    dc = _dev(nb, np.zeros((batch, M, N), np.float32))
    for prec in (0, 2):
        assert lib.nb200_copy_h2d(dc, np.zeros((batch, M, N), np.float32).ctypes.data, batch * M * N * 4) == 0
        assert lib.nb200_sgemm_batched(dc, da, db, batch, M, N, K, M * K, 0, M * N, prec) == 0, lib.nb200_last_error()
        got = _fetch(nb, dc, (batch, M, N))
        for i in range(batch):
            assert rel_err(got[i], ORACLE.matmul(a[i], b)).max() <= RTOL
    for p in (da, db, dc):
        lib.nb200_free(p)


def test_sgemm_batched_more_matrices_than_grid_z(nb):
    """70 000 tiny matrices (3x5 . 5x4: the fp32 SIMT path puts the batch on gridDim.z, limit 65 535): processed in two launches,
    every matrix equal to numpy's fp32 product of the same dyadic operands (exact)."""
    lib = nb.lib()
    r = _rng(65)
    batch, M, K, N = 70000, 3, 5, 4
    a = (r.integers(-8, 9, size=(batch, M, K)).astype(np.float32) / 8)
    b = (r.integers(-8, 9, size=(batch, K, N)).astype(np.float32) / 8)
    da, db, dc = _dev(nb, a), _dev(nb, b), _dev(nb, np.full((batch, M, N), -1.0, np.float32))
    assert lib.nb200_sgemm_batched(dc, da, db, batch, M, N, K, M * K, K * N, M * N, nb.GEMM_AUTO) == 0, lib.nb200_last_error()
    np.testing.assert_array_equal(_fetch(nb, dc, (batch, M, N)), np.matmul(a, b))
    for p in (da, db, dc):
        lib.nb200_free(p)


@pytest.mark.parametrize("prec", [0, 2, 4])
def test_sgemm_batched_in_several_workspace_chunks(nb, prec, monkeypatch):
    """A workspace budget of 1 MiB forces chunks of 2 + 2 + 1 matrices; B is shared (stride 0), i.e. split once with the
    first chunk and reused by the shorter last one."""
    lib = nb.lib()
    r = _rng(33)
    batch, M, N, K = 5, 256, 128, 512
    a, b = r.random((batch, M, K), dtype=np.float32), r.random((K, N), dtype=np.float32)
    da, db = _dev(nb, a), _dev(nb, b)
    dc = _dev(nb, np.zeros((batch, M, N), np.float32))
    monkeypatch.setenv("NB200_GEMM_WS_BUDGET_MB", "1")
    assert lib.nb200_sgemm_batched(dc, da, db, batch, M, N, K, M * K, 0, M * N, prec) == 0, lib.nb200_last_error()
    got = _fetch(nb, dc, (batch, M, N))
    for i in range(batch):
        assert rel_err(got[i], ORACLE.matmul(a[i], b)).max() <= RTOL
    for p in (da, db, dc):
        lib.nb200_free(p)


@pytest.mark.parametrize("prec", [4, 5, 3])
@pytest.mark.parametrize("shared", ["none", "B", "A"])
def test_sgemm_batched_many_chunks_repair_and_fallback_in_late_chunks(nb, prec, shared, monkeypatch):
    """13 matrices processed in 13 workspace chunks (1 MiB budget).  Every matrix against cblas_sgemm; an out-of-window element in a
    LATE chunk (sparse repair), a shared operand that carries a repaired element (prepared once with chunk 0: its records must
    survive for the whole call), and a late chunk with too many out-of-window elements (that chunk falls back: bit-identical to
    TF32X3; the flag is per call, so later chunks follow)."""
    monkeypatch.setenv("NB200_GEMM_WS_BUDGET_MB", "1")
    lib = nb.lib()
    r = _rng(700 + prec)
    batch, M, N, K = 13, 256, 264, 200
    a = r.random((batch, M, K) if shared != "A" else (M, K), dtype=np.float32) + 0.25
    b = r.random((batch, K, N) if shared != "B" else (K, N), dtype=np.float32) + 0.25
    A = lambda i: a[i] if shared != "A" else a
    Bm = lambda i: b[i] if shared != "B" else b
    # repair cases: an element 2^-40 below its row maximum, observable through a B column that selects only it
    if shared != "A":
        a[11, 17, 5] = a[11, 17].max() * np.float32(2.0 ** -40)
    else:
        a[17, 5] = a[17].max() * np.float32(2.0 ** -40)
    if shared != "B":
        b[11, :, 9] = 0.0; b[11, 5, 9] = 1.5
    else:
        b[:, 9] = 0.0; b[5, 9] = 1.5
    da, db = _dev(nb, a), _dev(nb, b)
    dc = _dev(nb, np.zeros((batch, M, N), np.float32))
    sa, sb = (0 if shared == "A" else M * K), (0 if shared == "B" else K * N)

    def run(p):
        assert lib.nb200_sgemm_batched(dc, da, db, batch, M, N, K, sa, sb, M * N, p) == 0, lib.nb200_last_error()
        return _fetch(nb, dc, (batch, M, N))
    got = run(prec)
    for i in range(batch):
        exp = ORACLE.matmul(A(i), Bm(i))
        assert rel_err(got[i], exp).max() <= RTOL, i
    # many calls back to back: workspace sets / events are reused across calls
    for _ in range(3):
        np.testing.assert_array_equal(run(prec), got)
    # too many out-of-window elements in a late chunk only -> TF32x3 fallback of the chunks that see the flag, 1e-5 everywhere
    if shared != "A":
        a2 = a.copy()
        a2[12, :40, :128] *= np.float32(2.0 ** -40)
        assert lib.nb200_copy_h2d(da, a2.ctypes.data, a2.nbytes) == 0
        got2 = run(prec)
        for i in range(batch):
            assert rel_err(got2[i], ORACLE.matmul(a2[i], Bm(i))).max() <= RTOL, i
        np.testing.assert_array_equal(got2[12], nb.nd.matmul(nb.NDArray.array(a2[12]).gpu(), nb.NDArray.array(Bm(12)).gpu(), nb.TF32X3).toArray())
    for p in (da, db, dc):
        lib.nb200_free(p)


@pytest.mark.parametrize("mkn", [(1024, 512, 768), (700, 260, 132), (64, 64, 64), (1025, 128, 128), (257, 256, 256), (2049, 512, 64), (4097, 256, 128)])
def test_sgemm_host_pipeline_matches_resident_call(nb, mkn):
    """nb200_sgemm_host (B once, A row blocks in; every block's split + GEMM + download on one of two worker streams) within 1e-5 of
    cblas_sgemm and of the resident call.  M % 256 == 1: the one-row tail joins the previous row block instead of being
    refused by the tensor path; 4097 rows = 16 blocks alternating between the worker streams."""
    lib = nb.lib()
    M, K, N = mkn
    r = _rng(M)
    a, b = r.random((M, K), dtype=np.float32), r.random((K, N), dtype=np.float32)
    c = np.empty((M, N), np.float32)
    for prec in (0, 2, 3):
        c[:] = -1.0
        assert lib.nb200_sgemm_host(c.ctypes.data, a.ctypes.data, b.ctypes.data, M, N, K, prec) == 0, lib.nb200_last_error()
        assert rel_err(c, ORACLE.matmul(a, b)).max() <= RTOL
        resident = nb.nd.matmul(nb.NDArray.array(a).gpu(), nb.NDArray.array(b).gpu(), prec).toArray()
        assert rel_err(c, resident).max() <= RTOL
        c2 = np.full((M, N), -1.0, np.float32)     # a second call reuses streams, events and staging buffers: same bits
        assert lib.nb200_sgemm_host(c2.ctypes.data, a.ctypes.data, b.ctypes.data, M, N, K, prec) == 0, lib.nb200_last_error()
        np.testing.assert_array_equal(c2, c)


def test_sgemm_batched_host_pipeline(nb):
    """nb200_sgemm_batched_host: chunks of matrices upload / multiply / download on three streams; every matrix within 1e-5 of
    cblas_sgemm, for a batch that does not divide into the chunk count and for a single matrix."""
    lib = nb.lib()
    r = _rng(91)
    for batch, M, K, N in ((11, 256, 200, 264), (1, 300, 136, 128), (3, 64, 64, 64)):
        a, b = r.random((batch, M, K), dtype=np.float32), r.random((batch, K, N), dtype=np.float32)
        c = np.full((batch, M, N), -1.0, np.float32)
        assert lib.nb200_sgemm_batched_host(c.ctypes.data, a.ctypes.data, b.ctypes.data, batch, M, N, K, nb.GEMM_AUTO) == 0, lib.nb200_last_error()
        for i in range(batch):
            assert rel_err(c[i], ORACLE.matmul(a[i], b[i])).max() <= RTOL, (batch, i)


def test_cuda_graph_capture_replays_a_sequence_of_calls(nb):
    """nb200_graph_begin / end / launch: a chain of small elementwise calls and one nd::matmul recorded once, replayed on new data;
    every replay matches the oracle (the recorded FP16x3 call carries its own control-block memset)."""
    import ctypes as C
    lib = nb.lib()
    r = _rng(55)
    n = 1 << 16
    a, b = r.random(n, dtype=np.float32), r.random(n, dtype=np.float32)
    m1, m2 = r.random((256, 384), dtype=np.float32), r.random((384, 264), dtype=np.float32)
    da, db, dt, do = _dev(nb, a), _dev(nb, b), _dev(nb, np.zeros(n, np.float32)), _dev(nb, np.zeros(n, np.float32))
    dm1, dm2, dmo = _dev(nb, m1), _dev(nb, m2), _dev(nb, np.zeros((256, 264), np.float32))
    shp, st = (C.c_int64 * 1)(n), (C.c_int64 * 1)(1)
    assert lib.nb200_sgemm(dmo, dm1, dm2, 256, 264, 384, 384, 264, 264, nb.GEMM_AUTO) == 0      # warm-up: workspace allocated outside the capture
    assert lib.nb200_synchronize() == 0
    assert lib.nb200_graph_begin() == 0, lib.nb200_last_error()
    assert lib.nb200_ew_binary(2, dt, da, db, 1, shp, st, st) == 0            # t = a * b
    assert lib.nb200_ew_binary(0, do, dt, da, 1, shp, st, st) == 0            # o = t + a
    assert lib.nb200_ew_unary(1, do, do, n, 0.0, 0.0) == 0                    # o = sqrt(o)
    assert lib.nb200_sgemm(dmo, dm1, dm2, 256, 264, 384, 384, 264, 264, nb.GEMM_AUTO) == 0
    g = C.c_void_p()
    assert lib.nb200_graph_end(C.byref(g)) == 0, lib.nb200_last_error()
    for rep in range(3):
        a2 = r.random(n, dtype=np.float32)
        m1b = r.random((256, 384), dtype=np.float32)
        assert lib.nb200_copy_h2d(da, a2.ctypes.data, a2.nbytes) == 0
        assert lib.nb200_copy_h2d(dm1, m1b.ctypes.data, m1b.nbytes) == 0
        assert lib.nb200_graph_launch(g) == 0, lib.nb200_last_error()
        assert lib.nb200_synchronize() == 0
        exp = ORACLE.unary("sqrt", ORACLE.binary("add", ORACLE.binary("mul", a2, b), a2))
        np.testing.assert_array_equal(_fetch(nb, do, (n,)), exp)
        assert rel_err(_fetch(nb, dmo, (256, 264)), ORACLE.matmul(m1b, m2)).max() <= RTOL
    assert lib.nb200_graph_destroy(g) == 0
    for p in (da, db, dt, do, dm1, dm2, dmo):
        lib.nb200_free(p)


# --------------------------------------------------------------------- comparisons (SURVEY §8 f, N2)
@pytest.mark.parametrize("op", ["equal", "not_equal", "greater", "greater_equal", "less", "less_equal"])
def test_comparisons_vs_oracle(nb, op):
    r = _rng(14)
    a = r.integers(-3, 4, size=(257, 130)).astype(np.float32)
    b = r.integers(-3, 4, size=(257, 130)).astype(np.float32)
    A, B = nb.NDArray.array(a).gpu(), nb.NDArray.array(b).gpu()
    np.testing.assert_array_equal(nb.nd.binary(op, A, B).toArray(), ORACLE.binary(op, a, b))
    # broadcast: the port (the reference's less / equal read out of bounds when shapes differ, logic.c:229-244, :534-553)
    np.testing.assert_array_equal(nb.nd.binary(op, A, nb.NDArray.array(b[0]).gpu()).toArray(), oracle.port.binary(op, a, b[0]))
    np.testing.assert_array_equal(nb.nd.binary(op, A, 1.0).toArray(), ORACLE.binary(op, a, np.float32(1.0)))
    # NaN operands: every predicate is ordered in the reference (incl. not_equal)
    a2 = a.copy().reshape(-1)[:4096]
    a2[::7] = np.nan
    got = nb.nd.binary(op, nb.NDArray.array(a2).gpu(), nb.NDArray.array(b.reshape(-1)[:4096]).gpu()).toArray()
    np.testing.assert_array_equal(got, oracle.port.binary(op, a2, b.reshape(-1)[:4096]))
    assert (got[::7] == 0).all()


def test_comparison_and_transpose_phpt_golden_vectors(nb):
    """tests/logic/003..008-*.phpt and tests/manipulation/001-ndarray-transpose.phpt (tests/golden/compare_vectors.json): the
    comparison kernels give the printed 0 / 1 masks; the 2-D transposes go through nb200_transpose2d (the 1-D case is the identity
    and the (1, 1, 4) case only permutes unit dimensions: no data movement, not kernel work)."""
    import ctypes as C
    import json
    import os
    lib = nb.lib()
    for rec in json.load(open(os.path.join(os.path.dirname(__file__), "golden", "compare_vectors.json")))["vectors"]:
        exp = np.asarray(rec["expect"], np.float32)
        if rec["op"] != "transpose":
            a, b = (nb.NDArray.array(np.asarray(v, np.float32)).gpu() for v in rec["args"])
            np.testing.assert_array_equal(nb.nd.binary(rec["op"], a, b).toArray(), exp, err_msg=str(rec))
            continue
        x = np.asarray(rec["args"][0], np.float32)
        if x.ndim != 2:
            assert exp.size == x.size and exp.ravel().tolist() == x.ravel().tolist()
            continue
        dx, do = _dev(nb, x), _dev(nb, np.zeros(exp.shape, np.float32))
        assert lib.nb200_transpose2d(do, dx, x.shape[0], x.shape[1]) == 0, lib.nb200_last_error()
        np.testing.assert_array_equal(_fetch(nb, do, exp.shape), exp, err_msg=str(rec))
        lib.nb200_free(dx); lib.nb200_free(do)


def test_array_equal(nb):
    a = _rng(1).random((100, 37), dtype=np.float32)
    A = nb.NDArray.array(a).gpu()
    assert nb.nd.array_equal(A, nb.NDArray.array(a.copy()).gpu())
    b = a.copy()
    b[50, 3] += 1
    assert not nb.nd.array_equal(A, nb.NDArray.array(b).gpu())
    assert not nb.nd.array_equal(A, nb.NDArray.array(a[:50]).gpu())


# --------------------------------------------------------------------- empty / degenerate inputs
def test_empty_and_degenerate_inputs(nb):
    e = nb.NDArray.array(np.zeros((0,), np.float32)).gpu()
    assert (e + e).shape == (0,) and (e + e).toArray().size == 0
    assert nb.nd.exp(e).toArray().size == 0
    assert nb.nd.sum(e) == 0.0 and nb.nd.prod(e) == 1.0          # empty loops of arithmetics.c:44-46, 66-68
    with pytest.raises(nb.BackendError, match="empty sequence"):  # calculation.c:169-172
        nb.nd.argmax(e)
    z = nb.NDArray.array(np.zeros((4, 0, 3), np.float32)).gpu()
    assert nb.nd.sum(z, 1).shape == (4, 3)
    one = nb.NDArray.array(np.array([[3.5]], np.float32)).gpu()
    assert nb.nd.argmax(one) == 0.0 and nb.nd.sum(one) == 3.5 and nb.nd.max(one) == 3.5
    with pytest.raises(nb.BackendError, match="out of bounds"):   # ndarray.c:534-538
        nb.nd.sum(one, 5)
    with pytest.raises(nb.BackendError, match="broadcast"):       # arithmetics.c:199-202
        nb.nd.add(nb.NDArray.array(np.ones((3, 4), np.float32)).gpu(), nb.NDArray.array(np.ones((5,), np.float32)).gpu())
    with pytest.raises(nb.BackendError, match="Device mismatch"):  # arithmetics.c:163-166
        nb.nd.add(one, nb.NDArray.array(np.ones((1, 1), np.float32)))
    # 0-dim op 0-dim, and a CPU scalar next to a GPU array (exempt from the device check)
    s = (one * 2.0).toArray()
    assert s[0, 0] == 7.0


def test_config5_batched_matmul_slice_properties(nb):
    """BASELINE configs[4] is 1024 x (2048x2048); one GPU's shard at N = 16 (64 matrices, 1/16 of the job) through
    the batched entry point: two sampled matrices against the reference's cblas_sgemm and every matrix through the
    fp64 checksum identity (1^T A_i) B_i == 1^T C_i."""
    import ctypes as C
    import torch
    lib = nb.lib()
    batch, n = 64, 2048
    g = torch.Generator(device="cuda").manual_seed(10)
    a = torch.rand(batch, n, n, device="cuda", generator=g)
    b = torch.rand(batch, n, n, device="cuda", generator=g)
    c = torch.empty(batch, n, n, device="cuda")
    torch.cuda.synchronize()
    assert lib.nb200_sgemm_batched(c.data_ptr(), a.data_ptr(), b.data_ptr(), batch, n, n, n, n * n, n * n, n * n, 0) == 0
    assert lib.nb200_synchronize() == 0
    for i in (0, 37):
        assert rel_err(c[i].cpu().numpy(), ORACLE.matmul(a[i].cpu().numpy(), b[i].cpu().numpy())).max() <= RTOL
    lhs = torch.bmm(a.double().sum(1, keepdim=True), b.double()).squeeze(1)   # (batch, n) fp64 checker on the GPU
    rhs = c.double().sum(1)
    assert float(((lhs - rhs).abs().amax(1) / lhs.abs().amax(1)).max()) <= RTOL


def test_mean_composition(nb):
    """nd::mean = sum / n (numpower.c:2642-2688), composed from the same kernels."""
    x = _set_p2((300, 40), 5)
    A = nb.NDArray.array(x).gpu()
    assert nb.nd.mean(A) == float(np.float32(ORACLE.reduce_full("sum", x)) / np.float32(x.size))
    for axis in (0, 1):
        exp = ORACLE.binary("div", oracle.port.reduce_axis("sum", x, axis), np.float32(x.shape[axis]))
        np.testing.assert_array_equal(nb.nd.mean(A, axis).toArray(), exp)


def test_statistics_compositions(nb):
    """variance / std / average = the reference's own compositions of hot-path ops (statistics.c:87-153)."""
    x = _set_p2((200, 50), 6)
    w = (_rng(7).integers(1, 9, size=(200, 50)).astype(np.float32) / 8)
    A, W = nb.NDArray.array(x).gpu(), nb.NDArray.array(w).gpu()
    n = np.float32(x.size)
    mean = np.float32(ORACLE.reduce_full("sum", x)) / n
    var_ref = np.float32(ORACLE.reduce_full("sum", ORACLE.binary("pow", ORACLE.unary("abs", ORACLE.binary("sub", x, mean)), np.float32(2.0)))) / n
    assert nb.nd.variance(A) == pytest.approx(float(var_ref), rel=RTOL)
    std_ref = np.sqrt(np.float32(ORACLE.reduce_full("sum", ORACLE.binary("pow", ORACLE.binary("sub", x, mean), np.float32(2.0)))) / n, dtype=np.float32)
    assert nb.nd.std(A) == pytest.approx(float(std_ref), rel=RTOL)
    avg_ref = np.float32(ORACLE.reduce_full("sum", ORACLE.binary("mul", x, w))) / np.float32(ORACLE.reduce_full("sum", w))
    assert nb.nd.average(A, W) == pytest.approx(float(avg_ref), rel=RTOL)
    assert nb.nd.average(A) == float(mean)


def test_arrays_beyond_2_pow_31_elements(nb):
    """Maximum sizes: 64-bit extents end to end.  The reference stops at 2^31 elements / 4 GiB (`int` shapes ndarray.h:61-74,
    `vmalloc(void **, unsigned int)` gpu_alloc.c:11), so there is no oracle run at this size: the expected values follow from the
    few elements planted at known positions of an otherwise constant 8.6 GB array (every sum below is exact in any order)."""
    import ctypes as C
    import torch
    lib = nb.lib()
    n = (1 << 31) + (1 << 20) + 3
    if torch.cuda.mem_get_info()[0] < 3 * n * 4:
        pytest.skip("needs 26 GB of free device memory")
    pa, po = C.c_void_p(), C.c_void_p()
    assert lib.nb200_alloc(C.byref(pa), n * 4) == 0, lib.nb200_last_error()
    assert lib.nb200_alloc(C.byref(po), n * 4) == 0, lib.nb200_last_error()

    def poke(p, i, v):
        x = np.array([v], np.float32)
        assert lib.nb200_copy_h2d(C.c_void_p(p.value + 4 * i), x.ctypes.data, 4) == 0

    def peek(p, i):
        x = np.empty(1, np.float32)
        assert lib.nb200_copy_d2h(x.ctypes.data, C.c_void_p(p.value + 4 * i), 4) == 0
        return float(x[0])

    def full(op, is_arg=False):
        v = C.c_float()
        fn = lib.nb200_argminmax_host if is_arg else lib.nb200_reduce_full_host
        assert fn(op, C.byref(v), pa, n) == 0, lib.nb200_last_error()
        return v.value

    try:
        assert lib.nb200_fill(pa, 0.0, n) == 0
        far = (1 << 31) + 5
        for i in (0, far, n - 1):
            poke(pa, i, 1.0)
        assert full(oracle.RED_OPS["sum"]) == 3.0
        poke(pa, n - 2, 5.0)
        poke(pa, n - 3, -2.0)
        assert full(oracle.RED_OPS["max"]) == 5.0 and full(oracle.RED_OPS["min"]) == -2.0
        assert full(1, True) == float(np.float32(n - 2)) and full(0, True) == float(np.float32(n - 3))   # (float) idx, calculation.c:160-190
        # elementwise over all n elements: out = a + 1, then exp in place
        assert lib.nb200_ew_binary_scalar(oracle.BIN_OPS["add"], po, pa, 1.0, 0, n) == 0, lib.nb200_last_error()
        assert [peek(po, i) for i in (7, far, n - 1, n - 2, n - 3)] == [1.0, 2.0, 2.0, 6.0, -1.0]
        assert lib.nb200_ew_unary(oracle.UN_OPS["negative"], po, po, n, 0.0, 0.0) == 0, lib.nb200_last_error()
        assert [peek(po, i) for i in (7, far, n - 2)] == [-1.0, -2.0, -6.0]
        # axis reductions of the (65537, 32768) view = 2^31 + 2^15 elements: ones at [0, 0] and [65536, 5]
        rows, cols = 65537, 32768
        assert rows * cols <= n - 3
        out0, out1 = np.empty(cols, np.float32), np.empty(rows, np.float32)
        pr = C.c_void_p()
        assert lib.nb200_alloc(C.byref(pr), rows * 4) == 0
        assert lib.nb200_reduce_axis(oracle.RED_OPS["sum"], pr, pa, 1, rows, cols, nb.ORDER_TREE) == 0, lib.nb200_last_error()
        assert lib.nb200_copy_d2h(out0.ctypes.data, pr, cols * 4) == 0
        e0 = np.zeros(cols, np.float32); e0[0] = 1; e0[5] = 1
        np.testing.assert_array_equal(out0, e0)
        assert lib.nb200_reduce_axis(oracle.RED_OPS["sum"], pr, pa, rows, cols, 1, nb.ORDER_TREE) == 0, lib.nb200_last_error()
        assert lib.nb200_copy_d2h(out1.ctypes.data, pr, rows * 4) == 0
        e1 = np.zeros(rows, np.float32); e1[0] = 1; e1[65536] = 1
        np.testing.assert_array_equal(out1, e1)
        lib.nb200_free(pr)
    finally:
        lib.nb200_free(pa); lib.nb200_free(po)


def test_matmul_gemv_broadcast_beyond_2_pow_31_elements(nb):
    """An operand of (2^20 + 1) x 2048 = 2^31 + 2048 elements through nd::matmul (AUTO and TF32X3), nd::dot (gemv), a row / column
    broadcast chain and the boolean reductions.  Dyadic values (k/64): every product and every sum is exact in fp32 whatever the
    order, so the expected result follows from numpy on the few distinct rows."""
    import ctypes as C
    import torch
    lib = nb.lib()
    M, K, N = (1 << 20) + 1, 2048, 128
    if torch.cuda.mem_get_info()[0] < 6 * M * K * 4:
        pytest.skip("needs ~52 GB of free device memory")
    r = _rng(2031)
    b = (r.integers(-64, 65, size=(K, N)).astype(np.float32) / 64)
    x = (r.integers(-64, 65, size=K).astype(np.float32) / 64)
    planted = {0: None, (1 << 19) + 7: None, M - 1: None}
    for i in planted:
        planted[i] = (r.integers(-64, 65, size=K).astype(np.float32) / 64)
    pa, pb, pc, px, py, pe = (C.c_void_p() for _ in range(6))
    for p, nbytes in ((pa, M * K * 4), (pb, K * N * 4), (pc, M * N * 4), (px, K * 4), (py, M * 4), (pe, M * K * 4)):
        assert lib.nb200_alloc(C.byref(p), nbytes) == 0, lib.nb200_last_error()
    try:
        assert lib.nb200_fill(pa, 0.5, M * K) == 0
        for i, row in planted.items():
            assert lib.nb200_copy_h2d(C.c_void_p(pa.value + 4 * i * K), row.ctypes.data, K * 4) == 0
        assert lib.nb200_copy_h2d(pb, b.ctypes.data, b.nbytes) == 0 and lib.nb200_copy_h2d(px, x.ctypes.data, x.nbytes) == 0
        exp_c = np.tile((0.5 * b.astype(np.float64).sum(axis=0)).astype(np.float32), (M, 1))
        exp_y = np.full(M, np.float32(0.5 * x.astype(np.float64).sum()), np.float32)
        for i, row in planted.items():
            exp_c[i] = (row.astype(np.float64) @ b.astype(np.float64)).astype(np.float32)
            exp_y[i] = np.float32(row.astype(np.float64) @ x.astype(np.float64))
        got = np.empty((M, N), np.float32)
        for prec in (nb.GEMM_AUTO, nb.TF32X3):
            assert lib.nb200_fill(pc, -1.0, M * N) == 0
            assert lib.nb200_sgemm(pc, pa, pb, M, N, K, K, N, N, prec) == 0, lib.nb200_last_error()
            assert lib.nb200_copy_d2h(got.ctypes.data, pc, got.nbytes) == 0
            np.testing.assert_array_equal(got, exp_c)
        goty = np.empty(M, np.float32)
        assert lib.nb200_gemv(py, pa, px, M, K) == 0, lib.nb200_last_error()
        assert lib.nb200_copy_d2h(goty.ctypes.data, py, goty.nbytes) == 0
        np.testing.assert_array_equal(goty, exp_y)
        # e = a * x (row vector) + y (column vector): fused broadcast chain over 2^31 + 2048 output elements
        shp = (C.c_int64 * 2)(M, K)
        s_full, s_row, s_col = (C.c_int64 * 2)(K, 1), (C.c_int64 * 2)(0, 1), (C.c_int64 * 2)(1, 0)
        assert lib.nb200_ew_mul_add(pe, pa, px, py, 2, shp, s_full, s_row, s_col) == 0, lib.nb200_last_error()
        rowbuf = np.empty(K, np.float32)
        for i in (0, 1, (1 << 19) + 7, (1 << 20) - 1, M - 1):
            assert lib.nb200_copy_d2h(rowbuf.ctypes.data, C.c_void_p(pe.value + 4 * i * K), K * 4) == 0
            arow = planted.get(i, np.full(K, 0.5, np.float32))
            np.testing.assert_array_equal(rowbuf, arow * x + exp_y[i])
        # boolean reductions over the whole operand: no zero in a (0.5 everywhere, planted rows may hold zeros -> overwrite them)
        flag = C.c_int(-1)
        for i in planted:
            assert lib.nb200_fill(C.c_void_p(pa.value + 4 * i * K), 0.25, K) == 0
        assert lib.nb200_all(C.byref(flag), pa, M * K) == 0 and flag.value == 1
        zero = np.zeros(1, np.float32)
        assert lib.nb200_copy_h2d(C.c_void_p(pa.value + 4 * (M * K - 1)), zero.ctypes.data, 4) == 0
        assert lib.nb200_all(C.byref(flag), pa, M * K) == 0 and flag.value == 0
        assert lib.nb200_copy_d2d(pe, pa, M * K * 4) == 0
        assert lib.nb200_allclose(C.byref(flag), pa, pe, M * K, 1e-5, 1e-8) == 0 and flag.value == 1
        one = np.ones(1, np.float32)
        assert lib.nb200_copy_h2d(C.c_void_p(pe.value + 4 * (M * K - 2)), one.ctypes.data, 4) == 0
        assert lib.nb200_allclose(C.byref(flag), pa, pe, M * K, 1e-5, 1e-8) == 0 and flag.value == 0
    finally:
        for p in (pa, pb, pc, px, py, pe):
            lib.nb200_free(p)


def test_outer_and_l1_norm_compositions(nb):
    """nd.outer (NDArray_Outer, linalg.c:724-751) and nd.norm(a, 1) (NDArray_L1Norm, linalg.c:423-447) as compositions of path
    kernels, against the reference's own functions; dyadic inputs make every product and sum exact."""
    r = _rng(77)
    a = (r.integers(-64, 65, size=1000).astype(np.float32) / 64)
    b = (r.integers(-64, 65, size=777).astype(np.float32) / 64)
    got = nb.nd.outer(nb.NDArray.array(a).gpu(), nb.NDArray.array(b).gpu()).toArray()
    assert got.shape == (1000, 777)
    exp = ORACLE.outer(a, b) if hasattr(ORACLE, "outer") else np.outer(a, b).astype(np.float32)
    assert (got == exp).all()
    # the reference's NDArray_L1Norm is only safe on square inputs (it sizes and scans its per-column results by the row count,
    # linalg.c:427, 437-441): square against the reference, rectangular against the definition
    sq = (r.integers(-64, 65, size=(129, 129)).astype(np.float32) / 64)
    exp_n = ORACLE.norm1(sq) if hasattr(ORACLE, "norm1") else np.abs(sq).sum(axis=0).max()
    assert nb.nd.norm(nb.NDArray.array(sq).gpu(), 1) == float(exp_n) == float(np.abs(sq).sum(axis=0).max())
    m = (r.integers(-64, 65, size=(300, 41)).astype(np.float32) / 64)
    assert nb.nd.norm(nb.NDArray.array(m).gpu(), 1) == float(np.abs(m).sum(axis=0).max())
    with pytest.raises(ValueError):
        nb.nd.outer(nb.NDArray.array(m).gpu(), nb.NDArray.array(b).gpu())


def test_matmul_inf_nan_propagation_and_dynamic_range(nb):
    """IEEE special values must propagate like cblas_sgemm (inf stays inf, inf*0 and NaN give NaN), and rows with a wide
    dynamic range keep fp32-class accuracy (TF32 has the full 8-bit exponent)."""
    r = _rng(41)
    a = r.random((256, 128), dtype=np.float32)
    b = r.random((128, 256), dtype=np.float32)
    a[3, 7] = np.inf
    a[100, 5] = -np.inf
    b[9, 200] = np.nan
    b[7, 50] = 0.0          # inf * 0 -> NaN in row 3, column 50
    exp = ORACLE.matmul(a, b)
    fin = np.isfinite(exp)
    for prec, tol in ((nb.TF32X3, 2e-3), (nb.BF16X3, 2e-2)):   # a flagged call runs at single-pass accuracy (documented)
        got = nb.nd.matmul(nb.NDArray.array(a).gpu(), nb.NDArray.array(b).gpu(), prec).toArray()
        np.testing.assert_array_equal(np.isnan(got), np.isnan(exp))
        np.testing.assert_array_equal(np.isposinf(got), np.isposinf(exp))
        np.testing.assert_array_equal(np.isneginf(got), np.isneginf(exp))
        assert rel_err(got[fin], exp[fin]).max() <= tol
    # wide dynamic range, finite: every row scaled by 2^k, k in [-60, 60]
    a2 = (r.random((256, 160), dtype=np.float32) * np.exp2(r.integers(-60, 61, size=(256, 1))).astype(np.float32)).astype(np.float32)
    b2 = (r.random((160, 128), dtype=np.float32) * np.exp2(r.integers(-60, 61, size=(1, 128))).astype(np.float32)).astype(np.float32)
    exp2 = ORACLE.matmul(a2, b2)
    ok = np.isfinite(exp2) & (np.abs(exp2) > 1e-30)
    for prec in (nb.TF32X3, nb.BF16X3):
        got2 = nb.nd.matmul(nb.NDArray.array(a2).gpu(), nb.NDArray.array(b2).gpu(), prec).toArray()
        assert rel_err(got2[ok], exp2[ok]).max() <= RTOL
    # values next to FLT_MAX round up to inf in bf16: the split must truncate instead (finite result expected)
    a3 = np.full((128, 128), 3.4e38, np.float32)
    b3 = np.full((128, 128), 2.0 ** -10, np.float32)
    got3 = nb.nd.matmul(nb.NDArray.array(a3).gpu(), nb.NDArray.array(b3).gpu(), nb.BF16X3).toArray()
    assert np.isfinite(got3).all() and rel_err(got3, ORACLE.matmul(a3, b3)).max() <= RTOL


# --------------------------------------------------------------------- nd::all / nd::allclose / transpose (SURVEY §8 f, N2)
@pytest.mark.parametrize("n", [1, 7, 8, 1000, 4099, (1 << 22) + 3])
def test_all_and_allclose_vs_oracle_port(nb, n):
    """Boolean reductions (nb200_all / nb200_allclose) against oracle/port.c, which restates the intended semantics of
    NDArray_All / float_allclose (logic.c:25-58, 718-738) and is pinned by the reference's two logic phpt tests."""
    r = _rng(n)
    x = (r.random(n, dtype=np.float32) + 0.5).astype(np.float32)
    X = nb.NDArray.array(x).gpu()
    assert nb.nd.all(X) == oracle.port.all(x) == 1
    for pos in {0, n // 2, n - 1}:
        y = x.copy()
        y[pos] = 0.0
        assert nb.nd.all(nb.NDArray.array(y).gpu()) == oracle.port.all(y) == 0
        y[pos] = -0.0
        assert nb.nd.all(nb.NDArray.array(y).gpu()) == oracle.port.all(y) == 0
        y[pos] = np.nan
        assert nb.nd.all(nb.NDArray.array(y).gpu()) == oracle.port.all(y) == 1
        z = x.copy()
        z[pos] *= np.float32(1.001)
        Z = nb.NDArray.array(z).gpu()
        for rtol, atol in ((1e-5, 1e-8), (1e-2, 0.0), (0.0, 1e-2), (0.0, 0.0)):
            assert nb.nd.allclose(X, Z, rtol, atol) == bool(oracle.port.allclose(x, z, rtol, atol)), (pos, rtol, atol)
            assert nb.nd.allclose(Z, X, rtol, atol) == bool(oracle.port.allclose(z, x, rtol, atol))
        z[pos] = np.nan
        assert nb.nd.allclose(X, nb.NDArray.array(z).gpu()) == bool(oracle.port.allclose(x, z))
    assert nb.nd.allclose(X, X) is True
    if n > 4:   # 4-byte aligned views ($a[i] rows): the scalar path
        v = nb.NDArray.array(np.stack([x, x])).gpu()
        assert nb.nd.all(v[1]) == 1


def test_all_allclose_golden_and_errors(nb):
    import json
    import os
    for rec in json.load(open(os.path.join(os.path.dirname(__file__), "golden", "logic_vectors.json")))["vectors"]:
        args = [nb.NDArray.array(np.asarray(a, np.float32)).gpu() for a in rec["args"]]
        assert int(getattr(nb.nd, rec["op"])(*args)) == rec["expect"], rec
    with pytest.raises(RuntimeError, match="Shape mismatch"):
        nb.nd.allclose(np.zeros((2, 3), np.float32), np.zeros((3, 2), np.float32))
    assert nb.nd.all(np.zeros((0,), np.float32)) == 1


@pytest.mark.parametrize("rc", [(1, 1), (3, 5), (32, 32), (33, 65), (257, 1031), (1000, 1), (1, 777), (2048, 4100), (70001, 3), (3, 2100001)])
def test_transpose2d_and_the_legacy_in_place_call(nb, rc):
    """nb200_transpose2d and cuda_float_transpose (cuda_math.h:77) as the host calls it: d_in == d_out (manipulation.c:124),
    (width, height) = (cols, rows).  Bit-exact (index work)."""
    import ctypes as C
    lib = nb.lib()
    rows, cols = rc
    x = _rng(rows * 31 + cols).random((rows, cols), dtype=np.float32)
    dx, do = _dev(nb, x), _dev(nb, np.zeros((cols, rows), np.float32))
    assert lib.nb200_transpose2d(do, dx, rows, cols) == 0, lib.nb200_last_error()
    np.testing.assert_array_equal(_fetch(nb, do, (cols, rows)), x.T)
    assert lib.nb200_transpose2d(dx, dx, rows, cols) != 0          # out == in is refused by the 64-bit entry point
    f = lib.cuda_float_transpose                                    # legacy symbol of the same library (include/nb200_legacy.h)
    f.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    f.restype = None
    f(0, 0, dx, dx, cols, rows)                                     # in place through the legacy symbol
    np.testing.assert_array_equal(_fetch(nb, dx, (cols, rows)), x.T)
    for p in (dx, do):
        lib.nb200_free(p)


# --------------------------------------------------------------------- residency: `$a->gpu()` / `->cpu()` from pageable memory (N3)
@pytest.mark.parametrize("nbytes", [(16 << 20) - 4, (16 << 20), (16 << 20) + 12, (40 << 20) + 4, (100 << 20) + 4092])
def test_pageable_copies_staged_through_pinned_slots_are_bit_exact(nb, nbytes):
    """nb200_copy_h2d / nb200_copy_d2h from pageable host memory: >= 16 MiB goes through the pinned staging ring with the
    multi-threaded slice copy (abi.cu), smaller copies and pinned pointers take the direct path; chunk tails, slot reuse."""
    import ctypes as C
    lib = nb.lib()
    n = nbytes // 4
    src = _rng(nbytes % 1000).integers(0, 1 << 31, size=n, dtype=np.int64).astype(np.uint32).view(np.float32)   # arbitrary bit patterns
    back = np.zeros(n, np.float32)
    p = C.c_void_p()
    assert lib.nb200_alloc(C.byref(p), n * 4) == 0
    assert lib.nb200_copy_h2d(p, src.ctypes.data, n * 4) == 0, lib.nb200_last_error()
    assert lib.nb200_copy_d2h(back.ctypes.data, p, n * 4) == 0, lib.nb200_last_error()
    np.testing.assert_array_equal(back.view(np.uint32), src.view(np.uint32))
    # through a pinned buffer (direct path) the device holds the same bits
    hp = C.c_void_p()
    assert lib.nb200_host_alloc(C.byref(hp), n * 4) == 0
    assert lib.nb200_copy_d2h(hp, p, n * 4) == 0
    pinned = np.ctypeslib.as_array(C.cast(hp, C.POINTER(C.c_uint32)), shape=(n,))
    np.testing.assert_array_equal(pinned, src.view(np.uint32))
    assert lib.nb200_host_free(hp) == 0 and lib.nb200_free(p) == 0
