"""Level-1 integration (GPU): the reference's host C code with the one-line glue calls of oracle/n1_patch.py
(integration/nb200_numpower_glue.c) — what unchanged PHP would reach after the host patches of INTEGRATION.md.
Checks results against the reference's CPU branch AND that the structural wins are real by counting the kernels
libnb200 launches (nb200_launch_count): one launch where the unpatched host needs O(len) launches / copies."""
import numpy as np
import pytest

import oracle
from helpers import rel_err

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not (oracle.dropin_n1.available and oracle.ref.available), reason="oracle/_ref N1 build absent")]


def _rng(s):
    return np.random.default_rng(s)


def _launches():
    import numpower_b200 as nb
    return nb.lib().nb200_launch_count()


def _count(fn):
    n0 = _launches()
    r = fn()
    return r, _launches() - n0


def eq(g, e):
    assert ((g == e) | (np.isnan(g) & np.isnan(e))).all()


def test_broadcast_and_scalar_operands_are_not_materialised():
    r = _rng(1)
    a = (r.random((96, 80), dtype=np.float32) + 0.5).astype(np.float32)
    row, col = (r.random(80, dtype=np.float32) + 0.5), (r.random((96, 1), dtype=np.float32) + 0.5)
    for op in ("add", "sub", "mul", "div"):
        for b in (row, col, np.float32(1.5)):
            got, n = _count(lambda: oracle.dropin_n1.binary(op, a, b))
            eq(got, oracle.ref.binary(op, a, b))
            assert n == 1, (op, np.shape(b), n)          # one kernel: no NDArray_Fill / NDArray_Broadcast temporaries
    # Level 0 (unpatched host) for comparison: the scalar is filled into a temporary first (cuda_fill_float + op)
    if oracle.dropin.available:
        _, n0 = _count(lambda: oracle.dropin.binary("add", a, np.float32(1.5)))
        assert n0 >= 2


def test_reduce_is_one_launch_instead_of_one_per_slice():
    x = (_rng(2).integers(-64, 65, size=(64, 48)).astype(np.float32) / 64)
    for axis in (0, 1):
        got, n = _count(lambda: oracle.dropin_n1.reduce_axis("sum", x, axis))
        np.testing.assert_array_equal(got, oracle.ref.reduce_axis("sum", x, axis))
        assert n <= 2, n
        if oracle.dropin.available:
            _, n0 = _count(lambda: oracle.dropin.reduce_axis("sum", x, axis))
            assert n0 >= x.shape[axis] - 1                 # the reference's slice loop: one NDArray_Add_Float per slice
    y = _rng(3).choice(np.array([1, 1, -1, 2, 0.5], np.float32), size=(16, 12))
    np.testing.assert_array_equal(oracle.dropin_n1.reduce_axis("prod", y, 0), oracle.ref.reduce_axis("prod", y, 0))


def test_argmax_argmin_now_run_on_the_gpu():
    x = _rng(4).integers(0, 40, size=(9, 70, 5)).astype(np.float32)
    for is_max in (True, False):
        np.testing.assert_array_equal(oracle.dropin_n1.argminmax(is_max, x), oracle.ref.argminmax(is_max, x))
        for axis in (0, 1, 2):
            for kd in (False, True):
                np.testing.assert_array_equal(oracle.dropin_n1.argminmax(is_max, x, axis, kd), oracle.ref.argminmax(is_max, x, axis, kd))


def test_matmul_2d_and_stacks():
    r = _rng(5)
    a, b = r.random((256, 160), dtype=np.float32), r.random((160, 192), dtype=np.float32)
    assert rel_err(oracle.dropin_n1.matmul(a, b), oracle.ref.matmul(a, b)).max() <= 1e-5
    a3, b3 = r.random((6, 128, 96), dtype=np.float32), r.random((6, 96, 160), dtype=np.float32)
    got, n = _count(lambda: oracle.dropin_n1.matmul_nd(a3, b3))
    for i in range(6):
        assert rel_err(got[i], oracle.ref.matmul(a3[i], b3[i])).max() <= 1e-5
    assert n <= 2          # lo-split + one batched GEMM launch for the whole stack
    with pytest.raises(RuntimeError, match="Stack of matrices not allowed"):
        oracle.ref.matmul_nd(a3, b3)     # the unpatched reference rejects stacks (linalg.c:240-243)


def test_maximum_minimum_max_axis_now_run_on_the_gpu():
    """ndarray.c:782-784, 853, 896 throw for device arrays in the reference; with the N1 lines they are one kernel each."""
    r = _rng(6)
    a = (r.random((48, 40), dtype=np.float32) * 4 - 2).astype(np.float32)
    for b in ((r.random((48, 40), dtype=np.float32) * 4 - 2).astype(np.float32), (r.random(40, dtype=np.float32) * 4 - 2).astype(np.float32)):
        for op in ("maximum", "minimum"):
            got, n = _count(lambda: oracle.dropin_n1.binary(op, a, b))
            eq(got, oracle.ref.binary(op, a, b))
            assert n == 1
    x = r.integers(-50, 50, size=(37, 29)).astype(np.float32)
    for axis in (0, 1):
        got, n = _count(lambda: oracle.dropin_n1.reduce_axis("max", x, axis))
        np.testing.assert_array_equal(got, oracle.ref.reduce_axis("max", x, axis))
        assert n <= 2
    if oracle.dropin.available:
        with pytest.raises(RuntimeError, match="not implemented"):
            oracle.dropin.binary("maximum", a, a)          # the unpatched host


def test_dot_nd_by_1d_covers_every_leading_row():
    """linalg.c:373-381: the reference's GPU branch hands only shape[ndim-2] rows to the gemv; the N1 line serves all of them."""
    r = _rng(7)
    a2, v = r.random((70, 96), dtype=np.float32), r.random(96, dtype=np.float32)
    assert rel_err(oracle.dropin_n1.dot(a2, v), oracle.ref.dot(a2, v)).max() <= 1e-5
    a3 = r.random((5, 30, 96), dtype=np.float32)
    got = oracle.dropin_n1.dot(a3, v)
    assert got.shape == (5, 30)
    exp = (a3.astype(np.float64) @ v.astype(np.float64))
    assert rel_err(got, exp.astype(np.float32)).max() <= 1e-5


def test_unary_methods_run_out_of_place_in_one_kernel():
    """numpower.c:1648-3348 calls NDArrayMathGPU_ElementWise(nda, cuda_float_<op>): the reference copies the operand and runs the op in
    place on the copy (two passes over the data); libnb200's driver recognises the op pointer and writes a fresh array directly."""
    x = (_rng(8).random(5000, dtype=np.float32) * 8 + 0.1).astype(np.float32)
    for op in ("abs", "sqrt", "exp", "log", "sin", "tanh", "floor", "sign", "reciprocal"):
        got, n = _count(lambda: oracle.dropin_n1.unary(op, x))
        assert rel_err(got, oracle.ref.unary(op, x)).max() <= 1e-5
        assert n == 1, (op, n)
    got, n = _count(lambda: oracle.dropin_n1.unary("clip", x, 1.0, 5.0))
    eq(got, oracle.ref.unary("clip", x, 1.0, 5.0))
    assert n == 1


def test_gpu_and_cpu_moves_go_through_the_staged_copies():
    """ndarray.c:1037-1093: NDArray_ToGPU / NDArray_ToCPU (every test of this file moves its operands through them).  A 24 MiB
    operand is above the staging threshold: pageable -> pinned slots -> device and back, bit-exact; the sum is checked too."""
    r = _rng(9)
    a = (r.integers(-64, 65, size=(3, 1 << 21)).astype(np.float32) / 64)      # 24 MiB, exactly summable
    b = (r.integers(-64, 65, size=(3, 1 << 21)).astype(np.float32) / 64)
    got = oracle.dropin_n1.binary("add", a, b)
    np.testing.assert_array_equal(got, oracle.ref.binary("add", a, b))
    assert oracle.dropin_n1.reduce_full("sum", a) == oracle.ref.reduce_full("sum", a)
