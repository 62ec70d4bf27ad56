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
