"""Drop-in proof (GPU): NumPower's UNMODIFIED host C code, compiled with HAVE_CUBLAS and linked against
libnb200.so instead of its own cuda_math.o/gpu_alloc.o (oracle/build_dropin.sh), drives the B200 kernels
through the legacy symbols (include/nb200_legacy.h) exactly as the PHP extension would:
NDArray_ToGPU -> NDArray_Add_Float(gpu, gpu) -> cuda_add_float(...) -> NDArray_ToCPU.
Results are compared with the same reference code's CPU branch (oracle.ref)."""
import numpy as np
import pytest

import oracle
from helpers import rel_err

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not (oracle.dropin.available and oracle.ref.available), reason="oracle/_ref drop-in build absent")]


def _rng(s):
    return np.random.default_rng(s)


def eq(g, e):
    ok = (g == e) | (np.isnan(g) & np.isnan(e))
    assert ok.all()


@pytest.mark.parametrize("op", ["add", "sub", "mul", "div", "pow"])
def test_reference_host_binary_ops_on_b200(op):
    r = _rng(1)
    a = (r.random((64, 48), dtype=np.float32) + 0.5).astype(np.float32)
    for b in (np.float32(2.0), (r.random((64, 48), dtype=np.float32) + 0.5), (r.random(48, dtype=np.float32) + 0.5),
              (r.random((64, 1), dtype=np.float32) + 0.5)):
        g, e = oracle.dropin.binary(op, a, b), oracle.ref.binary(op, a, b)
        if op == "pow":
            assert rel_err(g, e).max() <= 1e-5
        else:
            eq(g, e)


def test_reference_host_mod_and_chain_on_b200():
    r = _rng(2)
    a = (r.random((32, 64), dtype=np.float32) * 9 + 1).astype(np.float32)
    b = (r.random((32, 64), dtype=np.float32) * 3 + 0.5).astype(np.float32)
    eq(oracle.dropin.binary("mod", a, b), oracle.ref.binary("mod", a, b))
    c = r.random((32, 64), dtype=np.float32)
    eq(oracle.dropin.mul_add(a, b, c), oracle.ref.mul_add(a, b, c))


@pytest.mark.parametrize("op", ["abs", "sqrt", "exp", "log", "sin", "cos", "tanh", "floor", "ceil", "sign", "sinc",
                                "degrees", "reciprocal", "clip", "round", "square"])
def test_reference_host_unary_ops_on_b200(op):
    x = (_rng(3).random(5000, dtype=np.float32) * 8 + 0.1).astype(np.float32)
    p0, p1 = (1.0, 5.0) if op == "clip" else ((2.0, 0.0) if op == "round" else (0.0, 0.0))
    g, e = oracle.dropin.unary(op, x, p0, p1), oracle.ref.unary(op, x, p0, p1)
    assert rel_err(g, e).max() <= 1e-5


def test_reference_host_reductions_on_b200():
    x = (_rng(4).integers(-64, 65, size=(40, 33)).astype(np.float32) / 64)
    for op in ("sum", "min", "max"):
        assert oracle.dropin.reduce_full(op, x) == oracle.ref.reduce_full(op, x)
    # reduce(): the reference's slice loop issues one NDArray_Add_Float per slice -> O(len) kernel launches
    for axis in (0, 1):
        np.testing.assert_array_equal(oracle.dropin.reduce_axis("sum", x, axis), oracle.ref.reduce_axis("sum", x, axis))
    y = _rng(5).choice(np.array([1, 1, -1, 2, 0.5], np.float32), size=300)
    assert oracle.dropin.reduce_full("prod", y) == oracle.ref.reduce_full("prod", y)


def test_reference_host_matmul_reaches_tcgen05_through_cublas_shim():
    """linalg.c:54-72 calls cublasCreate/cublasSgemm/cublasDestroy; include/nb200_cublas_shim maps them
    onto nb200_sgemm (TF32x3), so the reference host's GPU matmul runs on the hand-written kernel."""
    r = _rng(6)
    a, b = r.random((256, 384), dtype=np.float32), r.random((384, 512), dtype=np.float32)
    assert rel_err(oracle.dropin.matmul(a, b), oracle.ref.matmul(a, b)).max() <= 1e-5
    a2, b2 = np.array([[1, 2], [3, 4]], np.float32), np.array([[5, 6], [7, 8]], np.float32)
    np.testing.assert_array_equal(oracle.dropin.matmul(a2, b2), [[19, 22], [43, 50]])   # tests/linalg/001-ndarray-matmul.phpt


def test_reference_gpu_argmax_still_unsupported_in_host():
    """calculation.c:75-78 throws for GPU arrays; reaching nb200_argminmax needs host patch N1."""
    with pytest.raises(RuntimeError, match="GPU not supported"):
        oracle.dropin.argminmax(True, np.arange(10, dtype=np.float32))


def test_reference_host_outer_and_l1_norm_on_b200():
    """NDArray_Outer (linalg.c:724-751) -> cuda_calculate_outer_product = one broadcast multiply; NDArray_Norm(a, 1)
    (linalg.c:423-447) = Transpose + Abs + one Sum per column + max, every step through the legacy symbols.  Dyadic inputs:
    products and sums are exact, so the GPU branch must give the CPU branch's bits."""
    r = _rng(21)
    a = (r.integers(-64, 65, size=300).astype(np.float32) / 64)
    b = (r.integers(-64, 65, size=517).astype(np.float32) / 64)
    g, e = oracle.dropin.outer(a, b), oracle.ref.outer(a, b)
    assert g.shape == (300, 517)
    eq(g, e)
    eq(g, np.outer(a, b).astype(np.float32))
    # square only: NDArray_L1Norm sizes its per-column results by shape[ndim-2] and scans that many entries (linalg.c:427, 437-441),
    # so a non-square input reads uninitialised entries (rows > cols) or writes past the buffer (rows < cols) in the reference
    m = (r.integers(-64, 65, size=(33, 33)).astype(np.float32) / 64)
    assert oracle.dropin.norm1(m) == oracle.ref.norm1(m) == np.abs(m).sum(axis=0).max()
