"""Python face of the host mirror: the PHP-facing ``NDArray`` class and ``nd::*`` static surface
(stubs/numpower.stubs.php in the reference: add :43, sum :401, matmul :983, argmax :1141) bound
with ctypes to include/nb200_host.h.  Used by the tests and bench.py the way a PHP script would
use the extension: build arrays, ``->gpu()``, operate, ``->toArray()``.

No arithmetic happens in Python or torch; every op is one call into libnb200.so.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L

BIN = {"add": 0, "sub": 1, "mul": 2, "div": 3, "mod": 4, "pow": 5, "maximum": 6, "minimum": 7, "arctan2": 8,
       "mod_trunc": 9, "equal": 10, "not_equal": 11, "greater": 12, "greater_equal": 13, "less": 14, "less_equal": 15}
UN = {
    "abs": 0, "sqrt": 1, "exp": 2, "exp2": 3, "expm1": 4, "log": 5, "log2": 6, "log10": 7, "log1p": 8,
    "logb": 9, "sin": 10, "cos": 11, "tan": 12, "arcsin": 13, "arccos": 14, "arctan": 15, "sinh": 16,
    "cosh": 17, "tanh": 18, "arcsinh": 19, "arccosh": 20, "arctanh": 21, "degrees": 22, "radians": 23,
    "rint": 24, "fix": 25, "trunc": 26, "floor": 27, "ceil": 28, "sinc": 29, "negative": 30,
    "positive": 31, "sign": 32, "reciprocal": 33, "rsqrt": 34, "clip": 35, "round": 36, "square": 37,
}
RED = {"sum": 0, "prod": 1, "min": 2, "max": 3}
TF32X3, TF32X1, BF16X3, GEMM_AUTO, FP16X3, FP16X3U = 0, 1, 2, 3, 4, 5   # include/nb200.h nb200_gemm_precision
ORDER_TREE, ORDER_SEQUENTIAL = 0, 1


def _shape_arr(shape):
    return (C.c_int64 * max(len(shape), 1))(*shape)


class NDArray:
    """Handle on an NB_NDArray (host mirror of struct NDArray, src/ndarray.h:61-74)."""

    __slots__ = ("_h", "_keep")

    def __init__(self, handle, keep=None):
        self._h = handle
        self._keep = keep  # parent kept alive for views

    # ---- construction / residency
    @staticmethod
    def array(values) -> "NDArray":
        """NDArray::array(): build a CPU NDArray from nested lists / numpy (float32)."""
        a = np.ascontiguousarray(np.asarray(values, dtype=np.float32)) if np.ndim(values) else np.asarray(values, np.float32).copy()
        h = L.check_ptr(L.lib().NB_NDArray_FromHost(a.ctypes.data, a.ndim, _shape_arr(a.shape)))
        return NDArray(h)

    def gpu(self) -> "NDArray":
        return NDArray(L.check_ptr(L.lib().NB_NDArray_ToGPU(self._h)))

    def cpu(self) -> "NDArray":
        return NDArray(L.check_ptr(L.lib().NB_NDArray_ToCPU(self._h)))

    def isGPU(self) -> bool:
        return self._h.contents.device == 1

    @property
    def shape(self):
        c = self._h.contents
        return tuple(c.shape[i] for i in range(c.ndim))

    @property
    def ndim(self):
        return self._h.contents.ndim

    @property
    def size(self):
        return self._h.contents.numel

    @property
    def data_ptr(self) -> int:
        return self._h.contents.data or 0

    def toArray(self) -> np.ndarray:
        out = np.empty(self.shape, dtype=np.float32)
        if L.lib().NB_NDArray_CopyToHost(self._h, out.ctypes.data) != 0:
            raise L.BackendError(-1, L.lib().NB_last_error().decode())
        return out

    def __getitem__(self, i: int) -> "NDArray":
        return NDArray(L.check_ptr(L.lib().NB_NDArray_Slice0(self._h, int(i))), keep=self)

    def reshape(self, *shape) -> "NDArray":
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        return NDArray(L.check_ptr(L.lib().NB_NDArray_Reshape(self._h, len(shape), _shape_arr(shape))), keep=self)

    def __del__(self):
        try:
            if self._h:
                L.lib().NB_NDArray_FREE(self._h)
                self._h = None
        except Exception:
            pass

    # ---- operator overloading (ndarray_do_operation_ex, numpower.c:193-229)
    def _coerce(self, other) -> "NDArray":
        if isinstance(other, NDArray):
            return other
        o = NDArray.array(other)            # ZVAL_TO_NDARRAY numpower.c:89-117: scalars -> 0-dim CPU, arrays -> CPU NDArray
        if o.ndim > 0 and self.isGPU():
            o = o.gpu()
        return o

    def _bin(self, name, other, swap=False):
        o = self._coerce(other)
        a, b = (o, self) if swap else (self, o)
        return NDArray(L.check_ptr(L.lib().NB_NDArray_Binary(BIN[name], a._h, b._h)))

    def __add__(self, o): return self._bin("add", o)
    def __radd__(self, o): return self._bin("add", o, True)
    def __sub__(self, o): return self._bin("sub", o)
    def __rsub__(self, o): return self._bin("sub", o, True)
    def __mul__(self, o): return self._bin("mul", o)
    def __rmul__(self, o): return self._bin("mul", o, True)
    def __truediv__(self, o): return self._bin("div", o)
    def __rtruediv__(self, o): return self._bin("div", o, True)
    def __mod__(self, o): return self._bin("mod", o)
    def __pow__(self, o): return self._bin("pow", o)
    def __matmul__(self, o): return nd.matmul(self, o)


class nd:
    """The static ``nd::`` / ``NDArray::`` method surface for the hot path."""

    array = staticmethod(NDArray.array)

    @staticmethod
    def _a(x) -> NDArray:
        return x if isinstance(x, NDArray) else NDArray.array(x).gpu()

    @staticmethod
    def binary(op: str, a, b) -> NDArray:
        a = nd._a(a)
        return a._bin(op, b)

    @staticmethod
    def add(a, b): return nd.binary("add", a, b)
    @staticmethod
    def subtract(a, b): return nd.binary("sub", a, b)
    @staticmethod
    def multiply(a, b): return nd.binary("mul", a, b)
    @staticmethod
    def divide(a, b): return nd.binary("div", a, b)
    @staticmethod
    def mod(a, b): return nd.binary("mod", a, b)
    @staticmethod
    def pow(a, b): return nd.binary("pow", a, b)
    @staticmethod
    def maximum(a, b): return nd.binary("maximum", a, b)
    @staticmethod
    def minimum(a, b): return nd.binary("minimum", a, b)
    @staticmethod
    def arctan2(a, b): return nd.binary("arctan2", a, b)

    # comparisons -> 0/1 float masks (src/logic.c:68-660; SURVEY §8 f, N2)
    @staticmethod
    def equal(a, b): return nd.binary("equal", a, b)
    @staticmethod
    def not_equal(a, b): return nd.binary("not_equal", a, b)
    @staticmethod
    def greater(a, b): return nd.binary("greater", a, b)
    @staticmethod
    def greater_equal(a, b): return nd.binary("greater_equal", a, b)
    @staticmethod
    def less(a, b): return nd.binary("less", a, b)
    @staticmethod
    def less_equal(a, b): return nd.binary("less_equal", a, b)

    @staticmethod
    def array_equal(a, b) -> bool:
        """NDArray_ArrayEqual (logic.c:703-716): same shape and every element pair equal."""
        a, b = nd._a(a), nd._a(b)
        if a.shape != b.shape:
            return False
        return a.size == 0 or nd.min(nd.equal(a, b)) == 1.0

    @staticmethod
    def all(a) -> int:
        """NDArray::all (NDArray_All, logic.c:25-58): 1 iff no element is zero (one boolean-reduction kernel, nb200_all)."""
        a = nd._a(a)
        out = C.c_int(0)
        L.check(L.lib().nb200_all(C.byref(out), a.data_ptr, a.size))
        return int(out.value)

    @staticmethod
    def allclose(a, b, rtol: float = 1e-5, atol: float = 1e-8) -> bool:
        """NDArray::allclose (NDArray_AllClose, logic.c:748-771; the reference refuses device arrays): |a - b| <= atol + rtol * |b|
        everywhere.  Same error behaviour: "Shape mismatch"."""
        a, b = nd._a(a), nd._a(b)
        if a.shape != b.shape:
            raise RuntimeError("Shape mismatch")
        out = C.c_int(0)
        L.check(L.lib().nb200_allclose(C.byref(out), a.data_ptr, b.data_ptr, a.size, float(rtol), float(atol)))
        return bool(out.value)

    @staticmethod
    def mul_add(a, b, c) -> NDArray:
        """Fused ``$a * $b + $c`` (nb200_ew_mul_add): one pass, same bits as the two nd:: calls."""
        a, b, c = nd._a(a), nd._a(b), nd._a(c)
        return NDArray(L.check_ptr(L.lib().NB_NDArray_MulAdd(a._h, b._h, c._h)))

    @staticmethod
    def unary(op: str, a, p0: float = 0.0, p1: float = 0.0) -> NDArray:
        a = nd._a(a)
        return NDArray(L.check_ptr(L.lib().NB_NDArray_Map(a._h, UN[op], p0, p1)))

    @staticmethod
    def clip(a, min: float, max: float): return nd.unary("clip", a, float(min), float(max))
    @staticmethod
    def round(a, precision: float = 0): return nd.unary("round", a, float(precision))

    @staticmethod
    def _full(fn, a) -> float:
        a = nd._a(a)
        out = C.c_float()
        if fn(a._h, C.byref(out)) != 0:
            raise L.BackendError(-1, L.lib().NB_last_error().decode())
        return float(out.value)

    @staticmethod
    def reduce(op: str, a, axis=None, order: int = ORDER_TREE):
        a = nd._a(a)
        if axis is None:
            fn = {"sum": L.lib().NB_NDArray_Sum_Float, "prod": L.lib().NB_NDArray_Float_Prod,
                  "min": L.lib().NB_NDArray_Min, "max": L.lib().NB_NDArray_Max}[op]
            return nd._full(fn, a)
        return NDArray(L.check_ptr(L.lib().NB_reduce(a._h, int(axis), RED[op], order)))

    @staticmethod
    def sum(a, axis=None, order: int = ORDER_TREE): return nd.reduce("sum", a, axis, order)
    @staticmethod
    def prod(a, axis=None, order: int = ORDER_TREE): return nd.reduce("prod", a, axis, order)
    @staticmethod
    def min(a, axis=None): return nd.reduce("min", a, axis)
    @staticmethod
    def max(a, axis=None): return nd.reduce("max", a, axis)

    @staticmethod
    def mean(a, axis=None):
        """nd::mean (numpower.c:2642-2688): no axis -> (float) sum / numElements; axis -> reduce(sum) then a
        float division by the axis length (NDArray_Divide_Float by a 0-dim scalar)."""
        a = nd._a(a)
        if axis is None:
            return float(np.float32(nd.sum(a)) / np.float32(a.size))
        return nd.sum(a, axis) / float(a.shape[int(axis)])

    # ---- statistics composed from the hot-path kernels (src/ndmath/statistics.c:87-153; SURVEY §8 f, N4).  The
    # reference's GPU branches throw ("NDArray::std not available for GPU."); the op sequence is the CPU one.
    @staticmethod
    def variance(a) -> float:
        """NDArray_Variance (statistics.c:110-124): mean(|a - mean|^2), the square taken with pow(.., 2)."""
        a = nd._a(a)
        m = np.float32(nd.sum(a)) / np.float32(a.size)
        p = nd.pow(nd.unary("abs", a - float(m)), 2.0)
        return float(np.float32(nd.sum(p)) / np.float32(p.size))

    @staticmethod
    def std(a) -> float:
        """NDArray_Std (statistics.c:86-101): sqrt(sum((a - mean)^2) / n)."""
        a = nd._a(a)
        m = np.float32(nd.sum(a)) / np.float32(a.size)
        d = a - float(m)
        return float(np.sqrt(np.float32(nd.sum(nd.pow(d, 2.0))) / np.float32(a.size), dtype=np.float32))

    @staticmethod
    def average(a, weights=None) -> float:
        """NDArray_Average (statistics.c:133-153): sum(a*w) / sum(w), or the mean without weights."""
        a = nd._a(a)
        if weights is None:
            return float(np.float32(nd.sum(a)) / np.float32(a.size))
        w = nd._a(weights)
        return float(np.float32(nd.sum(a * w)) / np.float32(nd.sum(w)))

    @staticmethod
    def _arg(a, axis, keepdims, is_max):
        a = nd._a(a)
        r = NDArray(L.check_ptr(L.lib().NB_NDArray_ArgMinMaxCommon(a._h, 128 if axis is None else int(axis), int(keepdims), int(is_max))))
        return float(r.toArray()) if r.ndim == 0 else r   # RETURN_NDARRAY: 0-dim -> PHP float (numpower.c:137-150)

    @staticmethod
    def argmax(a, axis=None, keepdims=False): return nd._arg(a, axis, keepdims, True)
    @staticmethod
    def argmin(a, axis=None, keepdims=False): return nd._arg(a, axis, keepdims, False)

    @staticmethod
    def matmul(a, b, precision: int = GEMM_AUTO) -> NDArray:
        a, b = nd._a(a), nd._a(b)
        return NDArray(L.check_ptr(L.lib().NB_NDArray_Matmul(a._h, b._h, precision)))

    @staticmethod
    def dot(a, b) -> NDArray:
        a, b = nd._a(a), nd._a(b)
        return NDArray(L.check_ptr(L.lib().NB_NDArray_Dot(a._h, b._h)))

    @staticmethod
    def outer(a, b) -> NDArray:
        """NDArray_Outer (linalg.c:724-751: two 1-D vectors): the broadcast multiply (m, 1) * (1, n), one launch."""
        a, b = nd._a(a), nd._a(b)
        if a.ndim != 1 or b.ndim != 1:
            raise ValueError("Invalid operation: NDArray::outer() requires both arrays to be 1-dimensional vectors.")
        return nd.multiply(a.reshape(a.size, 1), b.reshape(1, b.size))

    @staticmethod
    def norm(a, order: int = 1) -> float:
        """NDArray_Norm(a, 1) = NDArray_L1Norm (linalg.c:423-447): max over columns of the sum of absolute values.  The reference
        transposes and sums every column with its own call (and is only safe on square inputs: it sizes its per-column results by the
        row count); here abs + one axis-0 reduction + max, any 2-D shape.  Other orders need the SVD."""
        if order != 1:
            raise NotImplementedError("norm: only order 1 is on the elementwise / reduction path")
        a = nd._a(a)
        return float(nd.max(nd.sum(nd.unary("abs", a), 0)))


def _make_unary(name):
    def f(a):
        return nd.unary(name, a)
    f.__name__ = name
    return staticmethod(f)


for _n in UN:
    if _n not in ("clip", "round"):
        setattr(nd, _n, _make_unary(_n))


class GoldenBackend:
    """Adapter with the same method set as oracle.ref / oracle.port so the golden-vector replay
    (tests/helpers.run_golden_record) reads identically for the oracle and the GPU path."""

    @staticmethod
    def binary(op, a, b):
        a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
        A = NDArray.array(a).gpu() if a.ndim else NDArray.array(a)
        B = NDArray.array(b).gpu() if b.ndim else NDArray.array(b)
        return NDArray(L.check_ptr(L.lib().NB_NDArray_Binary(BIN[op], A._h, B._h))).toArray()

    @staticmethod
    def unary(op, x, p0=0.0, p1=0.0):
        return nd.unary(op, np.asarray(x, np.float32), p0, p1).toArray()

    @staticmethod
    def reduce_full(op, x):
        return np.float32(nd.reduce(op, np.asarray(x, np.float32)))

    @staticmethod
    def reduce_axis(op, x, axis):
        return nd.reduce(op, np.asarray(x, np.float32), axis).toArray()

    @staticmethod
    def matmul(a, b):
        return nd.matmul(np.asarray(a, np.float32), np.asarray(b, np.float32)).toArray()
