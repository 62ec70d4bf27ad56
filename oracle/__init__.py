"""TEST INFRASTRUCTURE — NOT PRODUCT CODE.

ctypes front-ends for the two CPU oracles of the NDArray hot path:

* ``ref``  — the reference's OWN object code (``oracle/_ref/libnumpower_ref.so``,
  compiled by ``oracle/build_ref.sh`` from /root/reference; see ref_entry.c).
* ``port`` — the plain-C restatement in ``oracle/port.c``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  ``numpower_b200`` never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libnumpower_ref.so")
PORT_SO = os.path.join(HERE, "liboracle_port.so")

# op ids shared with include/nb200.h
BIN_OPS = {"add": 0, "sub": 1, "mul": 2, "div": 3, "mod": 4, "pow": 5, "maximum": 6, "minimum": 7, "arctan2": 8,
           "equal": 10, "not_equal": 11, "greater": 12, "greater_equal": 13, "less": 14, "less_equal": 15}
UN_OPS = {
    "abs": 0, "sqrt": 1, "exp": 2, "exp2": 3, "expm1": 4, "log": 5, "log2": 6, "log10": 7, "log1p": 8,
    "logb": 9, "sin": 10, "cos": 11, "tan": 12, "arcsin": 13, "arccos": 14, "arctan": 15, "sinh": 16,
    "cosh": 17, "tanh": 18, "arcsinh": 19, "arccosh": 20, "arctanh": 21, "degrees": 22, "radians": 23,
    "rint": 24, "fix": 25, "trunc": 26, "floor": 27, "ceil": 28, "sinc": 29, "negative": 30,
    "positive": 31, "sign": 32, "reciprocal": 33, "rsqrt": 34, "clip": 35, "round": 36, "square": 37,
}
RED_OPS = {"sum": 0, "prod": 1, "min": 2, "max": 3}

_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int)
_dp = C.POINTER(C.c_double)


def _f(a: np.ndarray):
    return a.ctypes.data_as(_fp)


def _shape(a: np.ndarray):
    return (C.c_int * max(a.ndim, 1))(*a.shape)


def _c32(a) -> np.ndarray:
    a = np.asarray(a, dtype=np.float32)
    if a.ndim == 0:  # np.ascontiguousarray would promote a scalar to shape (1,)
        return a.copy()
    return np.ascontiguousarray(a)


def build_port(force: bool = False) -> str:
    """Compile oracle/port.c -> oracle/liboracle_port.so (same -march flags as the _ref build)."""
    src = os.path.join(HERE, "port.c")
    if force or not os.path.exists(PORT_SO) or os.path.getmtime(PORT_SO) < os.path.getmtime(src):
        subprocess.check_call(
            ["gcc", "-O2", "-mavx2", "-march=x86-64-v3", "-fopenmp", "-fPIC", "-shared", "-o", PORT_SO, src, "-lm"]
        )
    return PORT_SO


def build_ref() -> str:
    """Run oracle/build_ref.sh (needs /root/reference; on the GPU box the prebuilt .so is used)."""
    subprocess.check_call(["bash", os.path.join(HERE, "build_ref.sh")])
    return REF_SO


class _Ref:
    """The reference's own CPU code path (kind = "reference").

    With ``so_path`` = oracle/_ref/libnumpower_host_b200.so (oracle/build_dropin.sh) the SAME
    reference host objects are compiled with HAVE_CUBLAS and linked against libnb200.so: arrays are
    moved with the reference's NDArray_ToGPU and every op takes the reference's GPU branch — the
    drop-in proof used by tests/test_dropin_gpu.py.
    """

    def __init__(self, so_path: str = None):
        self._lib = None
        self._so = so_path or REF_SO

    @property
    def available(self) -> bool:
        return os.path.exists(self._so)

    @property
    def lib(self):
        if self._lib is None:
            if not os.path.exists(self._so) and self._so == REF_SO:
                build_ref()
            blas_txt = os.path.join(HERE, "_ref", "blas_path.txt")
            if os.path.exists(blas_txt):
                blas = open(blas_txt).read().strip()
                if os.path.exists(blas):
                    C.CDLL(blas, mode=C.RTLD_GLOBAL)
            lib = C.CDLL(self._so)
            lib.ref_last_error.restype = C.c_char_p
            lib.ref_binary.restype = C.c_long
            lib.ref_binary.argtypes = [C.c_int, _fp, _ip, C.c_int, _fp, _ip, C.c_int, _fp, C.c_long, _dp]
            lib.ref_mul_add.restype = C.c_long
            lib.ref_mul_add.argtypes = [_fp, _ip, C.c_int, _fp, _ip, C.c_int, _fp, _ip, C.c_int, _fp, C.c_long, _dp]
            lib.ref_unary.restype = C.c_long
            lib.ref_unary.argtypes = [C.c_int, _fp, C.c_long, _fp, C.c_float, C.c_float, _dp]
            lib.ref_reduce_full.restype = C.c_float
            lib.ref_reduce_full.argtypes = [C.c_int, _fp, C.c_long, _dp]
            lib.ref_reduce_axis.restype = C.c_long
            lib.ref_reduce_axis.argtypes = [C.c_int, _fp, _ip, C.c_int, C.c_int, _fp, C.c_long, _dp]
            lib.ref_argminmax.restype = C.c_long
            lib.ref_argminmax.argtypes = [C.c_int, _fp, _ip, C.c_int, C.c_int, C.c_int, _fp, C.c_long, _dp]
            lib.ref_matmul.restype = C.c_long
            lib.ref_matmul.argtypes = [_fp, _fp, C.c_int, C.c_int, C.c_int, _fp, _dp]
            lib.ref_dot.restype = C.c_long
            lib.ref_dot.argtypes = [_fp, _ip, C.c_int, _fp, _ip, C.c_int, _fp, C.c_long, _dp]
            lib.ref_matmul_nd.restype = C.c_long
            lib.ref_matmul_nd.argtypes = [_fp, _ip, C.c_int, _fp, _ip, C.c_int, _fp, C.c_long, _dp]
            if hasattr(lib, "ref_outer"):
                lib.ref_outer.restype = C.c_long
                lib.ref_outer.argtypes = [_fp, C.c_int, _fp, C.c_int, _fp]
                lib.ref_norm1.restype = C.c_long
                lib.ref_norm1.argtypes = [_fp, _ip, C.c_int, _fp]
            if hasattr(lib, "ref_all"):
                lib.ref_all.argtypes = [_fp, _ip, C.c_int]
                lib.ref_allclose.argtypes = [_fp, _fp, _ip, C.c_int, C.c_float, C.c_float]
            self._lib = lib
        return self._lib

    def _check(self, n, what):
        err = self.lib.ref_last_error().decode()
        self.lib.ref_clear_error()
        if n < 0 or err:
            raise RuntimeError(f"reference {what}: {err or 'returned NULL'}")

    last_seconds = 0.0

    def blas_info(self) -> dict:
        blas_txt = os.path.join(HERE, "_ref", "blas_path.txt")
        info = {}
        try:
            b = C.CDLL(open(blas_txt).read().strip())
            b.scipy_openblas_get_num_threads.restype = C.c_int
            b.scipy_openblas_get_corename.restype = C.c_char_p
            b.scipy_openblas_get_config.restype = C.c_char_p
            info = {
                "threads": b.scipy_openblas_get_num_threads(),
                "core": b.scipy_openblas_get_corename().decode(),
                "config": b.scipy_openblas_get_config().decode(),
            }
        except Exception as e:  # pragma: no cover - informational only
            info = {"error": str(e)}
        return info

    def set_blas_threads(self, n: int) -> int:
        """cblas_sgemm uses OpenBLAS's own threads (all host cores in the reference); launchers such as
        torchrun export OMP_NUM_THREADS=1, so the benchmark sets the count explicitly."""
        try:
            b = C.CDLL(open(os.path.join(HERE, "_ref", "blas_path.txt")).read().strip())
            b.scipy_openblas_set_num_threads(int(n))
            b.scipy_openblas_get_num_threads.restype = C.c_int
            return b.scipy_openblas_get_num_threads()
        except Exception:
            return 0

    def binary(self, op: str, a, b) -> np.ndarray:
        a, b = _c32(a), _c32(b)
        big = a if a.size >= b.size else b
        out = np.empty(big.shape if big.ndim else (), dtype=np.float32)
        sec = C.c_double()
        n = self.lib.ref_binary(BIN_OPS[op], _f(a), _shape(a), a.ndim, _f(b), _shape(b), b.ndim,
                                _f(out), max(out.size, 1), C.byref(sec))
        self._check(n, op)
        self.last_seconds = sec.value
        return out

    def mul_add(self, a, b, c) -> np.ndarray:
        a, b, c = _c32(a), _c32(b), _c32(c)
        big = max((a, b, c), key=lambda x: x.size)
        out = np.empty(big.shape, dtype=np.float32)
        sec = C.c_double()
        n = self.lib.ref_mul_add(_f(a), _shape(a), a.ndim, _f(b), _shape(b), b.ndim, _f(c), _shape(c), c.ndim,
                                 _f(out), out.size, C.byref(sec))
        self._check(n, "mul_add")
        self.last_seconds = sec.value
        return out

    def unary(self, op: str, x, p0: float = 0.0, p1: float = 0.0) -> np.ndarray:
        x = _c32(x)
        flat = x.reshape(-1)
        out = np.empty_like(flat)
        sec = C.c_double()
        n = self.lib.ref_unary(UN_OPS[op], _f(flat), flat.size, _f(out), p0, p1, C.byref(sec))
        self._check(n, op)
        self.last_seconds = sec.value
        return out.reshape(x.shape)

    def reduce_full(self, op: str, x) -> np.float32:
        x = _c32(x).reshape(-1)
        sec = C.c_double()
        v = self.lib.ref_reduce_full(RED_OPS[op], _f(x), x.size, C.byref(sec))
        self.last_seconds = sec.value
        return np.float32(v)

    def reduce_axis(self, op: str, x, axis: int) -> np.ndarray:
        x = _c32(x)
        oshape = x.shape[:axis] + x.shape[axis + 1:]
        out = np.empty(oshape, dtype=np.float32)
        sec = C.c_double()
        n = self.lib.ref_reduce_axis(RED_OPS[op], _f(x), _shape(x), x.ndim, axis, _f(out), max(out.size, 1), C.byref(sec))
        self._check(n, f"reduce_axis {op}")
        self.last_seconds = sec.value
        return out

    def argminmax(self, is_max: bool, x, axis=None, keepdims: bool = False) -> np.ndarray:
        x = _c32(x)
        if axis is None:
            oshape = (1,) * x.ndim if keepdims else ()
            ax = 128
        else:
            ax = axis
            oshape = tuple(1 if i == axis else s for i, s in enumerate(x.shape)) if keepdims else \
                x.shape[:axis] + x.shape[axis + 1:]
        out = np.empty(oshape, dtype=np.float32)
        sec = C.c_double()
        n = self.lib.ref_argminmax(int(is_max), _f(x), _shape(x), x.ndim, ax, int(keepdims), _f(out),
                                   max(out.size, 1), C.byref(sec))
        self._check(n, "argminmax")
        self.last_seconds = sec.value
        return out

    def matmul(self, a, b) -> np.ndarray:
        a, b = _c32(a), _c32(b)
        M, K = a.shape
        K2, N = b.shape
        assert K == K2
        out = np.empty((M, N), dtype=np.float32)
        sec = C.c_double()
        n = self.lib.ref_matmul(_f(a), _f(b), M, K, N, _f(out), C.byref(sec))
        self._check(n, "matmul")
        self.last_seconds = sec.value
        return out

    def matmul_nd(self, a, b) -> np.ndarray:
        a, b = _c32(a), _c32(b)
        out = np.empty(a.shape[:-1] + (b.shape[-1],), dtype=np.float32)
        sec = C.c_double()
        n = self.lib.ref_matmul_nd(_f(a), _shape(a), a.ndim, _f(b), _shape(b), b.ndim, _f(out), out.size, C.byref(sec))
        self._check(n, "matmul")
        self.last_seconds = sec.value
        return out

    def all(self, a) -> int:
        """NDArray_All on the reference's object code - only trustworthy for fewer than 8 elements (see oracle/port.c)."""
        a = _c32(a)
        r = self.lib.ref_all(_f(a), _shape(a), a.ndim)
        self._check(0, "all")
        return int(r)

    def allclose(self, a, b, rtol: float = 1e-5, atol: float = 1e-8) -> int:
        """NDArray_AllClose on the reference's object code - reads out of bounds beyond the first element (see oracle/port.c)."""
        a, b = _c32(a), _c32(b)
        r = self.lib.ref_allclose(_f(a), _f(b), _shape(a), a.ndim, rtol, atol)
        self._check(0, "allclose")
        return int(r)

    def outer(self, a, b) -> np.ndarray:
        """NDArray_Outer (linalg.c:724-751)."""
        a, b = _c32(a).ravel(), _c32(b).ravel()
        out = np.empty((a.size, b.size), dtype=np.float32)
        n = self.lib.ref_outer(_f(a), a.size, _f(b), b.size, _f(out))
        self._check(n, "outer")
        return out

    def norm1(self, a) -> np.float32:
        """NDArray_Norm(a, 1) = NDArray_L1Norm (linalg.c:423-447): max over columns of the sum of absolute values."""
        a = _c32(a)
        out = np.empty(1, dtype=np.float32)
        n = self.lib.ref_norm1(_f(a), _shape(a), a.ndim, _f(out))
        self._check(n, "norm1")
        return out[0]

    def dot(self, a, b) -> np.ndarray:
        a, b = _c32(a), _c32(b)
        if a.ndim == 2 and b.ndim == 2:
            oshape = (a.shape[0], b.shape[1])
        elif b.ndim == 1 and a.ndim > 1:
            oshape = a.shape[:-1]
        else:
            oshape = ()
        out = np.empty(oshape, dtype=np.float32)
        sec = C.c_double()
        n = self.lib.ref_dot(_f(a), _shape(a), a.ndim, _f(b), _shape(b), b.ndim, _f(out), max(out.size, 1), C.byref(sec))
        self._check(n, "dot")
        self.last_seconds = sec.value
        return out


class _Port:
    """The plain-C restatement (kind = "port")."""

    def __init__(self):
        self._lib = None

    @property
    def lib(self):
        if self._lib is None:
            build_port()
            lib = C.CDLL(PORT_SO)
            lib.port_binary.argtypes = [C.c_int, _fp, _fp, _fp, C.c_long]
            lib.port_mul_add.argtypes = [_fp, _fp, _fp, _fp, C.c_long]
            lib.port_unary.argtypes = [C.c_int, _fp, _fp, C.c_long, C.c_float, C.c_float]
            lib.port_reduce_full.restype = C.c_float
            lib.port_reduce_full.argtypes = [C.c_int, _fp, C.c_long]
            lib.port_reduce_axis.argtypes = [C.c_int, _fp, _fp, C.c_long, C.c_long, C.c_long]
            lib.port_argminmax.argtypes = [C.c_int, _fp, _fp, C.c_long, C.c_long, C.c_long]
            lib.port_matmul.argtypes = [_fp, _fp, _fp, C.c_long, C.c_long, C.c_long]
            lib.port_matmul_f64.argtypes = [_fp, _fp, _dp, C.c_long, C.c_long, C.c_long]
            lib.port_gemv.argtypes = [_fp, _fp, _fp, C.c_long, C.c_long]
            lib.port_all.argtypes = [_fp, C.c_long]
            lib.port_allclose.argtypes = [_fp, _fp, C.c_long, C.c_float, C.c_float]
            self._lib = lib
        return self._lib

    def binary(self, op: str, a, b) -> np.ndarray:
        a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
        shape = np.broadcast_shapes(a.shape, b.shape)
        if a.size >= b.size and a.size == int(np.prod(shape, dtype=np.int64)):
            shape = a.shape
        ab, bb = _c32(np.broadcast_to(a, shape)), _c32(np.broadcast_to(b, shape))
        out = np.empty(shape, dtype=np.float32)
        self.lib.port_binary(BIN_OPS[op], _f(ab), _f(bb), _f(out), out.size)
        return out

    def mul_add(self, a, b, c) -> np.ndarray:
        shape = np.broadcast_shapes(np.shape(a), np.shape(b), np.shape(c))
        aa, bb, cc = (_c32(np.broadcast_to(np.asarray(v, np.float32), shape)) for v in (a, b, c))
        out = np.empty(shape, dtype=np.float32)
        self.lib.port_mul_add(_f(aa), _f(bb), _f(cc), _f(out), out.size)
        return out

    def unary(self, op: str, x, p0: float = 0.0, p1: float = 0.0) -> np.ndarray:
        x = _c32(x)
        out = np.empty_like(x)
        self.lib.port_unary(UN_OPS[op], _f(x), _f(out), x.size, p0, p1)
        return out

    def reduce_full(self, op: str, x) -> np.float32:
        x = _c32(x).reshape(-1)
        return np.float32(self.lib.port_reduce_full(RED_OPS[op], _f(x), x.size))

    @staticmethod
    def _oli(shape, axis):
        outer = int(np.prod(shape[:axis], dtype=np.int64))
        inner = int(np.prod(shape[axis + 1:], dtype=np.int64))
        return outer, shape[axis], inner

    def reduce_axis(self, op: str, x, axis: int) -> np.ndarray:
        x = _c32(x)
        o, l, i = self._oli(x.shape, axis)
        out = np.empty(x.shape[:axis] + x.shape[axis + 1:], dtype=np.float32)
        self.lib.port_reduce_axis(RED_OPS[op], _f(x), _f(out), o, l, i)
        return out

    def argminmax(self, is_max: bool, x, axis=None, keepdims: bool = False) -> np.ndarray:
        x = _c32(x)
        if axis is None:
            out = np.empty((), dtype=np.float32)
            self.lib.port_argminmax(int(is_max), _f(x), _f(out), 1, x.size, 1)
            return out.reshape((1,) * x.ndim) if keepdims else out
        o, l, i = self._oli(x.shape, axis)
        out = np.empty(x.shape[:axis] + x.shape[axis + 1:], dtype=np.float32)
        self.lib.port_argminmax(int(is_max), _f(x), _f(out), o, l, i)
        return np.expand_dims(out, axis) if keepdims else out

    def matmul(self, a, b) -> np.ndarray:
        a, b = _c32(a), _c32(b)
        out = np.empty((a.shape[0], b.shape[1]), dtype=np.float32)
        self.lib.port_matmul(_f(a), _f(b), _f(out), a.shape[0], a.shape[1], b.shape[1])
        return out

    def matmul_f64(self, a, b) -> np.ndarray:
        a, b = _c32(a), _c32(b)
        out = np.empty((a.shape[0], b.shape[1]), dtype=np.float64)
        self.lib.port_matmul_f64(_f(a), _f(b), out.ctypes.data_as(_dp), a.shape[0], a.shape[1], b.shape[1])
        return out

    def all(self, a) -> int:
        a = _c32(a)
        return int(self.lib.port_all(_f(a), a.size))

    def allclose(self, a, b, rtol: float = 1e-5, atol: float = 1e-8) -> int:
        a, b = _c32(a), _c32(b)
        if a.shape != b.shape:
            raise RuntimeError("Shape mismatch")
        return int(self.lib.port_allclose(_f(a), _f(b), a.size, rtol, atol))

    def gemv(self, a, x) -> np.ndarray:
        a, x = _c32(a), _c32(x)
        rows = a.size // a.shape[-1]
        out = np.empty(a.shape[:-1], dtype=np.float32)
        self.lib.port_gemv(_f(a), _f(x), _f(out), rows, a.shape[-1])
        return out


DROPIN_SO = os.path.join(HERE, "_ref", "libnumpower_host_b200.so")
ref = _Ref()
port = _Port()
dropin = _Ref(DROPIN_SO)   # reference host objects (HAVE_CUBLAS) on top of libnb200.so; needs a GPU
N1_SO = os.path.join(HERE, "_ref", "libnumpower_host_b200_n1.so")
dropin_n1 = _Ref(N1_SO)    # same, with the Level-1 host patches (oracle/n1_patch.py + integration/nb200_numpower_glue.c)
