"""numpower_b200 — B200-native (sm_100a) device backend for NumPower's NDArray hot path.

Layout: csrc/ (hand-written CUDA kernels + the C-ABI of include/nb200.h, the legacy symbol layer and
the C++ host mirror), _lib.py (ctypes loader, no fallback), ndarray.py (NDArray / nd:: surface).
"""
from ._lib import BackendError, BackendMissing, lib  # noqa: F401
from .ndarray import NDArray, nd, GoldenBackend, TF32X1, TF32X3, BF16X3, GEMM_AUTO, FP16X3, FP16X3U, ORDER_TREE, ORDER_SEQUENTIAL  # noqa: F401
