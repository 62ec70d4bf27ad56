"""ctypes loader for libnb200.so — the hand-written sm_100a backend.  Fails loudly: there is no
CPU / eager fallback anywhere in this package."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libnb200.so")

i64 = C.c_int64
fp = C.c_void_p  # device / host float pointers travel as raw addresses
i64p = C.POINTER(C.c_int64)


class NB_NDArray(C.Structure):
    pass


NB_NDArray._fields_ = [
    ("data", C.c_void_p), ("ndim", C.c_int), ("shape", C.c_int64 * 8), ("numel", C.c_int64),
    ("device", C.c_int), ("refcount", C.c_int), ("base", C.POINTER(NB_NDArray)),
]
ndp = C.POINTER(NB_NDArray)

_lib = None

# name -> (restype, argtypes): exactly the declarations of include/nb200.h and include/nb200_host.h
ABI = {
    "nb200_init": (C.c_int, [C.c_int]), "nb200_shutdown": (C.c_int, []),
    "nb200_device_count": (C.c_int, [C.POINTER(C.c_int)]), "nb200_set_device": (C.c_int, [C.c_int]),
    "nb200_get_device": (C.c_int, [C.POINTER(C.c_int)]), "nb200_synchronize": (C.c_int, []),
    "nb200_last_error": (C.c_char_p, []), "nb200_stream": (C.c_void_p, []),
    "nb200_set_stream": (C.c_int, [C.c_void_p]), "nb200_launch_count": (i64, []),
    "nb200_trace_enable": (C.c_int, [C.c_void_p]),
    "nb200_graph_begin": (C.c_int, []), "nb200_graph_end": (C.c_int, [C.POINTER(C.c_void_p)]),
    "nb200_graph_launch": (C.c_int, [C.c_void_p]), "nb200_graph_destroy": (C.c_int, [C.c_void_p]),
    "nb200_poll_domain_error": (C.c_int, [C.POINTER(C.c_int)]),
    "nb200_alloc": (C.c_int, [C.POINTER(C.c_void_p), i64]), "nb200_free": (C.c_int, [C.c_void_p]),
    "nb200_copy_h2d": (C.c_int, [C.c_void_p, C.c_void_p, i64]),
    "nb200_copy_d2h": (C.c_int, [C.c_void_p, C.c_void_p, i64]),
    "nb200_copy_d2d": (C.c_int, [C.c_void_p, C.c_void_p, i64]),
    "nb200_memset_zero": (C.c_int, [C.c_void_p, i64]),
    "nb200_mem_stats": (C.c_int, [i64p, i64p]),
    "nb200_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), i64]), "nb200_host_free": (C.c_int, [C.c_void_p]),
    "nb200_ew_binary": (C.c_int, [C.c_int, fp, fp, fp, C.c_int, i64p, i64p, i64p]),
    "nb200_ew_binary_scalar": (C.c_int, [C.c_int, fp, fp, C.c_float, C.c_int, i64]),
    "nb200_ew_mul_add": (C.c_int, [fp, fp, fp, fp, C.c_int, i64p, i64p, i64p, i64p]),
    "nb200_ew_unary": (C.c_int, [C.c_int, fp, fp, i64, C.c_float, C.c_float]),
    "nb200_fill": (C.c_int, [fp, C.c_float, i64]),
    "nb200_reduce_full": (C.c_int, [C.c_int, fp, fp, i64]),
    "nb200_reduce_full_host": (C.c_int, [C.c_int, C.POINTER(C.c_float), fp, i64]),
    "nb200_reduce_axis": (C.c_int, [C.c_int, fp, fp, i64, i64, i64, C.c_int]),
    "nb200_argminmax": (C.c_int, [C.c_int, fp, fp, i64, i64, i64]),
    "nb200_argminmax_host": (C.c_int, [C.c_int, C.POINTER(C.c_float), fp, i64]),
    "nb200_sgemm": (C.c_int, [fp, fp, fp, i64, i64, i64, i64, i64, i64, C.c_int]),
    "nb200_sgemm_batched": (C.c_int, [fp, fp, fp, i64, i64, i64, i64, i64, i64, i64, C.c_int]),
    "nb200_sgemm_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, i64, i64, i64, C.c_int]),
    "nb200_sgemm_batched_host": (C.c_int, [fp, fp, fp, i64, i64, i64, i64, C.c_int]),
    "nb200_sgemm_workspace_bytes": (C.c_int, [i64, i64, i64, i64, C.c_int, i64p]),
    "nb200_gemm_resolve_precision": (C.c_int, [C.c_int, i64]),
    "nb200_gemv": (C.c_int, [fp, fp, fp, i64, i64]),
    "nb200_transpose2d": (C.c_int, [fp, fp, i64, i64]),
    "nb200_all": (C.c_int, [C.POINTER(C.c_int), fp, i64]),
    "nb200_allclose": (C.c_int, [C.POINTER(C.c_int), fp, fp, i64, C.c_float, C.c_float]),
    # multi-GPU shards: pointer arrays are (c_void_p * G)
    "nb200_shard_init": (C.c_int, [C.c_int, C.POINTER(C.c_int)]), "nb200_shard_finalize": (C.c_int, []),
    "nb200_shard_count": (C.c_int, [C.POINTER(C.c_int)]), "nb200_shard_device": (C.c_int, [C.c_int, C.POINTER(C.c_int)]),
    "nb200_shard_range": (C.c_int, [i64, C.c_int, i64p, i64p]),
    "nb200_shard_split": (C.c_int, [i64, C.c_int, C.c_int, i64p, i64p]), "nb200_shard_synchronize": (C.c_int, []),
    "nb200_shard_scatter": (C.c_int, [C.c_void_p, fp, i64, i64, C.c_int, C.c_int]),
    "nb200_shard_gather": (C.c_int, [fp, C.c_void_p, i64, i64, C.c_int, C.c_int]),
    "nb200_shard_upload": (C.c_int, [C.c_void_p, fp, i64, i64]),
    "nb200_shard_download": (C.c_int, [fp, C.c_void_p, i64, i64]),
    "nb200_shard_ew_binary": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, i64, i64]),
    "nb200_shard_ew_mul_add": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, i64, i64]),
    "nb200_shard_ew_unary": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, i64, i64, C.c_float, C.c_float]),
    "nb200_shard_reduce_full": (C.c_int, [C.c_int, C.POINTER(C.c_float), C.c_void_p, i64]),
    "nb200_shard_argminmax": (C.c_int, [C.c_int, C.POINTER(C.c_float), C.c_void_p, i64]),
    "nb200_sgemm_batched_sharded": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, i64, i64, i64, i64, C.c_int]),
    "nb200_sgemm_batched_scatter_gather": (C.c_int, [fp, fp, fp, i64, i64, i64, i64, C.c_int, C.c_int, C.c_int, i64, C.POINTER(C.c_float)]),
    # host mirror (include/nb200_host.h)
    "NB_last_error": (C.c_char_p, []),
    "NB_NDArray_FromHost": (ndp, [C.c_void_p, C.c_int, i64p]),
    "NB_NDArray_Empty": (ndp, [C.c_int, i64p, C.c_int]),
    "NB_NDArray_ToGPU": (ndp, [ndp]), "NB_NDArray_ToCPU": (ndp, [ndp]),
    "NB_NDArray_Slice0": (ndp, [ndp, i64]), "NB_NDArray_Reshape": (ndp, [ndp, C.c_int, i64p]),
    "NB_NDArray_FREE": (None, [ndp]), "NB_NDArray_CopyToHost": (C.c_int, [ndp, C.c_void_p]),
    "NB_NDArray_Add_Float": (ndp, [ndp, ndp]), "NB_NDArray_Subtract_Float": (ndp, [ndp, ndp]),
    "NB_NDArray_Multiply_Float": (ndp, [ndp, ndp]), "NB_NDArray_Divide_Float": (ndp, [ndp, ndp]),
    "NB_NDArray_Mod_Float": (ndp, [ndp, ndp]), "NB_NDArray_Pow_Float": (ndp, [ndp, ndp]),
    "NB_NDArray_Maximum": (ndp, [ndp, ndp]), "NB_NDArray_Minimum": (ndp, [ndp, ndp]),
    "NB_NDArray_Arctan2": (ndp, [ndp, ndp]), "NB_NDArray_Binary": (ndp, [C.c_int, ndp, ndp]),
    "NB_NDArray_MulAdd": (ndp, [ndp, ndp, ndp]),
    "NB_NDArray_Map": (ndp, [ndp, C.c_int, C.c_float, C.c_float]),
    "NB_NDArray_Sum_Float": (C.c_int, [ndp, C.POINTER(C.c_float)]),
    "NB_NDArray_Float_Prod": (C.c_int, [ndp, C.POINTER(C.c_float)]),
    "NB_NDArray_Min": (C.c_int, [ndp, C.POINTER(C.c_float)]),
    "NB_NDArray_Max": (C.c_int, [ndp, C.POINTER(C.c_float)]),
    "NB_reduce": (ndp, [ndp, C.c_int, C.c_int, C.c_int]),
    "NB_NDArray_ArgMinMaxCommon": (ndp, [ndp, C.c_int, C.c_int, C.c_int]),
    "NB_NDArray_Matmul": (ndp, [ndp, ndp, C.c_int]),
    "NB_NDArray_Dot": (ndp, [ndp, ndp]),
}


class BackendMissing(RuntimeError):
    pass


class BackendError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"nb200 error {code}: {message}")
        self.code = code
        self.message = message


def lib():
    """Load libnb200.so (built by `python -m numpower_b200.build`).  No fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise BackendMissing(
                f"{LIB_PATH} not found: build the CUDA backend with `python -m numpower_b200.build` "
                "(__graft_entry__.build()). numpower_b200 has no CPU or PyTorch fallback.")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in ABI.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc: int):
    if rc != 0:
        raise BackendError(rc, lib().nb200_last_error().decode())


def check_ptr(p):
    if not p:
        raise BackendError(-1, lib().NB_last_error().decode())
    return p
