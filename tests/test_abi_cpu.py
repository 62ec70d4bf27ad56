"""CPU tests of the boundary: libnb200.so loads, exports every symbol include/*.h declares, matches the
ctypes prototypes, and fails LOUDLY (no CPU fallback) when no GPU is present."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = set(re.findall(r"\b((?:nb200|NB|cuda|v(?:malloc|free|memcheck|memcpyd2d|memcpyh2d)|NDArray_VFLOAT\w*|NDArrayMathGPU)\w*)\s*\(", text))
    return {n for n in names if not n.startswith("NB200_") and n not in ("NB_NDArray",)}


@pytest.fixture(scope="module")
def lib():
    import numpower_b200 as nb
    return nb.lib()


@pytest.mark.parametrize("header", ["nb200.h", "nb200_host.h", "nb200_legacy.h"])
def test_every_declared_symbol_is_exported(lib, header):
    names = _declared(header)
    assert len(names) > 10
    missing = [n for n in sorted(names) if not hasattr(lib, n)]
    assert not missing, f"{header}: not exported by libnb200.so: {missing}"


def test_cublas_shim_symbols_exported(lib):
    for n in ("nb200_shim_cublasCreate", "nb200_shim_cublasSgemm", "nb200_shim_cublasDestroy"):
        assert hasattr(lib, n)


def test_ctypes_table_matches_header():
    from numpower_b200._lib import ABI
    declared = _declared("nb200.h") | _declared("nb200_host.h")
    assert declared <= set(ABI), sorted(declared - set(ABI))


def test_no_cpu_fallback_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import numpower_b200 as nb
    with pytest.raises(nb.BackendError, match="no CPU fallback"):
        nb.NDArray.array(np.ones((2, 2), np.float32)).gpu()
    a = nb.NDArray.array(np.ones((2, 2), np.float32))   # host container works without a device
    assert a.shape == (2, 2) and not a.isGPU()
    with pytest.raises(nb.BackendError, match="GPU only"):
        nb.nd.add(a, a)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under numpower_b200/ may reference it."""
    pkg = os.path.join(ROOT, "numpower_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in text and "from oracle" not in text and "liboracle" not in text, f


def test_gemm_auto_resolves_to_the_guaranteed_mode_by_default(monkeypatch):
    """NB200_GEMM_AUTO (what nd::matmul passes) = FP16X3 for K >= 128 (guaranteed bound at the kind::f16 rate, device-side
    repair / TF32X3 fallback) and TF32X3 below; explicit modes are returned unchanged.  Pure host logic: callable without a GPU."""
    monkeypatch.delenv("NB200_GEMM_AUTO_MODE", raising=False)
    import numpower_b200 as nb
    lib = nb.lib()
    assert lib.nb200_gemm_resolve_precision(nb.GEMM_AUTO, 4096) == nb.FP16X3U
    assert lib.nb200_gemm_resolve_precision(nb.GEMM_AUTO, 128) == nb.FP16X3U
    assert lib.nb200_gemm_resolve_precision(nb.GEMM_AUTO, 16) == nb.TF32X3
    for mode in (nb.TF32X3, nb.TF32X1, nb.BF16X3, nb.FP16X3):
        assert lib.nb200_gemm_resolve_precision(mode, 4096) == mode


@pytest.mark.skipif(not os.path.isdir(os.environ.get("NB200_REFERENCE_DIR", "/root/reference")), reason="reference tree absent")
def test_with_b200_build_patch_applies_to_the_reference_build_files(tmp_path):
    """integration/b200_build_patch.py: the --with-b200 switch lands after PHP_ARG_WITH(cuda) (config.m4:7-8), gpu_alloc.c leaves the
    source list in favour of the glue file, and Makefile.frag gains build-b200 / install-b200 next to install-cuda.  (The rule
    itself is exercised by oracle/build_dropin_n1.sh, which builds the Level-1 drop-in library through it.)"""
    import subprocess
    import sys
    ref = os.environ.get("NB200_REFERENCE_DIR", "/root/reference")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.check_call([sys.executable, os.path.join(root, "integration", "b200_build_patch.py"), ref, str(tmp_path)], stdout=subprocess.DEVNULL)
    m4 = (tmp_path / "config.m4").read_text()
    assert m4.index("PHP_ARG_WITH(cuda") < m4.index("PHP_ARG_WITH(b200") < m4.index('if test "$PHP_CUDA" != "no"')
    assert "PHP_ADD_LIBRARY_WITH_PATH(nb200" in m4 and "AC_DEFINE(HAVE_CUBLAS,1" in m4 and "PHP_CUDA=no" in m4
    assert "src/gpu_alloc.c \\" not in m4 and "$NDARRAY_GPU_ALLOC_SRC $NDARRAY_NB200_GLUE_SRC \\" in m4
    frag = (tmp_path / "Makefile.frag").read_text()
    assert frag.index("install-cuda:") < frag.index("build-b200:") < frag.index("install-b200: build-b200")
    assert "-lnb200" in frag and "cuda_math.cu" not in frag[frag.index("build-b200:"):]
    out = subprocess.run(["make", "-n", "-f", str(tmp_path / "Makefile.frag"), "build-b200", "TARGET_SIZE=64", "NB200_HOST_SRCS=src/types.c"],
                         capture_output=True, text=True)
    assert out.returncode == 0 and "-DHAVE_CUBLAS=1" in out.stdout and "-lnb200" in out.stdout, out.stderr
