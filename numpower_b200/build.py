"""Build libnb200.so (hand-written sm_100a CUDA + the C-ABI) in-tree with nvcc.

    python -m numpower_b200.build            # incremental
    python -m numpower_b200.build --force

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels with gpurun snapshots.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libnb200.so")
SOURCES = ["abi.cu", "ew.cu", "reduce.cu", "sgemm_tcgen05.cu", "misc.cu", "host_pipeline.cu", "shard.cu", "legacy.cu", "host/ndarray_host.cpp"]
# bring-up probes (scripts/tcgen05_probe.py): a separate library on top of libnb200.so, never loaded by the product
DEBUG_LIB = os.path.join(HERE, "libnb200_debug.so")
DEBUG_SOURCES = ["sgemm_debug.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=default"]


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, "common.cuh"), os.path.join(HERE, "..", "include", "nb200.h"),
               os.path.join(HERE, "..", "include", "nb200_legacy.h"), os.path.join(HERE, "..", "include", "nb200_host.h")]
    jobs = []
    objs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, os.path.basename(src).rsplit(".", 1)[0] + ".o")
        objs.append(o)
        if force or _stale(o, [s] + headers):
            jobs.append([NVCC, *FLAGS, "-c", s, "-o", o])

    def run(cmd):
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)

    with ThreadPoolExecutor(max_workers=min(6, max(1, len(jobs)))) as ex:
        list(ex.map(run, jobs))
    if force or jobs or _stale(LIB, objs):
        run([NVCC, "-shared", "-o", LIB, *objs, "-cudart", "static", "-lrt", "-lpthread", "-ldl"])
    dobjs = []
    for src in DEBUG_SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, os.path.basename(src).rsplit(".", 1)[0] + ".o")
        dobjs.append(o)
        if force or _stale(o, [s] + headers):
            run([NVCC, *FLAGS, "-c", s, "-o", o])
    if force or _stale(DEBUG_LIB, dobjs + [LIB]):
        run([NVCC, "-shared", "-o", DEBUG_LIB, *dobjs, "-cudart", "static", "-L", HERE, "-lnb200", "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
