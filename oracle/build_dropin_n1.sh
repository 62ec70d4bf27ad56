#!/usr/bin/env bash
# TEST INFRASTRUCTURE — Level-1 drop-in: the reference host built as in build_dropin.sh, but with the
# one-line glue calls of oracle/n1_patch.py inserted into copies of arithmetics.c / ndarray.c / calculation.c / linalg.c
# (copies live only under oracle/_ref, git-ignored) and integration/nb200_numpower_glue.c compiled in - built THROUGH the
# `build-b200` rule that integration/b200_build_patch.py appends to the reference's Makefile.frag (the --with-b200 plumbing).
# Output: oracle/_ref/libnumpower_host_b200_n1.so
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(dirname "$HERE")"
REF="${NB200_REFERENCE_DIR:-/root/reference}"
OUT="$HERE/_ref"
if [ ! -d "$REF/src" ]; then
  [ -f "$OUT/libnumpower_host_b200_n1.so" ] || exit 1
  exit 0
fi
PY="${PYTHON:-python}"
BLAS_DIR="$($PY -c 'import os, scipy; print(os.path.join(os.path.dirname(os.path.dirname(scipy.__file__)), "scipy.libs"))')"
BLAS_SO="$(ls "$BLAS_DIR"/libscipy_openblas-*.so | head -1)"
CUDA="${CUDA_HOME:-/usr/local/cuda}"
# ---- the tree a maintainer would build in: the reference's sources (symlinks), the four patched copies, the glue file as
# src/nb200_glue.c, config.h where configure would put it, and the PATCHED Makefile.frag (integration/b200_build_patch.py)
TREE="$OUT/n1_tree"
rm -rf "$TREE" "$OUT/n1_src"; mkdir -p "$TREE" "$OUT/n1_src"
cp -rs "$REF/src" "$TREE/src"
$PY "$HERE/n1_patch.py" "$REF" "$OUT/n1_src"
for f in src/ndarray src/ndmath/arithmetics src/ndmath/calculation src/ndmath/linalg; do
  rm -f "$TREE/$f.c"; cp "$OUT/n1_src/$f.c" "$TREE/$f.c"
done
cp "$ROOT/integration/nb200_numpower_glue.c" "$TREE/src/nb200_glue.c"
cp "$HERE/zend_shim_gpu/config.h" "$TREE/config.h"
$PY "$ROOT/integration/b200_build_patch.py" "$REF" "$TREE"
REN=""
for s in cblas_sgemm cblas_sgemv cblas_sasum cblas_sdot cblas_sger cblas_snrm2 \
         LAPACKE_sgesdd LAPACKE_sgetrf LAPACKE_sgetri LAPACKE_sgeqrf LAPACKE_sorgqr LAPACKE_sgeev \
         LAPACKE_sgels LAPACKE_sgelsd LAPACKE_sgesv LAPACKE_spotrf LAPACKE_sgesvd sgetrf_ sgetri_; do
  REN="$REN -D$s=scipy_$s"
done
G="$HERE/zend_shim_gpu"
INC="-I$G -I$G/a/b -I$G/x -I$HERE/zend_shim -I$CUDA/include -I$TREE -I$TREE/src -I$TREE/src/ndmath"
XFLAGS="-O2 -mavx2 -march=x86-64-v3 -w $REN -DREF_ENTRY_GPU"
# the test harness entry points (ref_entry.c) ride along as an extra object; numpower.c (the Zend binding) needs real PHP headers
gcc $XFLAGS -fPIC -I"$ROOT/include/nb200_cublas_shim" -I"$ROOT/include" $INC -c "$HERE/ref_entry.c" -o "$TREE/ref_entry.o"
# ---- `make build-b200` of the patched Makefile.frag: host files through $(CC), link against libnb200.so
make -s -C "$TREE" -f Makefile.frag build-b200 builddir="$TREE/" CC=gcc TARGET_SIZE=64 \
    NB200_INCLUDE="$ROOT/include" NB200_LIBDIR="$ROOT/numpower_b200" \
    NB200_HOST_SRCS="src/types.c src/buffer.c src/iterators.c src/initializers.c src/manipulation.c src/indexing.c src/logic.c src/ndmath/double_math.c src/ndarray.c src/ndmath/arithmetics.c src/ndmath/calculation.c src/ndmath/linalg.c src/nb200_glue.c" \
    COMMON_FLAGS="$INC" EXTRA_CFLAGS="$XFLAGS" NB200_EXTRA_OBJS="$TREE/ref_entry.o" \
    NB200_EXTRA_LIBS="-Wl,-Bsymbolic $BLAS_SO -L$CUDA/lib64 -lcudart -Wl,-rpath,$BLAS_DIR -Wl,-rpath,$CUDA/lib64 -lm" \
    NB200_OUT="$OUT/libnumpower_host_b200_n1.so"
rm -rf "$TREE" "$OUT/n1_src"
echo "built $OUT/libnumpower_host_b200_n1.so"
