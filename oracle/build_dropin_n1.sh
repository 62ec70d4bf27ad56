#!/usr/bin/env bash
# TEST INFRASTRUCTURE — Level-1 drop-in: the reference host built as in build_dropin.sh, but with the
# one-line glue calls of oracle/n1_patch.py inserted into copies of arithmetics.c / ndarray.c / calculation.c / linalg.c
# (copies live only in oracle/_ref/n1_src, git-ignored) and integration/nb200_numpower_glue.c compiled in.
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
rm -rf "$OUT/n1_src" "$OUT/obj_n1"; mkdir -p "$OUT/n1_src" "$OUT/obj_n1"
$PY "$HERE/n1_patch.py" "$REF" "$OUT/n1_src"
REN=""
for s in cblas_sgemm cblas_sgemv cblas_sasum cblas_sdot cblas_sger cblas_snrm2 \
         LAPACKE_sgesdd LAPACKE_sgetrf LAPACKE_sgetri LAPACKE_sgeqrf LAPACKE_sorgqr LAPACKE_sgeev \
         LAPACKE_sgels LAPACKE_sgelsd LAPACKE_sgesv LAPACKE_spotrf LAPACKE_sgesvd sgetrf_ sgetri_; do
  REN="$REN -D$s=scipy_$s"
done
G="$HERE/zend_shim_gpu"
INC="-I$ROOT/include/nb200_cublas_shim -I$ROOT/include -I$G -I$G/a/b -I$G/x -I$HERE/zend_shim -I$CUDA/include -I$REF -I$REF/src -I$REF/src/ndmath"
CFLAGS="-O2 -mavx2 -march=x86-64-v3 -fPIC -w $REN -DREF_ENTRY_GPU $INC"
OBJS=""
for f in src/types src/buffer src/iterators src/initializers src/manipulation src/indexing src/logic src/ndmath/double_math; do
  o="$OUT/obj_n1/$(basename $f).o"; gcc $CFLAGS -c "$REF/$f.c" -o "$o"; OBJS="$OBJS $o"
done
for f in src/ndarray src/ndmath/arithmetics src/ndmath/calculation src/ndmath/linalg; do
  o="$OUT/obj_n1/$(basename $f).o"
  # patched copy; quoted includes ("../config.h", "iterators.h", ...) resolve through -I to the reference tree
  gcc $CFLAGS -I"$REF/$(dirname $f)" -c "$OUT/n1_src/$f.c" -o "$o"; OBJS="$OBJS $o"
done
gcc $CFLAGS -c "$ROOT/integration/nb200_numpower_glue.c" -o "$OUT/obj_n1/glue.o"
gcc $CFLAGS -c "$HERE/ref_entry.c" -o "$OUT/obj_n1/ref_entry.o"
gcc -shared -Wl,-Bsymbolic -o "$OUT/libnumpower_host_b200_n1.so" $OBJS "$OUT/obj_n1/glue.o" "$OUT/obj_n1/ref_entry.o" \
    "$ROOT/numpower_b200/libnb200.so" "$BLAS_SO" -L"$CUDA/lib64" -lcudart \
    -Wl,-rpath,"$ROOT/numpower_b200" -Wl,-rpath,"$BLAS_DIR" -Wl,-rpath,"$CUDA/lib64" -lm
rm -rf "$OUT/n1_src"
echo "built $OUT/libnumpower_host_b200_n1.so"
