#!/usr/bin/env bash
# TEST INFRASTRUCTURE — builds oracle/_ref/libnumpower_ref.so: the reference's
# OWN CPU hot path (src/*.c, src/ndmath/*.c), compiled UNMODIFIED from where the
# sources lie under /root/reference against the Zend shim in oracle/zend_shim and
# linked to the scipy-bundled OpenBLAS (0.3.31.dev, LP64, `scipy_` symbol prefix).
# The reference's own build system (phpize/config.m4) is not run.
#
# Flags: the reference's CFLAGS are `-mavx2 -march=native` (config.m4:50); we use
# -march=x86-64-v3 (AVX2+FMA, no AVX-512) so the .so also runs on the GPU box's
# host CPU.  Like -march=native on any AVX2+FMA machine this lets GCC contract
# the AVX mod body `a - floor(a/b)*b` into one vfnmadd (arithmetics.c:787-806).
#
# No reference SOURCE is copied into the repo; outputs go to oracle/_ref/ only
# (git-ignored, NOT gpurun-ignored: the .so travels to the GPU box).
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${NB200_REFERENCE_DIR:-/root/reference}"
OUT="$HERE/_ref"
SHIM="$HERE/zend_shim"
if [ ! -d "$REF/src" ]; then
  echo "build_ref.sh: $REF not present (GPU box): using prebuilt $OUT" >&2
  [ -f "$OUT/libnumpower_ref.so" ] || { echo "no prebuilt oracle/_ref" >&2; exit 1; }
  exit 0
fi
PY="${PYTHON:-python}"
BLAS_DIR="$($PY - <<'PY'
import os, scipy
print(os.path.join(os.path.dirname(os.path.dirname(scipy.__file__)), "scipy.libs"))
PY
)"
BLAS_SO="$(ls "$BLAS_DIR"/libscipy_openblas-*.so | head -1)"
mkdir -p "$OUT/obj"
REN=""
for s in cblas_sgemm cblas_sgemv cblas_sasum cblas_sdot cblas_sger cblas_snrm2 \
         LAPACKE_sgesdd LAPACKE_sgetrf LAPACKE_sgetri LAPACKE_sgeqrf LAPACKE_sorgqr LAPACKE_sgeev \
         LAPACKE_sgels LAPACKE_sgelsd LAPACKE_sgesv LAPACKE_spotrf LAPACKE_sgesvd sgetrf_ sgetri_; do
  REN="$REN -D$s=scipy_$s"
done
CFLAGS="-O2 -mavx2 -march=x86-64-v3 -fPIC -w $REN -I$SHIM -I$SHIM/a/b -I$SHIM/x -I$REF -I$REF/src"
OBJS=""
for f in src/types src/buffer src/iterators src/initializers src/ndarray src/manipulation src/indexing src/logic \
         src/ndmath/double_math src/ndmath/arithmetics src/ndmath/calculation src/ndmath/linalg; do
  o="$OUT/obj/$(basename $f).o"
  gcc $CFLAGS -c "$REF/$f.c" -o "$o"
  OBJS="$OBJS $o"
done
gcc $CFLAGS -c "$HERE/ref_entry.c" -o "$OUT/obj/ref_entry.o"
gcc -shared -Wl,-Bsymbolic -o "$OUT/libnumpower_ref.so" $OBJS "$OUT/obj/ref_entry.o" \
    "$BLAS_SO" -Wl,-rpath,"$BLAS_DIR" -lm
echo "$BLAS_SO" > "$OUT/blas_path.txt"
echo "built $OUT/libnumpower_ref.so (BLAS: $BLAS_SO)"
