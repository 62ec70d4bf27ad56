#!/usr/bin/env bash
# TEST INFRASTRUCTURE — the drop-in proof (SURVEY.md §8 b): NumPower's UNMODIFIED host sources are
# compiled with HAVE_CUBLAS (its --with-cuda configuration) where they lie under /root/reference and
# linked against numpower_b200/libnb200.so INSTEAD OF the reference's cuda_math.o + gpu_alloc.o.
# Output: oracle/_ref/libnumpower_host_b200.so (git-ignored; travels to the GPU box).
# Include order: include/nb200_cublas_shim (cublas_v2.h -> nb200_sgemm), the HAVE_CUBLAS config.h,
# the Zend shim, CUDA runtime headers (the host calls cudaMemcpy/cudaMemset/cudaDeviceSynchronize itself).
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(dirname "$HERE")"
REF="${NB200_REFERENCE_DIR:-/root/reference}"
OUT="$HERE/_ref"
if [ ! -d "$REF/src" ]; then
  echo "build_dropin.sh: $REF not present: using prebuilt $OUT/libnumpower_host_b200.so" >&2
  [ -f "$OUT/libnumpower_host_b200.so" ] || exit 1
  exit 0
fi
[ -f "$ROOT/numpower_b200/libnb200.so" ] || { echo "build libnb200.so first" >&2; exit 1; }
PY="${PYTHON:-python}"
BLAS_DIR="$($PY -c 'import os, scipy; print(os.path.join(os.path.dirname(os.path.dirname(scipy.__file__)), "scipy.libs"))')"
BLAS_SO="$(ls "$BLAS_DIR"/libscipy_openblas-*.so | head -1)"
CUDA="${CUDA_HOME:-/usr/local/cuda}"
mkdir -p "$OUT/obj_gpu"
REN=""
for s in cblas_sgemm cblas_sgemv cblas_sasum cblas_sdot cblas_sger cblas_snrm2 \
         LAPACKE_sgesdd LAPACKE_sgetrf LAPACKE_sgetri LAPACKE_sgeqrf LAPACKE_sorgqr LAPACKE_sgeev \
         LAPACKE_sgels LAPACKE_sgelsd LAPACKE_sgesv LAPACKE_spotrf LAPACKE_sgesvd sgetrf_ sgetri_; do
  REN="$REN -D$s=scipy_$s"
done
G="$HERE/zend_shim_gpu"
CFLAGS="-O2 -mavx2 -march=x86-64-v3 -fPIC -w $REN -DREF_ENTRY_GPU -I$ROOT/include/nb200_cublas_shim -I$G -I$G/a/b -I$G/x -I$HERE/zend_shim -I$CUDA/include -I$REF -I$REF/src"
OBJS=""
for f in src/types src/buffer src/iterators src/initializers src/ndarray src/manipulation src/indexing src/logic \
         src/ndmath/double_math src/ndmath/arithmetics src/ndmath/calculation src/ndmath/linalg; do
  o="$OUT/obj_gpu/$(basename $f).o"
  gcc $CFLAGS -c "$REF/$f.c" -o "$o"
  OBJS="$OBJS $o"
done
gcc $CFLAGS -c "$HERE/ref_entry.c" -o "$OUT/obj_gpu/ref_entry.o"
gcc -shared -Wl,-Bsymbolic -o "$OUT/libnumpower_host_b200.so" $OBJS "$OUT/obj_gpu/ref_entry.o" \
    "$ROOT/numpower_b200/libnb200.so" "$BLAS_SO" -L"$CUDA/lib64" -lcudart \
    -Wl,-rpath,"$ROOT/numpower_b200" -Wl,-rpath,"$BLAS_DIR" -Wl,-rpath,"$CUDA/lib64" -lm
echo "built $OUT/libnumpower_host_b200.so"
nm -D -u "$OUT/libnumpower_host_b200.so" | grep -E " (cuda_|v(malloc|free|mem)|NDArray_VFLOAT|NDArrayMathGPU|nb200_shim)" | sed 's/^ *U //' | tr '\n' ' '
echo
