"""--with-b200: the build plumbing a NumPower maintainer adds so that the extension links libnb200.so as its device backend
(SURVEY.md §8 f, N3; replaces the `--with-cuda` block of config.m4:7-34 and the `install-cuda` rule of Makefile.frag:65-90).

    python integration/b200_build_patch.py <numpower_dir> <out_dir>

writes patched COPIES of config.m4 and Makefile.frag to <out_dir> (nothing of the reference is stored in this repo).  The blocks
below are the whole patch: two insertions + one substitution in config.m4, one appended section in Makefile.frag.
oracle/build_dropin_n1.sh builds the Level-1 drop-in library THROUGH the patched Makefile.frag (`make build-b200`), so the rule is
exercised by tests/test_dropin_n1_gpu.py; config.m4 needs phpize/autoconf, which this image does not have - the patched file is
only checked for its anchors (tests/test_abi_cpu.py).
"""
import os
import re
import sys

# ---- config.m4 -------------------------------------------------------------------------------------------------------
# (1) right after the PHP_ARG_WITH(cuda, ...) declaration (config.m4:7-8): the new switch.  HAVE_CUBLAS is the extension's
#     existing meaning of "a device backend is linked" (every GPU branch of src/ is guarded by it); libnb200.so provides the
#     symbols of gpu_alloc.c / cuda_math.cu / cuda_dnn.cu, so --with-b200 switches --with-cuda's nvcc path off.
M4_SWITCH = r'''
PHP_ARG_WITH(b200, for the B200 device backend (libnb200),
[  --with-b200=DIR       Use DIR/lib/libnb200.so (headers in DIR/include) as the device backend], [no], [no])

NDARRAY_GPU_ALLOC_SRC="src/gpu_alloc.c"
NDARRAY_NB200_GLUE_SRC=""
if test "$PHP_B200" != "no"; then
    if test ! -f "$PHP_B200/include/nb200.h"; then
      AC_MSG_ERROR([nb200.h not found under $PHP_B200/include])
    fi
    AC_DEFINE(HAVE_CUBLAS,1,[a device backend is linked (libnb200 provides the cuda_* / vmalloc symbols)])
    AC_DEFINE(HAVE_NB200,1,[B200 backend])
    PHP_ADD_INCLUDE($PHP_B200/include/nb200_cublas_shim)
    PHP_ADD_INCLUDE($PHP_B200/include)
    PHP_ADD_LIBRARY_WITH_PATH(nb200, $PHP_B200/lib, NDARRAY_SHARED_LIBADD)
    PHP_ADD_MAKEFILE_FRAGMENT($abs_srcdir/Makefile.frag, $abs_builddir)
    NDARRAY_GPU_ALLOC_SRC=""
    NDARRAY_NB200_GLUE_SRC="src/nb200_glue.c"
    PHP_CUDA=no
    AC_MSG_RESULT([B200 backend: $PHP_B200])
fi
'''
M4_ANCHOR = r"\[  --with-cuda           Include CUDA support\], \[no\], \[no\]\)\n"
# (2) the source list of PHP_NEW_EXTENSION: gpu_alloc.c only without the B200 backend, the glue file only with it
M4_SRC_OLD = "      src/gpu_alloc.c \\\n"
M4_SRC_NEW = "      $NDARRAY_GPU_ALLOC_SRC $NDARRAY_NB200_GLUE_SRC \\\n"

# ---- Makefile.frag ---------------------------------------------------------------------------------------------------
FRAG = r'''
################################## B200 backend (libnb200.so) ##################################
# The counterpart of install-cuda for --with-b200=DIR: every host file is compiled with the C compiler - no nvcc - with HAVE_CUBLAS
# defined (the extension's switch for "a device backend exists") and the cuBLAS shim header first on the include path; the link
# pulls the device code from libnb200.so.  src/gpu_alloc.c, src/ndmath/cuda/cuda_math.cu and src/ndmath/cuda/cuda_dnn.cu are NOT
# compiled: libnb200.so exports their symbols (include/nb200_legacy.h).
NB200_DIR ?= /usr/local/nb200
NB200_INCLUDE ?= $(NB200_DIR)/include
NB200_LIBDIR ?= $(NB200_DIR)/lib
NB200_HOST_SRCS ?= numpower.c src/initializers.c src/ndmath/double_math.c src/ndarray.c src/debug.c src/buffer.c src/logic.c \
    src/ndmath/linalg.c src/manipulation.c src/dnn.c src/iterators.c src/indexing.c src/ndmath/arithmetics.c \
    src/ndmath/calculation.c src/ndmath/statistics.c src/ndmath/signal.c src/types.c src/nb200_glue.c
NB200_CFLAGS = -DHAVE_CUBLAS=1 -DHAVE_NB200=1 -I$(NB200_INCLUDE)/nb200_cublas_shim -I$(NB200_INCLUDE)
NB200_OUT ?= .libs/ndarray.so

build-b200:
	rm -rf ./.libs
	mkdir ./.libs
	for f in $(NB200_HOST_SRCS); do \
	  $(CC) $(NB200_CFLAGS) -I. -I$(builddir)./src $(COMMON_FLAGS) $(CFLAGS_CLEAN) $(EXTRA_CFLAGS) -fPIC -c $(builddir)./$$f -o .libs/`basename $$f .c`.o || exit 1; \
	done
	$(CC) -shared .libs/*.o $(NB200_EXTRA_OBJS) -L$(NB200_LIBDIR) -lnb200 -Wl,-rpath,$(NB200_LIBDIR) $(NB200_EXTRA_LIBS) -o $(NB200_OUT)

install-b200: build-b200
	cp $(NB200_OUT) $(phplibdir)/ndarray.so
	cp $(NB200_OUT) $(EXTENSION_DIR)/ndarray.so
'''


def patch_config_m4(text: str) -> str:
    m = re.search(M4_ANCHOR, text)
    if not m:
        raise SystemExit("config.m4: PHP_ARG_WITH(cuda, ...) anchor not found")
    text = text[:m.end()] + M4_SWITCH + text[m.end():]
    if text.count(M4_SRC_OLD) != 1:
        raise SystemExit("config.m4: source-list anchor (src/gpu_alloc.c) not found")
    return text.replace(M4_SRC_OLD, M4_SRC_NEW)


def patch_makefile_frag(text: str) -> str:
    if "install-cuda:" not in text:
        raise SystemExit("Makefile.frag: install-cuda rule not found")
    return text.rstrip("\n") + "\n" + FRAG


def main(ref: str, out: str) -> None:
    os.makedirs(out, exist_ok=True)
    open(os.path.join(out, "config.m4"), "w").write(patch_config_m4(open(os.path.join(ref, "config.m4")).read()))
    open(os.path.join(out, "Makefile.frag"), "w").write(patch_makefile_frag(open(os.path.join(ref, "Makefile.frag")).read()))
    print(f"patched config.m4 and Makefile.frag -> {out}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
