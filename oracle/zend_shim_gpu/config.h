/* TEST INFRASTRUCTURE — config.h for the DROP-IN build: the reference's host sources compiled with
 * HAVE_CUBLAS (what `--with-cuda` defines, config.m4:7-34), so that every device branch
 * (`#ifdef HAVE_CUBLAS ... cuda_*_float(...)`) is active and resolves against libnb200.so. */
#define HAVE_AVX2 1
#define HAVE_CBLAS 1
#define HAVE_LAPACKE 1
#define HAVE_CUBLAS 1
