/* TEST INFRASTRUCTURE — the flags config.m4:46-57,67-87,89-106 would define on
 * an AVX2 host with cblas + lapacke.  HAVE_CUBLAS is deliberately NOT set:
 * the oracle is the reference's CPU path. */
#define HAVE_AVX2 1
#define HAVE_CBLAS 1
#define HAVE_LAPACKE 1
