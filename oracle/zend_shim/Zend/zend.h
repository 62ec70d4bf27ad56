/* TEST INFRASTRUCTURE — Zend shim (see ../php.h). */
#ifndef NB200_ORACLE_ZEND_SHIM_H
#define NB200_ORACLE_ZEND_SHIM_H
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <stdio.h>
#include <string.h>
#include <limits.h>
#include <float.h>
#include <math.h>
#include <stdbool.h>
#include <stdarg.h>
#include <assert.h>
#include <errno.h>
#include <ctype.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef unsigned long zend_ulong;
typedef long zend_long;
typedef unsigned char zend_uchar;

typedef struct _zend_string { size_t len; char val[1]; } zend_string;
typedef struct _zend_class_entry { int unused; } zend_class_entry;
typedef struct _zend_object { zend_class_entry *ce; void *properties_table; } zend_object;
struct _zval_struct;
typedef struct _zend_array {
    uint32_t nNumUsed;
    uint32_t nNumOfElements;
    struct _zval_struct *arPacked;
} zend_array;
typedef zend_array HashTable;

typedef union _zend_value {
    zend_long lval;
    double dval;
    zend_array *arr;
    zend_object *obj;
    void *ptr;
} zend_value;

typedef struct _zval_struct {
    zend_value value;
    uint32_t type;
} zval;

#define IS_UNDEF 0
#define IS_NULL 1
#define IS_FALSE 2
#define IS_TRUE 3
#define IS_LONG 4
#define IS_DOUBLE 5
#define IS_STRING 6
#define IS_ARRAY 7
#define IS_OBJECT 8

#define Z_TYPE_P(z) ((z)->type)
#define Z_TYPE(z) ((z).type)
#define Z_ARRVAL_P(z) ((z)->value.arr)
#define Z_ARR_P(z) ((z)->value.arr)
#define Z_OBJ_P(z) ((z)->value.obj)
#define Z_LVAL_P(z) ((z)->value.lval)
#define Z_DVAL_P(z) ((z)->value.dval)
#define Z_LVAL(z) ((z).value.lval)
#define Z_DVAL(z) ((z).value.dval)
#define ZVAL_DEREF(z) do { } while (0)
#define ZVAL_LONG(z, l) do { (z)->value.lval = (l); (z)->type = IS_LONG; } while (0)
#define ZVAL_DOUBLE(z, d) do { (z)->value.dval = (d); (z)->type = IS_DOUBLE; } while (0)

#define ZEND_HASH_FOREACH_VAL(ht, _val) do { \
    zend_array *__ht = (ht); uint32_t __i; \
    for (__i = 0; __i < __ht->nNumUsed; __i++) { \
        _val = &__ht->arPacked[__i];
#define ZEND_HASH_FOREACH_END() } } while (0)

static inline uint32_t zend_array_count(zend_array *ht) { return ht->nNumOfElements; }
static inline zval *zend_hash_index_find(const zend_array *ht, zend_ulong h) {
    return h < ht->nNumUsed ? &ht->arPacked[h] : NULL;
}
static inline zend_long zval_get_long(zval *z) {
    return z->type == IS_DOUBLE ? (zend_long) z->value.dval : z->value.lval;
}
static inline double zval_get_double(zval *z) {
    return z->type == IS_DOUBLE ? z->value.dval : (double) z->value.lval;
}
static inline void convert_to_long(zval *z) { zend_long v = zval_get_long(z); ZVAL_LONG(z, v); }
static inline void convert_to_double(zval *z) { double v = zval_get_double(z); ZVAL_DOUBLE(z, v); }

/* PHP-array export is out of scope for the oracle: stubs keep the files linking. */
static inline void array_init_size(zval *z, uint32_t n) { (void) n; z->type = IS_ARRAY; z->value.arr = NULL; }
static inline void array_init(zval *z) { z->type = IS_ARRAY; z->value.arr = NULL; }
static inline int add_index_zval(zval *arr, zend_ulong i, zval *v) { (void) arr; (void) i; (void) v; return 0; }
static inline int add_index_double(zval *arr, zend_ulong i, double d) { (void) arr; (void) i; (void) d; return 0; }
static inline int add_next_index_double(zval *arr, double d) { (void) arr; (void) d; return 0; }
static inline int add_next_index_zval(zval *arr, zval *v) { (void) arr; (void) v; return 0; }
static inline int add_next_index_long(zval *arr, zend_long d) { (void) arr; (void) d; return 0; }

#define emalloc(n) malloc(n)
#define efree(p) free(p)
#define ecalloc(n, s) calloc((n), (s))
#define erealloc(p, n) realloc((p), (n))
#define safe_emalloc(n, s, o) malloc((size_t)(n) * (size_t)(s) + (size_t)(o))
#define estrdup(s) strdup(s)

#define zend_always_inline inline __attribute__((always_inline))
#ifndef XtOffsetOf
#define XtOffsetOf(s, m) offsetof(s, m)
#endif
#define ZEND_API
#define SUCCESS 0
#define FAILURE -1
#define E_WARNING 2
#define E_NOTICE 8
#define E_ERROR 1

/* provided by oracle/ref_entry.c */
void zend_throw_error(zend_class_entry *ce, const char *format, ...);
void zend_error(int type, const char *format, ...);
void php_error_docref(const char *docref, int type, const char *format, ...);
#define php_printf printf

#ifdef __cplusplus
}
#endif
#endif
