/* TEST INFRASTRUCTURE — minimal stand-in for PHP's <php.h> so the reference's
 * hot-path C files (src/*.c, src/ndmath/*.c under /root/reference) compile
 * UNMODIFIED outside a PHP build.  Only the Zend identifiers those files use
 * are provided; allocation maps to libc, zend_throw_error records a message.
 * Nothing here is product code. */
#ifndef NB200_ORACLE_PHP_SHIM_H
#define NB200_ORACLE_PHP_SHIM_H
#include "Zend/zend.h"
#endif
