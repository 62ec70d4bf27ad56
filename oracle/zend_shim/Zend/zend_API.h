/* TEST INFRASTRUCTURE — Zend shim. */
#include "zend.h"
