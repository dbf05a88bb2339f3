/* tests/r_api_stub: declarations of the small subset of R's public C API that
 * ebpmf.alpha_b200/r/src/r_glue.c uses, written from R's documented signatures ("Writing R
 * Extensions").  NOT R, and never linked: tests/test_r_glue.py runs `gcc -fsyntax-only` against it so
 * that typos and type errors in the glue are caught in an image that has no R installation. */
#ifndef EBPMF_TEST_R_STUB_R_H
#define EBPMF_TEST_R_STUB_R_H
#include <stddef.h>
char *R_alloc(size_t nelem, int eltsize);
#endif
