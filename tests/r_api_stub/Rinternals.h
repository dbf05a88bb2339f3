/* see R.h in this directory: syntax-check stub, not R */
#ifndef EBPMF_TEST_R_STUB_RINTERNALS_H
#define EBPMF_TEST_R_STUB_RINTERNALS_H
#include <stddef.h>
typedef struct SEXPREC *SEXP;
typedef ptrdiff_t R_xlen_t;
typedef enum { FALSE = 0, TRUE } Rboolean;
typedef unsigned int SEXPTYPE;
#define NILSXP 0
#define LGLSXP 10
#define INTSXP 13
#define REALSXP 14
#define STRSXP 16
#define VECSXP 19
#define EXTPTRSXP 22
extern SEXP R_NilValue, R_DimSymbol, R_DimNamesSymbol;
int TYPEOF(SEXP x);
double *REAL(SEXP x);
int *INTEGER(SEXP x);
int *LOGICAL(SEXP x);
R_xlen_t XLENGTH(SEXP x);
const char *CHAR(SEXP x);
SEXP STRING_ELT(SEXP x, R_xlen_t i);
SEXP VECTOR_ELT(SEXP x, R_xlen_t i);
SEXP SET_VECTOR_ELT(SEXP x, R_xlen_t i, SEXP v);
SEXP Rf_protect(SEXP s);
void Rf_unprotect(int n);
#define PROTECT(s) Rf_protect(s)
#define UNPROTECT(n) Rf_unprotect(n)
void Rf_error(const char *fmt, ...) __attribute__((noreturn, format(printf, 1, 2)));
double Rf_asReal(SEXP x);
int Rf_asInteger(SEXP x);
int Rf_asLogical(SEXP x);
int Rf_ncols(SEXP x);
int Rf_nrows(SEXP x);
SEXP Rf_install(const char *name);
SEXP R_do_slot(SEXP obj, SEXP name);
SEXP Rf_allocVector(SEXPTYPE type, R_xlen_t len);
SEXP Rf_allocMatrix(SEXPTYPE type, int nrow, int ncol);
typedef struct R_allocator R_allocator_t;
SEXP Rf_allocVector3(SEXPTYPE type, R_xlen_t len, R_allocator_t *allocator);
SEXP Rf_ScalarReal(double x);
SEXP Rf_ScalarInteger(int x);
SEXP Rf_ScalarLogical(int x);
SEXP Rf_mkNamed(SEXPTYPE type, const char **names);
SEXP Rf_setAttrib(SEXP x, SEXP name, SEXP val);
SEXP Rf_getAttrib(SEXP x, SEXP name);
Rboolean Rf_isMatrix(SEXP x);
Rboolean Rf_isS4(SEXP x);
Rboolean Rf_inherits(SEXP x, const char *klass);
SEXP Rf_coerceVector(SEXP x, SEXPTYPE type);
typedef void (*R_CFinalizer_t)(SEXP);
void R_RegisterCFinalizerEx(SEXP s, R_CFinalizer_t fun, Rboolean onexit);
SEXP R_MakeExternalPtr(void *p, SEXP tag, SEXP prot);
void *R_ExternalPtrAddr(SEXP s);
void R_ClearExternalPtr(SEXP s);
#endif
