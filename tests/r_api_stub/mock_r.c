/* mock_r.c -- a FUNCTIONAL mock of the small subset of R's C API that ebpmf.alpha_b200/r/src/r_glue.c uses
 * (declared in this directory's R.h / Rinternals.h / R_ext).  TEST INFRASTRUCTURE, NOT R.
 *
 * R is installed neither in the build image nor on the GPU box, so the `.Call` shim could only be syntax-checked.
 * With this file it RUNS: tests/r_mock.py builds r_glue.c + mock_r.c into one shared object next to
 * libebpmf_b200.so, constructs SEXPs from NumPy arrays, dispatches `.Call("C_...", ...)` through the glue's own
 * registration table (R_registerRoutines, arity checked as R checks it) and reads the results back.  The
 * semantics follow "Writing R Extensions": vectors carry a length, a type and an attribute list; S4 slots are
 * attributes; Rf_error() does not return (longjmp to the .Call trampoline); PROTECT/UNPROTECT must balance over a
 * successful call -- the trampoline checks that, which a real R session only reports as a warning.
 * What it does not model: garbage collection (objects live until mockr_free), ALTREP, encodings, NA_INTEGER. */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <R_ext/Rallocators.h>
#include <setjmp.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define SYMSXP 1
#define CHARSXP 9

struct attr { SEXP name, value; struct attr *next; };
struct SEXPREC {
    SEXPTYPE type;
    R_xlen_t len;
    void *data;                 /* double / int / SEXP elements, or the characters of a CHARSXP / SYMSXP */
    struct attr *attrib;
    int s4;
    void *ext;                  /* EXTPTRSXP address */
    R_CFinalizer_t fin;
    R_allocator_t alloc;        /* Rf_allocVector3 */
    void *alloc_block;
};

static struct SEXPREC nil_rec = {NILSXP, 0, NULL, NULL, 0, NULL, NULL, {NULL, NULL, NULL, NULL}, NULL};
SEXP R_NilValue = &nil_rec;
SEXP R_DimSymbol, R_DimNamesSymbol;
static SEXP sym_names, sym_class;

static jmp_buf g_jmp;
static int g_in_call = 0, g_protect = 0, g_max_protect = 0;
static char g_err[2048];
static const R_CallMethodDef *g_routines = NULL;
static long g_live = 0;

static SEXP new_rec(SEXPTYPE type, R_xlen_t len, size_t elt)
{
    SEXP s = (SEXP)calloc(1, sizeof *s);
    s->type = type; s->len = len;
    if (elt) s->data = calloc((size_t)(len > 0 ? len : 1), elt);
    if (type == VECSXP || type == STRSXP) for (R_xlen_t i = 0; i < len; ++i) ((SEXP *)s->data)[i] = R_NilValue;
    ++g_live;
    return s;
}

static size_t elt_size(SEXPTYPE t)
{
    switch (t) {
    case REALSXP: return sizeof(double);
    case INTSXP: case LGLSXP: return sizeof(int);
    case VECSXP: case STRSXP: return sizeof(SEXP);
    default: return 0;
    }
}

/* ---- symbols, strings ------------------------------------------------------------------------------------ */
#define MAX_SYMS 256
static SEXP g_syms[MAX_SYMS];
static int g_nsyms = 0;
SEXP Rf_install(const char *name)
{
    for (int i = 0; i < g_nsyms; ++i) if (!strcmp((const char *)g_syms[i]->data, name)) return g_syms[i];
    SEXP s = new_rec(SYMSXP, (R_xlen_t)strlen(name), 0);
    s->data = strdup(name);
    if (g_nsyms < MAX_SYMS) g_syms[g_nsyms++] = s;
    return s;
}

SEXP mockr_mkchar(const char *s)
{
    SEXP c = new_rec(CHARSXP, (R_xlen_t)strlen(s), 0);
    c->data = strdup(s);
    return c;
}

const char *CHAR(SEXP x) { return (const char *)x->data; }

/* ---- accessors ------------------------------------------------------------------------------------------- */
int TYPEOF(SEXP x) { return (int)x->type; }
R_xlen_t XLENGTH(SEXP x) { return x->len; }
static void need(SEXP x, SEXPTYPE t, const char *what)
{
    if (x->type != t) Rf_error("mock R: %s() applied to a SEXP of type %u", what, x->type);
}
double *REAL(SEXP x) { need(x, REALSXP, "REAL"); return (double *)x->data; }
int *INTEGER(SEXP x)
{
    if (x->type != INTSXP && x->type != LGLSXP) Rf_error("mock R: INTEGER() applied to a SEXP of type %u", x->type);
    return (int *)x->data;
}
int *LOGICAL(SEXP x) { need(x, LGLSXP, "LOGICAL"); return (int *)x->data; }
SEXP STRING_ELT(SEXP x, R_xlen_t i)
{
    need(x, STRSXP, "STRING_ELT");
    if (i < 0 || i >= x->len) Rf_error("mock R: STRING_ELT index out of range");
    return ((SEXP *)x->data)[i];
}
SEXP VECTOR_ELT(SEXP x, R_xlen_t i)
{
    need(x, VECSXP, "VECTOR_ELT");
    if (i < 0 || i >= x->len) Rf_error("mock R: VECTOR_ELT index out of range");
    return ((SEXP *)x->data)[i];
}
SEXP SET_VECTOR_ELT(SEXP x, R_xlen_t i, SEXP v)
{
    need(x, VECSXP, "SET_VECTOR_ELT");
    if (i < 0 || i >= x->len) Rf_error("mock R: SET_VECTOR_ELT index out of range");
    ((SEXP *)x->data)[i] = v;
    return v;
}
void mockr_set_string_elt(SEXP x, R_xlen_t i, SEXP c) { ((SEXP *)x->data)[i] = c; }

/* ---- protection stack (counted, not used for collection) --------------------------------------------------- */
SEXP Rf_protect(SEXP s)
{
    if (++g_protect > g_max_protect) g_max_protect = g_protect;
    return s;
}
void Rf_unprotect(int n)
{
    g_protect -= n;
    if (g_protect < 0) { g_protect = 0; Rf_error("mock R: unprotect(): only %d protected item(s)", g_protect + n); }
}

/* ---- errors ------------------------------------------------------------------------------------------------ */
void Rf_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    if (g_in_call) longjmp(g_jmp, 1);
    fprintf(stderr, "mock R: Rf_error outside .Call: %s\n", g_err);
    abort();
}

/* ---- allocation --------------------------------------------------------------------------------------------- */
SEXP Rf_allocVector(SEXPTYPE type, R_xlen_t len)
{
    if (type == NILSXP) return R_NilValue;
    if (!elt_size(type) || len < 0) Rf_error("mock R: allocVector(type %u, length %ld) is not supported", type, (long)len);
    return new_rec(type, len, elt_size(type));
}

/* R puts the vector header and the data into ONE block obtained from the allocator and releases it with
 * mem_free from the garbage collector; the mock keeps the header separate but asks for the same number of bytes. */
#define R_VECTOR_HEADER_BYTES 48
SEXP Rf_allocVector3(SEXPTYPE type, R_xlen_t len, R_allocator_t *allocator)
{
    if (!allocator) return Rf_allocVector(type, len);
    if (type != REALSXP) Rf_error("mock R: allocVector3 is only modelled for REALSXP");
    SEXP s = new_rec(type, len, 0);
    s->alloc = *allocator;
    s->alloc_block = allocator->mem_alloc(&s->alloc, R_VECTOR_HEADER_BYTES + (size_t)len * sizeof(double));
    if (!s->alloc_block) { free(s); --g_live; Rf_error("cannot allocate vector of size %.1f Gb", (double)len * 8 / 1073741824.0); }
    s->data = (char *)s->alloc_block + R_VECTOR_HEADER_BYTES;
    return s;
}

SEXP Rf_allocMatrix(SEXPTYPE type, int nrow, int ncol)
{
    SEXP s = Rf_allocVector(type, (R_xlen_t)nrow * ncol);
    SEXP dim = Rf_allocVector(INTSXP, 2);
    INTEGER(dim)[0] = nrow; INTEGER(dim)[1] = ncol;
    Rf_setAttrib(s, R_DimSymbol, dim);
    return s;
}

char *R_alloc(size_t nelem, int eltsize) { return (char *)calloc(nelem ? nelem : 1, (size_t)eltsize); }   /* leaked: transient in R */

SEXP Rf_ScalarReal(double x) { SEXP s = Rf_allocVector(REALSXP, 1); REAL(s)[0] = x; return s; }
SEXP Rf_ScalarInteger(int x) { SEXP s = Rf_allocVector(INTSXP, 1); INTEGER(s)[0] = x; return s; }
SEXP Rf_ScalarLogical(int x) { SEXP s = Rf_allocVector(LGLSXP, 1); LOGICAL(s)[0] = x != 0; return s; }

SEXP Rf_mkNamed(SEXPTYPE type, const char **names)
{
    R_xlen_t n = 0;
    while (names[n][0] != '\0') ++n;
    SEXP s = Rf_allocVector(type, n);
    SEXP nm = Rf_allocVector(STRSXP, n);
    for (R_xlen_t i = 0; i < n; ++i) mockr_set_string_elt(nm, i, mockr_mkchar(names[i]));
    Rf_setAttrib(s, sym_names, nm);
    return s;
}

/* ---- attributes (S4 slots are attributes, as in R) ------------------------------------------------------------ */
SEXP Rf_getAttrib(SEXP x, SEXP name)
{
    for (struct attr *a = x->attrib; a; a = a->next) if (a->name == name) return a->value;
    return R_NilValue;
}
SEXP Rf_setAttrib(SEXP x, SEXP name, SEXP val)
{
    if (x == R_NilValue) Rf_error("attempt to set an attribute on NULL");
    for (struct attr *a = x->attrib; a; a = a->next) if (a->name == name) { a->value = val; return val; }
    struct attr *a = (struct attr *)calloc(1, sizeof *a);
    a->name = name; a->value = val; a->next = x->attrib; x->attrib = a;
    return val;
}
SEXP R_do_slot(SEXP obj, SEXP name)
{
    SEXP v = Rf_getAttrib(obj, name);
    if (v == R_NilValue) Rf_error("no slot of name \"%s\" for this object", (const char *)name->data);
    return v;
}
Rboolean Rf_isS4(SEXP x) { return x->s4 ? TRUE : FALSE; }
Rboolean Rf_inherits(SEXP x, const char *klass)
{
    SEXP c = Rf_getAttrib(x, sym_class);
    if (c == R_NilValue || c->type != STRSXP) return FALSE;
    for (R_xlen_t i = 0; i < c->len; ++i) if (!strcmp(CHAR(((SEXP *)c->data)[i]), klass)) return TRUE;
    return FALSE;
}
Rboolean Rf_isMatrix(SEXP x)
{
    SEXP d = Rf_getAttrib(x, R_DimSymbol);
    return (d != R_NilValue && d->len == 2) ? TRUE : FALSE;
}
int Rf_nrows(SEXP x)
{
    SEXP d = Rf_getAttrib(x, R_DimSymbol);
    if (d != R_NilValue) return INTEGER(d)[0];
    if (!elt_size(x->type)) Rf_error("object is not a matrix");
    return (int)x->len;
}
int Rf_ncols(SEXP x)
{
    SEXP d = Rf_getAttrib(x, R_DimSymbol);
    if (d != R_NilValue) return d->len >= 2 ? INTEGER(d)[1] : 1;
    if (!elt_size(x->type)) Rf_error("object is not a matrix");
    return 1;
}

/* ---- coercion ---------------------------------------------------------------------------------------------------- */
double Rf_asReal(SEXP x)
{
    if (x->len >= 1) {
        if (x->type == REALSXP) return ((double *)x->data)[0];
        if (x->type == INTSXP || x->type == LGLSXP) return (double)((int *)x->data)[0];
    }
    Rf_error("mock R: asReal() of a SEXP of type %u and length %ld", x->type, (long)x->len);
    return 0.0;
}
int Rf_asInteger(SEXP x)
{
    if (x->len >= 1) {
        if (x->type == REALSXP) return (int)((double *)x->data)[0];
        if (x->type == INTSXP || x->type == LGLSXP) return ((int *)x->data)[0];
    }
    Rf_error("mock R: asInteger() of a SEXP of type %u and length %ld", x->type, (long)x->len);
    return 0;
}
int Rf_asLogical(SEXP x)
{
    if (x->len >= 1) {
        if (x->type == REALSXP) return ((double *)x->data)[0] != 0.0;
        if (x->type == INTSXP || x->type == LGLSXP) return ((int *)x->data)[0] != 0;
    }
    Rf_error("mock R: asLogical() of a SEXP of type %u and length %ld", x->type, (long)x->len);
    return 0;
}
SEXP Rf_coerceVector(SEXP x, SEXPTYPE type)
{
    if (x->type == type) return x;
    if (type == REALSXP && (x->type == INTSXP || x->type == LGLSXP)) {
        SEXP s = Rf_allocVector(REALSXP, x->len);
        for (R_xlen_t i = 0; i < x->len; ++i) REAL(s)[i] = (double)((int *)x->data)[i];
        return s;
    }
    if (type == INTSXP && x->type == REALSXP) {
        SEXP s = Rf_allocVector(INTSXP, x->len);
        for (R_xlen_t i = 0; i < x->len; ++i) INTEGER(s)[i] = (int)((double *)x->data)[i];
        return s;
    }
    Rf_error("mock R: coerceVector(%u -> %u) is not modelled", x->type, type);
    return R_NilValue;
}

/* ---- external pointers --------------------------------------------------------------------------------------------- */
SEXP R_MakeExternalPtr(void *p, SEXP tag, SEXP prot)
{
    (void)tag; (void)prot;
    SEXP s = new_rec(EXTPTRSXP, 1, 0);
    s->ext = p;
    return s;
}
void *R_ExternalPtrAddr(SEXP s) { return s->type == EXTPTRSXP ? s->ext : NULL; }
void R_ClearExternalPtr(SEXP s) { s->ext = NULL; }
void R_RegisterCFinalizerEx(SEXP s, R_CFinalizer_t fun, Rboolean onexit) { (void)onexit; s->fin = fun; }

/* ---- registration ---------------------------------------------------------------------------------------------------- */
int R_registerRoutines(DllInfo *info, const void *c, const R_CallMethodDef *call, const void *f, const void *e)
{
    (void)info; (void)c; (void)f; (void)e;
    g_routines = call;
    return 1;
}
Rboolean R_useDynamicSymbols(DllInfo *info, Rboolean value) { (void)info; (void)value; return TRUE; }

/* ================================================================================================================
 * the test-side interface (tests/r_mock.py)
 * ================================================================================================================ */
void R_init_ebpmf_alpha(DllInfo *dll);

void mockr_init(void)
{
    if (R_DimSymbol) return;
    R_DimSymbol = Rf_install("dim");
    R_DimNamesSymbol = Rf_install("dimnames");
    sym_names = Rf_install("names");
    sym_class = Rf_install("class");
    R_init_ebpmf_alpha(NULL);                        /* what R does when it loads the package's shared object */
}

int mockr_n_routines(void)
{
    int n = 0;
    while (g_routines && g_routines[n].name) ++n;
    return n;
}
const char *mockr_routine_name(int i) { return g_routines[i].name; }
int mockr_routine_nargs(int i) { return g_routines[i].numArgs; }

typedef SEXP (*fn0)(void);
#define A(i) args[i]
/* .Call(name, ...): looked up in the glue's registration table, arity checked as R does, errors caught */
SEXP mockr_dot_call(const char *name, int nargs, SEXP *args)
{
    const R_CallMethodDef *r = g_routines;
    g_err[0] = '\0';
    while (r && r->name && strcmp(r->name, name)) ++r;
    if (!r || !r->name) { snprintf(g_err, sizeof g_err, "C symbol name \"%s\" not in the registration table", name); return NULL; }
    if (r->numArgs != nargs) {
        snprintf(g_err, sizeof g_err, "Incorrect number of arguments (%d), expecting %d for '%s'", nargs, r->numArgs, name);
        return NULL;
    }
    void *f = (void *)r->fun;
    SEXP out = NULL;
    g_protect = 0; g_max_protect = 0;
    g_in_call = 1;
    if (setjmp(g_jmp) == 0) {
        switch (nargs) {
        case 0: out = ((fn0)f)(); break;
        case 1: out = ((SEXP(*)(SEXP))f)(A(0)); break;
        case 2: out = ((SEXP(*)(SEXP, SEXP))f)(A(0), A(1)); break;
        case 3: out = ((SEXP(*)(SEXP, SEXP, SEXP))f)(A(0), A(1), A(2)); break;
        case 4: out = ((SEXP(*)(SEXP, SEXP, SEXP, SEXP))f)(A(0), A(1), A(2), A(3)); break;
        case 5: out = ((SEXP(*)(SEXP, SEXP, SEXP, SEXP, SEXP))f)(A(0), A(1), A(2), A(3), A(4)); break;
        case 6: out = ((SEXP(*)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP))f)(A(0), A(1), A(2), A(3), A(4), A(5)); break;
        case 7: out = ((SEXP(*)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP))f)(A(0), A(1), A(2), A(3), A(4), A(5), A(6)); break;
        case 8: out = ((SEXP(*)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP))f)(A(0), A(1), A(2), A(3), A(4), A(5), A(6), A(7)); break;
        case 9: out = ((SEXP(*)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP))f)(A(0), A(1), A(2), A(3), A(4), A(5), A(6), A(7), A(8)); break;
        case 10: out = ((SEXP(*)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP))f)(A(0), A(1), A(2), A(3), A(4), A(5), A(6), A(7), A(8), A(9)); break;
        case 11: out = ((SEXP(*)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP))f)(A(0), A(1), A(2), A(3), A(4), A(5), A(6), A(7), A(8), A(9), A(10)); break;
        case 12: out = ((SEXP(*)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP))f)(A(0), A(1), A(2), A(3), A(4), A(5), A(6), A(7), A(8), A(9), A(10), A(11)); break;
        case 13: out = ((SEXP(*)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP))f)(A(0), A(1), A(2), A(3), A(4), A(5), A(6), A(7), A(8), A(9), A(10), A(11), A(12)); break;
        case 14: out = ((SEXP(*)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP))f)(A(0), A(1), A(2), A(3), A(4), A(5), A(6), A(7), A(8), A(9), A(10), A(11), A(12), A(13)); break;
        default: snprintf(g_err, sizeof g_err, "mock R: .Call with %d arguments is not modelled", nargs); break;
        }
        if (out && g_protect != 0) {                   /* R: "stack imbalance in '.Call'" */
            snprintf(g_err, sizeof g_err, "stack imbalance in '.Call' of %s: %d item(s) left protected", name, g_protect);
            out = NULL;
        }
    } else {
        out = NULL;                                    /* Rf_error(): message in g_err */
    }
    g_in_call = 0;
    g_protect = 0;
    return out;
}
const char *mockr_last_error(void) { return g_err; }
int mockr_max_protect(void) { return g_max_protect; }
long mockr_live_objects(void) { return g_live; }

/* constructors / accessors for the Python side */
SEXP mockr_nil(void) { return R_NilValue; }
SEXP mockr_alloc(int type, long len) { return new_rec((SEXPTYPE)type, (R_xlen_t)len, elt_size((SEXPTYPE)type)); }
void *mockr_data(SEXP x) { return x->data; }
long mockr_length(SEXP x) { return (long)x->len; }
int mockr_type(SEXP x) { return (int)x->type; }
void mockr_set_attr(SEXP x, const char *name, SEXP v) { Rf_setAttrib(x, Rf_install(name), v); }
SEXP mockr_get_attr(SEXP x, const char *name) { return Rf_getAttrib(x, Rf_install(name)); }
void mockr_set_s4(SEXP x, int s4) { x->s4 = s4; }
SEXP mockr_new_s4(void) { SEXP s = new_rec(25 /* S4SXP */, 0, 0); s->s4 = 1; return s; }
SEXP mockr_vector_elt(SEXP x, long i) { return ((SEXP *)x->data)[i]; }
void mockr_set_vector_elt(SEXP x, long i, SEXP v) { ((SEXP *)x->data)[i] = v; }
const char *mockr_char(SEXP c) { return (const char *)c->data; }
int mockr_used_allocator(SEXP x) { return x->alloc_block != NULL; }

/* what the garbage collector would do with an unreachable object: finalizer, custom allocator, memory */
void mockr_free(SEXP x)
{
    if (!x || x == R_NilValue || x->type == SYMSXP) return;
    if (x->type == EXTPTRSXP && x->fin) x->fin(x);
    if (x->type == STRSXP) for (R_xlen_t i = 0; i < x->len; ++i) mockr_free(((SEXP *)x->data)[i]);
    if (x->alloc_block) x->alloc.mem_free(&x->alloc, x->alloc_block);
    else free(x->data);
    for (struct attr *a = x->attrib; a;) {             /* attribute values belong to the object (never shared here) */
        struct attr *n = a->next;
        mockr_free(a->value);
        free(a);
        a = n;
    }
    free(x);
    --g_live;
}
