/*
 * r_glue.c -- the thin `.Call` shim between R and the plain C ABI of include/ebpmf_b200.h.
 *
 * NOT COMPILED IN THIS REPOSITORY'S CI: R (R.h, Rinternals.h, libR) is not installed in the build
 * image or on the GPU box, so this file is compile-checked only where R exists
 * (R CMD INSTALL with src/Makevars).  It is deliberately tiny: every line of arithmetic lives
 * behind the C ABI, which IS tested (from Python/ctypes, tests/).  What is checked here without R:
 * tests/test_r_glue.py syntax-checks this file against declarations of the R API subset it uses
 * (tests/r_api_stub, never linked) and checks the .Call arities against R/vga_cuda.R.
 *
 * Conventions
 *   - arguments are never written (R shares them, copy-on-modify); outputs are fresh R objects, PROTECTed
 *     until return, with dim copied from M.  Result MATRICES of 64 MB and more are allocated with
 *     allocVector3() and an R_allocator_t backed by the library's warm host pool (ebpmf_host_alloc /
 *     ebpmf_host_free): ordinary R numeric matrices in every respect, but their pages are already mapped
 *     when the result is copied out, and R's garbage collector hands them back to the pool;
 *   - errors: the C ABI returns a status; we copy the message to a stack buffer and call
 *     Rf_error() (a longjmp) only after nothing needs freeing;
 *   - one context per R session lives in an external pointer with a finalizer.
 *
 * Replaces (paths relative to the reference root):
 *   R/vga_solver_mat.R:6-41, :44-65, :71-109 ; R/ebpmf_log.R:302-332, :406-410, :413-447, :470-472
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <R_ext/Rallocators.h>
#include <stdlib.h>
#include <string.h>

#include "ebpmf_b200.h"

/* ---- helpers ---------------------------------------------------------------------------- */
static ebpmf_ctx *get_ctx(SEXP xp)
{
    if (TYPEOF(xp) != EXTPTRSXP || R_ExternalPtrAddr(xp) == NULL)
        Rf_error("ebpmf_b200: invalid or closed GPU context");
    return (ebpmf_ctx *)R_ExternalPtrAddr(xp);
}

static void ctx_finalizer(SEXP xp)
{
    ebpmf_ctx *ctx = (ebpmf_ctx *)R_ExternalPtrAddr(xp);
    if (ctx) { ebpmf_ctx_destroy(ctx); R_ClearExternalPtr(xp); }
}

static void check(ebpmf_ctx *ctx, int rc)
{
    if (rc == EBPMF_OK) return;
    char msg[1024];
    strncpy(msg, ebpmf_last_error(ctx), sizeof msg - 1);
    msg[sizeof msg - 1] = '\0';
    if (rc == EBPMF_ERR_NAN) Rf_error("missing value where TRUE/FALSE needed");   /* R/vga_solver_mat.R:26 */
    if (rc == EBPMF_ERR_VAR_TYPE) Rf_error("Non-supported var type");               /* R/ebpmf_log.R:444,483 */
    Rf_error("ebpmf_b200: %s", msg);
}

static int var_type_code(SEXP vt)
{
    const char *s = CHAR(STRING_ELT(vt, 0));
    if (!strcmp(s, "constant")) return EBPMF_VAR_CONSTANT;
    if (!strcmp(s, "by_row")) return EBPMF_VAR_BY_ROW;
    if (!strcmp(s, "by_col")) return EBPMF_VAR_BY_COL;
    return -1;
}

/* scalar / length-n vector / n x p matrix, as R would recycle it against an n x p matrix */
static int operand_kind(SEXP a, R_xlen_t n, R_xlen_t p, const char *what)
{
    R_xlen_t len = XLENGTH(a);
    if (len == 1) return EBPMF_OP_SCALAR;
    if (len == n * p) return EBPMF_OP_MATRIX;
    if (len == n) return EBPMF_OP_ROWVEC;
    Rf_error("%s: non-conformable arrays", what);
    return -1;
}

static int is_dgC(SEXP Y) { return Rf_isS4(Y) && Rf_inherits(Y, "dgCMatrix"); }

/* R_allocator_t over the library's host pool.  R asks for `size` bytes (vector header + data), keeps the
 * allocator alongside and calls mem_free from the garbage collector. */
static void *pool_alloc(R_allocator_t *allocator, size_t size) { (void)allocator; return ebpmf_host_alloc(size); }
static void pool_free(R_allocator_t *allocator, void *ptr) { (void)allocator; (void)ebpmf_host_free(ptr); }

/* a fresh n x p numeric matrix: from the pool when it is large, from R's heap otherwise */
static SEXP alloc_result_matrix(R_xlen_t n, R_xlen_t p)
{
    const R_xlen_t len = n * p;
    SEXP out;
    if ((double)len * 8.0 >= 64.0 * 1048576.0) {
        R_allocator_t al;
        al.mem_alloc = pool_alloc; al.mem_free = pool_free; al.res = NULL; al.data = NULL;
        out = PROTECT(Rf_allocVector3(REALSXP, len, &al));
    } else {
        out = PROTECT(Rf_allocVector(REALSXP, len));
    }
    SEXP dim = PROTECT(Rf_allocVector(INTSXP, 2));
    INTEGER(dim)[0] = (int)n; INTEGER(dim)[1] = (int)p;
    Rf_setAttrib(out, R_DimSymbol, dim);
    UNPROTECT(2);
    return out;
}

static SEXP alloc_like(SEXP M)
{
    SEXP dim = Rf_getAttrib(M, R_DimSymbol);
    SEXP out = PROTECT(alloc_result_matrix(INTEGER(dim)[0], INTEGER(dim)[1]));
    SEXP dn = Rf_getAttrib(M, R_DimNamesSymbol);
    if (dn != R_NilValue) Rf_setAttrib(out, R_DimNamesSymbol, dn);
    UNPROTECT(1);
    return out;
}

static SEXP info_list(const ebpmf_solve_info *info, SEXP M, SEXP V, SEXP v_sum)
{
    const char *names[] = {"M", "V", "v_sum", "sym", "ssexp", "slogv", "n_updates", "converged", ""};
    SEXP out = PROTECT(Rf_mkNamed(VECSXP, names));
    SET_VECTOR_ELT(out, 0, M);
    SET_VECTOR_ELT(out, 1, V);
    SET_VECTOR_ELT(out, 2, v_sum);
    SET_VECTOR_ELT(out, 3, Rf_ScalarReal(info->sym));
    SET_VECTOR_ELT(out, 4, Rf_ScalarReal(info->ssexp));
    SET_VECTOR_ELT(out, 5, Rf_ScalarReal(info->slogv));
    SET_VECTOR_ELT(out, 6, Rf_ScalarInteger(info->n_updates));
    SET_VECTOR_ELT(out, 7, Rf_ScalarLogical(info->converged));
    UNPROTECT(1);
    return out;
}

/* ---- context ----------------------------------------------------------------------------- */
SEXP C_ebpmf_ctx_new(SEXP device)
{
    ebpmf_ctx *ctx = NULL;
    int rc = ebpmf_ctx_create(Rf_asInteger(device), &ctx);
    if (rc != EBPMF_OK) Rf_error("ebpmf_b200: %s", ebpmf_last_error(NULL));   /* no CPU fallback */
    SEXP xp = PROTECT(R_MakeExternalPtr(ctx, R_NilValue, R_NilValue));
    R_RegisterCFinalizerEx(xp, ctx_finalizer, TRUE);
    UNPROTECT(1);
    return xp;
}

/* Y: numeric matrix (as.matrix(Y), R/ebpmf_log.R:202) or dgCMatrix (slots p, i, x, Dim) */
SEXP C_ebpmf_ctx_set_y(SEXP xp, SEXP Y)
{
    ebpmf_ctx *ctx = get_ctx(xp);
    if (is_dgC(Y)) {
        SEXP Dim = R_do_slot(Y, Rf_install("Dim"));
        check(ctx, ebpmf_ctx_set_y_csc(ctx, INTEGER(Dim)[0], INTEGER(Dim)[1], INTEGER(R_do_slot(Y, Rf_install("p"))),
                                       INTEGER(R_do_slot(Y, Rf_install("i"))), REAL(R_do_slot(Y, Rf_install("x")))));
    } else {
        if (!Rf_isMatrix(Y) || TYPEOF(Y) != REALSXP) Rf_error("Y must be a numeric matrix or a dgCMatrix");
        check(ctx, ebpmf_ctx_set_y_dense(ctx, Rf_nrows(Y), Rf_ncols(Y), REAL(Y)));
    }
    return R_NilValue;
}

/* const = -sum_lfactorial_sparseMat(Y)  R/ebpmf_log.R:172-176, :470-472 */
SEXP C_ebpmf_sum_lfactorial(SEXP xp, SEXP x)
{
    ebpmf_ctx *ctx = get_ctx(xp);
    double out = 0.0;
    if (x == R_NilValue) check(ctx, ebpmf_ctx_sum_lfactorial(ctx, &out));          /* resident Y */
    else check(ctx, ebpmf_sum_lfactorial(ctx, (int64_t)XLENGTH(x), REAL(x), &out));
    return Rf_ScalarReal(out);
}

/* ---- the fused VGA step of the outer loop, R/ebpmf_log.R:302-327 ----------------------------
 * list(M, V|NULL, v_sum, sym, ssexp, slogv, n_updates, converged) */
SEXP C_ebpmf_vga_step(SEXP xp, SEXP M, SEXP EL, SEXP EF, SEXP sigma2, SEXP var_type, SEXP maxiter, SEXP tol,
                      SEXP want_V)
{
    ebpmf_ctx *ctx = get_ctx(xp);
    const int vt = var_type_code(var_type);
    if (vt < 0) Rf_error("Non-supported var type");
    const R_xlen_t n = Rf_nrows(M), p = Rf_ncols(M);
    if (Rf_nrows(EL) != n || Rf_nrows(EF) != p || Rf_ncols(EL) != Rf_ncols(EF)) Rf_error("non-conformable arguments");
    const R_xlen_t want = vt == EBPMF_VAR_BY_COL ? p : vt == EBPMF_VAR_BY_ROW ? n : 1;
    if (XLENGTH(sigma2) != want) Rf_error("sigma2 has the wrong length for this var_type");
    const int wv = Rf_asLogical(want_V);
    SEXP Mo = PROTECT(alloc_like(M));
    SEXP Vo = PROTECT(wv ? alloc_like(M) : R_NilValue);
    SEXP vs = PROTECT(Rf_allocVector(REALSXP, want));
    ebpmf_solve_info info;
    int rc = ebpmf_vga_step_host(ctx, (int64_t)n, (int64_t)p, REAL(M), Rf_ncols(EL), REAL(EL), REAL(EF), vt, REAL(sigma2),
                                 Rf_asInteger(maxiter), Rf_asReal(tol), REAL(Mo), wv ? REAL(Vo) : NULL, REAL(vs), &info);
    if (rc != EBPMF_OK) { UNPROTECT(3); check(ctx, rc); }
    SEXP out = info_list(&info, Mo, Vo, vs);
    UNPROTECT(3);
    return out;
}

/* M_init -> device, once per ebpmf_log() call (the latent M then stays resident: every later step's
 * input is the previous step's output, R/ebpmf_log.R:309) */
SEXP C_ebpmf_ctx_set_m(SEXP xp, SEXP M)
{
    ebpmf_ctx *ctx = get_ctx(xp);
    if (!Rf_isMatrix(M) || TYPEOF(M) != REALSXP) Rf_error("M must be a numeric matrix");
    check(ctx, ebpmf_ctx_set_m(ctx, REAL(M)));
    return R_NilValue;
}

/* The per-iteration VGA step on the resident M: factors and sigma2 in, list(M, V|NULL, v_sum, sym,
 * ssexp, slogv, n_updates, converged) out.  dims = c(n, p) of the resident matrices. */
SEXP C_ebpmf_vga_step_resident(SEXP xp, SEXP dims, SEXP EL, SEXP EF, SEXP sigma2, SEXP var_type, SEXP maxiter,
                               SEXP tol, SEXP want_V)
{
    ebpmf_ctx *ctx = get_ctx(xp);
    const int vt = var_type_code(var_type);
    if (vt < 0) Rf_error("Non-supported var type");
    const int n = INTEGER(dims)[0], p = INTEGER(dims)[1];
    if (Rf_nrows(EL) != n || Rf_nrows(EF) != p || Rf_ncols(EL) != Rf_ncols(EF)) Rf_error("non-conformable arguments");
    const R_xlen_t want = vt == EBPMF_VAR_BY_COL ? p : vt == EBPMF_VAR_BY_ROW ? n : 1;
    if (XLENGTH(sigma2) != want) Rf_error("sigma2 has the wrong length for this var_type");
    const int wv = Rf_asLogical(want_V);
    SEXP Mo = PROTECT(alloc_result_matrix(n, p));
    SEXP Vo = PROTECT(wv ? alloc_result_matrix(n, p) : R_NilValue);
    SEXP vs = PROTECT(Rf_allocVector(REALSXP, want));
    ebpmf_solve_info info;
    int rc = ebpmf_vga_step_resident_host(ctx, (int64_t)n, (int64_t)p, Rf_ncols(EL), REAL(EL), REAL(EF), vt, REAL(sigma2), Rf_asInteger(maxiter),
                                          Rf_asReal(tol), REAL(Mo), wv ? REAL(Vo) : NULL, REAL(vs), &info);
    if (rc != EBPMF_OK) { UNPROTECT(3); check(ctx, rc); }
    SEXP out = info_list(&info, Mo, Vo, vs);
    UNPROTECT(3);
    return out;
}

/* ---- vga_pois_solver_mat_newton(M,X,S,Beta,Sigma2,maxiter,tol,return_V)  R/vga_solver_mat.R:6 */
SEXP C_vga_pois_solver_mat_newton(SEXP xp, SEXP M, SEXP X, SEXP S, SEXP Beta, SEXP Sigma2, SEXP maxiter, SEXP tol,
                                  SEXP return_V)
{
    ebpmf_ctx *ctx = get_ctx(xp);
    const R_xlen_t n = Rf_nrows(M), p = Rf_ncols(M);
    const int rv = Rf_asLogical(return_V);
    const int ks = operand_kind(S, n, p, "S"), kb = operand_kind(Beta, n, p, "Beta"),
              kv = operand_kind(Sigma2, n, p, "Sigma2");
    SEXP Mo = PROTECT(alloc_like(M));
    SEXP Vo = PROTECT(rv ? alloc_like(M) : R_NilValue);
    ebpmf_solve_info info;
    int rc;
    if (is_dgC(X))
        rc = ebpmf_vga_newton(ctx, n, p, REAL(M), NULL, INTEGER(R_do_slot(X, Rf_install("p"))),
                              INTEGER(R_do_slot(X, Rf_install("i"))), REAL(R_do_slot(X, Rf_install("x"))), REAL(S), ks,
                              REAL(Beta), kb, REAL(Sigma2), kv, Rf_asInteger(maxiter), Rf_asReal(tol), REAL(Mo),
                              rv ? REAL(Vo) : NULL, &info);
    else
        rc = ebpmf_vga_newton(ctx, n, p, REAL(M), REAL(X), NULL, NULL, NULL, REAL(S), ks, REAL(Beta), kb, REAL(Sigma2),
                              kv, Rf_asInteger(maxiter), Rf_asReal(tol), REAL(Mo), rv ? REAL(Vo) : NULL, &info);
    if (rc != EBPMF_OK) { UNPROTECT(2); check(ctx, rc); }
    SEXP out;
    if (rv) {                                                   /* list(M=M, V=Sigma2/temp)  :36 */
        const char *names[] = {"M", "V", ""};
        out = PROTECT(Rf_mkNamed(VECSXP, names));
        SET_VECTOR_ELT(out, 0, Mo);
        SET_VECTOR_ELT(out, 1, Vo);
        UNPROTECT(1);
    } else {
        out = Mo;                                               /* return(M)  :38 */
    }
    UNPROTECT(2);
    return out;
}

/* ---- vga_pois_solver_mat_newton_fixed_iter(M,V,Y,S,Beta,Sigma2,maxiter)  R/vga_solver_mat.R:44 */
SEXP C_vga_pois_solver_mat_newton_fixed_iter(SEXP xp, SEXP M, SEXP V, SEXP Y, SEXP S, SEXP Beta, SEXP Sigma2,
                                             SEXP maxiter)
{
    ebpmf_ctx *ctx = get_ctx(xp);
    const R_xlen_t n = Rf_nrows(M), p = Rf_ncols(M);
    const int ks = operand_kind(S, n, p, "S"), kb = operand_kind(Beta, n, p, "Beta"),
              kv = operand_kind(Sigma2, n, p, "Sigma2");
    SEXP Mo = PROTECT(alloc_like(M));
    SEXP Vo = PROTECT(alloc_like(M));
    int rc = ebpmf_vga_fixed_iter(ctx, n, p, REAL(M), V == R_NilValue ? NULL : REAL(V), REAL(Y), REAL(S), ks, REAL(Beta),
                                  kb, REAL(Sigma2), kv, Rf_asInteger(maxiter), REAL(Mo), REAL(Vo));
    if (rc != EBPMF_OK) { UNPROTECT(2); check(ctx, rc); }
    const char *names[] = {"M", "V", ""};
    SEXP out = PROTECT(Rf_mkNamed(VECSXP, names));
    SET_VECTOR_ELT(out, 0, Mo);
    SET_VECTOR_ELT(out, 1, Vo);
    UNPROTECT(3);
    return out;
}

/* ---- vga_pois_solver_mat_newton_low_memory(M,Y,l0,f0,EL,EF,sigma2,var_type,maxiter,tol)  :71 */
SEXP C_vga_pois_solver_mat_newton_low_memory(SEXP xp, SEXP M, SEXP Y, SEXP l0, SEXP f0, SEXP EL, SEXP EF, SEXP sigma2,
                                             SEXP var_type, SEXP maxiter, SEXP tol)
{
    ebpmf_ctx *ctx = get_ctx(xp);
    const int vt = var_type_code(var_type);
    if (vt < 0) Rf_error("Non-supported var type");
    const R_xlen_t n = Rf_nrows(M), p = Rf_ncols(M);
    SEXP Mo = PROTECT(alloc_like(M));
    ebpmf_solve_info info;
    int rc;
    if (is_dgC(Y))
        rc = ebpmf_vga_low_memory(ctx, n, p, Rf_ncols(EL), REAL(M), NULL, INTEGER(R_do_slot(Y, Rf_install("p"))),
                                  INTEGER(R_do_slot(Y, Rf_install("i"))), REAL(R_do_slot(Y, Rf_install("x"))), REAL(l0),
                                  REAL(f0), REAL(EL), REAL(EF), REAL(sigma2), vt, Rf_asInteger(maxiter), Rf_asReal(tol),
                                  REAL(Mo), &info);
    else
        rc = ebpmf_vga_low_memory(ctx, n, p, Rf_ncols(EL), REAL(M), REAL(Y), NULL, NULL, NULL, REAL(l0), REAL(f0),
                                  REAL(EL), REAL(EF), REAL(sigma2), vt, Rf_asInteger(maxiter), Rf_asReal(tol), REAL(Mo),
                                  &info);
    if (rc != EBPMF_OK) { UNPROTECT(1); check(ctx, rc); }
    UNPROTECT(1);
    return Mo;                                                  /* return(as.matrix(M))  :108 */
}

/* ---- ebpmf_log_update_sigma2(fit_flash,sigma2,v_sum,var_type,cap,a0,b0,n,p)  R/ebpmf_log.R:413
 * The R wrapper passes R2 = fit_flash$flash_fit$R2 (summed for 'constant') and, when cap > 0,
 * EL = flash_fit$EF[[1]], EF = flash_fit$EF[[2]] (else NULL). */
SEXP C_ebpmf_log_update_sigma2(SEXP xp, SEXP sigma2, SEXP v_sum, SEXP R2, SEXP var_type, SEXP cap, SEXP a0, SEXP b0,
                               SEXP n, SEXP p, SEXP EL, SEXP EF)
{
    ebpmf_ctx *ctx = get_ctx(xp);
    const int vt = var_type_code(var_type);
    if (vt < 0) Rf_error("Non-supported var type");
    SEXP out = PROTECT(Rf_allocVector(REALSXP, XLENGTH(sigma2)));
    const int have_f = EL != R_NilValue && EF != R_NilValue;
    int rc = ebpmf_update_sigma2(ctx, vt, (int64_t)Rf_asReal(n), (int64_t)Rf_asReal(p), REAL(sigma2), REAL(v_sum),
                                 REAL(R2), Rf_asReal(a0), Rf_asReal(b0), Rf_asReal(cap), have_f ? Rf_ncols(EL) : 0,
                                 have_f ? REAL(EL) : NULL, have_f ? REAL(EF) : NULL, REAL(out));
    if (rc != EBPMF_OK) { UNPROTECT(1); check(ctx, rc); }
    UNPROTECT(1);
    return out;
}

/* ---- calc_ebpmf_log_obj(n,p,sym,ssexp,slogv,sv,sigma2,R2,KL_LF,const,ss,a0,b0)  R/ebpmf_log.R:406
 * sv / sigma2 / R2 keep their own lengths: R takes each sum() of :407 at the length of its own operands. */
SEXP C_calc_ebpmf_log_obj(SEXP xp, SEXP n, SEXP p, SEXP sym, SEXP ssexp, SEXP slogv, SEXP sv, SEXP sigma2, SEXP R2,
                          SEXP KL_LF, SEXP const_, SEXP ss, SEXP a0, SEXP b0)
{
    ebpmf_ctx *ctx = get_ctx(xp);
    double obj = 0.0;
    check(ctx, ebpmf_calc_obj(ctx, (int64_t)Rf_asReal(n), (int64_t)Rf_asReal(p), Rf_asReal(sym), Rf_asReal(ssexp),
                              Rf_asReal(slogv), (int64_t)XLENGTH(sv), REAL(sv), (int64_t)XLENGTH(sigma2), REAL(sigma2),
                              (int64_t)XLENGTH(R2), REAL(R2), Rf_asReal(KL_LF), Rf_asReal(const_), Rf_asReal(ss),
                              Rf_asReal(a0), Rf_asReal(b0), &obj));
    return Rf_ScalarReal(obj);
}

/* ---- calc_split_PMF_obj_flashier(Y,S,sigma2,M,V,fit_flash,KL_LF,const,var_type)  inst/code/Poisson_MF_splitting.R:373
 * (the archived driver's un-fused ELBO).  The R wrapper passes R2 = fit_flash$flash.fit$R2. */
SEXP C_calc_split_PMF_obj_flashier(SEXP xp, SEXP Y, SEXP S, SEXP sigma2, SEXP M, SEXP V, SEXP R2, SEXP KL_LF, SEXP const_,
                                   SEXP var_type)
{
    ebpmf_ctx *ctx = get_ctx(xp);
    const int vt = var_type_code(var_type);
    if (vt < 0) Rf_error("Non-supported var type");
    const R_xlen_t n = Rf_nrows(M), p = Rf_ncols(M);
    if (Rf_nrows(Y) != n || Rf_ncols(Y) != p || Rf_nrows(V) != n || Rf_ncols(V) != p) Rf_error("non-conformable arrays");
    const int ks = operand_kind(S, n, p, "S");
    double obj = 0.0;
    check(ctx, ebpmf_calc_split_obj(ctx, (int64_t)n, (int64_t)p, REAL(Y), REAL(S), ks, REAL(M), REAL(V), vt, (int64_t)XLENGTH(sigma2),
                                    REAL(sigma2), (int64_t)XLENGTH(R2), REAL(R2), Rf_asReal(KL_LF), Rf_asReal(const_), &obj));
    return Rf_ScalarReal(obj);
}

/* ---- N1: what flashier needs from the RESIDENT M (R/ebpmf_log_flash_update.R:35,55-64) ---------
 * list(MF = M %*% EF_w, MtL = t(M) %*% EL_w, r2_col = colSums((M - EL_w EF_w')^2), r2_row = rowSums(...)) */
SEXP C_ebpmf_m_products(SEXP xp, SEXP dims, SEXP EL_w, SEXP EF_w)
{
    ebpmf_ctx *ctx = get_ctx(xp);
    const int n = INTEGER(dims)[0], p = INTEGER(dims)[1];
    const int Kw = Rf_ncols(EL_w);
    if (Rf_nrows(EL_w) != n || Rf_nrows(EF_w) != p || Rf_ncols(EF_w) != Kw) Rf_error("non-conformable arguments");
    SEXP MF = PROTECT(Rf_allocMatrix(REALSXP, n, Kw));
    SEXP MtL = PROTECT(Rf_allocMatrix(REALSXP, p, Kw));
    SEXP rc_ = PROTECT(Rf_allocVector(REALSXP, p));
    SEXP rr_ = PROTECT(Rf_allocVector(REALSXP, n));
    int rc = ebpmf_ctx_m_products(ctx, Kw, REAL(EL_w), REAL(EF_w), REAL(MF), REAL(MtL), REAL(rc_), REAL(rr_), NULL);
    if (rc != EBPMF_OK) { UNPROTECT(4); check(ctx, rc); }
    const char *names[] = {"MF", "MtL", "r2_col", "r2_row", ""};
    SEXP out = PROTECT(Rf_mkNamed(VECSXP, names));
    SET_VECTOR_ELT(out, 0, MF);
    SET_VECTOR_ELT(out, 1, MtL);
    SET_VECTOR_ELT(out, 2, rc_);
    SET_VECTOR_ELT(out, 3, rr_);
    UNPROTECT(5);
    return out;
}

/* ---- N4: on-device sim_data_log (R/simulate_poisson_matrix.R:9-61) into the context's resident state ------
 * returns list(i, p, x, Dim, EL, EF, sigma2): the dgCMatrix slots of Y and the solver state that was generated;
 * the start M stays on the device (it is the resident M). */
SEXP C_ebpmf_sim_data_log(SEXP xp, SEXP n_, SEXP p_, SEXP K_, SEXP seed, SEXP log_offset, SEXP bg_range, SEXP pi0,
                          SEXP noise, SEXP warm_start)
{
    ebpmf_ctx *ctx = get_ctx(xp);
    const int64_t n = (int64_t)Rf_asReal(n_), p = (int64_t)Rf_asReal(p_), K = (int64_t)Rf_asReal(K_);
    int64_t nnz = 0;
    check(ctx, ebpmf_ctx_sim_data_log(ctx, n, p, K, 0, (uint64_t)Rf_asReal(seed), Rf_asReal(log_offset), REAL(bg_range)[0],
                                      REAL(bg_range)[1], Rf_asReal(pi0), Rf_asReal(noise), Rf_asLogical(warm_start), &nnz));
    SEXP si = PROTECT(Rf_allocVector(INTSXP, nnz > 0 ? nnz : 1));
    SEXP sp = PROTECT(Rf_allocVector(INTSXP, p + 1));
    SEXP sx = PROTECT(Rf_allocVector(REALSXP, nnz > 0 ? nnz : 1));
    SEXP EL = PROTECT(Rf_allocMatrix(REALSXP, (int)n, (int)K + 2));
    SEXP EF = PROTECT(Rf_allocMatrix(REALSXP, (int)p, (int)K + 2));
    SEXP s2 = PROTECT(Rf_allocVector(REALSXP, p));
    int rc = ebpmf_ctx_get_y_csc(ctx, INTEGER(sp), INTEGER(si), REAL(sx));
    if (rc == EBPMF_OK) rc = ebpmf_ctx_get_factors(ctx, REAL(EL), REAL(EF));
    if (rc == EBPMF_OK) rc = ebpmf_ctx_get_sigma2(ctx, REAL(s2));
    if (rc != EBPMF_OK) { UNPROTECT(6); check(ctx, rc); }
    SEXP Dim = PROTECT(Rf_allocVector(INTSXP, 2));
    INTEGER(Dim)[0] = (int)n; INTEGER(Dim)[1] = (int)p;
    const char *names[] = {"i", "p", "x", "Dim", "EL", "EF", "sigma2", "nnz", ""};
    SEXP out = PROTECT(Rf_mkNamed(VECSXP, names));
    SET_VECTOR_ELT(out, 0, si);
    SET_VECTOR_ELT(out, 1, sp);
    SET_VECTOR_ELT(out, 2, sx);
    SET_VECTOR_ELT(out, 3, Dim);
    SET_VECTOR_ELT(out, 4, EL);
    SET_VECTOR_ELT(out, 5, EF);
    SET_VECTOR_ELT(out, 6, s2);
    SET_VECTOR_ELT(out, 7, Rf_ScalarReal((double)nnz));
    UNPROTECT(8);
    return out;
}

/* ---- one R session, several GPUs: rows sharded over the devices of an ebpmf_group ---------------------
 * (north star: "Rows are sharded across the GPUs of one 8xB200 box").  vga_cuda.R selects this path when
 * EBPMF_B200_NGPU > 1; the calls mirror the single-GPU ones with the FULL host matrices. */
static void group_finalizer(SEXP xp)
{
    ebpmf_group *g = (ebpmf_group *)R_ExternalPtrAddr(xp);
    if (g) { ebpmf_group_destroy(g); R_ClearExternalPtr(xp); }
}

static ebpmf_group *get_group(SEXP xp)
{
    if (TYPEOF(xp) != EXTPTRSXP || R_ExternalPtrAddr(xp) == NULL)
        Rf_error("ebpmf_b200: invalid or closed multi-GPU group");
    return (ebpmf_group *)R_ExternalPtrAddr(xp);
}

static void check_group(ebpmf_group *g, int rc)
{
    if (rc == EBPMF_OK) return;
    char msg[1024];
    strncpy(msg, ebpmf_group_last_error(g), sizeof msg - 1);
    msg[sizeof msg - 1] = '\0';
    if (rc == EBPMF_ERR_NAN) Rf_error("missing value where TRUE/FALSE needed");   /* R/vga_solver_mat.R:26 */
    if (rc == EBPMF_ERR_VAR_TYPE) Rf_error("Non-supported var type");
    Rf_error("ebpmf_b200: %s", msg);
}

/* devices: integer vector of device indices (length = number of GPUs) */
SEXP C_ebpmf_group_new(SEXP devices)
{
    ebpmf_group *g = NULL;
    int rc = ebpmf_group_create((int)XLENGTH(devices), INTEGER(devices), &g);
    if (rc != EBPMF_OK) Rf_error("ebpmf_b200: %s", ebpmf_group_last_error(NULL));
    SEXP xp = PROTECT(R_MakeExternalPtr(g, R_NilValue, R_NilValue));
    R_RegisterCFinalizerEx(xp, group_finalizer, TRUE);
    UNPROTECT(1);
    return xp;
}

SEXP C_ebpmf_group_set_y(SEXP xp, SEXP Y)
{
    ebpmf_group *g = get_group(xp);
    if (!is_dgC(Y)) Rf_error("Y must be a dgCMatrix");
    SEXP Dim = R_do_slot(Y, Rf_install("Dim"));
    check_group(g, ebpmf_group_set_y_csc(g, INTEGER(Dim)[0], INTEGER(Dim)[1], INTEGER(R_do_slot(Y, Rf_install("p"))),
                                         INTEGER(R_do_slot(Y, Rf_install("i"))), REAL(R_do_slot(Y, Rf_install("x")))));
    return R_NilValue;
}

SEXP C_ebpmf_group_sum_lfactorial(SEXP xp)
{
    ebpmf_group *g = get_group(xp);
    double out = 0.0;
    check_group(g, ebpmf_group_sum_lfactorial(g, &out));
    return Rf_ScalarReal(out);
}

SEXP C_ebpmf_group_set_m(SEXP xp, SEXP M)
{
    ebpmf_group *g = get_group(xp);
    if (!Rf_isMatrix(M) || TYPEOF(M) != REALSXP) Rf_error("M must be a numeric matrix");
    check_group(g, ebpmf_group_set_m(g, Rf_nrows(M), Rf_ncols(M), REAL(M)));
    return R_NilValue;
}

/* M: the n x p input matrix, or NULL when the shards already hold it (C_ebpmf_group_set_m once, then every
 * step's output); dims = c(n, p).  Same result list as C_ebpmf_vga_step. */
SEXP C_ebpmf_group_vga_step(SEXP xp, SEXP M, SEXP dims, SEXP EL, SEXP EF, SEXP sigma2, SEXP var_type, SEXP maxiter,
                            SEXP tol, SEXP want_V)
{
    ebpmf_group *g = get_group(xp);
    const int vt = var_type_code(var_type);
    if (vt < 0) Rf_error("Non-supported var type");
    const int n = INTEGER(dims)[0], p = INTEGER(dims)[1];
    if (M != R_NilValue && (Rf_nrows(M) != n || Rf_ncols(M) != p)) Rf_error("non-conformable arrays");
    if (Rf_nrows(EL) != n || Rf_nrows(EF) != p || Rf_ncols(EL) != Rf_ncols(EF)) Rf_error("non-conformable arguments");
    const R_xlen_t want = vt == EBPMF_VAR_BY_COL ? p : vt == EBPMF_VAR_BY_ROW ? n : 1;
    if (XLENGTH(sigma2) != want) Rf_error("sigma2 has the wrong length for this var_type");
    const int wv = Rf_asLogical(want_V);
    SEXP Mo = PROTECT(alloc_result_matrix(n, p));
    SEXP Vo = PROTECT(wv ? alloc_result_matrix(n, p) : R_NilValue);
    SEXP vs = PROTECT(Rf_allocVector(REALSXP, want));
    ebpmf_solve_info info;
    int rc = ebpmf_group_vga_step_host(g, (int64_t)n, (int64_t)p, M == R_NilValue ? NULL : REAL(M), Rf_ncols(EL), REAL(EL),
                                       REAL(EF), vt, REAL(sigma2), Rf_asInteger(maxiter), Rf_asReal(tol), REAL(Mo),
                                       wv ? REAL(Vo) : NULL, REAL(vs), &info);
    if (rc != EBPMF_OK) { UNPROTECT(3); check_group(g, rc); }
    SEXP out = info_list(&info, Mo, Vo, vs);
    UNPROTECT(3);
    return out;
}

/* ---- registration ------------------------------------------------------------------------ */
/* ---- the batched branch of the outer loop, R/ebpmf_log.R:195-203, :254-300 --------------------
 * C_ebpmf_batched_new(Y dgCMatrix, batch_row0 integer/double vector of 0-based batch starts, device, nslots):
 * external pointer; C_ebpmf_batched_vga_step(...): list(M, V|NULL, v_sum, sym, ssexp, slogv, n_updates
 * (max over batches), converged (all batches)) with attribute "n_updates_per_batch". */
static void batched_finalizer(SEXP xp)
{
    ebpmf_batched *h = (ebpmf_batched *)R_ExternalPtrAddr(xp);
    if (h) { ebpmf_batched_destroy(h); R_ClearExternalPtr(xp); }
}

static ebpmf_batched *get_batched(SEXP xp)
{
    if (TYPEOF(xp) != EXTPTRSXP || R_ExternalPtrAddr(xp) == NULL)
        Rf_error("ebpmf_b200: invalid or closed batched GPU handle");
    return (ebpmf_batched *)R_ExternalPtrAddr(xp);
}

static void check_batched(ebpmf_batched *h, int rc)
{
    if (rc == EBPMF_OK) return;
    char msg[1024];
    strncpy(msg, ebpmf_batched_last_error(h), sizeof msg - 1);
    msg[sizeof msg - 1] = '\0';
    if (rc == EBPMF_ERR_NAN) Rf_error("missing value where TRUE/FALSE needed");   /* R/vga_solver_mat.R:26 */
    if (rc == EBPMF_ERR_VAR_TYPE) Rf_error("Non-supported var type");
    Rf_error("ebpmf_b200: %s", msg);
}

SEXP C_ebpmf_batched_new(SEXP Y, SEXP batch_row0, SEXP device, SEXP nslots)
{
    if (!is_dgC(Y)) Rf_error("Y must be a dgCMatrix");
    ebpmf_batched *h = NULL;
    int rc = ebpmf_batched_create(Rf_asInteger(device), Rf_asInteger(nslots), &h);
    if (rc != EBPMF_OK) Rf_error("ebpmf_b200: %s", ebpmf_batched_last_error(NULL));
    SEXP xp = PROTECT(R_MakeExternalPtr(h, R_NilValue, R_NilValue));
    R_RegisterCFinalizerEx(xp, batched_finalizer, TRUE);
    const R_xlen_t nb1 = XLENGTH(batch_row0);
    int64_t *row0 = (int64_t *)R_alloc((size_t)nb1, sizeof(int64_t));
    SEXP r0 = PROTECT(Rf_coerceVector(batch_row0, REALSXP));
    for (R_xlen_t b = 0; b < nb1; ++b) row0[b] = (int64_t)REAL(r0)[b];
    SEXP Dim = R_do_slot(Y, Rf_install("Dim"));
    rc = ebpmf_batched_set_y_csc(h, INTEGER(Dim)[0], INTEGER(Dim)[1], INTEGER(R_do_slot(Y, Rf_install("p"))),
                                 INTEGER(R_do_slot(Y, Rf_install("i"))), REAL(R_do_slot(Y, Rf_install("x"))),
                                 (int64_t)nb1 - 1, row0);
    UNPROTECT(2);
    check_batched(h, rc);
    return xp;
}

SEXP C_ebpmf_batched_vga_step(SEXP xp, SEXP M, SEXP EL, SEXP EF, SEXP sigma2, SEXP var_type, SEXP maxiter, SEXP tol,
                              SEXP want_V, SEXP nbatch)
{
    ebpmf_batched *h = get_batched(xp);
    const int vt = var_type_code(var_type);
    if (vt < 0) Rf_error("Non-supported var type");
    const R_xlen_t n = Rf_nrows(M), p = Rf_ncols(M);
    if (Rf_nrows(EL) != n || Rf_nrows(EF) != p || Rf_ncols(EL) != Rf_ncols(EF)) Rf_error("non-conformable arguments");
    const R_xlen_t want = vt == EBPMF_VAR_BY_COL ? p : vt == EBPMF_VAR_BY_ROW ? n : 1;
    if (XLENGTH(sigma2) != want) Rf_error("sigma2 has the wrong length for this var_type");
    const int wv = Rf_asLogical(want_V);
    SEXP Mo = PROTECT(alloc_like(M));
    SEXP Vo = PROTECT(wv ? alloc_like(M) : R_NilValue);
    SEXP vs = PROTECT(Rf_allocVector(REALSXP, want));
    SEXP nu = PROTECT(Rf_allocVector(INTSXP, Rf_asInteger(nbatch)));
    ebpmf_solve_info info;
    int rc = ebpmf_batched_vga_step_host(h, (int64_t)n, (int64_t)p, REAL(M), Rf_ncols(EL), REAL(EL), REAL(EF), vt, REAL(sigma2),
                                         Rf_asInteger(maxiter), Rf_asReal(tol), REAL(Mo), wv ? REAL(Vo) : NULL, REAL(vs),
                                         &info, INTEGER(nu), NULL);
    if (rc != EBPMF_OK) { UNPROTECT(4); check_batched(h, rc); }
    SEXP out = PROTECT(info_list(&info, Mo, Vo, vs));
    Rf_setAttrib(out, Rf_install("n_updates_per_batch"), nu);
    UNPROTECT(5);
    return out;
}

static const R_CallMethodDef call_methods[] = {
    {"C_ebpmf_ctx_new", (DL_FUNC)&C_ebpmf_ctx_new, 1},
    {"C_ebpmf_ctx_set_y", (DL_FUNC)&C_ebpmf_ctx_set_y, 2},
    {"C_ebpmf_sum_lfactorial", (DL_FUNC)&C_ebpmf_sum_lfactorial, 2},
    {"C_ebpmf_vga_step", (DL_FUNC)&C_ebpmf_vga_step, 9},
    {"C_ebpmf_ctx_set_m", (DL_FUNC)&C_ebpmf_ctx_set_m, 2},
    {"C_ebpmf_vga_step_resident", (DL_FUNC)&C_ebpmf_vga_step_resident, 9},
    {"C_ebpmf_batched_new", (DL_FUNC)&C_ebpmf_batched_new, 4},
    {"C_ebpmf_batched_vga_step", (DL_FUNC)&C_ebpmf_batched_vga_step, 10},
    {"C_vga_pois_solver_mat_newton", (DL_FUNC)&C_vga_pois_solver_mat_newton, 9},
    {"C_vga_pois_solver_mat_newton_fixed_iter", (DL_FUNC)&C_vga_pois_solver_mat_newton_fixed_iter, 8},
    {"C_vga_pois_solver_mat_newton_low_memory", (DL_FUNC)&C_vga_pois_solver_mat_newton_low_memory, 11},
    {"C_ebpmf_log_update_sigma2", (DL_FUNC)&C_ebpmf_log_update_sigma2, 12},
    {"C_calc_ebpmf_log_obj", (DL_FUNC)&C_calc_ebpmf_log_obj, 14},
    {"C_calc_split_PMF_obj_flashier", (DL_FUNC)&C_calc_split_PMF_obj_flashier, 10},
    {"C_ebpmf_m_products", (DL_FUNC)&C_ebpmf_m_products, 4},
    {"C_ebpmf_sim_data_log", (DL_FUNC)&C_ebpmf_sim_data_log, 10},
    {"C_ebpmf_group_new", (DL_FUNC)&C_ebpmf_group_new, 1},
    {"C_ebpmf_group_set_y", (DL_FUNC)&C_ebpmf_group_set_y, 2},
    {"C_ebpmf_group_sum_lfactorial", (DL_FUNC)&C_ebpmf_group_sum_lfactorial, 1},
    {"C_ebpmf_group_set_m", (DL_FUNC)&C_ebpmf_group_set_m, 2},
    {"C_ebpmf_group_vga_step", (DL_FUNC)&C_ebpmf_group_vga_step, 10},
    {NULL, NULL, 0}};

void R_init_ebpmf_alpha(DllInfo *dll)
{
    R_registerRoutines(dll, NULL, call_methods, NULL, NULL);
    R_useDynamicSymbols(dll, FALSE);
}
