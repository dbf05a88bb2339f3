/*
 * CPU ORACLE (C transcription) for the ebpmf_log VGA Newton hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it, and only as the
 * checker or the timed CPU baseline.  Nothing under ebpmf.alpha_b200/ links or calls it.
 *
 * PARITY: the reference (DongyueXie/ebpmf.alpha) is 100 % R, ships no golden vectors or
 * assertions for this path, and GNU R is not installed here ("parity unpinned" against GNU R
 * output).  What this file is held to instead: the outputs of the reference's own R source
 * as executed by the restricted R evaluator oracle/mini_r.py (tests/golden/minir_reference.npz,
 * tests/test_minir_reference.py).  It is a second, independent transcription of the R source
 * (the first is oracle/vga_numpy.py); tests require the two to agree to <= 1e-13.
 *
 * Citations are relative to /root/reference.  All matrices are column-major (R layout):
 * element (i,j) of an n x p matrix is a[i + n*j].  Operands that R would recycle
 * (scalar / length-n vector / n x p matrix) are passed as (pointer, row stride, col stride):
 *   scalar: (p,0,0)   length-n vector: (p,1,0)   length-p by-column: (p,0,1)   matrix: (p,1,n)
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp; no -ffast-math so each
 * R operator stays one IEEE operation, in R's order).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_OK 0
#define ORACLE_ERR_NAN 1      /* R: "missing value where TRUE/FALSE needed" (R/vga_solver_mat.R:26) */
#define ORACLE_ERR_ARG 2
#define ORACLE_ERR_MEM 3

#define SLOGV_CONST 0.9189385 /* literal, R/ebpmf_log.R:276,318 */

typedef long long i64;

int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

static void set_threads(int nthreads) {
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#else
    (void)nthreads;
#endif
}

/* ------------------------------------------------------------------------------------
 * a1  vga_pois_solver_mat_newton   R/vga_solver_mat.R:6-41
 * One pass per Newton iteration ("fused" loop order); the arithmetic per entry is R's,
 * operator by operator.  The global stop (:26) is evaluated over all n*p entries before
 * any entry is updated, as in R.  M1 is scratch (n*p) for the tentative update.
 * ------------------------------------------------------------------------------------ */
int oracle_vga_newton(i64 n, i64 p, const double *M0, const double *X,
                      const double *S, i64 S_rs, i64 S_cs,
                      const double *Beta, i64 B_rs, i64 B_cs,
                      const double *Sigma2, i64 V_rs, i64 V_cs,
                      int maxiter, double tol,
                      double *Mout, double *Vout /* may be NULL */,
                      int *n_updates, int *converged, double *maxabs_trace /* maxiter or NULL */,
                      int nthreads)
{
    if (maxiter < 1 || n < 1 || p < 1) return ORACLE_ERR_ARG;
    set_threads(nthreads);
    const i64 N = n * p;
    double *A = (double *)malloc(sizeof(double) * (size_t)N);
    double *B = (double *)malloc(sizeof(double) * (size_t)N);
    if (!A || !B) { free(A); free(B); return ORACLE_ERR_MEM; }

    /* :9-16  constants are recomputed per entry each pass (same values as R's stored arrays) */
    #pragma omp parallel for schedule(static)
    for (i64 j = 0; j < p; ++j)
        for (i64 i = 0; i < n; ++i) {
            const i64 e = i + n * j;
            const double s2 = Sigma2[i * V_rs + j * V_cs], be = Beta[i * B_rs + j * B_cs];
            const double const0 = (s2 * X[e] + be + 1);          /* :9  */
            const double c0m1 = const0 - 1;
            A[e] = (c0m1 < M0[e]) ? c0m1 : M0[e];                /* :16 Rfast::Pmin = std::min(M, const0-1): a NaN in M stays a NaN */
        }

    double *cur = A, *nxt = B;
    int u = 0, conv = 0, nan_seen = 0;
    const double *Vsrc = cur;   /* the M that the last evaluated temp was computed from */
    for (int it = 1; it <= maxiter; ++it) {                       /* :21 */
        double mx = 0.0; int has_nan = 0;
        #pragma omp parallel for schedule(static) reduction(max:mx) reduction(|:has_nan)
        for (i64 j = 0; j < p; ++j)
            for (i64 i = 0; i < n; ++i) {
                const i64 e = i + n * j;
                const double s2 = Sigma2[i * V_rs + j * V_cs], be = Beta[i * B_rs + j * B_cs];
                const double s = S[i * S_rs + j * S_cs], x = X[e], m = cur[e];
                const double const0 = (s2 * x + be + 1);          /* :9  */
                const double const1 = 1 / s2;                     /* :10 */
                const double const2 = s2 / 2;                     /* :11 */
                const double const3 = be / s2;                    /* :12 */
                const double temp = (const0 - m);                 /* :22 */
                const double sexp = s * exp(m + const2 / temp);   /* :23 */
                const double f = x - sexp - m * const1 + const3;  /* :25 */
                const double af = fabs(f);
                if (af != af) has_nan = 1; else if (af > mx) mx = af;
                nxt[e] = m - f / (-sexp * (1 + const2 / (temp * temp)) - const1);   /* :31 */
            }
        if (maxabs_trace) maxabs_trace[it - 1] = has_nan ? NAN : mx;
        Vsrc = cur;
        if (has_nan) { nan_seen = 1; break; }                     /* R errors in if() */
        if (mx < tol) { conv = 1; break; }                        /* :26-28 */
        double *t = cur; cur = nxt; nxt = t;                      /* :31 commit */
        ++u;
    }
    if (nan_seen) { free(A); free(B); return ORACLE_ERR_NAN; }

    #pragma omp parallel for schedule(static)
    for (i64 j = 0; j < p; ++j)
        for (i64 i = 0; i < n; ++i) {
            const i64 e = i + n * j;
            Mout[e] = cur[e];
            if (Vout) {
                const double s2 = Sigma2[i * V_rs + j * V_cs], be = Beta[i * B_rs + j * B_cs];
                const double const0 = (s2 * X[e] + be + 1);
                const double temp = (const0 - Vsrc[e]);           /* stale on exhaustion (:22 vs :31) */
                Vout[e] = s2 / temp;                              /* :36 */
            }
        }
    *n_updates = u; *converged = conv;
    free(A); free(B);
    return ORACLE_OK;
}

/* ------------------------------------------------------------------------------------
 * a1, "R-style" execution model: one full-array pass and one freshly allocated temporary
 * per R operator, single thread -- what the R interpreter does with R/vga_solver_mat.R:6-41.
 * Dense n x p Beta / Sigma2 and scalar S (the live call, R/ebpmf_log.R:302-308).
 * Used only as the timed "what R does" CPU baseline; results are identical to the above.
 * ------------------------------------------------------------------------------------ */
#define NEWV(name) double *name = (double *)malloc(sizeof(double) * (size_t)N); if (!name) return ORACLE_ERR_MEM
#define LOOP for (i64 e = 0; e < N; ++e)
int oracle_vga_newton_rstyle(i64 n, i64 p, const double *M0, const double *X, double S,
                             const double *Beta, const double *Sigma2,
                             int maxiter, double tol, double *Mout, double *Vout,
                             int *n_updates, int *converged)
{
    if (maxiter < 1) return ORACLE_ERR_ARG;
    const i64 N = n * p;
    NEWV(t1); LOOP t1[e] = Sigma2[e] * X[e];
    NEWV(t2); LOOP t2[e] = t1[e] + Beta[e]; free(t1);
    NEWV(const0); LOOP const0[e] = t2[e] + 1; free(t2);               /* :9  */
    NEWV(const1); LOOP const1[e] = 1 / Sigma2[e];                      /* :10 */
    NEWV(const2); LOOP const2[e] = Sigma2[e] / 2;                      /* :11 */
    NEWV(const3); LOOP const3[e] = Beta[e] / Sigma2[e];                /* :12 */
    NEWV(c0m1); LOOP c0m1[e] = const0[e] - 1;
    NEWV(M); LOOP M[e] = (M0[e] < c0m1[e]) ? M0[e] : c0m1[e]; free(c0m1);   /* :16 */
    double *temp = NULL;
    int u = 0, conv = 0;
    for (int it = 1; it <= maxiter; ++it) {
        free(temp);
        temp = (double *)malloc(sizeof(double) * (size_t)N); LOOP temp[e] = const0[e] - M[e];   /* :22 */
        NEWV(a1); LOOP a1[e] = const2[e] / temp[e];
        NEWV(a2); LOOP a2[e] = M[e] + a1[e]; free(a1);
        NEWV(a3); LOOP a3[e] = exp(a2[e]); free(a2);
        NEWV(sexp); LOOP sexp[e] = S * a3[e]; free(a3);               /* :23 */
        NEWV(b1); LOOP b1[e] = X[e] - sexp[e];
        NEWV(b2); LOOP b2[e] = M[e] * const1[e];
        NEWV(b3); LOOP b3[e] = b1[e] - b2[e]; free(b1); free(b2);
        NEWV(f); LOOP f[e] = b3[e] + const3[e]; free(b3);             /* :25 */
        NEWV(af); LOOP af[e] = fabs(f[e]);
        double mx = -INFINITY; int has_nan = 0;
        LOOP { if (af[e] != af[e]) has_nan = 1; else if (af[e] > mx) mx = af[e]; }
        free(af);
        if (has_nan) { free(f); free(sexp); free(temp); free(M); free(const0); free(const1); free(const2); free(const3); return ORACLE_ERR_NAN; }
        if (mx < tol) { conv = 1; free(f); free(sexp); break; }      /* :26 */
        NEWV(d1); LOOP d1[e] = temp[e] * temp[e];
        NEWV(d2); LOOP d2[e] = const2[e] / d1[e]; free(d1);
        NEWV(d3); LOOP d3[e] = 1 + d2[e]; free(d2);
        NEWV(d4); LOOP d4[e] = -sexp[e]; free(sexp);
        NEWV(d5); LOOP d5[e] = d4[e] * d3[e]; free(d3); free(d4);
        NEWV(d6); LOOP d6[e] = d5[e] - const1[e]; free(d5);
        NEWV(d7); LOOP d7[e] = f[e] / d6[e]; free(d6); free(f);
        NEWV(Mn); LOOP Mn[e] = M[e] - d7[e]; free(d7); free(M); M = Mn;   /* :31 */
        ++u;
    }
    LOOP Mout[e] = M[e];
    if (Vout) LOOP Vout[e] = Sigma2[e] / temp[e];                      /* :36 */
    free(temp); free(M); free(const0); free(const1); free(const2); free(const3);
    *n_updates = u; *converged = conv;
    return ORACLE_OK;
}
#undef NEWV
#undef LOOP

/* ------------------------------------------------------------------------------------
 * a9  vga_pois_solver_mat_newton_fixed_iter   R/vga_solver_mat.R:44-65
 * V0 == NULL  <=>  is.null(V)  ->  V = Sigma2/2  (:47-49)
 * ------------------------------------------------------------------------------------ */
int oracle_vga_fixed_iter(i64 n, i64 p, const double *M0, const double *V0, const double *Y,
                          const double *S, i64 S_rs, i64 S_cs,
                          const double *Beta, i64 B_rs, i64 B_cs,
                          const double *Sigma2, i64 V_rs, i64 V_cs,
                          int maxiter, double *Mout, double *Vout, int nthreads)
{
    if (maxiter < 1) return ORACLE_ERR_ARG;
    set_threads(nthreads);
    #pragma omp parallel for schedule(static)
    for (i64 j = 0; j < p; ++j)
        for (i64 i = 0; i < n; ++i) {
            const i64 e = i + n * j;
            const double s2 = Sigma2[i * V_rs + j * V_cs], be = Beta[i * B_rs + j * B_cs];
            const double s = S[i * S_rs + j * S_cs], y = Y[e];
            double m = M0[e];
            double v = V0 ? V0[e] : s2 / 2;                                         /* :48 */
            for (int it = 0; it < maxiter; ++it) {                                  /* :52 */
                double temp = s * exp(m + v / 2);                                   /* :54 */
                m = m - (y / temp - 1 - (m - be) / s2 / temp) / (-1 - 1 / s2 / temp);   /* :56 */
                temp = s * exp(m + v / 2) * v / 2;                                  /* :58 */
                v = v / exp((-1 - v / 2 / s2 / temp + 0.5 / temp)
                            / (-(1 + v / 2) - v / 2 / s2 / temp));                  /* :62 */
            }
            Mout[e] = m; Vout[e] = v;
        }
    return ORACLE_OK;
}

/* ------------------------------------------------------------------------------------
 * a10  vga_pois_solver_mat_newton_low_memory   R/vga_solver_mat.R:71-109
 * var_type: 0 constant, 1 by_row, 2 by_col (the reference's own coding, R/ebpmf_log.R:178-192);
 * any other value reproduces R's fall-through: clamp only (:78) and return.
 * sigma2 has length p (by_col), n (by_row) or 1 (constant).  l0 has length n, f0 length p.
 * ------------------------------------------------------------------------------------ */
int oracle_vga_low_memory(i64 n, i64 p, i64 K, const double *M0, const double *Y,
                          const double *l0, const double *f0,
                          const double *EL /* n x K */, const double *EF /* p x K */,
                          const double *sigma2, int var_type, int maxiter, double tol,
                          double *Mout, int *n_updates, int *converged, int nthreads)
{
    if (maxiter < 1) return ORACLE_ERR_ARG;
    set_threads(nthreads);
    const i64 N = n * p;
    double *Beta = (double *)malloc(sizeof(double) * (size_t)N);
    double *A = (double *)malloc(sizeof(double) * (size_t)N);
    double *B = (double *)malloc(sizeof(double) * (size_t)N);
    if (!Beta || !A || !B) { free(Beta); free(A); free(B); return ORACLE_ERR_MEM; }
    #pragma omp parallel for schedule(static)
    for (i64 j = 0; j < p; ++j)
        for (i64 i = 0; i < n; ++i) {
            double b = 0.0;
            for (i64 k = 0; k < K; ++k) b += EL[i + n * k] * EF[j + p * k];          /* :77 tcrossprod */
            Beta[i + n * j] = b;
            A[i + n * j] = (b < M0[i + n * j]) ? b : M0[i + n * j];                   /* :78 std::min(M, Beta) */
        }
    double *cur = A, *nxt = B;
    int u = 0, conv = 0, nan_seen = 0;
    if (var_type == 0 || var_type == 1 || var_type == 2) {
        for (int it = 1; it <= maxiter; ++it) {
            double mx = 0.0; int has_nan = 0;
            #pragma omp parallel for schedule(static) reduction(max:mx) reduction(|:has_nan)
            for (i64 j = 0; j < p; ++j)
                for (i64 i = 0; i < n; ++i) {
                    const i64 e = i + n * j;
                    const double y = Y[e], be = Beta[e], m = cur[e];
                    double sexp, f, g;
                    if (var_type == 2) {
                        const double s2 = sigma2[j];
                        const double const0 = y * s2 + be + 1;                               /* :81 */
                        sexp = l0[i] * (exp(m + (1 / (const0 - m)) * (s2 / 2)) * f0[j]);     /* :83 */
                        f = y - sexp + (be - m) * (1 / s2);                                  /* :84 */
                        g = -sexp * (1 + 0.5 * pow(const0 - m, -2.0)) - (1 / s2);            /* :90 */
                    } else {
                        const double s2 = (var_type == 1) ? sigma2[i] : sigma2[0];
                        const double const0 = s2 * y + be + 1;                               /* :95 */
                        sexp = l0[i] * (exp(m + (1 / (const0 - m)) * s2 / 2) * f0[j]);       /* :97 */
                        f = y - sexp + (be - m) / s2;                                        /* :98 */
                        g = -sexp * (1 + 0.5 * pow(const0 - m, -2.0)) - 1 / s2;              /* :105 */
                    }
                    const double af = fabs(f);
                    if (af != af) has_nan = 1; else if (af > mx) mx = af;
                    nxt[e] = m - f / g;
                }
            if (has_nan) { nan_seen = 1; break; }
            if (mx < tol) { conv = 1; break; }                                               /* :85,:100 */
            double *t = cur; cur = nxt; nxt = t;
            ++u;
        }
    }
    if (!nan_seen) memcpy(Mout, cur, sizeof(double) * (size_t)N);                            /* :108 */
    free(Beta); free(A); free(B);
    if (nan_seen) return ORACLE_ERR_NAN;
    *n_updates = u; *converged = conv;
    return ORACLE_OK;
}

/* ------------------------------------------------------------------------------------
 * a4/a5  ELBO partial sums and v_sum   R/ebpmf_log.R:316-327
 * long double accumulators = base R's sum()/colSums()/rowSums().
 * col_v (length p) and row_v (length n) may each be NULL.
 * ------------------------------------------------------------------------------------ */
int oracle_partials(i64 n, i64 p, const double *Y, const double *M, const double *V,
                    double *sym, double *ssexp, double *slogv, double *sum_v,
                    double *col_v, double *row_v, int nthreads)
{
    set_threads(nthreads);
    long double a = 0, b = 0, c = 0, d = 0;
    #pragma omp parallel for schedule(static) reduction(+:a,b,c,d)
    for (i64 j = 0; j < p; ++j) {
        long double cv = 0;
        for (i64 i = 0; i < n; ++i) {
            const i64 e = i + n * j;
            a += Y[e] * M[e];                             /* :316 */
            b += exp(M[e] + V[e] / 2);                    /* :317 */
            c += log(V[e]) / 2 + SLOGV_CONST;             /* :318 */
            cv += V[e];
        }
        d += cv;
        if (col_v) col_v[j] = (double)cv;                 /* :324 */
    }
    if (row_v) {
        #pragma omp parallel for schedule(static)
        for (i64 i = 0; i < n; ++i) {
            long double rv = 0;
            for (i64 j = 0; j < p; ++j) rv += V[i + n * j];
            row_v[i] = (double)rv;                        /* :326 */
        }
    }
    *sym = (double)a; *ssexp = (double)b; *slogv = (double)c; *sum_v = (double)d;        /* :322 */
    return ORACLE_OK;
}

/* a8  sum_lfactorial_sparseMat   R/ebpmf_log.R:470-472  (lfactorial(x) == lgamma(x+1)) */
double oracle_sum_lfactorial(i64 nnz, const double *x)
{
    long double a = 0;
    for (i64 e = 0; e < nnz; ++e) a += lgamma(x[e] + 1.0);
    return (double)a;
}

/* a2  Beta = tcrossprod(EL, EF)   R/ebpmf_log.R:266,305 */
void oracle_tcrossprod(i64 n, i64 p, i64 K, const double *EL, const double *EF, double *Beta, int nthreads)
{
    set_threads(nthreads);
    #pragma omp parallel for schedule(static)
    for (i64 j = 0; j < p; ++j)
        for (i64 i = 0; i < n; ++i) {
            double b = 0.0;
            for (i64 k = 0; k < K; ++k) b += EL[i + n * k] * EF[j + p * k];
            Beta[i + n * j] = b;
        }
}

/* ------------------------------------------------------------------------------------
 * The un-batched VGA step of ebpmf_log's outer loop, R/ebpmf_log.R:302-327, for the
 * timed CPU baseline: Beta = fitted(fit_flash) (a2), Sigma2 = adjust_var_shape (a3, not
 * materialised -- strides), a1 with S = 1, then a4/a5.  Y is dense (R densifies, :202).
 * var_type: 0 constant, 1 by_row, 2 by_col.  v_sum has length 1 / n / p accordingly.
 * ------------------------------------------------------------------------------------ */
int oracle_vga_step(i64 n, i64 p, i64 K, const double *M0, const double *Y,
                    const double *EL, const double *EF, const double *sigma2, int var_type,
                    int maxiter, double tol, double *Mout, double *Vout,
                    double *sums /* sym, ssexp, slogv */, double *v_sum,
                    int *n_updates, int *converged, int nthreads)
{
    const i64 N = n * p;
    double *Beta = (double *)malloc(sizeof(double) * (size_t)N);
    if (!Beta) return ORACLE_ERR_MEM;
    oracle_tcrossprod(n, p, K, EL, EF, Beta, nthreads);
    const double one = 1.0;
    i64 rs = 0, cs = 0;
    if (var_type == 1) rs = 1; else if (var_type == 2) cs = 1; else if (var_type != 0) { free(Beta); return ORACLE_ERR_ARG; }
    int rc = oracle_vga_newton(n, p, M0, Y, &one, 0, 0, Beta, 1, n, sigma2, rs, cs,
                               maxiter, tol, Mout, Vout, n_updates, converged, NULL, nthreads);
    free(Beta);
    if (rc) return rc;
    double sv;
    rc = oracle_partials(n, p, Y, Mout, Vout, &sums[0], &sums[1], &sums[2], &sv,
                         var_type == 2 ? v_sum : NULL, var_type == 1 ? v_sum : NULL, nthreads);
    if (var_type == 0) v_sum[0] = sv;
    return rc;
}
