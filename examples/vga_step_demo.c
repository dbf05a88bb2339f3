/* vga_step_demo.c -- the C ABI of include/ebpmf_b200.h from plain C, as a non-R host would use it.
 *
 *   gcc -std=c99 -Iinclude examples/vga_step_demo.c -Lebpmf.alpha_b200/lib -lebpmf_b200 \
 *       -Wl,-rpath,$PWD/ebpmf.alpha_b200/lib -lm -o /tmp/vga_step_demo && /tmp/vga_step_demo
 *
 * Builds a small Poisson count matrix as a dgCMatrix (@p, @i, @x), runs one VGA step of ebpmf_log's
 * outer loop (R/ebpmf_log.R:302-327) and prints the update count and the ELBO partial sums.  There is no
 * CPU fallback: without a B200 the first call fails and the program exits 2 with the library's message. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "ebpmf_b200.h"

static uint64_t rng = 0x9E3779B97F4A7C15ull;
static double unif(void)
{
    rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17;
    return (double)(rng >> 11) / 9007199254740992.0;
}

int main(void)
{
    const int64_t n = 600, p = 70, K = 3;
    ebpmf_ctx *ctx = NULL;
    if (ebpmf_ctx_create(0, &ctx) != EBPMF_OK) {
        fprintf(stderr, "ebpmf_b200: %s\n", ebpmf_last_error(NULL));
        return 2;
    }
    /* factors and a sparse count matrix: lambda_ij = exp(sum_k EL_ik EF_jk), EL[,0] = offset, EF[,0] = 1 */
    double *EL = malloc(sizeof(double) * n * K), *EF = malloc(sizeof(double) * p * K);
    for (int64_t i = 0; i < n; ++i) { EL[i] = -2.0 + 0.5 * unif(); EL[i + n] = unif() - 0.5; EL[i + 2 * n] = unif() - 0.5; }
    for (int64_t j = 0; j < p; ++j) { EF[j] = 1.0; EF[j + p] = 2.0 * unif() - 1.0; EF[j + 2 * p] = 2.0 * unif() - 1.0; }
    int32_t *colptr = malloc(sizeof(int32_t) * (p + 1)), *rowidx = malloc(sizeof(int32_t) * n * p);
    double *x = malloc(sizeof(double) * n * p), *M = malloc(sizeof(double) * n * p), *Mo = malloc(sizeof(double) * n * p);
    double *sigma2 = malloc(sizeof(double) * p), *v_sum = malloc(sizeof(double) * p);
    int32_t nnz = 0;
    for (int64_t j = 0; j < p; ++j) {
        colptr[j] = nnz;
        sigma2[j] = 0.1 + 0.9 * unif();
        for (int64_t i = 0; i < n; ++i) {
            double beta = 0.0;
            for (int64_t k = 0; k < K; ++k) beta += EL[i + k * n] * EF[j + k * p];
            /* Poisson(lambda) by inversion (lambda is small here) */
            double lam = exp(beta), u = unif(), pr = exp(-lam), cdf = pr;
            int y = 0;
            while (u > cdf && y < 50) { ++y; pr *= lam / y; cdf += pr; }
            if (y > 0) { rowidx[nnz] = (int32_t)i; x[nnz] = y; ++nnz; }
            M[i + j * n] = log(y + 0.5);                       /* cold start */
        }
    }
    colptr[p] = nnz;

    ebpmf_solve_info info;
    int rc = ebpmf_ctx_set_y_csc(ctx, n, p, colptr, rowidx, x);
    if (rc == EBPMF_OK)
        rc = ebpmf_vga_step_host(ctx, n, p, M, K, EL, EF, EBPMF_VAR_BY_COL, sigma2, 10, 1e-5, Mo, NULL, v_sum, &info);
    if (rc != EBPMF_OK) {
        fprintf(stderr, "ebpmf_b200: %s\n", ebpmf_last_error(ctx));
        ebpmf_ctx_destroy(ctx);
        return 2;
    }
    printf("%s: %lld x %lld, nnz %d: %d Newton updates (converged %d), sym %.6f ssexp %.6f slogv %.6f v_sum[0] %.6f\n",
           ebpmf_version(), (long long)n, (long long)p, nnz, info.n_updates, info.converged, info.sym, info.ssexp,
           info.slogv, v_sum[0]);
    ebpmf_ctx_destroy(ctx);
    free(EL); free(EF); free(colptr); free(rowidx); free(x); free(M); free(Mo); free(sigma2); free(v_sum);
    return 0;
}
