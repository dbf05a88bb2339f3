/*
 * ebpmf_b200.h -- plain C ABI of the B200-native VGA Newton hot path of ebpmf.alpha.
 *
 * The reference (DongyueXie/ebpmf.alpha) is 100 % R with no native code (DESCRIPTION:42,
 * NAMESPACE has no useDynLib), so there is no existing FFI for this path.  The entry points
 * below are what a `.Call` shim binds; each cites the R function / lines it replaces
 * (paths relative to the reference root).  The R-side glue that calls them is
 * ebpmf.alpha_b200/r/src/r_glue.c, the binding a maintainer adds is shown in INTEGRATION.md.
 *
 * Conventions
 *   - every pointer is a HOST pointer unless the name says `dev`; matrices are column-major
 *     (R layout): element (i,j) of an n x p matrix is a[i + n*j]; doubles are IEEE fp64;
 *   - inputs are never written (R shares its arguments, copy-on-modify);
 *   - every function returns an int status (EBPMF_OK == 0); the message of the last failure on
 *     a context is returned by ebpmf_last_error();
 *   - there is NO CPU fallback: without a usable sm_100 GPU every call fails with
 *     EBPMF_ERR_CUDA;
 *   - a context is bound to one GPU and is not thread-safe; use one context per host thread
 *     (R is single-threaded).  Calls are synchronous: results are valid on return.
 *
 * var_type coding (the reference's own `var.type`, R/ebpmf_log.R:178-192):
 *   0 = 'constant', 1 = 'by_row', 2 = 'by_col'; anything else -> EBPMF_ERR_VAR_TYPE
 *   ("Non-supported var type", R/ebpmf_log.R:444,483).
 */
#ifndef EBPMF_B200_H
#define EBPMF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)   /* the library is built with -fvisibility=hidden: only this ABI is exported */
#endif

#define EBPMF_OK            0
#define EBPMF_ERR_ARG       1   /* bad argument (NULL, maxiter outside [1,65534], shapes that do not conform) */
#define EBPMF_ERR_CUDA      2   /* CUDA runtime failure / no usable GPU */
#define EBPMF_ERR_NAN       3   /* a NaN reached the stop test: R stops with "missing value where
                                   TRUE/FALSE needed" at R/vga_solver_mat.R:26,85,100 */
#define EBPMF_ERR_VAR_TYPE  4   /* "Non-supported var type" R/ebpmf_log.R:444,483 */
#define EBPMF_ERR_STATE     5   /* call order: Y / M / factors / sigma2 not set */
#define EBPMF_ERR_NCCL      6   /* NCCL missing or a collective failed */
#define EBPMF_ERR_NOMEM     7   /* host or device allocation failed */

#define EBPMF_VAR_CONSTANT  0
#define EBPMF_VAR_BY_ROW    1
#define EBPMF_VAR_BY_COL    2

/* How an operand that R would recycle against an n x p matrix is laid out. */
#define EBPMF_OP_SCALAR     0   /* one double                                         */
#define EBPMF_OP_ROWVEC     1   /* length n, element i applies to row i (R recycling) */
#define EBPMF_OP_COLVEC     2   /* length p, element j applies to column j (byrow=TRUE matrix) */
#define EBPMF_OP_MATRIX     3   /* n x p column-major                                 */

typedef struct ebpmf_ctx ebpmf_ctx;

/* Outcome of one solver call.  n_updates is the GLOBAL Newton update count u that
 * R/vga_solver_mat.R:21-33 would have performed (break at iteration u+1, or u == maxiter when
 * the loop is exhausted, in which case V is one iteration stale exactly as in R, :22,:31,:36). */
typedef struct ebpmf_solve_info {
    int32_t n_updates;      /* u */
    int32_t converged;      /* 1: max(abs(f)) < tol held at iteration u+1 ; 0: maxiter exhausted */
    int32_t passes;         /* kernel passes over the matrix this call needed (diagnostic) */
    int32_t reserved;
    double  sym;            /* sum(Y*M)                      R/ebpmf_log.R:274,316 */
    double  ssexp;          /* sum(exp(M+V/2))               R/ebpmf_log.R:275,317 */
    double  slogv;          /* sum(log(V)/2+0.9189385)       R/ebpmf_log.R:276,318 */
    double  kernel_ms;      /* device time of the kernels of this call (CUDA events) */
    double  pass1_ms;       /* first pass (every unit to its own convergence)        */
    double  pass2_ms;       /* top-up pass(es) to the global count                    */
    double  reduce_ms;      /* partial-sum reductions (+ NCCL combine when sharded)   */
} ebpmf_solve_info;

/* ---- host memory for RESULT matrices ------------------------------------------------------
 * R allocates a fresh n x p matrix for every result (Rf_allocMatrix).  At the sizes of this path that is
 * millions of first-touch page faults per call, which halves the device->host rate.  The .Call glue instead
 * allocates result matrices with allocVector3() and an R_allocator_t whose mem_alloc / mem_free are these
 * two functions: blocks come from a process-wide pool and return to it when R's garbage collector frees the
 * matrix, so from the second outer iteration on a result lands in pages that are already mapped and
 * page-locked (blocks >= 64 MB are cudaHostRegister'ed once, so the copy engine writes them directly;
 * EBPMF_B200_POOL_PIN=0 keeps them pageable; EBPMF_B200_POOL_KEEP_GB, default 96, bounds what the pool keeps).  Any host pointer remains valid as an
 * output argument of every entry point; the pool is an optimisation, not a requirement. */
void *ebpmf_host_alloc(size_t bytes);          /* NULL when the system is out of memory */
int   ebpmf_host_free(void *ptr);              /* EBPMF_ERR_ARG when ptr did not come from ebpmf_host_alloc */
void  ebpmf_host_pool_trim(void);              /* give the kept blocks back to the system */
void  ebpmf_host_pool_stats(double *out4);     /* fresh allocations, reuses, bytes kept, blocks handed out */

/* ---- library / context ---------------------------------------------------------------- */
const char *ebpmf_version(void);
int  ebpmf_device_count(int *count);
int  ebpmf_ctx_create(int device, ebpmf_ctx **out);
int  ebpmf_ctx_destroy(ebpmf_ctx *ctx);
const char *ebpmf_last_error(const ebpmf_ctx *ctx);       /* ctx may be NULL (creation errors) */

/* ---- multi-GPU: one context per GPU, rows sharded ------------------------
 * Each rank holds a contiguous ROW shard of Y / M / EL (and of sigma2, v_sum for 'by_row');
 * EF and the by_col/constant sigma2 are replicated.  After ebpmf_ctx_comm_init every
 * ebpmf_ctx_vga_step combines, over NCCL (NVLink/NVSwitch): the global update count (MAX),
 * then v_sum[p] || sym,ssexp,slogv || flags (one fp64 SUM all-reduce).  id128 is a 128-byte
 * ncclUniqueId produced by rank 0 with ebpmf_nccl_unique_id and shipped to the other ranks by
 * the caller (any side channel).  NCCL is dlopen'ed on first use; single-GPU use needs none. */
int  ebpmf_nccl_unique_id(void *id128);
int  ebpmf_ctx_comm_init(ebpmf_ctx *ctx, int rank, int world, const void *id128);

/* ---- resident state of the fused ebpmf_log VGA step --------------------------------------
 * Y is constant across outer iterations and is uploaded once (R/ebpmf_log.R:195-203).
 * set_y_csc takes the dgCMatrix slots @p (int32[p+1]), @i (int32[nnz], 0-based, sorted within
 * a column), @x (double[nnz]) of this rank's row shard; set_y_dense takes as.matrix(Y). */
int  ebpmf_ctx_set_y_csc(ebpmf_ctx *ctx, int64_t n, int64_t p,
                         const int32_t *colptr, const int32_t *rowidx, const double *x);
int  ebpmf_ctx_set_y_dense(ebpmf_ctx *ctx, int64_t n, int64_t p, const double *Y);

/* sum(lfactorial(Y@x)) over the resident Y -- sum_lfactorial_sparseMat, R/ebpmf_log.R:470-472
 * (call site :172-176; `const` is its negative).  Rank-local; the caller sums over ranks. */
int  ebpmf_ctx_sum_lfactorial(ebpmf_ctx *ctx, double *out);
/* same for an arbitrary host vector x (e.g. a dense Y) */
int  ebpmf_sum_lfactorial(ebpmf_ctx *ctx, int64_t len, const double *x, double *out);

/* flash_fit$Y (the latent M, n x p), flash_fit$EF[[1]] (EL, n x K), flash_fit$EF[[2]]
 * (EF, p x K) with K = K_flash + 2 offset columns (R/ebpmf_log_flash_init.R:31-45), sigma2 of
 * length 1 / n / p per var_type.  Beta = EL %*% t(EF) = fitted(fit_flash) is never formed. */
int  ebpmf_ctx_set_m(ebpmf_ctx *ctx, const double *M);
int  ebpmf_ctx_set_factors(ebpmf_ctx *ctx, int64_t K, const double *EL, const double *EF);
int  ebpmf_ctx_set_sigma2(ebpmf_ctx *ctx, int var_type, const double *sigma2);

/* One VGA step on the resident state, R/ebpmf_log.R:302-327 (un-batched branch):
 *   res   = vga_pois_solver_mat_newton(M, Y, 1, fitted(fit_flash), adjust_var_shape(sigma2,..),
 *                                      maxiter, tol, return_V=TRUE)       R/vga_solver_mat.R:6-41
 *   M    <- res$M ; sym, ssexp, slogv (:316-318) ; v_sum = col/row/total sums of res$V (:321-327)
 * The resident M is replaced by res$M.  V is materialised on the device only if want_v != 0. */
int  ebpmf_ctx_vga_step(ebpmf_ctx *ctx, int maxiter, double tol, int want_v, ebpmf_solve_info *info);

int  ebpmf_ctx_get_m(ebpmf_ctx *ctx, double *M);
/* number of peer GPUs whose published update count this context can read over NVLink during
 * pass 1 (0: none; the ranks then only meet in the all-reduce between the passes) */
int  ebpmf_ctx_peer_count(ebpmf_ctx *ctx, int *npeers);

/* undo the M <- res$M of the last step: the resident M is again the step's input (the input
 * buffer is never written by a step).  Lets a benchmark repeat the same step without a copy. */
int  ebpmf_ctx_rewind_m(ebpmf_ctx *ctx);
int  ebpmf_ctx_get_v(ebpmf_ctx *ctx, double *V);                  /* needs want_v at the step */
int  ebpmf_ctx_get_v_sum(ebpmf_ctx *ctx, double *v_sum);          /* length 1 / n / p */

/* read-back of the resident state (used by the parity tests on device-generated inputs) */
int  ebpmf_ctx_get_shape(ebpmf_ctx *ctx, int64_t *n, int64_t *p, int64_t *K, int64_t *nnz);
int  ebpmf_ctx_get_y_csc(ebpmf_ctx *ctx, int32_t *colptr, int32_t *rowidx, double *x);
int  ebpmf_ctx_get_factors(ebpmf_ctx *ctx, double *EL, double *EF);
int  ebpmf_ctx_get_sigma2(ebpmf_ctx *ctx, double *sigma2);
/* rows [row0, row0+nrows) of the resident M (which = 0), of V (1), or of the M the last step started from (2:
 * what ebpmf_ctx_rewind_m would restore; a step never writes its input) as an nrows x p matrix */
int  ebpmf_ctx_get_rows(ebpmf_ctx *ctx, int which, int64_t row0, int64_t nrows, double *out);
/* the block rows [row0, row0+nrows) x columns [col0, col0+ncols) of the resident M (or V), packed nrows x ncols */
int  ebpmf_ctx_get_block(ebpmf_ctx *ctx, int which, int64_t row0, int64_t nrows, int64_t col0, int64_t ncols, double *out);

/* ebpmf_log_update_sigma2, R/ebpmf_log.R:413-447, on the device from the v_sum of the last step.
 * R2 = fit_flash$flash_fit$R2 (length 1|n|p; for 'constant' the caller passes sum(R2)).
 * n_total / p_total are the FULL matrix dims (n of all ranks).  cap_var_mean_ratio > 0 enables
 * the gate of :416,:425,:435 on fitted(fit_flash), recomputed on device from EL, EF: the global max
 * ('constant'), matrixStats::colMaxs = the column maxima ('by_col'), and for 'by_row' what the reference
 * really compares -- Rfast::rowMaxs(x), whose default value = FALSE returns the 1-based POSITION of each
 * row's maximum, not the maximum (NAMESPACE: importFrom(Rfast,rowMaxs)).  The new sigma2 becomes the
 * resident one and is copied to sigma2_out. */
int  ebpmf_ctx_update_sigma2(ebpmf_ctx *ctx, const double *R2, double a0, double b0,
                             double cap_var_mean_ratio, int64_t n_total, int64_t p_total,
                             double *sigma2_out);

/* Stand-alone ebpmf_log_update_sigma2(fit_flash,sigma2,v_sum,var_type,cap_var_mean_ratio,a0,b0,n,p),
 * R/ebpmf_log.R:413-447, all operands given by the caller (host pointers): sigma2 / v_sum / R2 of
 * length 1|n|p ('constant': R2 is the vector fit_flash$flash_fit$R2 summed by the caller -> length 1).
 * EL (n x K), EF (p x K) are only read when cap_var_mean_ratio > 0 (the gate needs fitted(fit_flash)). */
int  ebpmf_update_sigma2(ebpmf_ctx *ctx, int var_type, int64_t n, int64_t p, const double *sigma2,
                         const double *v_sum, const double *R2, double a0, double b0,
                         double cap_var_mean_ratio, int64_t K, const double *EL, const double *EF,
                         double *sigma2_out);

/* calc_ebpmf_log_obj, R/ebpmf_log.R:406-410.  sv, sigma2 and R2 come with their OWN lengths: R takes each
 * sum() of :407 at the length of its own operands (the sigma2-only terms over length(sigma2); sum(sv/2/sigma2)
 * and sum(R2/2/sigma2) over the longer operand, the shorter one recycled), e.g. var_type 'constant' has a
 * scalar sigma2 and may be handed the vector fit_flash$flash_fit$R2. */
int  ebpmf_calc_obj(ebpmf_ctx *ctx, int64_t n, int64_t p, double sym, double ssexp, double slogv,
                    int64_t len_sv, const double *sv, int64_t len_sigma2, const double *sigma2,
                    int64_t len_r2, const double *R2,
                    double KL_LF, double const_, double ss, double a0, double b0, double *obj);

/* calc_split_PMF_obj_flashier(Y,S,sigma2,M,V,fit_flash,KL_LF,const,var_type), inst/code/Poisson_MF_splitting.R:373-391
 * (the archived driver's un-fused ELBO): dense n x p Y, M, V; S by EBPMF_OP_* kind; R2 = fit_flash$flash.fit$R2.
 * One streaming pass on the device (32 B per entry) + the sigma2 terms, every sum() at its own length. */
int  ebpmf_calc_split_obj(ebpmf_ctx *ctx, int64_t n, int64_t p, const double *Y, const double *S, int S_kind,
                          const double *M, const double *V, int var_type, int64_t len_sigma2, const double *sigma2,
                          int64_t len_r2, const double *R2, double KL_LF, double const_, double *obj);

/* The whole step with CALLER-OWNED HOST buffers in one call (what the .Call shim uses each outer iteration):
 * factors + sigma2 + M in, step, M [+ V] + v_sum out.  n, p are the dims of the host matrices and must equal
 * those of the resident Y (else EBPMF_ERR_ARG, R's "non-conformable arrays").  The host matrices may be
 * pageable (REAL() of an R matrix): for matrices >= 32 MB the library moves them through an internal ring of
 * pinned slots drained / filled by a few host threads, by COLUMN CHUNKS that overlap the kernels -- the M
 * input of chunk c+1 goes up while the first pass runs on chunk c, and the result of a chunk comes down as
 * soon as its first pass is done (speculatively: the tiles are taken hardest-first by the previous step's
 * per-tile counts, so normally only the first chunk is still changed by the top-up pass; chunks the top-up
 * pass did touch are copied again).  Results are bit-identical to set_m + vga_step + get_m.  V_out may be NULL.
 * Measured rates and the probe behind the design: profiles/r02_probe_host_copy.txt, DESIGN.md section 7b. */
int  ebpmf_vga_step_host(ebpmf_ctx *ctx, int64_t n, int64_t p, const double *M_in, int64_t K, const double *EL,
                         const double *EF, int var_type, const double *sigma2,
                         int maxiter, double tol,
                         double *M_out, double *V_out, double *v_sum, ebpmf_solve_info *info);

/* The same step when M already lives on the device: ebpmf_log feeds res$M back unchanged
 * (fit_flash$flash_fit$Y = res$M, R/ebpmf_log.R:309; flash_init(fit_flash$flash_fit$Y, ...),
 * R/ebpmf_log_flash_update.R:35), so after one ebpmf_ctx_set_m(M_init) the input of every later step is
 * the resident M left by the previous one.  In: factors and sigma2.  Out: M [, V], v_sum, info. */
int  ebpmf_vga_step_resident_host(ebpmf_ctx *ctx, int64_t n, int64_t p, int64_t K, const double *EL, const double *EF,
                                  int var_type, const double *sigma2, int maxiter, double tol,
                                  double *M_out, double *V_out, double *v_sum, ebpmf_solve_info *info);

/* ---- N1: what flashier needs FROM M, computed where M lives -----------------------------------
 * flashier treats the latent M as its data matrix (fit_flash$flash_fit$Y = res$M, R/ebpmf_log.R:270,309;
 * flash_init(fit_flash$flash_fit$Y, ...), R/ebpmf_log_flash_update.R:35) and touches it only through
 * products with the current loadings / factors and through the residual sums of squares R2
 * (R/ebpmf_log_flash_update.R:55-64).  One streaming pass over the RESIDENT M (8 B per entry, both products on
 * the FP64 tensor path -- mma.sync.m8n8k4.f64 --; nothing crosses PCIe but the small results):
 *   MF  (n x Kw) = M %*% EF_w                      MtL (p x Kw) = t(M) %*% EL_w
 *   r2_col (p)   = colSums((M - EL_w %*% t(EF_w))^2)    r2_row (n) = rowSums(...)     (either may be NULL)
 * EL_w is n x Kw, EF_w is p x Kw (column-major host pointers; this rank's rows of EL_w).  With a
 * communicator MtL and r2_col are summed over the ranks (rows are sharded); MF and r2_row stay sharded. */
int  ebpmf_ctx_m_products(ebpmf_ctx *ctx, int64_t Kw, const double *EL_w, const double *EF_w,
                          double *MF, double *MtL, double *r2_col, double *r2_row,
                          double *kernel_ms /* NULL or 3 doubles: whole call, the streaming products kernel(s), the partial-sum kernels */);

/* ---- one process, several GPUs (an R session is one process) -------------------------------
 * A group owns one context per GPU and one NCCL communicator over them; every group call runs one
 * host thread per GPU.  set_y takes the FULL dgCMatrix and deals its rows to the GPUs in 256-row
 * blocks; vga_step_host takes the FULL n x p / n x K host matrices (column-major) and each GPU
 * copies its row shard with strided transfers.  Results are those of ebpmf_vga_step_host. */
typedef struct ebpmf_group ebpmf_group;
int  ebpmf_group_create(int ndev, const int *devices /* NULL: 0..ndev-1 */, ebpmf_group **out);
int  ebpmf_group_destroy(ebpmf_group *g);
const char *ebpmf_group_last_error(const ebpmf_group *g);
int  ebpmf_group_size(const ebpmf_group *g);
int  ebpmf_group_set_option(ebpmf_group *g, const char *name, double value);
int  ebpmf_group_set_y_csc(ebpmf_group *g, int64_t n, int64_t p,
                           const int32_t *colptr, const int32_t *rowidx, const double *x);
int  ebpmf_group_sum_lfactorial(ebpmf_group *g, double *out);
/* n, p: dims of the host matrices (must equal the resident Y's).  M_in == NULL: the input is the M the shards
 * already hold (ebpmf_group_set_m once, then every step's output), as in ebpmf_vga_step_resident_host. */
int  ebpmf_group_set_m(ebpmf_group *g, int64_t n, int64_t p, const double *M);
int  ebpmf_group_vga_step_host(ebpmf_group *g, int64_t n, int64_t p, const double *M_in, int64_t K, const double *EL,
                               const double *EF, int var_type, const double *sigma2,
                               int maxiter, double tol,
                               double *M_out, double *V_out, double *v_sum, ebpmf_solve_info *info);

/* ---- batch_size-faithful, host-streamed step (R/ebpmf_log.R:195-203, :254-300) -------------
 * With general_control$batch_size < n, ebpmf_log cuts the rows into n_batch contiguous batches
 * (split(1:n, cut(...)), :195-199) and runs vga_pois_solver_mat_newton once per batch: the stop test
 * max(abs(f)) < tol is scoped to the batch (:263-268), M_b is written back (:270), sym/ssexp/slogv
 * add up in batch order (:274-276), v_sum adds up ('by_col', 'constant') or concatenates ('by_row')
 * (:281-293).  ebpmf_batched reproduces that branch: every batch's Y slab stays resident as its own
 * CSC, NO M is kept on the device between calls, and the batches are pipelined through `nslots`
 * worker contexts (one stream and one host thread each), so the H2D copy of one batch overlaps the
 * solve of another and the D2H copy of a third.  Device memory holds nslots batches of M, so the
 * matrix may be larger than HBM.  batch_row0 has nbatch+1 increasing entries, [0] = 0, [nbatch] = n
 * (R's cut() gives contiguous ranges; the caller computes them).
 * vga_step_host: full n x p / n x K column-major host matrices in and out (M_out may alias M_in);
 * v_sum has length p / n / 1; info carries the sums over batches, n_updates = max and converged =
 * all over batches; n_updates / converged (length nbatch, may be NULL) give the per-batch values. */
typedef struct ebpmf_batched ebpmf_batched;
int  ebpmf_batched_create(int device, int nslots, ebpmf_batched **out);
/* several GPUs, one process: batches are independent, so there is no collective -- slot s lives on
 * devices[s % ndev] (NULL: 0..ndev-1) and batch b runs on slot b % (ndev * nslots_per_device) */
int  ebpmf_batched_create_multi(int ndev, const int *devices, int nslots_per_device, ebpmf_batched **out);
int  ebpmf_batched_destroy(ebpmf_batched *h);
const char *ebpmf_batched_last_error(const ebpmf_batched *h);
int  ebpmf_batched_set_option(ebpmf_batched *h, const char *name, double value);
int  ebpmf_batched_set_y_csc(ebpmf_batched *h, int64_t n, int64_t p, const int32_t *colptr,
                             const int32_t *rowidx, const double *x,
                             int64_t nbatch, const int64_t *batch_row0);
int  ebpmf_batched_sum_lfactorial(ebpmf_batched *h, double *out);
int  ebpmf_batched_vga_step_host(ebpmf_batched *h, int64_t n, int64_t p, const double *M_in, int64_t K, const double *EL,
                                 const double *EF, int var_type, const double *sigma2,
                                 int maxiter, double tol,
                                 double *M_out, double *V_out, double *v_sum, ebpmf_solve_info *info,
                                 int32_t *n_updates, int32_t *converged);

/* ---- drop-in solver signatures (dense operands, as R passes them) ------------------------ */

/* vga_pois_solver_mat_newton(M,X,S,Beta,Sigma2,maxiter,tol,return_V)  R/vga_solver_mat.R:6-41
 * X: dense n x p (X != NULL) or dgCMatrix slots (X == NULL; test_memory.R:24-47 passes sparse
 * counts).  S / Beta / Sigma2 are described by an EBPMF_OP_* kind.  V_out == NULL <=> return_V=FALSE. */
int  ebpmf_vga_newton(ebpmf_ctx *ctx, int64_t n, int64_t p, const double *M,
                      const double *X, const int32_t *X_colptr, const int32_t *X_rowidx, const double *X_x,
                      const double *S, int S_kind, const double *Beta, int Beta_kind,
                      const double *Sigma2, int Sigma2_kind, int maxiter, double tol,
                      double *M_out, double *V_out, ebpmf_solve_info *info);

/* vga_pois_solver_mat_newton_fixed_iter(M,V,Y,S,Beta,Sigma2,maxiter)  R/vga_solver_mat.R:44-65
 * V == NULL <=> is.null(V) -> V = Sigma2/2 (:47-49). */
int  ebpmf_vga_fixed_iter(ebpmf_ctx *ctx, int64_t n, int64_t p, const double *M, const double *V,
                          const double *Y, const double *S, int S_kind, const double *Beta, int Beta_kind,
                          const double *Sigma2, int Sigma2_kind, int maxiter,
                          double *M_out, double *V_out);

/* vga_pois_solver_mat_newton_low_memory(M,Y,l0,f0,EL,EF,sigma2,var_type,maxiter,tol)
 * R/vga_solver_mat.R:71-109.  Y dense (Y != NULL) or dgCMatrix slots.  l0 length n, f0 length p,
 * EL n x K, EF p x K, sigma2 length p ('by_col'), n ('by_row') or 1 ('constant').  Returns M only
 * (:108).  The stray debug print(range(f)) of :99 is not reproduced. */
int  ebpmf_vga_low_memory(ebpmf_ctx *ctx, int64_t n, int64_t p, int64_t K, const double *M,
                          const double *Y, const int32_t *Y_colptr, const int32_t *Y_rowidx, const double *Y_x,
                          const double *l0, const double *f0, const double *EL, const double *EF,
                          const double *sigma2, int var_type, int maxiter, double tol,
                          double *M_out, ebpmf_solve_info *info);

/* ---- benchmark support: on-device synthetic sim_data_log inputs --------------------------
 * Restates the DGP of sim_data_log (R/simulate_poisson_matrix.R:9-61) with a counter-based RNG
 * keyed on (seed, global row, column), directly into the context's resident state for THIS
 * rank's rows [row0, row0+n): Y (CSC), EL = [log l0 + c, 1, L~], EF = [1, log f0, F~],
 * sigma2 ~ U(0.05,1) by column, M0 = log(Y+0.5) (cold) or Beta (warm).  Not part of the
 * reference's API; exists so that 1M x 30k inputs never cross PCIe.  density_out / nnz_out report
 * what was generated. */
int  ebpmf_ctx_sim_data_log(ebpmf_ctx *ctx, int64_t n, int64_t p, int64_t K, int64_t row0,
                            uint64_t seed, double log_offset, double bg_lo, double bg_hi,
                            double pi0, double noise, int warm_start,
                            int64_t *nnz_out);

/* Options of the fused step (defaults reproduce R's update count for EVERY entry):
 *   "delta"    0 (default): exact mode.  d > 0 (<= 1e-10): delta mode -- a unit whose entries all
 *              satisfy |f| < tol/16 and whose next Newton update |f/f'| <= d is frozen: its remaining
 *              updates are numerical no-ops and are skipped, so M and V differ from exact mode by at
 *              most ~d (absolute); the global update count n_updates is still R's.  DESIGN.md §3b.
 *   "freeze"   1 (default): once a converged unit's iterates repeat with period <= 2 (a floating-point fixed
 *              point or a 2-cycle of adjacent doubles) its update count is advanced by an even number at once;
 *              results are BIT-IDENTICAL to "freeze" 0, which performs every update (tests).
 *   "history"  1 (default): the first pass takes the tiles hardest-first by the previous step's per-tile counts
 *              (kept per context and shape), so K* is known within the first wave and the top-up pass is
 *              nearly empty; 0: natural order.  Results are bit-identical either way.
 *   "forget_history" (any value): drop the kept per-tile counts (the next step runs in natural order).
 *   "speculate" 1 (default) / 0, "pipe_chunks" (0 = automatic): the chunked host-buffer step described at
 *              ebpmf_vga_step_host.
 *   "debug_verify_scale" (0,1], default 1: TEST HOOK -- the top-up pass verifies against tol*scale, which
 *              forces the host's K*+1 retry loop to run (tests/test_gpu_parity.py::test_retry_loop_...). */
int  ebpmf_ctx_set_option(ebpmf_ctx *ctx, const char *name, double value);

/* diagnostics of the last fused solve: units (warp x 2 columns) that needed second-pass work and
 * the total number of unit evaluations (each = 64 entry evaluations of f) */
int  ebpmf_ctx_last_diag(ebpmf_ctx *ctx, uint64_t *topup_units, uint64_t *unit_evals);

/* counters of the context / its last fused solve: out[0] units that needed second-pass work, [1] unit
 * evaluations, [2] tiles the second pass restaged, [3] kernels launched by this context so far, [4] bytes the
 * last host-buffer step copied out speculatively, [5] bytes it had to copy again, [6] tiles per pass,
 * [7] 1 if per-tile counts of a previous step are kept, [8] device ms of the last ebpmf_vga_fixed_iter /
 * ebpmf_calc_split_obj kernel.
 * count <= 9. */
int  ebpmf_ctx_counters(ebpmf_ctx *ctx, double *out, int count);

/* device-time micro-probe of the FP64 pipe (DFMA chains), used by bench.py for the FP64 line */
int  ebpmf_probe_fp64_tflops(ebpmf_ctx *ctx, double *tflops);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* EBPMF_B200_H */
