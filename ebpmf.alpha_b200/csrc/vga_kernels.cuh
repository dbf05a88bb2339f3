// Hand-written sm_100a kernels for the VGA Newton hot path of ebpmf_log.
//
// What they replace (paths relative to the reference root):
//   a1  vga_pois_solver_mat_newton            R/vga_solver_mat.R:6-41
//   a2  Beta = fitted(fit_flash) = EL EF^T    R/ebpmf_log.R:266,305      (rebuilt in registers)
//   a3  adjust_var_shape                      R/ebpmf_log.R:475-486      (never materialised)
//   a4  sym / ssexp / slogv                   R/ebpmf_log.R:316-318
//   a5  v_sum = col/row/total sums of V       R/ebpmf_log.R:321-327
//   a10 vga_pois_solver_mat_newton_low_memory R/vga_solver_mat.R:71-109
//
// Layout.  M is column-major n x p (R), so consecutive THREADS own consecutive ROWS: a CTA
// of 256 threads owns a row block of 256 rows and walks a span of columns; every global
// access to M/V is a 2 KB contiguous segment per column.  Per 16-column tile the CTA stages,
// in shared memory, the dense image of the sparse Y tile (scattered from the CSC arrays through
// a per-row-block offset table) and the Beta tile: each thread loads its row of EL
// (K_tot <= KTP doubles, registers, only for this phase) and forms Beta_ij = sum_k EL_ik EF_jk
// against EF rows broadcast from shared memory -- K_tot FMAs per entry, an outer-product
// rebuild, not a tensor-core contraction.
//
// The GLOBAL stop rule.  R stops when max|f| < tol over ALL entries, so every entry receives
// the same number u of Newton updates.  The kernels reproduce u exactly with
// register-resident iterations and no per-iteration grid sync:
//   unit      = one warp x E columns (32*E entries), iterates in lock step;
//   pass 1    (vga_fused_kernel) each unit iterates until ITS entries all satisfy |f_k| < tol at
//             some k >= t0, where t0 is the largest such k any finished unit had published when
//             this unit started (never above u, because at k = u every unit is converged); it
//             finalises speculatively (M_k, const0, partial sums) and records k_unit; K* = max;
//   pass 2    (vga_topup_kernel, streaming: reads M_k and const0 back, no tile staging) units with
//             k_unit < K* continue the SAME iteration sequence up to K* updates, test
//             |f_K*| < tol and re-finalise.  If every unit passes, u = K* (it is the first k at
//             which all entries are below tol).  Otherwise the host repeats pass 2 with K*+1
//             (rare: needs an entry to bounce back above tol).
//   exhausted if a unit reaches maxiter updates, u = maxiter and V keeps the reciprocal of the
//             LAST EVALUATED temp, i.e. it is one iteration stale exactly as in R (F6).
// A NaN in f never converges, so its unit runs to maxiter and the NaN is seen in the last
// evaluated f: R's `if()` would have failed at the first NaN it tested, which is <= u.
#pragma once
#include "vga_math.cuh"

namespace ebpmf {

constexpr int kThreads = 256;
constexpr int kWarps = 8;
constexpr int kRowsPerCta = 256;
constexpr int kTileCols = 16;
constexpr double kSlogvConst = 0.9189385;   // literal of R/ebpmf_log.R:276,318

enum { FORMULA_A1 = 0, FORMULA_A10 = 1 };
enum { MODE_FIRST = 0, MODE_TOPUP = 1 };

struct Status {           // device words shared by the passes of one solve
    int kstar;            // max over units of k_unit (pass 1), target of pass 2
    int nan;              // a NaN reached the stop test
    int verify_fail;      // pass 2: some unit is not below tol at K*
    int exhausted;        // some unit hit maxiter
    unsigned long long topup_units;   // diagnostics: units that needed pass 2 work
    unsigned long long unit_iters;    // diagnostics: sum over units of evaluations performed
    int notconv0;         // ordered pipeline: some entry fails |f_0| < tol (so iteration 0's update counts)
    int pad0;
    // --- not reset per solve: (epoch << 32 | kstar), raised with atomicMax; what the OTHER GPUs read
    // over NVLink during pass 1 so that every rank learns the global count while it is still working.
    unsigned long long pub;
};
constexpr size_t kStatusResetBytes = 40;    // the words before `pub`

struct FusedArgs {
    long long n, p;                 // rows of this shard, columns
    int K;                          // K_tot
    int var_type;                   // 0 constant, 1 by_row, 2 by_col
    const double *M_in;             // pass 1 source
    double *M_out;                  // pass 1 destination, pass 2 in place
    double *V_out;                  // nullable
    double *C0;                     // n x p: const0 = (Sigma2*X + Beta) + 1, written by pass 1
    const double *EL;               // n x K column-major
    const double *EFt;              // p x ldk row-major, zero padded (ldk = KTP, or a multiple of 32 when K > 32)
    int ldk;
    const double *sigma2;           // length 1 | n | p
    const double *sc1, *sc2;        // 1/sigma2 (const1), sigma2/2 (const2), same length
    const double *l0, *f0;          // A10 only
    const int *rowidx;              // CSC row indices (0-based within the shard)
    const double *yx;               // CSC values
    const int *rbptr;               // (nRB+1) x p: first nnz of column j with row >= rb*256
    unsigned short *kunit;          // [nJG][nRW]
    long long nRW, nJG;             // ceil(n/32), ceil(p/E)
    double *colV;                   // [nRW][p]      sum over the warp's 32 rows of V
    double *unitS;                  // [nRW][nJG][2] sum exp(M+V/2), sum(log(V)/2+c) of the unit
    double *rowpart;                // [nJG][n] sum over the unit's columns of V (by_row only), else null
    float *score;                   // ordered / scored pipelines: per-tile max |f_0/f_0'| (stage A or vga_score_kernel)
    const int *perm;                // ordered / scored pipelines: tile ids by decreasing score (null: natural order)
    int nspans;                     // column spans per row block (grid.x of the natural decomposition)
    Status *status;
    const Status *const *peers;     // the other ranks' Status words, mapped peer memory (NVLink); may be null
    int npeers;
    unsigned epoch;                 // solve counter, identical on every rank
    const MathTables *tables;       // exp2 / log tables (device copy of the host-built tables)
    int maxiter;
    double tol;
    double delta;                   // 0: exact update count for every entry; > 0: delta mode
    double tol_verify;              // == tol; a test hook scales it to exercise the K*+1 retry path
    long long col0, col1;           // columns of this launch
    int cols_per_cta;               // multiple of kTileCols
};

// publish a unit's first-converged count locally and to the peers' view
__device__ __forceinline__ void publish_kstar(Status *st, unsigned epoch, int kb) {
    atomicMax(&st->kstar, kb);
    atomicMax(&st->pub, ((unsigned long long)epoch << 32) | (unsigned)kb);
}
// one peer's published word (a plain load over NVLink; 0 when there are no peers)
__device__ __forceinline__ unsigned long long peek_peer(const FusedArgs &a, unsigned which) {
    if (a.npeers <= 0) return 0ull;
    return *reinterpret_cast<const volatile unsigned long long *>(&a.peers[which % (unsigned)a.npeers]->pub);
}
// fold a peeked word into the local count if it belongs to this solve: a peer's K* is <= the global u
__device__ __forceinline__ void absorb_peer(Status *st, unsigned epoch, unsigned long long pv) {
    if ((unsigned)(pv >> 32) == epoch) {
        const int k = (int)(unsigned)pv;
        if (k > *reinterpret_cast<volatile int *>(&st->kstar)) atomicMax(&st->kstar, k);
    }
}

__device__ __forceinline__ double ld_stream(const double *p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream(double *p, double v) { __stcs(p, v); }

// ---------------------------------------------------------------------------------------
// Sum S (4 or 8) per-lane values over the 32 lanes with a transposing butterfly: each step
// halves the number of values a lane carries, so S values cost S-1+log2(32/S)... shuffles of a
// double instead of 5*S.  On return every lane holds the total of slot warp_slot<S>(lane).
// Fixed tree => deterministic.
// ---------------------------------------------------------------------------------------
template <int S> __device__ __forceinline__ int warp_slot(int lane) {
    return (S == 8) ? (((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1))
                    : (((lane >> 4) & 1) * 2 + ((lane >> 3) & 1));
}

template <int S> __device__ __forceinline__ double warp_multi_sum(const double (&v)[S], int lane) {
    static_assert(S == 4 || S == 8, "4 or 8 slots");
    const unsigned FULL = 0xffffffffu;
    double a[4];
    if (S == 8) {
        const bool hi = lane & 16;
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            const double keep = hi ? v[s + 4] : v[s], send = hi ? v[s] : v[s + 4];
            a[s] = keep + __shfl_xor_sync(FULL, send, 16);
        }
        const bool h8 = lane & 8;
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const double keep = h8 ? a[s + 2] : a[s], send = h8 ? a[s] : a[s + 2];
            a[s] = keep + __shfl_xor_sync(FULL, send, 8);
        }
        const bool h4 = lane & 4;
        const double keep = h4 ? a[1] : a[0], send = h4 ? a[0] : a[1];
        double t = keep + __shfl_xor_sync(FULL, send, 4);
        t += __shfl_xor_sync(FULL, t, 2);
        t += __shfl_xor_sync(FULL, t, 1);
        return t;
    } else {
        const bool hi = lane & 16;
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const double keep = hi ? v[s + 2] : v[s], send = hi ? v[s] : v[s + 2];
            a[s] = keep + __shfl_xor_sync(FULL, send, 16);
        }
        const bool h8 = lane & 8;
        const double keep = h8 ? a[1] : a[0], send = h8 ? a[0] : a[1];
        double t = keep + __shfl_xor_sync(FULL, send, 8);
        t += __shfl_xor_sync(FULL, t, 4);
        t += __shfl_xor_sync(FULL, t, 2);
        t += __shfl_xor_sync(FULL, t, 1);
        return t;
    }
}

// ---------------------------------------------------------------------------------------
// One unit: E entries per lane, all 32 lanes in lock step.  Per entry the loop carries
//   m, c0 = (s2*y + beta) + 1, w = (c0 - 1)/s2 = y + beta/s2, c1 = 1/s2, c2 = s2/2 [, sc = S]
// so that  f = x - sexp - M*const1 + const3  (R/vga_solver_mat.R:25) is  fma(-m, c1, w - sexp).
// tol_lane is tol for a live row and +Inf for a padding row (row >= n); a padding COLUMN (p not a
// multiple of E) is given the values of the lane's previous column, so it converges with it.
// `target` is t0 in the first pass (stop at the first k >= t0 at which the unit is converged)
// and K* in the top-up pass (stop exactly at k == K*).
// On return k is the number of updates the unit has received, r[] the reciprocal of the last
// evaluated temp, sexp[] / f[] the last evaluated S*exp(.) and f.
// Returns a bit mask: 1 converged at k, 2 exhausted, 8 frozen (delta mode only).
// ---------------------------------------------------------------------------------------
constexpr int kFrozen = 0xFFFF;     // kunit marker: unit is at its numerical fixed point (delta mode)

template <int E, int FORMULA, int MODE, bool SCALED, bool DELTA>
__device__ __forceinline__ int newton_unit(double (&m)[E], const double (&c0)[E], double (&w)[E],
                                           const double (&c1)[E], const double (&c2)[E], const double (&sc)[E],
                                           double tol_lane, int &k, int &kb, int target, int maxiter, double delta,
                                           unsigned exptab, double (&r)[E], double (&sexp)[E], double (&f)[E])
{
    int flags = 0;
    kb = -1;                        // first k at which the unit was converged (what K* is the max of)
    // keep w (and the lane tolerance) in registers: without this the compiler re-derives
    // (c0-1)*c1 in every iteration -- 2 extra FP64-pipe instructions per entry per iteration
#pragma unroll
    for (int e = 0; e < E; ++e) asm volatile("" : "+d"(w[e]));
    asm volatile("" : "+d"(tol_lane));
    asm volatile("" : "+r"(exptab));        // likewise the table's shared-window address
#pragma unroll 1
    for (;;) {
        bool lane_conv = true;
        double a[E], xx[E], ex[E];
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const double temp = __dadd_rn(c0[e], -m[e]);                       // :22
            r[e] = rcp_fast(temp);
            a[e] = __dmul_rn(c2[e], r[e]);                                     // const2/temp
            xx[e] = __dadd_rn(m[e], a[e]);
        }
        exp_fast_n<E>(xx, ex, exptab);
#pragma unroll
        for (int e = 0; e < E; ++e) {
            sexp[e] = SCALED ? __dmul_rn(sc[e], ex[e]) : ex[e];                // :23
            f[e] = __fma_rn(-m[e], c1[e], __dadd_rn(w[e], -sexp[e]));          // :25 | :84,:98
            lane_conv = lane_conv && (fabs(f[e]) < tol_lane);                  // :26 (strict <)
        }
        const bool conv = __all_sync(0xffffffffu, lane_conv);
        if (conv && kb < 0) kb = k;
        const bool stop = (MODE == MODE_FIRST) ? (conv && (k >= target)) : (k == target);
        if (!DELTA || !conv) {
            if (stop) { if (conv) flags |= 1; break; }
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const double q = (FORMULA == FORMULA_A1) ? a[e] : __dmul_rn(0.5, r[e]);   // const2/temp^2 | 0.5*temp^-2
                const double g = __fma_rn(-sexp[e], __fma_rn(q, r[e], 1.0), -c1[e]);      // :31 | :90,:105
                m[e] = __fma_rn(-f[e], rcp_fast(g), m[e]);
            }
        } else {
            // delta mode, unit converged: form the update and FREEZE the unit if every entry's update
            // is a numerical no-op (|f/g| <= delta, |f| < tol/16): all later iterates equal this one
            // to within delta, so the unit is final for any global count.
            double mn[E];
            bool lane_small = true;
            const double tol16 = tol_lane * 0.0625;
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const double q = (FORMULA == FORMULA_A1) ? a[e] : __dmul_rn(0.5, r[e]);
                const double g = __fma_rn(-sexp[e], __fma_rn(q, r[e], 1.0), -c1[e]);
                const double rg = rcp_fast(g);
                mn[e] = __fma_rn(-f[e], rg, m[e]);
                lane_small = lane_small && (fabs(__dmul_rn(f[e], rg)) <= delta) && (fabs(f[e]) < tol16);
            }
            if (__all_sync(0xffffffffu, lane_small)) { flags |= 1 | 8; break; }
            if (stop) { flags |= 1; break; }
#pragma unroll
            for (int e = 0; e < E; ++e) m[e] = mn[e];
        }
        ++k;
        if (k >= maxiter) { flags |= 2; break; }
    }
    return flags;
}

// ---------------------------------------------------------------------------------------
// Finalise one unit of the fused path: store M (and V), and for the A1 step the unit's
// partial sums: sum_rows V per column -> colV[rw][j]; sum exp(M+V/2), sum(log(V)/2+c) ->
// unitS[rw][jg][0..1]; by_row: sum_cols V -> rowpart[jg][i].  sum(Y*M) is taken over the
// non-zeros only by sym_nnz_kernel once M is final.  Returns true if a live entry's f is NaN.
// ---------------------------------------------------------------------------------------
template <int E, int FORMULA>
__device__ __forceinline__ bool finalize_unit(const FusedArgs &a, const double (&m)[E], const double (&c2)[E],
                                              const double (&r)[E], const double (&sexp)[E], const double (&f)[E],
                                              const bool (&ok)[E], int flags,
                                              size_t off,      // i + n*jb : entry (i, jb) of the n x p matrices
                                              size_t cv_idx,   // rw*p + jb : slot of (rw, jb) in colV
                                              size_t us_idx,   // (rw*nJG + jg)*2 : slot of the unit in unitS
                                              size_t rp_idx,   // jg*n + i : slot in rowpart
                                              int cols_left,   // p - jb
                                              int lane, bool row_ok, unsigned exptab, unsigned logtab)
{
    constexpr int S = (E + 2 <= 4) ? 4 : 8;
    static_assert(E + 2 <= 8, "at most 6 columns per unit");
    double vals[S];
#pragma unroll
    for (int q = 0; q < S; ++q) vals[q] = 0.0;
    double rowv = 0.0, vprod = 1.0;
    int nlive = 0;
    bool lane_nan = false;
    const size_t n = (size_t)a.n;
#pragma unroll
    for (int e = 0; e < E; ++e) {
        lane_nan = lane_nan || (ok[e] && (f[e] != f[e]));
        if (ok[e]) st_stream(a.M_out + off + e * n, m[e]);
        if (FORMULA == FORMULA_A1) {
            const double v = __dmul_rn(__dadd_rn(c2[e], c2[e]), r[e]);          // :36 Sigma2/temp (stale if exhausted)
            if (ok[e] && a.V_out) st_stream(a.V_out + off + e * n, v);
            double ex = sexp[e];                                                // == exp(M+V/2): S == 1 and the loop broke
            if (flags & 2) ex = exp_fast(__fma_rn(0.5, v, m[e]), exptab);       // exhausted: M moved after temp
            if (ok[e]) { vals[e] = v; vals[E] += ex; rowv += v; vprod *= v; ++nlive; }
        }
    }
    if (FORMULA == FORMULA_A1) {
        // sum_e (log(V_e)/2 + c) = log(prod_e V_e)/2 + nlive*c : one log per lane instead of E
        // (V in (0, sigma2]; a product of E <= 4 such values stays far from underflow)
        vals[E + 1] = (nlive > 0) ? fma(0.5, log_fast(vprod, logtab), kSlogvConst * nlive) : 0.0;   // R/ebpmf_log.R:318
    }
    if (FORMULA == FORMULA_A1) {
        const double tot = warp_multi_sum<S>(vals, lane);
        const int slot = warp_slot<S>(lane);
        if ((lane & ((S == 8) ? 3 : 7)) == 0) {                                 // one writer per slot
            if (slot < E) { if (slot < cols_left) a.colV[cv_idx + slot] = tot; }
            else if (slot < E + 2) a.unitS[us_idx + (slot - E)] = tot;
        }
        if (a.rowpart && row_ok) a.rowpart[rp_idx] = rowv;
    }
    return __any_sync(0xffffffffu, lane_nan);
}

// stage EF rows [j0, j0+16) x factor columns [k0, k0+KTP) into efs[16][KTP] (zero beyond jc1)
template <int KTP>
__device__ __forceinline__ void stage_ef_chunk(double *efs, const double *EFt, int ldk, int j0, int jc1, int k0, int tid) {
    for (int idx = tid; idx < kTileCols * KTP; idx += kThreads) {
        const int c = idx / KTP, k = idx - c * KTP;
        efs[idx] = (j0 + c < jc1) ? __ldg(EFt + (size_t)(j0 + c) * ldk + k0 + k) : 0.0;
    }
}

// shared memory of the first-pass kernel (dynamic): Y image, Beta image, EF rows, column constants
template <int KTP>
struct FusedSmem {
    double ys[kTileCols][kRowsPerCta];
    double bs[kTileCols][kRowsPerCta];
    double efs[kTileCols][KTP];
    double c1s[kTileCols], c2s[kTileCols], f0s[kTileCols];
    double exptab[EBPMF_EXP_TABLE_SIZE];
    double2 logtab[EBPMF_LOG_TABLE_SIZE];
    float wmax[kWarps];
    int wnot[kWarps];
};

#ifndef EBPMF_MINB
#define EBPMF_MINB 3          // resident CTAs per SM the register allocation is bounded for
#endif

// ---------------------------------------------------------------------------------------
// K1 / K4, pass 1: fused VGA step on CSC Y with Beta rebuilt from the flash factors.
// grid = (column spans, row blocks), block = 256 threads, dynamic smem = sizeof(FusedSmem<KTP>).
// ---------------------------------------------------------------------------------------
// ITER0_ONLY (ordered pipeline, stage A): the same tile staging and prologue, but each unit performs
// exactly Newton iteration 0 -- evaluate f_0, record |f_0| < tol, form and apply the update -- and
// writes M_1 and const0; no finalisation.  The CTA's largest first step max|f_0/f_0'| goes to
// a.score: it predicts where the slowest entries are (the CTA holding the entry that sets the global
// count ranks in the first few % by this score on every case tried, cold or warm start).
template <int KTP, int E, int FORMULA, bool DELTA, bool ITER0_ONLY = false, bool PERM = false>
__global__ void __launch_bounds__(kThreads, EBPMF_MINB) vga_fused_kernel(const FusedArgs a)
{
    static_assert(kTileCols % E == 0, "E must divide the tile width");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FusedSmem<KTP> &sm = *reinterpret_cast<FusedSmem<KTP> *>(smem_raw);

    // rows of a shard and columns fit 32 bits (dgCMatrix Dim is int); only n*j products are 64-bit
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = (int)a.n, p = (int)a.p;
    // scored pipeline: the CTAs take the (column span, row block) tiles by decreasing predicted
    // hardness (a.perm, from vga_score_kernel); CTAs are dispatched in linear order, x fastest
    // (PERM is a template parameter so that the default instantiation keeps its code unchanged)
    int span = blockIdx.x, rb = blockIdx.y;
    if (PERM) {
        const int id = a.perm[blockIdx.y * gridDim.x + blockIdx.x];
        rb = id / (int)gridDim.x;
        span = id - rb * (int)gridDim.x;
    }
    const int r0 = rb * kRowsPerCta;
    const int i = r0 + tid;
    const bool row_ok = i < n;
    const int rw = rb * kWarps + warp;
    const bool warp_ok = rw < (int)a.nRW;            // false: every row of this warp is beyond n
    const int jc0 = (int)a.col0 + span * a.cols_per_cta;
    const int jc1 = min((int)a.col1, jc0 + a.cols_per_cta);

    exp_table_init(sm.exptab, a.tables, tid, kThreads);
    log_table_init(sm.logtab, a.tables, tid, kThreads);
    const unsigned exptab_s = smem_addr(sm.exptab), logtab_s = smem_addr(sm.logtab);

    // row-constant pieces (by_row / constant sigma2, A10's l0)
    double c1_row = 1.0, c2_row = 0.5, l0_i = 1.0;
    if (a.var_type != 2) {
        const int si = (a.var_type == 1) ? (row_ok ? i : 0) : 0;
        c1_row = a.sc1[si];                                                    // :10
        c2_row = a.sc2[si];                                                    // :11
    }
    if (FORMULA == FORMULA_A10) l0_i = row_ok ? a.l0[i] : 1.0;

    const volatile int *kstar_live = &a.status->kstar;
    int any_exhausted = 0, any_nan = 0;
    unsigned my_iters = 0;
    const size_t ns = (size_t)n;
    const bool by_col = a.var_type == 2;
    int t0_next = ITER0_ONLY ? 0 : *kstar_live;
    double it0_step = 0.0;                           // ITER0_ONLY: max |f_0/f_0'| of this thread's entries
    int it0_notconv = 0;

    for (int j0 = jc0; j0 < jc1; j0 += kTileCols) {
        __syncthreads();                             // previous tile fully consumed
        // one thread per CTA peeks at one peer GPU's published count (round robin); the word is
        // consumed after the tile's units, so the NVLink latency is never waited for
        unsigned long long peer_word = 0ull;
        if (tid == 0 && !ITER0_ONLY) peer_word = peek_peer(a, (unsigned)(j0 / kTileCols) + blockIdx.x + blockIdx.y);
        // ---- stage the tile: zero Y image, EF rows, per-column constants --------------------
#pragma unroll
        for (int c = 0; c < kTileCols; ++c) sm.ys[c][tid] = 0.0;
        stage_ef_chunk<KTP>(&sm.efs[0][0], a.EFt, a.ldk, j0, jc1, 0, tid);
        if (tid < kTileCols) {
            const int j = j0 + tid;
            const bool col = by_col && (j < jc1);
            sm.c1s[tid] = col ? a.sc1[j] : 1.0;
            sm.c2s[tid] = col ? a.sc2[j] : 0.5;
            sm.f0s[tid] = (FORMULA == FORMULA_A10 && j < jc1) ? a.f0[j] : 1.0;
        }
        // first unit's M values: in flight while the tile is staged
        size_t off = (size_t)i + ns * (size_t)j0;    // entry (i, j0)
        double mnext[E];
#pragma unroll
        for (int e = 0; e < E; ++e)
            mnext[e] = (row_ok && (j0 + e < jc1)) ? ld_stream(a.M_in + off + e * ns) : 0.0;
        __syncthreads();
        // scatter the CSC non-zeros of rows [r0, r0+256) of each tile column
        {
            const int *rp0 = a.rbptr + (size_t)rb * p + j0, *rp1 = rp0 + p;
            for (int c = warp; c < kTileCols; c += kWarps) {
                if (j0 + c < jc1) {
                    const int lo = rp0[c], hi = rp1[c];
                    for (int q = lo + lane; q < hi; q += 32)
                        sm.ys[c][a.rowidx[q] - r0] = __ldg(a.yx + q);
                }
            }
        }
        // a2: this row's Beta for the 16 tile columns, K_tot FMAs each (EL row lives in registers
        // only here; EF rows are broadcast reads from shared memory).  K_tot > KTP (only possible
        // with KTP = 32) walks the factor columns in chunks of KTP, restaging the EF rows.
        for (int k0 = 0; k0 < a.K; k0 += KTP) {
            if (k0 > 0) {
                __syncthreads();
                stage_ef_chunk<KTP>(&sm.efs[0][0], a.EFt, a.ldk, j0, jc1, k0, tid);
                __syncthreads();
            }
            double el[KTP];
            const double *elp = a.EL + i + (size_t)k0 * ns;
#pragma unroll
            for (int k = 0; k < KTP; ++k) el[k] = (row_ok && k0 + k < a.K) ? __ldg(elp + k * ns) : 0.0;
#pragma unroll 4
            for (int c = 0; c < kTileCols; ++c) {
                double b = (k0 > 0) ? sm.bs[c][tid] : 0.0;
#pragma unroll
                for (int k = 0; k < KTP; ++k) b = fma(el[k], sm.efs[c][k], b);
                sm.bs[c][tid] = b;
            }
        }
        __syncthreads();

        // ---- units: running offsets, no 64-bit multiplies in the loop ------------------------
        const int jg0 = j0 / E;
        size_t ku_idx = (size_t)jg0 * (size_t)a.nRW + rw;          // kunit[jg][rw]
        size_t cv_idx = (size_t)rw * p + j0;                        // colV[rw][j]
        size_t us_idx = ((size_t)rw * (size_t)a.nJG + jg0) * 2;     // unitS[rw][jg][2]
        size_t rp_idx = (size_t)jg0 * ns + i;                       // rowpart[jg][i]
#pragma unroll 1
        for (int g = 0; g < kTileCols / E; ++g) {
            const int jb = j0 + g * E;
            if (jb >= jc1 || !warp_ok) break;
            // t0: the largest converged-at count any finished unit had published when the PREVIOUS
            // unit of this warp started (read one unit ahead so the L2 latency is off the critical
            // path).  It never exceeds the global count u, so running this unit at least that far
            // cannot overshoot.
            const int t0 = t0_next;
            if (!ITER0_ONLY) t0_next = *kstar_live;
            double m[E], c0[E], w[E], c1[E], c2[E], sc[E], r[E], sexp[E], f[E];
            bool ok[E];
            const bool more = (g + 1 < kTileCols / E);
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int c = g * E + e;
                const bool colok = jb + e < jc1;
                ok[e] = row_ok && colok;
                m[e] = mnext[e];
                mnext[e] = (row_ok && more && (jb + E + e < jc1)) ? ld_stream(a.M_in + off + (E + e) * ns) : 0.0;
                const double y = sm.ys[c][tid], b = sm.bs[c][tid];
                if (by_col) { c1[e] = sm.c1s[c]; c2[e] = sm.c2s[c]; }
                else { c1[e] = c1_row; c2[e] = c2_row; }
                c0[e] = __dadd_rn(__fma_rn(__dadd_rn(c2[e], c2[e]), y, b), 1.0);   // :9 | :81,:95
                const double c0m1 = __dadd_rn(c0[e], -1.0);
                w[e] = __dmul_rn(c0m1, c1[e]);                                     // x + const3
                sc[e] = (FORMULA == FORMULA_A10) ? __dmul_rn(l0_i, sm.f0s[c]) : 1.0;   // :83,:97 | S = 1
                const double cap = (FORMULA == FORMULA_A1) ? c0m1 : b;             // :16 | :78
                m[e] = (m[e] < cap) ? m[e] : cap;
                if (ok[e]) st_stream(a.C0 + off + e * ns, c0[e]);
                if (e > 0 && !colok) {          // padding column (p % E != 0): shadow the previous one
                    m[e] = m[e - 1]; c0[e] = c0[e - 1]; w[e] = w[e - 1]; c1[e] = c1[e - 1]; c2[e] = c2[e - 1];
                    sc[e] = sc[e - 1];
                }
            }
            if (ITER0_ONLY) {
                // iteration 0 with newton_unit's instruction sequence (bit-identical iterates)
                double aa[E], xx[E], ex[E];
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    r[e] = rcp_fast(__dadd_rn(c0[e], -m[e]));                      // :22
                    aa[e] = __dmul_rn(c2[e], r[e]);
                    xx[e] = __dadd_rn(m[e], aa[e]);
                }
                exp_fast_n<E>(xx, ex, exptab_s);
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    f[e] = __fma_rn(-m[e], c1[e], __dadd_rn(w[e], -ex[e]));        // :25 (S = 1)
                    const double g = __fma_rn(-ex[e], __fma_rn(aa[e], r[e], 1.0), -c1[e]);   // :31
                    const double rg = rcp_fast(g);
                    if (ok[e]) {
                        if (!(fabs(f[e]) < a.tol)) it0_notconv = 1;                // :26
                        it0_step = fmax(it0_step, fabs(__dmul_rn(f[e], rg)));
                        st_stream(a.M_out + off + e * ns, __fma_rn(-f[e], rg, m[e]));       // M_1
                    }
                }
                off += E * ns; ku_idx += (size_t)a.nRW; cv_idx += E; us_idx += 2; rp_idx += ns;
                continue;
            }
            int k = 0, kb;
            const int flags = newton_unit<E, FORMULA, MODE_FIRST, FORMULA == FORMULA_A10, DELTA>(
                m, c0, w, c1, c2, sc, row_ok ? a.tol : INFINITY, k, kb, t0, a.maxiter, a.delta, exptab_s, r, sexp, f);
            my_iters += (unsigned)(k + 1);
            if (flags & 2) { any_exhausted = 1; kb = k; }
            if (finalize_unit<E, FORMULA>(a, m, c2, r, sexp, f, ok, flags, off, cv_idx, us_idx, rp_idx, p - jb, lane,
                                          row_ok, exptab_s, logtab_s))
                any_nan = 1;
            if (lane == 0) {
                // K* = max over units of the count at which the unit stopped (its first converged k >= t0, which
                // is <= u), or of the first converged k for a frozen unit: every live unit then has k_unit <= K*
                const int kpub = (flags & 8) ? kb : k;
                a.kunit[ku_idx] = (unsigned short)((flags & 8) ? kFrozen : k);
                if (kpub > t0) publish_kstar(a.status, a.epoch, kpub);
            }
            off += E * ns; ku_idx += (size_t)a.nRW; cv_idx += E; us_idx += 2; rp_idx += ns;
        }
        if (tid == 0 && !ITER0_ONLY) absorb_peer(a.status, a.epoch, peer_word);
    }
    if (ITER0_ONLY) {
        float smx = (float)fmin(it0_step, 3.0e38);
        int nc = it0_notconv;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            smx = fmaxf(smx, __shfl_xor_sync(0xffffffffu, smx, o));
            nc |= __shfl_xor_sync(0xffffffffu, nc, o);
        }
        __syncthreads();
        if (lane == 0) { sm.wmax[warp] = smx; sm.wnot[warp] = nc; }
        __syncthreads();
        if (tid == 0) {
            float mx = 0.f; int any = 0;
            for (int q = 0; q < kWarps; ++q) { mx = fmaxf(mx, sm.wmax[q]); any |= sm.wnot[q]; }
            a.score[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = mx;
            if (any) a.status->notconv0 = 1;
        }
        return;
    }
    if (lane == 0) {
        if (any_exhausted) a.status->exhausted = 1;
        if (any_nan) a.status->nan = 1;
        if (my_iters) atomicAdd(&a.status->unit_iters, (unsigned long long)my_iters);
    }
}

// ---------------------------------------------------------------------------------------
// Scored pipeline, stage 0: rank the (column span, row block) tiles of the first pass by predicted
// hardness.  The score is the tile's largest FIRST Newton step max|f_0/f_0'| (R/vga_solver_mat.R:22-31
// evaluated once), computed in FP32 from the high words of M: it only decides the ORDER in which the
// fused kernel takes the tiles, never a result, so single precision and approximate MUFU functions
// are enough and the kernel runs near the rate at which M streams from HBM.  Same grid and tile
// decomposition as vga_fused_kernel; writes a.score[row block * spans + span].  FORMULA_A1 only.
// ---------------------------------------------------------------------------------------
constexpr int kMaxSpanCols = 256;      // cols_per_cta never exceeds this (cols_per_cta_for in ebpmf_api.cu)

constexpr int kScorePad = 4;            // row padding (floats) that keeps the fragment-shaped smem reads conflict-free

template <int KTP>
struct ScoreSmem {
    float efs[kMaxSpanCols][KTP + kScorePad];   // EF rows of the CTA's whole column span, rounded to tf32
    float c1s[kMaxSpanCols], c2s[kMaxSpanCols];
    float ys[2][kTileCols][kRowsPerCta + kScorePad];   // double-buffered Y image of a 256-row x 16-column tile
    float wmax[kWarps];
};

// float from the high word of a double (20 mantissa bits kept; |x| < 2^-126 -> 0, huge -> garbage/Inf,
// which the caller's clamp turns into "hard")
__device__ __forceinline__ float float_from_hi(unsigned hi) {
    const unsigned e = hi & 0x7fffffffu;
    const unsigned bits = (hi & 0x80000000u) | ((e - 0x38000000u) << 3);
    return (e < 0x38100000u) ? 0.f : __uint_as_float(bits);
}

__device__ __forceinline__ float rcp_approx_f32(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ float ex2_approx_f32(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ unsigned to_tf32(float x) {
    unsigned r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}

// D(16x8) += A(16x8, row) * B(8x8, col), tf32 inputs, f32 accumulate.  Fragment layout (g = lane >> 2,
// t = lane & 3): a0 (g, t) a1 (g+8, t) a2 (g, t+4) a3 (g+8, t+4); b0 (k = t, n = g) b1 (k = t+4, n = g);
// d0 (g, 2t) d1 (g, 2t+1) d2 (g+8, 2t) d3 (g+8, 2t+1).
__device__ __forceinline__ void mma_tf32_16x8x8(float (&d)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Scoring is a heuristic (it decides an ORDER, never a result), so this is the one place where the
// factor product runs on the tensor cores: per tile each warp forms Beta for its 32 rows x 16 columns
// with 2 x 2 x KTP/8 tf32 MMAs (EL fragments live in registers for the whole CTA, EF fragments come from
// shared memory) instead of 16 x KTP FFMAs per thread.  A lane then owns the 16 entries of the
// accumulator layout: rows 16*mb + 8*h + g, columns 8*nb + 2*t + w.  The Y image is double-buffered: the
// next tile's non-zeros are loaded before the arithmetic and scattered after it (two columns per warp),
// one barrier per tile; a thread zeroes each Y element right after reading it (it is the only reader).
// K_tot <= KTP (the host skips scoring otherwise).
#ifndef EBPMF_SCORE_MINB
#define EBPMF_SCORE_MINB 3     // measured: 2 CTAs/SM with a tile-ahead register prefetch of M is slower (9.6 vs 8.6 ms)
#endif
template <int KTP, bool BYCOL>
__global__ void __launch_bounds__(kThreads, EBPMF_SCORE_MINB) vga_score_kernel(const FusedArgs a)
{
    constexpr int KB = KTP / 8;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ScoreSmem<KTP> &sm = *reinterpret_cast<ScoreSmem<KTP> *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int n = (int)a.n, p = (int)a.p;
    const int rb = blockIdx.y;
    const int r0 = rb * kRowsPerCta;
    const int jc0 = (int)a.col0 + blockIdx.x * a.cols_per_cta;
    const int jc1 = min((int)a.col1, jc0 + a.cols_per_cta);
    const int span = jc1 - jc0;
    const int span16 = (span + kTileCols - 1) / kTileCols * kTileCols;
    const size_t ns = (size_t)n;

    for (int idx = tid; idx < span16 * KTP; idx += kThreads) {
        const int c = idx / KTP, k = idx - c * KTP;
        const float v = (c < span && k < a.K) ? (float)__ldg(a.EFt + (size_t)(jc0 + c) * a.ldk + k) : 0.f;
        sm.efs[c][k] = __uint_as_float(to_tf32(v));
    }
    if (BYCOL) {
        for (int c = tid; c < span16; c += kThreads) {
            sm.c1s[c] = (c < span) ? (float)a.sc1[jc0 + c] : 1.f;
            sm.c2s[c] = (c < span) ? (float)a.sc2[jc0 + c] : 0.5f;
        }
    }
#pragma unroll
    for (int c = 0; c < kTileCols; ++c) { sm.ys[0][c][tid] = 0.f; sm.ys[1][c][tid] = 0.f; }

    // this lane's four rows: lr[mb][h] = warp*32 + 16*mb + 8*h + g (within the row block)
    int lrow[2][2];
    bool rok[2][2];
    float c1r[2][2], c2r[2][2];
    unsigned afrag[2][KB][4];
#pragma unroll
    for (int mb = 0; mb < 2; ++mb) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            lrow[mb][h] = warp * 32 + 16 * mb + 8 * h + g;
            rok[mb][h] = r0 + lrow[mb][h] < n;
            c1r[mb][h] = 1.f; c2r[mb][h] = 0.5f;
            if (!BYCOL) {
                const int si = (a.var_type == 1) ? (rok[mb][h] ? r0 + lrow[mb][h] : 0) : 0;
                c1r[mb][h] = (float)a.sc1[si];
                c2r[mb][h] = (float)a.sc2[si];
            }
        }
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int h = q & 1, k = 8 * kb + t + 4 * (q >> 1);
                const float v = (rok[mb][h] && k < a.K) ? (float)__ldg(a.EL + (size_t)(r0 + lrow[mb][h]) + (size_t)k * ns) : 0.f;
                afrag[mb][kb][q] = to_tf32(v);
            }
        }
    }
    __syncthreads();

    const int *rpb = a.rbptr + (size_t)rb * p;            // CSC offsets of this row block, per column
    auto scatter_range = [&](int j, int &lo, int &hi) {
        lo = hi = 0;
        if (j < jc1) { lo = rpb[j]; hi = rpb[p + j]; }
    };
    {   // tile 0 -> buffer 0; this warp scatters tile columns `warp` and `warp + 8`
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            int lo, hi;
            scatter_range(jc0 + warp + 8 * h, lo, hi);
            for (int q = lo + lane; q < hi; q += 32) sm.ys[0][warp + 8 * h][a.rowidx[q] - r0] = (float)__ldg(a.yx + q);
        }
    }
    __syncthreads();

    // high words of M: column (8*nb + 2*t + w) of the tile, rows lrow[mb][h] -> 4 column pointers per tile,
    // row offsets are compile-time multiples of 8 rows
    const unsigned *mrow = reinterpret_cast<const unsigned *>(a.M_in + (size_t)r0 + (size_t)(warp * 32 + g)) + 1;
    float best = 0.f;
    int cur = 0;
    for (int j0 = jc0; j0 < jc1; j0 += kTileCols, cur ^= 1) {
        unsigned mhi[2][2][2][2];                         // [nb][w][mb][h]
        bool cok[2][2];
#pragma unroll
        for (int nb = 0; nb < 2; ++nb) {
#pragma unroll
            for (int w = 0; w < 2; ++w) {
                const int col = j0 + 8 * nb + 2 * t + w;
                cok[nb][w] = col < jc1;
                const unsigned *mc = mrow + 2 * ns * (size_t)col;
#pragma unroll
                for (int mb = 0; mb < 2; ++mb) {
#pragma unroll
                    for (int h = 0; h < 2; ++h)
                        mhi[nb][w][mb][h] = (cok[nb][w] && rok[mb][h]) ? __ldcs(mc + 2 * (16 * mb + 8 * h)) : 0u;
                }
            }
        }
        // next tile's non-zeros: the first 32 of each of this warp's two columns go through registers
        int nlo[2], nhi[2], nrow[2];
        float nval[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            scatter_range(j0 + kTileCols + warp + 8 * h, nlo[h], nhi[h]);
            const int q = nlo[h] + lane;
            nrow[h] = -1; nval[h] = 0.f;
            if (q < nhi[h]) { nrow[h] = a.rowidx[q] - r0; nval[h] = (float)__ldg(a.yx + q); }
        }
        const int cs0 = j0 - jc0;
#pragma unroll
        for (int nb = 0; nb < 2; ++nb) {
            float acc[2][4];
#pragma unroll
            for (int mb = 0; mb < 2; ++mb) { acc[mb][0] = acc[mb][1] = acc[mb][2] = acc[mb][3] = 0.f; }
#pragma unroll
            for (int kb = 0; kb < KB; ++kb) {
                const unsigned b0 = __float_as_uint(sm.efs[cs0 + 8 * nb + g][8 * kb + t]);
                const unsigned b1 = __float_as_uint(sm.efs[cs0 + 8 * nb + g][8 * kb + t + 4]);
                mma_tf32_16x8x8(acc[0], afrag[0][kb], b0, b1);
                mma_tf32_16x8x8(acc[1], afrag[1][kb], b0, b1);
            }
#pragma unroll
            for (int w = 0; w < 2; ++w) {
                const int c = 8 * nb + 2 * t + w;                 // column within the tile
                const float c1c = BYCOL ? sm.c1s[cs0 + c] : 1.f, c2c = BYCOL ? sm.c2s[cs0 + c] : 0.5f;
#pragma unroll
                for (int mb = 0; mb < 2; ++mb) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const float c1 = BYCOL ? c1c : c1r[mb][h], c2 = BYCOL ? c2c : c2r[mb][h];
                        const float y = sm.ys[cur][c][lrow[mb][h]];
                        sm.ys[cur][c][lrow[mb][h]] = 0.f;                          // clean for the tile after next
                        const float cap = fmaf(c2 + c2, y, acc[mb][2 * h + w]);    // const0 - 1
                        const float m = fminf(float_from_hi(mhi[nb][w][mb][h]), cap);   // :16
                        const float r = rcp_approx_f32((cap + 1.f) - m);           // :22
                        const float aa = c2 * r;
                        const float ex = ex2_approx_f32((m + aa) * 1.4426950408889634f);
                        const float f = fmaf(-m, c1, fmaf(cap, c1, -ex));          // :25
                        const float gg = fmaf(-ex, fmaf(aa, r, 1.f), -c1);         // :31
                        float step = fabsf(f * rcp_approx_f32(gg));
                        step = (step <= 3.0e38f) ? step : 3.0e38f;                 // Inf/NaN: take this tile early
                        if (cok[nb][w] && rok[mb][h]) best = fmaxf(best, step);
                    }
                }
            }
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            if (nrow[h] >= 0) sm.ys[cur ^ 1][warp + 8 * h][nrow[h]] = nval[h];
            for (int q = nlo[h] + lane + 32; q < nhi[h]; q += 32)                  // columns with > 32 non-zeros here
                sm.ys[cur ^ 1][warp + 8 * h][a.rowidx[q] - r0] = (float)__ldg(a.yx + q);
        }
        __syncthreads();
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, o));
    if (lane == 0) sm.wmax[warp] = best;
    __syncthreads();
    if (tid == 0) {
        float mx = 0.f;
        for (int q = 0; q < kWarps; ++q) mx = fmaxf(mx, sm.wmax[q]);
        a.score[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = mx;
    }
}

// ---------------------------------------------------------------------------------------
// Split pipeline, stage A: const0 = (Sigma2*Y + Beta) + 1 for every entry (R/vga_solver_mat.R:9,
// :81, :95), Beta rebuilt from the factors against the staged CSC tile.  Writes C0 (n x p).  For
// the low_memory formula it also applies the clamp M = Pmin(M, Beta) (:78), M_in -> M_out, because
// Beta alone cannot be recovered from const0 later.
// ---------------------------------------------------------------------------------------
template <int KTP>
struct PrepSmem {
    double ys[kTileCols][kRowsPerCta];
    double efs[kTileCols][KTP];
    double s2s[kTileCols];
};

template <int KTP, int FORMULA>
__global__ void __launch_bounds__(kThreads, 3) vga_prep_kernel(const FusedArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PrepSmem<KTP> &sm = *reinterpret_cast<PrepSmem<KTP> *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = (int)a.n, p = (int)a.p;
    const int rb = blockIdx.y;
    const int r0 = rb * kRowsPerCta;
    const int i = r0 + tid;
    const bool row_ok = i < n;
    const int jc0 = (int)a.col0 + blockIdx.x * a.cols_per_cta;
    const int jc1 = min((int)a.col1, jc0 + a.cols_per_cta);
    const size_t ns = (size_t)n;
    const bool by_col = a.var_type == 2;
    double s2_row = 1.0;
    if (!by_col) {
        const int si = (a.var_type == 1) ? (row_ok ? i : 0) : 0;
        s2_row = __dadd_rn(a.sc2[si], a.sc2[si]);
    }
    for (int j0 = jc0; j0 < jc1; j0 += kTileCols) {
        __syncthreads();
#pragma unroll
        for (int c = 0; c < kTileCols; ++c) sm.ys[c][tid] = 0.0;
        stage_ef_chunk<KTP>(&sm.efs[0][0], a.EFt, a.ldk, j0, jc1, 0, tid);
        if (tid < kTileCols) {
            const int j = j0 + tid;
            sm.s2s[tid] = (by_col && j < jc1) ? __dadd_rn(a.sc2[j], a.sc2[j]) : 1.0;    // sigma2 = 2*const2, exact
        }
        __syncthreads();
        {
            const int *rp0 = a.rbptr + (size_t)rb * p + j0, *rp1 = rp0 + p;
            for (int c = warp; c < kTileCols; c += kWarps) {
                if (j0 + c < jc1) {
                    const int lo = rp0[c], hi = rp1[c];
                    for (int q = lo + lane; q < hi; q += 32)
                        sm.ys[c][a.rowidx[q] - r0] = __ldg(a.yx + q);
                }
            }
        }
        // Beta accumulates in registers over chunks of KTP factor columns (one chunk when K_tot <= KTP)
        double beta[kTileCols];
#pragma unroll
        for (int c = 0; c < kTileCols; ++c) beta[c] = 0.0;
        for (int k0 = 0; k0 < a.K; k0 += KTP) {
            if (k0 > 0) {
                __syncthreads();
                stage_ef_chunk<KTP>(&sm.efs[0][0], a.EFt, a.ldk, j0, jc1, k0, tid);
                __syncthreads();
            }
            double el[KTP];
            const double *elp = a.EL + i + (size_t)k0 * ns;
#pragma unroll
            for (int k = 0; k < KTP; ++k) el[k] = (row_ok && k0 + k < a.K) ? __ldg(elp + k * ns) : 0.0;
#pragma unroll
            for (int c = 0; c < kTileCols; ++c) {
#pragma unroll
                for (int k = 0; k < KTP; ++k) beta[c] = fma(el[k], sm.efs[c][k], beta[c]);          // a2: Beta_ij
            }
        }
        __syncthreads();
        size_t off = (size_t)i + ns * (size_t)j0;
#pragma unroll
        for (int c = 0; c < kTileCols; ++c, off += ns) {
            const double b = beta[c];
            if (row_ok && j0 + c < jc1) {
                const double s2 = by_col ? sm.s2s[c] : s2_row;
                const double c0 = __dadd_rn(__fma_rn(s2, sm.ys[c][tid], b), 1.0);         // :9 | :81,:95
                st_stream(a.C0 + off, c0);
                if (FORMULA == FORMULA_A10) {
                    const double m0 = ld_stream(a.M_in + off);
                    st_stream(a.M_out + off, (m0 < b) ? m0 : b);                     // :78
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// Split pipeline, stage B (MODE_FIRST) and the top-up pass of both pipelines (MODE_TOPUP):
// streaming Newton.  Reads M and const0, no tile staging, no barrier after the table set-up.
// Same (column span, row block) decomposition as the other kernels so unit indices agree.
// ---------------------------------------------------------------------------------------
#ifndef EBPMF_STREAM_MINB
#define EBPMF_STREAM_MINB 3
#endif
template <int E, int FORMULA, int MODE, bool DELTA>
__global__ void __launch_bounds__(kThreads, EBPMF_STREAM_MINB) vga_stream_kernel(const FusedArgs a)
{
    __shared__ double exptab[EBPMF_EXP_TABLE_SIZE];
    __shared__ double2 logtab[EBPMF_LOG_TABLE_SIZE];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = (int)a.n, p = (int)a.p;
    // ordered pipeline: a 1-D grid walks the CTAs by decreasing predicted hardness (a.perm)
    int bx = blockIdx.x, rb = blockIdx.y;
    if (MODE == MODE_FIRST && a.perm) {
        const int id = a.perm[blockIdx.x];
        rb = id / a.nspans;
        bx = id - rb * a.nspans;
    }
    const int i = rb * kRowsPerCta + tid;
    const bool row_ok = i < n;
    const int rw = rb * kWarps + warp;
    exp_table_init(exptab, a.tables, tid, kThreads);
    log_table_init(logtab, a.tables, tid, kThreads);
    const unsigned exptab_s = smem_addr(exptab), logtab_s = smem_addr(logtab);
    __syncthreads();
    if (rw >= (int)a.nRW) return;                    // no block-wide barrier below this point
    const int jc0 = (int)a.col0 + bx * a.cols_per_cta;
    const int jc1 = min((int)a.col1, jc0 + a.cols_per_cta);
    const volatile int *kstar_live = &a.status->kstar;
    const int target = (MODE == MODE_TOPUP) ? a.status->kstar : 0;
    // ordered pipeline: the const0 kernel has already applied iteration 0 (M_1 is in M_out) unless every
    // entry of every rank passed |f_0| < tol, in which case R breaks at i = 1 and M_0 (clamped) is final
    const bool from_prep = (MODE == MODE_FIRST) && a.perm && (a.status->notconv0 != 0);
    const bool by_col = a.var_type == 2;
    double c1_row = 1.0, c2_row = 0.5, l0_i = 1.0;
    if (!by_col) {
        const int si = (a.var_type == 1) ? (row_ok ? i : 0) : 0;
        c1_row = a.sc1[si];
        c2_row = a.sc2[si];
    }
    if (FORMULA == FORMULA_A10) l0_i = row_ok ? a.l0[i] : 1.0;
    // first pass of A1 reads the caller's M and clamps it; A10's prep kernel already wrote the
    // clamped M to M_out; the top-up pass continues in place on M_out
    const double *Msrc = (MODE == MODE_FIRST && FORMULA == FORMULA_A1 && !from_prep) ? a.M_in : a.M_out;
    const int k_first = from_prep ? 1 : 0;
    int any_exhausted = 0, any_nan = 0, any_fail = 0;
    unsigned my_topups = 0, my_iters = 0;
    const size_t ns = (size_t)n;
    const int jg0 = jc0 / E;
    size_t off = (size_t)i + ns * (size_t)jc0;
    size_t ku_idx = (size_t)jg0 * (size_t)a.nRW + rw;
    size_t cv_idx = (size_t)rw * p + jc0;
    size_t us_idx = ((size_t)rw * (size_t)a.nJG + jg0) * 2;
    size_t rp_idx = (size_t)jg0 * ns + i;
    int t0_next = (MODE == MODE_FIRST) ? *kstar_live : 0;
    const double tol_lane = row_ok ? ((MODE == MODE_TOPUP) ? a.tol_verify : a.tol) : INFINITY;

    // register prefetch of the next unit's (M, const0); the top-up pass also reads the units' counts two
    // units ahead, so that it knows one unit in advance whether the next unit's operands are needed
    double mnext[E], cnext[E];
    int kn1 = 0, kn2 = 0;                            // top-up: k_unit of this unit and of the next one
    if (MODE == MODE_TOPUP) {
        kn1 = a.kunit[ku_idx];
        kn2 = (jc0 + E < jc1) ? a.kunit[ku_idx + (size_t)a.nRW] : kFrozen;
    }
    {
        const bool want = (MODE == MODE_FIRST) || (kn1 < target);
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const bool okn = want && row_ok && (jc0 + e < jc1);
            mnext[e] = okn ? ld_stream(Msrc + off + e * ns) : 0.0;
            cnext[e] = okn ? ld_stream(a.C0 + off + e * ns) : 2.0;
        }
    }

    unsigned long long peer_word = 0ull;
#pragma unroll 1
    for (int jb = jc0; jb < jc1; jb += E, off += E * ns, ku_idx += (size_t)a.nRW, cv_idx += E, us_idx += 2, rp_idx += ns) {
        int k = k_first;
        int t0 = 0;
        if (MODE == MODE_FIRST && tid == 0 && ((jb - jc0) & (8 * E - 1)) == 0) {
            absorb_peer(a.status, a.epoch, peer_word);           // the word peeked 8 units ago
            peer_word = peek_peer(a, (unsigned)(jb / E) + bx + rb);
        }
        double m[E], c0[E], w[E], c1[E], c2[E], sc[E], r[E], sexp[E], f[E];
        bool ok[E];
        bool want_next = true;
        if (MODE == MODE_TOPUP) {
            k = kn1;
            kn1 = kn2;
            kn2 = (jb + 2 * E < jc1) ? a.kunit[ku_idx + 2 * (size_t)a.nRW] : kFrozen;
            want_next = kn1 < target;
        } else {
            t0 = t0_next;                            // read one unit ahead: off the critical path, still <= u
            t0_next = *kstar_live;
        }
#pragma unroll
        for (int e = 0; e < E; ++e) {
            m[e] = mnext[e]; c0[e] = cnext[e];
            const bool okn = want_next && row_ok && (jb + E + e < jc1);
            mnext[e] = okn ? ld_stream(Msrc + off + (E + e) * ns) : 0.0;
            cnext[e] = okn ? ld_stream(a.C0 + off + (E + e) * ns) : 2.0;
        }
        if (MODE == MODE_TOPUP) {
            if (k >= target) continue;               // at K* already (k > K* cannot happen), or frozen (0xFFFF)
            ++my_topups;
        }
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const bool colok = jb + e < jc1;
            ok[e] = row_ok && colok;
            const int jj = colok ? jb + e : jc0;
            if (by_col) { c1[e] = __ldg(a.sc1 + jj); c2[e] = __ldg(a.sc2 + jj); }
            else { c1[e] = c1_row; c2[e] = c2_row; }
            const double c0m1 = __dadd_rn(c0[e], -1.0);
            w[e] = __dmul_rn(c0m1, c1[e]);
            sc[e] = (FORMULA == FORMULA_A10) ? __dmul_rn(l0_i, __ldg(a.f0 + jj)) : 1.0;
            if (MODE == MODE_FIRST && FORMULA == FORMULA_A1 && !from_prep) m[e] = (m[e] < c0m1) ? m[e] : c0m1;   // :16
            if (e > 0 && !colok) {                   // padding column (p % E != 0): shadow the previous one
                m[e] = m[e - 1]; c0[e] = c0[e - 1]; w[e] = w[e - 1]; c1[e] = c1[e - 1]; c2[e] = c2[e - 1];
                sc[e] = sc[e - 1];
            }
        }
        const int k_in = k;
        int kb;
        const int flags = newton_unit<E, FORMULA, MODE, FORMULA == FORMULA_A10, DELTA>(
            m, c0, w, c1, c2, sc, tol_lane, k, kb, (MODE == MODE_FIRST) ? t0 : target, a.maxiter, a.delta, exptab_s, r,
            sexp, f);
        my_iters += (unsigned)(k - k_in + 1);
        if (flags & 2) { any_exhausted = 1; kb = k; }
        if (MODE == MODE_TOPUP && !(flags & 3)) any_fail = 1;      // stopped at K* but not below tol
        if (finalize_unit<E, FORMULA>(a, m, c2, r, sexp, f, ok, flags, off, cv_idx, us_idx, rp_idx, p - jb, lane, row_ok,
                                      exptab_s, logtab_s))
            any_nan = 1;
        if (lane == 0) {
            const int kpub = (flags & 8) ? kb : k;
            a.kunit[ku_idx] = (unsigned short)((flags & 8) ? kFrozen : k);
            if (MODE == MODE_FIRST && kpub > t0) publish_kstar(a.status, a.epoch, kpub);
        }
    }
    if (lane == 0) {
        if (any_exhausted) a.status->exhausted = 1;
        if (any_nan) a.status->nan = 1;
        if (any_fail) a.status->verify_fail = 1;
        if (my_topups) atomicAdd(&a.status->topup_units, (unsigned long long)my_topups);
        if (my_iters) atomicAdd(&a.status->unit_iters, (unsigned long long)my_iters);
    }
}

// sum(Y*M) over the non-zeros (R/ebpmf_log.R:316): one warp per column, colsym[j] = sum_q x_q M[i_q, j]
__global__ void __launch_bounds__(256) sym_nnz_kernel(const int *colptr, const int *rowidx, const double *yx,
                                                      const double *M, long long n, long long p, double *colsym)
{
    const long long j = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (j >= p) return;
    double acc = 0.0;
    const double *Mc = M + n * j;
    for (int q = colptr[j] + lane; q < colptr[j + 1]; q += 32) acc = fma(__ldg(yx + q), Mc[rowidx[q]], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) colsym[j] = acc;
}

// 1/sigma2 and sigma2/2 (const1, const2 of R/vga_solver_mat.R:10-11) for every sigma2 element
__global__ void sigma_consts_kernel(const double *sigma2, long long len, double *sc1, double *sc2)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= len) return;
    const double s2 = sigma2[t];
    sc1[t] = 1.0 / s2;
    sc2[t] = s2 / 2;
}

// ---------------------------------------------------------------------------------------
// K2: drop-in dense-operand solver  vga_pois_solver_mat_newton(M,X,S,Beta,Sigma2,...)
// Operands recycle through (row stride, col stride).  Same two-pass global-stop scheme; both
// passes re-derive the per-entry constants from the operands.
// ---------------------------------------------------------------------------------------
struct DenseArgs {
    long long n, p;
    const double *M_in; double *M_out; double *V_out;
    const double *X;                       // dense n x p
    const double *S; long long S_rs, S_cs;
    const double *Beta; long long B_rs, B_cs;
    const double *Sigma2; long long V_rs, V_cs;
    unsigned short *kunit; long long nRW;
    Status *status;
    const MathTables *tables;
    int maxiter; double tol;
    int cols_per_cta;
};

template <int E, int MODE>
__global__ void __launch_bounds__(kThreads, 2) vga_dense_kernel(const DenseArgs a)
{
    __shared__ double exptab[EBPMF_EXP_TABLE_SIZE];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long rb = blockIdx.y;
    const long long i = rb * kRowsPerCta + tid;
    const bool row_ok = i < a.n;
    const long long rw = rb * kWarps + warp;
    exp_table_init(exptab, a.tables, tid, kThreads);
    const unsigned exptab_s = smem_addr(exptab);
    __syncthreads();
    if (rw >= a.nRW) return;                         // no block-wide barrier below this point
    const long long jc0 = (long long)blockIdx.x * a.cols_per_cta;
    const long long jc1 = min(a.p, jc0 + (long long)a.cols_per_cta);
    const int target = (MODE == MODE_TOPUP) ? a.status->kstar : 0;
    const volatile int *kstar_live = &a.status->kstar;
    const double *Msrc = (MODE == MODE_FIRST) ? a.M_in : a.M_out;
    int any_exhausted = 0, any_nan = 0, any_fail = 0;
    const long long ii = row_ok ? i : 0;

#pragma unroll 1
    for (long long jb = jc0; jb < jc1; jb += E) {
        const long long jg = jb / E;
        int k = 0;
        const int t0 = (MODE == MODE_FIRST) ? *kstar_live : 0;
        if (MODE == MODE_TOPUP) {
            k = a.kunit[jg * a.nRW + rw];
            if (k >= target) continue;
        }
        double m[E], c0[E], w[E], c1[E], c2[E], sc[E], s2[E], r[E], sexp[E], f[E];
        bool ok[E];
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const long long j = jb + e;
            ok[e] = row_ok && (j < jc1);
            const long long jj = (j < jc1) ? j : jb;                               // padding column shadows column jb
            m[e] = row_ok ? ld_stream(Msrc + i + a.n * jj) : 0.0;
            const double y = row_ok ? ld_stream(a.X + i + a.n * jj) : 0.0;
            const double b = row_ok ? a.Beta[ii * a.B_rs + jj * a.B_cs] : 0.0;
            s2[e] = row_ok ? a.Sigma2[ii * a.V_rs + jj * a.V_cs] : 1.0;
            sc[e] = row_ok ? a.S[ii * a.S_rs + jj * a.S_cs] : 1.0;
            c1[e] = 1.0 / s2[e];                                                   // :10
            c2[e] = s2[e] / 2;                                                     // :11
            c0[e] = __dadd_rn(__fma_rn(s2[e], y, b), 1.0);                         // :9
            const double c0m1 = __dadd_rn(c0[e], -1.0);
            w[e] = __dmul_rn(c0m1, c1[e]);                                         // x + const3
            if (MODE == MODE_FIRST) m[e] = (m[e] < c0m1) ? m[e] : c0m1;            // :16
        }
        int kb;
        const int flags = newton_unit<E, FORMULA_A1, MODE, true, false>(m, c0, w, c1, c2, sc, row_ok ? a.tol : INFINITY,
                                                                         k, kb, (MODE == MODE_FIRST) ? t0 : target,
                                                                         a.maxiter, 0.0, exptab_s, r, sexp, f);
        if (flags & 2) kb = k;
        if (flags & 2) any_exhausted = 1;
        if (MODE == MODE_TOPUP && !(flags & 3)) any_fail = 1;
        bool lane_nan = false;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const long long j = jb + e;
            lane_nan = lane_nan || (ok[e] && (f[e] != f[e]));
            if (ok[e]) {
                st_stream(a.M_out + i + a.n * j, m[e]);
                if (a.V_out) st_stream(a.V_out + i + a.n * j, __dmul_rn(s2[e], r[e]));   // :36
            }
        }
        if (__any_sync(0xffffffffu, lane_nan)) any_nan = 1;
        if (lane == 0) {
            a.kunit[jg * a.nRW + rw] = (unsigned short)k;
            if (MODE == MODE_FIRST && k > t0) atomicMax(&a.status->kstar, k);
        }
    }
    if (lane == 0) {
        if (any_exhausted) a.status->exhausted = 1;
        if (any_nan) a.status->nan = 1;
        if (any_fail) a.status->verify_fail = 1;
    }
}

// ---------------------------------------------------------------------------------------
// K3: vga_pois_solver_mat_newton_fixed_iter  R/vga_solver_mat.R:44-65  (purely per entry)
// ---------------------------------------------------------------------------------------
struct FixedArgs {
    long long n, p;
    const double *M_in, *V_in;            // V_in null <=> is.null(V)
    const double *Y;
    const double *S; long long S_rs, S_cs;
    const double *Beta; long long B_rs, B_cs;
    const double *Sigma2; long long V_rs, V_cs;
    double *M_out, *V_out;
    const MathTables *tables;
    int maxiter;
};

__global__ void __launch_bounds__(kThreads) vga_fixed_iter_kernel(const FixedArgs a)
{
    __shared__ double exptab[EBPMF_EXP_TABLE_SIZE];
    exp_table_init(exptab, a.tables, threadIdx.x, kThreads);
    const unsigned exptab_s = smem_addr(exptab);
    __syncthreads();
    const long long i = (long long)blockIdx.y * kRowsPerCta + threadIdx.x;
    if (i >= a.n) return;
    for (long long j = blockIdx.x; j < a.p; j += gridDim.x) {
        const long long e = i + a.n * j;
        const double s2 = a.Sigma2[i * a.V_rs + j * a.V_cs];
        const double be = a.Beta[i * a.B_rs + j * a.B_cs];
        const double s = a.S[i * a.S_rs + j * a.S_cs];
        const double y = ld_stream(a.Y + e);
        double m = ld_stream(a.M_in + e);
        double v = a.V_in ? ld_stream(a.V_in + e) : s2 / 2;                         // :48
        const double is2 = 1.0 / s2;
        for (int it = 0; it < a.maxiter; ++it) {                                    // :52
            double temp = s * exp_fast(fma(0.5, v, m), exptab_s);                     // :54
            double rt = rcp_fast(temp);
            // :56  M - (Y/temp - 1 - (M-Beta)/Sigma2/temp) / (-1 - 1/Sigma2/temp)
            const double num = fma(-(m - be) * is2, rt, fma(y, rt, -1.0));
            const double den = fma(-is2, rt, -1.0);
            m = fma(-num, rcp_fast(den), m);
            temp = s * exp_fast(fma(0.5, v, m), exptab_s) * v * 0.5;                  // :58
            rt = rcp_fast(temp);
            const double h = 0.5 * v * is2 * rt;                                    // V/2/Sigma2/temp
            const double numv = fma(0.5, rt, -1.0 - h);                             // :62 numerator
            const double denv = -(1.0 + 0.5 * v) - h;                               // :62 denominator
            v = v * exp_fast(-numv * rcp_fast(denv), exptab_s);                       // V / exp(.)
        }
        st_stream(a.M_out + e, m);
        st_stream(a.V_out + e, v);
    }
}

// ---------------------------------------------------------------------------------------
// reductions of the per-unit partials (fixed order => run-to-run reproducible)
// ---------------------------------------------------------------------------------------
// src is [R][C][W]; dst[c*W + w] = sum_r src[r][c][w].  block (32, 8): x = column, y = slice of r.
template <int W>
__global__ void __launch_bounds__(256) reduce_over_rows_kernel(const double *src, long long R, long long C, double *dst)
{
    __shared__ double sm[8][32][W];
    const long long c = (long long)blockIdx.x * 32 + threadIdx.x;
    double acc[W];
#pragma unroll
    for (int w = 0; w < W; ++w) acc[w] = 0.0;
    if (c < C) {
        for (long long r = threadIdx.y; r < R; r += 8) {
#pragma unroll
            for (int w = 0; w < W; ++w) acc[w] += src[(r * C + c) * W + w];
        }
    }
#pragma unroll
    for (int w = 0; w < W; ++w) sm[threadIdx.y][threadIdx.x][w] = acc[w];
    __syncthreads();
    if (threadIdx.y == 0 && c < C) {
#pragma unroll
        for (int w = 0; w < W; ++w) {
            double t = 0.0;
            for (int s = 0; s < 8; ++s) t += sm[s][threadIdx.x][w];
            dst[c * W + w] = t;
        }
    }
}

// one block: scalars[0] = sum colsym (sym), [1] = ssexp, [2] = slogv (from unit sums [nJG][2]), [3] = sum(V)
__global__ void __launch_bounds__(1024) reduce_scalars_kernel(const double *colsym, const double *v_sum, long long p,
                                                              const double *unitsum, long long nJG, double *scalars)
{
    __shared__ double sm[4][1024];
    double acc[4] = {0, 0, 0, 0};
    for (long long j = threadIdx.x; j < p; j += 1024) { acc[0] += colsym[j]; acc[3] += v_sum[j]; }
    for (long long g = threadIdx.x; g < nJG; g += 1024) { acc[1] += unitsum[g * 2]; acc[2] += unitsum[g * 2 + 1]; }
    for (int q = 0; q < 4; ++q) sm[q][threadIdx.x] = acc[q];
    __syncthreads();
    for (int s = 512; s > 0; s >>= 1) {
        if (threadIdx.x < s)
            for (int q = 0; q < 4; ++q) sm[q][threadIdx.x] += sm[q][threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x < 4) scalars[threadIdx.x] = sm[threadIdx.x][0];
}

// by_row: v_sum[i] = sum over column groups of rowpart[jg][i]
__global__ void __launch_bounds__(256) reduce_rows_kernel(const double *rowpart, long long n, long long nJG,
                                                          double *v_sum)
{
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    double acc = 0.0;
    for (long long jg = 0; jg < nJG; ++jg) acc += rowpart[jg * n + i];
    v_sum[i] = acc;
}

// ---------------------------------------------------------------------------------------
// set-up kernels
// ---------------------------------------------------------------------------------------
// rbptr[rb*p + j] = first nnz q of column j with rowidx[q] >= rb*256  (rb = 0..nRB)
__global__ void build_rbptr_kernel(const int *colptr, const int *rowidx, long long p, long long nRB, int *rbptr)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (nRB + 1) * p) return;
    const long long rb = t / p, j = t - rb * p;
    int lo = colptr[j], hi = colptr[j + 1];
    const long long key = rb * kRowsPerCta;
    while (lo < hi) {
        const int mid = lo + ((hi - lo) >> 1);
        if ((long long)rowidx[mid] < key) lo = mid + 1; else hi = mid;
    }
    rbptr[t] = lo;
}

// EFt[j*KTP + k] = EF[j + p*k] (zero padded to KTP)
__global__ void transpose_ef_kernel(const double *EF, long long p, int K, int KTP, double *EFt)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= p * KTP) return;
    const long long j = t / KTP; const int k = (int)(t - j * KTP);
    EFt[t] = (k < K) ? EF[j + p * k] : 0.0;
}

// dense -> CSC: count non-zeros per column (one warp per column)
__global__ void count_nnz_kernel(const double *Y, long long n, long long p, int *counts)
{
    const long long j = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (j >= p) return;
    int c = 0;
    for (long long i = lane; i < n; i += 32) c += (Y[i + n * j] != 0.0);
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0) counts[j] = c;
}

// single-block exclusive scan of counts[0..p) into colptr[0..p]
__global__ void __launch_bounds__(1024) scan_counts_kernel(const int *counts, long long p, int *colptr)
{
    __shared__ long long sm[1024];
    __shared__ long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (long long base = 0; base < p; base += 1024) {
        const long long j = base + threadIdx.x;
        const long long v = (j < p) ? counts[j] : 0;
        sm[threadIdx.x] = v;
        __syncthreads();
        for (int s = 1; s < 1024; s <<= 1) {
            long long t = (threadIdx.x >= s) ? sm[threadIdx.x - s] : 0;
            __syncthreads();
            sm[threadIdx.x] += t;
            __syncthreads();
        }
        if (j < p) colptr[j] = (int)(carry + sm[threadIdx.x] - v);
        __syncthreads();
        if (threadIdx.x == 1023) carry += sm[1023];
        __syncthreads();
    }
    if (threadIdx.x == 0) colptr[p] = (int)carry;
}

// dense -> CSC: ordered fill (one warp per column, ballot prefix keeps rows sorted)
__global__ void fill_csc_kernel(const double *Y, long long n, long long p, const int *colptr, int *rowidx, double *x)
{
    const long long j = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (j >= p) return;
    int pos = colptr[j];
    for (long long i0 = 0; i0 < n; i0 += 32) {
        const long long i = i0 + lane;
        const double v = (i < n) ? Y[i + n * j] : 0.0;
        const unsigned mask = __ballot_sync(0xffffffffu, v != 0.0);
        if (v != 0.0) {
            const int off = __popc(mask & ((1u << lane) - 1u));
            rowidx[pos + off] = (int)i;
            x[pos + off] = v;
        }
        pos += __popc(mask);
    }
}

// CSC -> dense (X of the drop-in solver given as dgCMatrix); dst must be zeroed first
__global__ void scatter_csc_kernel(const int *colptr, const int *rowidx, const double *x, long long n, long long p,
                                   double *dst)
{
    const long long j = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (j >= p) return;
    for (int q = colptr[j] + lane; q < colptr[j + 1]; q += 32) dst[rowidx[q] + n * j] = x[q];
}

// a8  sum(lfactorial(x)) = sum(lgamma(x+1))  R/ebpmf_log.R:470-472: per-block partials, fixed order
__global__ void __launch_bounds__(256) lfactorial_partial_kernel(const double *x, long long len, double *partials)
{
    __shared__ double sm[256];
    double acc = 0.0;
    for (long long q = (long long)blockIdx.x * 256 + threadIdx.x; q < len; q += (long long)gridDim.x * 256)
        acc += lgamma(x[q] + 1.0);
    sm[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) sm[threadIdx.x] += sm[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) partials[blockIdx.x] = sm[0];
}

__global__ void __launch_bounds__(1024) sum_partials_kernel(const double *partials, int count, double *out)
{
    __shared__ double sm[1024];
    double acc = 0.0;
    for (int q = threadIdx.x; q < count; q += 1024) acc += partials[q];
    sm[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 512; s > 0; s >>= 1) {
        if (threadIdx.x < s) sm[threadIdx.x] += sm[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = sm[0];
}

// ---------------------------------------------------------------------------------------
// K5: ebpmf_log_update_sigma2  R/ebpmf_log.R:413-447
// new = ((v_sum + R2)/2 + b0) / (m/2 + a0 + 1);  gate (cap > 0): fitmax > log(cap) - log(exp(s2)-1) - s2/2
// ---------------------------------------------------------------------------------------
__global__ void update_sigma2_kernel(const double *v_sum, const double *R2, long long len, double half_m,
                                     double a0, double b0, double cap, const double *fitmax, double *sigma2)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= len) return;
    const double nw = ((v_sum[t] + R2[t]) / 2 + b0) / (half_m + a0 + 1);
    if (cap > 0) {
        const double s2 = sigma2[t];
        const double thr = log(cap) - log(exp(s2) - 1) - s2 / 2;
        if (fitmax[t] > thr) sigma2[t] = nw;
    } else {
        sigma2[t] = nw;
    }
}

// col / row / global max of fitted(fit_flash) = EL EF^T for the cap gate (:416,:425,:435).
// mode 2: out[j] = max_i ; mode 1: out[i] = max_j ; (mode 0 reduces mode-2 output on the host side kernel)
__global__ void __launch_bounds__(256) fitted_colmax_kernel(const double *EL, const double *EF, long long n,
                                                            long long p, int K, double *out)
{
    __shared__ double sm[256];
    const long long j = blockIdx.x;
    double mx = -INFINITY;
    for (long long i = threadIdx.x; i < n; i += 256) {
        double b = 0.0;
        for (int k = 0; k < K; ++k) b = fma(EL[i + n * k], EF[j + p * k], b);
        mx = fmax(mx, b);
    }
    sm[threadIdx.x] = mx;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) sm[threadIdx.x] = fmax(sm[threadIdx.x], sm[threadIdx.x + s]);
        __syncthreads();
    }
    if (threadIdx.x == 0) out[j] = sm[0];
}

__global__ void __launch_bounds__(256) fitted_rowmax_kernel(const double *EL, const double *EF, long long n,
                                                            long long p, int K, double *out)
{
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    double mx = -INFINITY;
    for (long long j = 0; j < p; ++j) {
        double b = 0.0;
        for (int k = 0; k < K; ++k) b = fma(EL[i + n * k], EF[j + p * k], b);
        mx = fmax(mx, b);
    }
    out[i] = mx;
}

__global__ void __launch_bounds__(1024) max_reduce_kernel(const double *in, long long len, double *out)
{
    __shared__ double sm[1024];
    double mx = -INFINITY;
    for (long long q = threadIdx.x; q < len; q += 1024) mx = fmax(mx, in[q]);
    sm[threadIdx.x] = mx;
    __syncthreads();
    for (int s = 512; s > 0; s >>= 1) {
        if (threadIdx.x < s) sm[threadIdx.x] = fmax(sm[threadIdx.x], sm[threadIdx.x + s]);
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = sm[0];
}

// ---------------------------------------------------------------------------------------
// a7: calc_ebpmf_log_obj  R/ebpmf_log.R:406-410 -- the O(len) sums; one block, fixed order.
// terms[0] = sum(ss*log(2*pi*sigma2)/2), [1] = sum(sv/2/sigma2), [2] = sum(R2/2/sigma2),
// terms[3] = sum((a0+1)*log(sigma2)),   [4] = sum(b0/sigma2)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) obj_terms_kernel(const double *sv, const double *sigma2, const double *R2,
                                                         long long len, double ss, double a0, double b0,
                                                         double *terms)
{
    __shared__ double sm[5][1024];
    double acc[5] = {0, 0, 0, 0, 0};
    const double two_pi = 6.283185307179586;
    for (long long q = threadIdx.x; q < len; q += 1024) {
        const double s2 = sigma2[q];
        acc[0] += ss * log(two_pi * s2) / 2;
        acc[1] += sv[q] / 2 / s2;
        acc[2] += R2[q] / 2 / s2;
        acc[3] += (a0 + 1) * log(s2);
        acc[4] += b0 / s2;
    }
    for (int q = 0; q < 5; ++q) sm[q][threadIdx.x] = acc[q];
    __syncthreads();
    for (int s = 512; s > 0; s >>= 1) {
        if (threadIdx.x < s)
            for (int q = 0; q < 5; ++q) sm[q][threadIdx.x] += sm[q][threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x < 5) terms[threadIdx.x] = sm[threadIdx.x][0];
}

// FP64 pipe probe: 8 independent DFMA chains per thread
__global__ void __launch_bounds__(256) fp64_probe_kernel(double *out, int iters, double seed)
{
    double a0 = seed, a1 = seed + 1, a2 = seed + 2, a3 = seed + 3, a4 = seed + 4, a5 = seed + 5, a6 = seed + 6,
           a7 = seed + 7;
    const double b = 1.0000001, c = 1e-9;
    for (int it = 0; it < iters; ++it) {
        a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
        a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
    const double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == 123.456) out[0] = s;
}

}  // namespace ebpmf
