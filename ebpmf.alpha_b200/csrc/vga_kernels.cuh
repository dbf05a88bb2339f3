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
// Layout.  M is column-major n x p (R), so consecutive THREADS own consecutive ROWS: a CTA of 256
// threads owns one TILE = a row block of 256 rows x a span of columns, and walks the span in slabs of
// 16 columns.  Per slab, in shared memory:
//   ms  the M slab: 16 column segments of 2 KB, brought in by ONE bulk async copy per column
//       (cp.async.bulk ... mbarrier::complete_tx, issued by 16 lanes of warp 0 while the CTA stages the
//       rest); every thread reads its own row, writes the final m back IN PLACE, and the slab leaves
//       with one bulk async store per column (cp.async.bulk.global.shared::cta);
//   cs  the dense image of the sparse Y slab (scattered from the CSC arrays through a per-row-block
//       offset table), overwritten in place by const0 = (Sigma2*Y + Beta) + 1 once this thread's Beta
//       row is known (K_tot FMAs per entry against EF rows broadcast from shared memory -- an
//       outer-product rebuild, not a tensor-core contraction), and finally by V for the column sums.
// No const0 / Beta / V matrix ever goes to HBM: the step reads M once and writes M once.
//
// The GLOBAL stop rule.  R stops when max|f| < tol over ALL entries, so every entry receives the same
// number u of Newton updates.  The kernels reproduce u exactly with register-resident
// iterations and no per-iteration grid sync:
//   unit      = one warp x E columns (32*E entries), iterates in lock step;
//   pass 1    (MODE_FIRST) each unit iterates until ITS entries all satisfy |f_k| < tol at some k >= t0,
//             where t0 is the largest such k any finished unit had published when this unit started
//             (never above u, because at k = u every unit is converged); it finalises speculatively and
//             records k_unit; K* = max.  Tiles are taken in the order of a.perm: hardest first, by the
//             previous step's per-tile counts (tile_hard), so K* reaches u within the first wave and
//             nearly every unit runs straight to u inside pass 1;
//   pass 2    (MODE_TOPUP, same kernel) only tiles holding a unit with k_unit < K* do anything: they
//             restage, continue the SAME iteration sequence up to exactly K* updates, test
//             |f_K*| < tol and re-finalise the whole tile.  If every unit passes, u = K*.  Otherwise
//             the host repeats pass 2 with K*+1 (rare: an entry bounced back above tol).
//   exhausted pass 1 never performs update number maxiter: a unit that gets there parks at
//             k = maxiter-1 and publishes K* = maxiter; the top-up pass performs the last update for
//             every unit and leaves V with the reciprocal of the LAST EVALUATED temp, i.e. one
//             iteration stale exactly as in R (F6).
//   periodic  (FREEZE_PERIODIC, the default; bit-identical to iterating on) once a unit is converged
//             and every entry's update maps m_k to m_k itself or back to m_{k-1} (in floating point the
//             iteration ends in a fixed point or a 2-cycle of adjacent doubles), all later iterates
//             repeat with period 2, so the unit advances its count by an EVEN number of updates at once.
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
enum { FREEZE_NONE = 0, FREEZE_PERIODIC = 1, FREEZE_DELTA = 2 };

struct Status {           // device words shared by the passes of one solve
    int kstar;            // max over units of k_unit (pass 1), target of pass 2
    int nan;              // a NaN reached the stop test
    int verify_fail;      // pass 2: some unit is not below tol at K*
    int exhausted;        // some unit hit maxiter
    unsigned long long topup_units;   // diagnostics: units that needed pass 2 work
    unsigned long long unit_iters;    // diagnostics: sum over units of evaluations performed
    int topup_tiles;      // diagnostics: tiles the top-up pass had to restage
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
    const double *EL;               // n x K column-major
    const double *EFt;              // p x ldk row-major, zero padded (ldk = KTP, or a multiple of 32 when K > 32)
    int ldk;
    const double *sc1, *sc2;        // 1/sigma2 (const1), sigma2/2 (const2): length 1 | n | p
    const double *l0, *f0;          // A10 only
    const int *rowidx;              // CSC row indices (0-based within the shard)
    const double *yx;               // CSC values
    const int *rbptr;               // (nRB+1) x p: first nnz of column j with row >= rb*256
    unsigned short *kunit;          // [nRW][nJG] updates each unit has received (kFrozen: delta-mode fixed point)
    long long nRW, nJG;             // ceil(n/32), ceil(p/E)
    double *colV;                   // [nRB][p]        sum over the row block's 256 rows of V
    double *tileS;                  // [nRB*nspans][3] sum exp(M+V/2), sum(log(V)/2+c), sum(Y*M) of the tile
    double *rowpart;                // [nspans][n] sum over the tile's columns of V (by_row only), else null
    unsigned short *tile_hard;      // [nRB*nspans] max over the tile's units of the first converged count (pass 1)
    unsigned short *tile_kmin;      // [nRB*nspans] min over the tile's units of k_unit
    const int *perm;                // tile ids in launch order (null: natural order)
    int tile_begin;                 // first position of this launch in the order
    int nspans;                     // column spans per row block; tile id = rb * nspans + span
    int cols_per_cta;               // span width, a multiple of kTileCols
    int *chunk_dirty;               // [nchunks] set by the top-up pass where it changed M (host pipeline); may be null
    const int *span_rank;           // [nspans] span -> chunk
    int bulk;                       // 1: n is even, so every column segment is 16-byte aligned -> bulk async copies
    Status *status;
    const Status *const *peers;     // the other ranks' Status words, mapped peer memory (NVLink); may be null
    int npeers;
    unsigned epoch;                 // solve counter, identical on every rank
    const MathTables *tables;       // exp2 / log tables (device copy of the host-built tables)
    int maxiter;
    double tol;
    double delta;                   // FREEZE_DELTA only
    double tol_verify;              // == tol; a test hook scales it to exercise the K*+1 retry path
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
// bulk async copies (TMA unit, 1-D) and their mbarrier
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(unsigned mbar_s, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar_s), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned mbar_s, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar_s), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned mbar_s, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(mbar_s), "r"(parity) : "memory");
}
// global -> shared, completion signalled on the mbarrier (bytes: multiple of 16, both addresses 16-byte aligned)
__device__ __forceinline__ void bulk_load(unsigned dst_s, const void *src, unsigned bytes, unsigned mbar_s) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_s), "l"(src), "r"(bytes), "r"(mbar_s) : "memory");
}
// shared -> global, tracked by the issuing thread's bulk async-group
__device__ __forceinline__ void bulk_store(void *dst, unsigned src_s, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src_s), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// the issuing thread's bulk stores have finished READING shared memory (the buffer may be reused)
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// generic-proxy writes to shared memory become visible to the async proxy (before a bulk store reads them)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---------------------------------------------------------------------------------------
// One unit: E entries per lane, all 32 lanes in lock step.  Per entry the loop carries
//   m, c0 = (s2*y + beta) + 1, w = (c0 - 1)/s2 = y + beta/s2, c1 = 1/s2, c2 = s2/2 [, sc = S]
// so that  f = x - sexp - M*const1 + const3  (R/vga_solver_mat.R:25) is  fma(-m, c1, w - sexp).
// tol_lane is tol for a live row and +Inf for a padding row (row >= n); a padding COLUMN (p not a
// multiple of E) carries a copy of the lane's previous column, so it converges with it.
// `target` is t0 in the first pass (stop at the first k >= t0 at which the unit is converged)
// and K* in the top-up pass (stop exactly at k == K*).
// On return k is the number of updates the unit has received, r[] the reciprocal of the last
// evaluated temp, sexp[] / f[] the last evaluated S*exp(.) and f; evals counts the evaluations.
// Returns a bit mask: 1 converged at k, 2 exhausted (maxiter updates done, top-up pass only),
// 4 parked at maxiter-1 (first pass only: the last update belongs to the top-up pass), 8 frozen (delta mode).
// ---------------------------------------------------------------------------------------
constexpr int kFrozen = 0xFFFF;     // kunit marker: unit is at its numerical fixed point (delta mode)

// 1/g for the Newton step: MUFU seed + ONE quadratic correction (rel. error ~2^-40).  The step is
// self-correcting, so the extra cubic term of rcp_fast buys nothing here.
__device__ __forceinline__ double rcp_step(double x) {
#if EBPMF_EXACT_MATH
    return 1.0 / x;
#else
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    const double e = __fma_rn(-x, r, 1.0);
    return __fma_rn(r, e, r);
#endif
}

__device__ __forceinline__ bool same_bits(double a, double b) { return __double_as_longlong(a) == __double_as_longlong(b); }

// reciprocal of temp = const0 - M (:22).  MUFU seed + one quadratic correction (rel. error ~2^-40 = 9e-13): V =
// Sigma2/temp and the exp argument const2/temp inherit it, four orders inside the 1e-8 parity budget.
// -DEBPMF_RCP_CUBIC=1 restores the cubic step (~1.5 ulp) for A/B runs.
#ifndef EBPMF_RCP_CUBIC
#define EBPMF_RCP_CUBIC 0
#endif
#ifndef EBPMF_PIN_A
#define EBPMF_PIN_A 1
#endif
__device__ __forceinline__ double rcp_temp(double x) {
#if EBPMF_RCP_CUBIC
    return rcp_fast(x);
#else
    return rcp_step(x);
#endif
}

// one evaluation of f at m (R/vga_solver_mat.R:22-25); returns "every entry of the unit has |f| < tol"
template <int E, bool SCALED>
__device__ __forceinline__ bool newton_eval(const double (&m)[E], const double (&c0)[E], const double (&w)[E],
                                            const double (&c1)[E], const double (&c2)[E], const double (&sc)[E],
                                            double tol_lane, unsigned exptab, double (&r)[E], double (&a)[E],
                                            double (&sexp)[E], double (&f)[E])
{
    bool lane_conv = true;
    double xx[E], ex[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const double temp = __dadd_rn(c0[e], -m[e]);                       // :22
        r[e] = rcp_temp(temp);
        a[e] = __dmul_rn(c2[e], r[e]);                                     // const2/temp
        xx[e] = __dadd_rn(m[e], a[e]);
#if EBPMF_PIN_A
        // keep a and the exp argument in registers: otherwise the compiler recomputes both (2 FP64-pipe
        // instructions each) in the exp fast path and again in the update
        asm volatile("" : "+d"(a[e]));
        asm volatile("" : "+d"(xx[e]));
#endif
    }
    exp_fast_n<E>(xx, ex, exptab);
#pragma unroll
    for (int e = 0; e < E; ++e) {
        sexp[e] = SCALED ? __dmul_rn(sc[e], ex[e]) : ex[e];                // :23
        f[e] = __fma_rn(-m[e], c1[e], __dadd_rn(w[e], -sexp[e]));          // :25 | :84,:98
        lane_conv = lane_conv && (fabs(f[e]) < tol_lane);                  // :26 (strict <)
    }
    return __all_sync(0xffffffffu, lane_conv);
}

// the Newton update (R/vga_solver_mat.R:31 | :90,:105) from the last evaluation
template <int E, int FORMULA>
__device__ __forceinline__ void newton_step(const double (&m)[E], const double (&c1)[E], const double (&r)[E],
                                            const double (&a)[E], const double (&sexp)[E], const double (&f)[E],
                                            double (&mn)[E])
{
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const double q = (FORMULA == FORMULA_A1) ? a[e] : __dmul_rn(0.5, r[e]);   // const2/temp^2 | 0.5*temp^-2
        const double g = __fma_rn(-sexp[e], __fma_rn(q, r[e], 1.0), -c1[e]);
        mn[e] = __fma_rn(-f[e], rcp_step(g), m[e]);
    }
}

template <int E, int FORMULA, int MODE, bool SCALED, int FREEZE>
__device__ __forceinline__ int newton_unit(double (&m)[E], const double (&c0)[E], double (&w)[E],
                                           const double (&c1)[E], const double (&c2)[E], const double (&sc)[E],
                                           double tol_lane, int &k, int &kb, int target, int maxiter, double delta,
                                           unsigned exptab, double (&r)[E], double (&sexp)[E], double (&f)[E],
                                           unsigned &evals)
{
    int flags = 0;
    kb = -1;                        // first k at which the unit was converged (what K* is the max of)
    // keep w (and the lane tolerance) in registers: without this the compiler re-derives
    // (c0-1)*c1 in every iteration -- 2 extra FP64-pipe instructions per entry per iteration
#pragma unroll
    for (int e = 0; e < E; ++e) asm volatile("" : "+d"(w[e]));
    asm volatile("" : "+d"(tol_lane));
    asm volatile("" : "+r"(exptab));        // likewise the table's shared-window address
    const int k_in = k;
    int not_evaluated = 0;                  // counts skipped by the fast-forward, and the exhausted exit's last update
    double a[E];
    if constexpr (FREEZE != FREEZE_DELTA) {
        int klast = maxiter - 1;            // at k == klast the next update is number maxiter
        asm volatile("" : "+r"(klast));
        bool conv;
        // ---- phase 1: plain iteration until the unit is converged for the first time (nothing else to track)
#pragma unroll 1
        for (;;) {
            conv = newton_eval<E, SCALED>(m, c0, w, c1, c2, sc, tol_lane, exptab, r, a, sexp, f);
            if (conv || k >= klast || (MODE == MODE_TOPUP && k == target)) break;
            newton_step<E, FORMULA>(m, c1, r, a, sexp, f, m);
            ++k;
        }
        if (conv) kb = k;
        // ---- phase 2: the unit has been converged; run on to the target count, fast-forwarding once the
        // iterates repeat.  Entering an iteration, (r, a, sexp, f, conv) belong to m at index k.
        double mprev[E];
#pragma unroll
        for (int e = 0; e < E; ++e) mprev[e] = m[e];
        bool prev_conv = false;
#pragma unroll 1
        for (;;) {
            const bool stop = (MODE == MODE_FIRST) ? (conv && (k >= target)) : (k == target);
            if (stop) { if (conv) flags |= 1; break; }
            if (k >= klast) {               // update number maxiter: the top-up pass performs it, stale V and all (F6)
                if (MODE == MODE_FIRST) { flags |= 4; }
                else { newton_step<E, FORMULA>(m, c1, r, a, sexp, f, m); ++k; ++not_evaluated; flags |= 2; }
                break;
            }
            double mn[E];
            newton_step<E, FORMULA>(m, c1, r, a, sexp, f, mn);
            if (FREEZE == FREEZE_PERIODIC && conv && prev_conv) {
                // converged now and one iteration ago, and every entry's next iterate is m itself or the
                // previous iterate: from here on the sequence has period (at most) 2 and every state is
                // converged.  The state at index k + d (iterate, and everything an evaluation of it would
                // produce: r, a, sexp, f, conv) is therefore THIS state for every even d -- and for every d
                // when all entries sit on a true fixed point (mn == m bit for bit).
                bool lane_per = true, lane_fix = true;
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const bool fx = same_bits(mn[e], m[e]);
                    lane_fix = lane_fix && fx;
                    lane_per = lane_per && (fx || same_bits(mn[e], mprev[e]));
                }
                if (__all_sync(0xffffffffu, lane_per)) {
                    const int lim = min(target, klast);      // where the stop / exhaustion tests next fire
                    const int d = lim - k;
                    if (d > 0) {
                        if (!(d & 1) || __all_sync(0xffffffffu, lane_fix)) {
                            // index lim carries the current state: no evaluation left to do
                            k = lim; not_evaluated += d;
                            continue;
                        }
                        // odd distance on a 2-cycle: jump an even amount, the update below lands on lim
                        k += d - 1; not_evaluated += d - 1;
                    }
                }
            }
#pragma unroll
            for (int e = 0; e < E; ++e) { mprev[e] = m[e]; m[e] = mn[e]; }
            prev_conv = conv;
            ++k;
            conv = newton_eval<E, SCALED>(m, c0, w, c1, c2, sc, tol_lane, exptab, r, a, sexp, f);
            if (conv && kb < 0) kb = k;
        }
    } else {
        // delta mode: a converged unit FREEZES as soon as every entry's update is a numerical no-op
        // (|f/g| <= delta, |f| < tol/16): all later iterates equal this one to within delta, so the unit is
        // final for any global count.
#pragma unroll 1
        for (;;) {
            const bool conv = newton_eval<E, SCALED>(m, c0, w, c1, c2, sc, tol_lane, exptab, r, a, sexp, f);
            if (conv && kb < 0) kb = k;
            const bool stop = (MODE == MODE_FIRST) ? (conv && (k >= target)) : (k == target);
            if (!conv) {
                if (stop) break;
                if (MODE == MODE_FIRST && k >= maxiter - 1) { flags |= 4; break; }
                newton_step<E, FORMULA>(m, c1, r, a, sexp, f, m);
            } else {
                double mn[E];
                bool lane_small = true;
                const double tol16 = tol_lane * 0.0625;
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const double q = (FORMULA == FORMULA_A1) ? a[e] : __dmul_rn(0.5, r[e]);
                    const double g = __fma_rn(-sexp[e], __fma_rn(q, r[e], 1.0), -c1[e]);
                    const double rg = rcp_step(g);
                    mn[e] = __fma_rn(-f[e], rg, m[e]);
                    lane_small = lane_small && (fabs(__dmul_rn(f[e], rg)) <= delta) && (fabs(f[e]) < tol16);
                }
                if (__all_sync(0xffffffffu, lane_small)) { flags |= 1 | 8; break; }
                if (stop) { flags |= 1; break; }
                if (MODE == MODE_FIRST && k >= maxiter - 1) { flags |= 4; break; }
#pragma unroll
                for (int e = 0; e < E; ++e) m[e] = mn[e];
            }
            ++k;
            if (k >= maxiter) { flags |= 2; ++not_evaluated; break; }
        }
    }
    evals += (unsigned)(k - k_in + 1 - not_evaluated);
    return flags;
}

// D(8x8) += A(8x4) * B(4x8) on the FP64 tensor path; lane holds A[lane/4][lane%4], B[lane%4][lane/4],
// D[lane/4][2*(lane%4) + {0,1}]
__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

#ifndef EBPMF_BETA_DMMA
#define EBPMF_BETA_DMMA 1     // Beta = EL EF' of a slab on the FP64 tensor path (0: K_tot DFMAs per entry, thread = row)
#endif

// Layout of the slab's EF rows in shared memory, efs[c][k]: the B-fragment loads of the Beta DMMAs
// (c = 8nt + lane/4, k = 4ks + lane%4) must hit 16 distinct bank pairs per half warp.  Row stride KTP + 4 does
// that for KTP = 8, 16, 24; KTP = 32 keeps stride 32 (shared memory is full) and XOR-swizzles k by 4*(c & 3).
template <int KTP> struct EfLayout {
    static constexpr int kStride = (KTP == 32) ? 32 : KTP + 4;
    __device__ __forceinline__ static int at(int c, int k) { return c * kStride + ((KTP == 32) ? (k ^ ((c & 3) << 2)) : k); }
};

// stage EF rows [j0, j0+16) x factor columns [k0, k0+KTP) into efs (EfLayout); a padding column
// (j >= jc1) repeats the last real one
template <int KTP>
__device__ __forceinline__ void stage_ef_chunk(double *efs, const double *EFt, int ldk, int j0, int jc1, int k0, int tid) {
    for (int idx = tid; idx < kTileCols * KTP; idx += kThreads) {
        const int c = idx / KTP, k = idx - c * KTP;
        efs[EfLayout<KTP>::at(c, k)] = __ldg(EFt + (size_t)min(j0 + c, jc1 - 1) * ldk + k0 + k);
    }
}

// shared memory of the fused kernel (dynamic)
template <int KTP, int FORMULA>
struct FusedSmem {
    double ms[kTileCols][kRowsPerCta];      // M slab: bulk-load target, final M in place, bulk-store source
    double cs[kTileCols][kRowsPerCta];      // Y image -> const0 image -> V image
    double bs[FORMULA == FORMULA_A10 ? kTileCols : 1][FORMULA == FORMULA_A10 ? kRowsPerCta : 2];   // Beta (A10 clamp, :78)
    double efs[kTileCols * EfLayout<KTP>::kStride];
    double c1s[kTileCols], c2s[kTileCols], f0s[kTileCols];
    double exptab[EBPMF_EXP_TABLE_SIZE];
    double2 logtab[EBPMF_LOG_TABLE_SIZE];
    double wsum[3][kWarps];
    int wk[2][kWarps];
    unsigned long long mbar;
    // the slab's non-zeros (value, column << 8 | row) as the scatter met them: sum(Y * M) is taken against the FINAL
    // M while it is still in shared memory, without a second walk of the CSC arrays.  Sized to keep 3 CTAs per SM;
    // a slab with more non-zeros re-walks the CSC entries beyond the list.
    static constexpr int kYList = (FORMULA == FORMULA_A1) ? (KTP == 32 ? 224 : 208) : 1;
    double ylv[kYList];
    int ybase[kTileCols + 1];
    unsigned short ylp[kYList];
};

#ifndef EBPMF_MINB
#define EBPMF_MINB 3          // resident CTAs per SM the register allocation is bounded for
#endif

// ---------------------------------------------------------------------------------------
// K1 / K4: fused VGA step on CSC Y with Beta rebuilt from the flash factors; both passes.
// grid = tiles of this launch (1-D), block = 256 threads, dynamic smem = sizeof(FusedSmem).
// ---------------------------------------------------------------------------------------
template <int KTP, int E, int FORMULA, int MODE, int FREEZE>
__global__ void __launch_bounds__(kThreads, (FORMULA == FORMULA_A10) ? 2 : EBPMF_MINB) vga_fused_kernel(const FusedArgs a)
{
    static_assert(kTileCols % E == 0, "E must divide the slab width");
    static_assert(kTileCols / E <= 32, "one lane per unit of a slab");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    FusedSmem<KTP, FORMULA> &sm = *reinterpret_cast<FusedSmem<KTP, FORMULA> *>(smem_raw);
    constexpr int UNITS = kTileCols / E;

    // rows of a shard and columns fit 32 bits (dgCMatrix Dim is int); only n*j products are 64-bit
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = (int)a.n, p = (int)a.p;
    const int lin = a.tile_begin + (int)blockIdx.x;
    const int tile = a.perm ? a.perm[lin] : lin;
    const int rb = tile / a.nspans, span = tile - rb * a.nspans;
    int target = 0;
    if (MODE == MODE_TOPUP) {
        target = a.status->kstar;
        if ((int)a.tile_kmin[tile] >= target) return;        // nothing below K* here (block-uniform)
    }
    const int r0 = rb * kRowsPerCta;
    const int i = r0 + tid;
    const bool row_ok = i < n;
    const int rows_valid = min(kRowsPerCta, n - r0);
    const int rw = rb * kWarps + warp;
    const bool warp_ok = (r0 + warp * 32) < n;       // false: every row of this warp is beyond n
    const int jc0 = span * a.cols_per_cta;
    const int jc1 = min(p, jc0 + a.cols_per_cta);
    const size_t ns = (size_t)n;
    const bool bulk = a.bulk != 0;
    const double *Msrc = (MODE == MODE_FIRST) ? a.M_in : a.M_out;

    exp_table_init(sm.exptab, a.tables, tid, kThreads);
    log_table_init(sm.logtab, a.tables, tid, kThreads);
    const unsigned exptab_s = smem_addr(sm.exptab), logtab_s = smem_addr(sm.logtab);
    const unsigned mbar_s = smem_addr(&sm.mbar);
    if (tid == 0) {
        mbar_init(mbar_s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_async_smem();
    }

    // row-constant pieces (by_row / constant sigma2, A10's l0)
    const bool by_col = a.var_type == 2;
    double c1_row = 1.0, c2_row = 0.5, l0_i = 1.0;
    if (!by_col) {
        const int si = (a.var_type == 1) ? (row_ok ? i : 0) : 0;
        c1_row = a.sc1[si];                                                    // :10
        c2_row = a.sc2[si];                                                    // :11
    }
    if (FORMULA == FORMULA_A10) l0_i = row_ok ? a.l0[i] : 1.0;

    const volatile int *kstar_live = &a.status->kstar;
    int any_exhausted = 0, any_nan = 0, any_fail = 0;
    unsigned my_evals = 0, my_topups = 0;
    int wkb = 0, wkmin = kFrozen;                    // this warp's max first-converged count / min k_unit
    int t0_next = (MODE == MODE_FIRST) ? *kstar_live : 0;
    double acc_exp = 0.0, acc_log = 0.0, acc_rowv = 0.0;     // this row's sums over the tile's columns
    double acc_sym = 0.0;                                    // this thread's share of sum(Y * M) over the tile's non-zeros
    constexpr int kYList = FusedSmem<KTP, FORMULA>::kYList;
    const double tol_lane = row_ok ? ((MODE == MODE_TOPUP) ? a.tol_verify : a.tol) : INFINITY;
    unsigned phase = 0;

    for (int j0 = jc0; j0 < jc1; j0 += kTileCols) {
        const int ncv = min(kTileCols, jc1 - j0);    // real columns of this slab
        __syncthreads();                             // previous slab fully consumed (and the mbarrier initialised)
        // one thread per CTA peeks at one peer GPU's published count, once per tile (round robin over the peers by
        // tile; per slab it cost 3 % of the pass: the pointer fetch and the remote load are a dependent chain in
        // front of warp 0's bulk copies); the word is consumed after the slab's units, so the NVLink latency is
        // never waited for
        unsigned long long peer_word = 0ull;
        if (MODE == MODE_FIRST && tid == 0 && j0 == jc0) peer_word = peek_peer(a, (unsigned)tile);
        // ---- the M slab: one bulk async copy per column (a padding column repeats the last real one)
        if (bulk) {
            if (warp == 0) {
                if (lane < kTileCols) bulk_wait_read();          // my store of the previous slab has left shared memory
                __syncwarp();
                if (lane == 0) mbar_expect_tx(mbar_s, (unsigned)(kTileCols * rows_valid * 8));
                __syncwarp();
                if (lane < kTileCols)
                    bulk_load(smem_addr(&sm.ms[lane][0]), Msrc + (size_t)r0 + ns * (size_t)min(j0 + lane, jc1 - 1),
                              (unsigned)(rows_valid * 8), mbar_s);
            }
            if (!row_ok) {                                       // rows beyond n: benign operands
#pragma unroll
                for (int c = 0; c < kTileCols; ++c) sm.ms[c][tid] = 0.0;
            }
        } else {
#pragma unroll
            for (int c = 0; c < kTileCols; ++c)
                sm.ms[c][tid] = row_ok ? ld_stream(Msrc + (size_t)i + ns * (size_t)min(j0 + c, jc1 - 1)) : 0.0;
        }
        // ---- stage the slab: zero Y image, EF rows, per-column constants --------------------
#pragma unroll
        for (int c = 0; c < kTileCols; ++c) sm.cs[c][tid] = 0.0;
        stage_ef_chunk<KTP>(sm.efs, a.EFt, a.ldk, j0, jc1, 0, tid);
        if (tid < kTileCols) {
            const int j = min(j0 + tid, jc1 - 1);
            sm.c1s[tid] = by_col ? a.sc1[j] : 1.0;
            sm.c2s[tid] = by_col ? a.sc2[j] : 0.5;
            sm.f0s[tid] = (FORMULA == FORMULA_A10) ? a.f0[j] : 1.0;
        }
        if (FORMULA == FORMULA_A1 && warp == 1) {    // list offsets of the slab's columns: exclusive scan of their counts
            const int *rp = a.rbptr + (size_t)rb * p + j0;
            int cnt = (lane < ncv) ? rp[p + lane] - rp[lane] : 0, inc = cnt;
#pragma unroll
            for (int o = 1; o < kTileCols; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
            if (lane < kTileCols) sm.ybase[lane] = inc - cnt;
            if (lane == kTileCols - 1) sm.ybase[kTileCols] = inc;
        }
        __syncthreads();
        // scatter the CSC non-zeros of rows [r0, r0+256) of each slab column
        {
            const int *rp0 = a.rbptr + (size_t)rb * p + j0, *rp1 = rp0 + p;
            for (int c = warp; c < ncv; c += kWarps) {
                const int lo = rp0[c], hi = rp1[c];
                const int base = (FORMULA == FORMULA_A1) ? sm.ybase[c] - lo : 0;
                for (int q = lo + lane; q < hi; q += 32) {
                    const int rr = a.rowidx[q] - r0;
                    const double y = __ldg(a.yx + q);
                    sm.cs[c][rr] = y;
                    if (FORMULA == FORMULA_A1 && base + q < kYList) { sm.ylv[base + q] = y; sm.ylp[base + q] = (unsigned short)((c << 8) | rr); }
                }
            }
        }
        __syncthreads();
        // a2: Beta = EL EF' for this CTA's 256 rows x the 16 slab columns, folded straight into
        // const0 = (Sigma2*Y + Beta) + 1 (:9 | :81,:95) over the Y image.  K_tot > KTP (only possible with
        // KTP = 32) walks the factor columns in chunks of KTP, restaging the EF rows.
#if EBPMF_BETA_DMMA
        {
            // FP64 tensor path: warp w owns rows [32w, 32w+32) = 4 row tiles x 2 column tiles of 8 x 8, K_tot/4 DMMAs
            // each (an accumulating DFMA has three register sources and issues in 3 cycles; DMMA does not, and it
            // leaves the issue port to the other warps).  A fragments = EL straight from global memory (L1/L2),
            // B fragments = the staged EF rows.  Lane (l8 = lane/4, l4 = lane%4) ends up with Beta of rows
            // 32w + 8rt + l8, columns 8nt + 2*l4 + {0,1}.
            const int l4 = lane & 3, l8 = lane >> 2, wrow = 32 * warp;
            double d[4][2][2];
#pragma unroll
            for (int rt = 0; rt < 4; ++rt)
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) { d[rt][nt][0] = 0.0; d[rt][nt][1] = 0.0; }
            for (int k0 = 0; k0 < a.K; k0 += KTP) {
                if (k0 != 0) {
                    __syncthreads();
                    stage_ef_chunk<KTP>(sm.efs, a.EFt, a.ldk, j0, jc1, k0, tid);
                    __syncthreads();
                }
#pragma unroll
                for (int ks = 0; ks < KTP / 4; ++ks) {
                    if (k0 + 4 * ks >= a.K) break;               // factor columns past K_tot are zero padding (uniform)
                    const int kk = k0 + 4 * ks + l4;
                    const double bf0 = sm.efs[EfLayout<KTP>::at(l8, 4 * ks + l4)], bf1 = sm.efs[EfLayout<KTP>::at(8 + l8, 4 * ks + l4)];
#pragma unroll
                    for (int rt = 0; rt < 4; ++rt) {
                        const int ir = r0 + wrow + 8 * rt + l8;
                        const double af = (ir < n && kk < a.K) ? __ldg(a.EL + (size_t)ir + ns * (size_t)kk) : 0.0;
                        dmma884(d[rt][0][0], d[rt][0][1], af, bf0);
                        dmma884(d[rt][1][0], d[rt][1][1], af, bf1);
                    }
                }
            }
#pragma unroll
            for (int rt = 0; rt < 4; ++rt) {
                const int row = wrow + 8 * rt + l8;
                double c2r = c2_row;                     // by_row: sigma2/2 of THIS fragment row, not of row tid
                if (a.var_type == 1) c2r = a.sc2[min(r0 + row, n - 1)];
#pragma unroll
                for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int c = 8 * nt + 2 * l4 + e;
                        const double c2c = by_col ? sm.c2s[c] : c2r;
                        const double v = __dadd_rn(__fma_rn(__dadd_rn(c2c, c2c), sm.cs[c][row], d[rt][nt][e]), 1.0);
                        if (c < ncv) {
                            sm.cs[c][row] = v;
                            if constexpr (FORMULA == FORMULA_A10) sm.bs[c][row] = d[rt][nt][e];
                        }
                    }
            }
            __syncwarp();
            if (ncv < kTileCols) {                       // padding columns: copies of the last real one
                for (int c = ncv; c < kTileCols; ++c) {
                    sm.cs[c][tid] = sm.cs[ncv - 1][tid];
                    if constexpr (FORMULA == FORMULA_A10) sm.bs[c][tid] = sm.bs[ncv - 1][tid];
                }
                __syncwarp();
            }
        }
#else
        for (int k0 = 0; k0 < a.K; k0 += KTP) {
            const bool first = (k0 == 0), last = (k0 + KTP >= a.K);
            if (!first) {
                __syncthreads();
                stage_ef_chunk<KTP>(sm.efs, a.EFt, a.ldk, j0, jc1, k0, tid);
                __syncthreads();
            }
            double el[KTP];
            const double *elp = a.EL + i + (size_t)k0 * ns;
#pragma unroll
            for (int k = 0; k < KTP; ++k) el[k] = (row_ok && k0 + k < a.K) ? __ldg(elp + k * ns) : 0.0;
#pragma unroll 4
            for (int c = 0; c < kTileCols; ++c) {
                double b = 0.0;
#pragma unroll
                for (int k = 0; k < KTP; ++k) b = fma(el[k], sm.efs[EfLayout<KTP>::at(c, k)], b);
                const double c2c = by_col ? sm.c2s[c] : c2_row;
                double v = first ? __fma_rn(__dadd_rn(c2c, c2c), sm.cs[c][tid], b) : __dadd_rn(sm.cs[c][tid], b);
                if (last) v = __dadd_rn(v, 1.0);
                sm.cs[c][tid] = (c < ncv) ? v : sm.cs[ncv - 1][tid];           // padding column: copy of the last real one
                if constexpr (FORMULA == FORMULA_A10) {
                    const double bb = first ? b : __dadd_rn(sm.bs[c][tid], b);
                    sm.bs[c][tid] = (c < ncv) ? bb : sm.bs[ncv - 1][tid];
                }
            }
        }
#endif
        if (bulk) mbar_wait(mbar_s, phase);          // the M slab has landed
        phase ^= 1u;

        // ---- units ---------------------------------------------------------------------------
        const int jg0 = j0 / E;
        int kmine = kFrozen;                          // lane g keeps the count of unit g of this slab
        unsigned short *kup = a.kunit + (size_t)rw * (size_t)a.nJG + jg0;
        const bool k_lane = warp_ok && lane < UNITS && (j0 + lane * E < jc1);
        if (MODE == MODE_TOPUP && k_lane) kmine = kup[lane];
        if (!warp_ok) {                               // every row of this warp is beyond n: no units, zero V image
#pragma unroll
            for (int c = 0; c < kTileCols; ++c) sm.cs[c][tid] = 0.0;
        }
#pragma unroll 1
        for (int g = 0; g < UNITS; ++g) {
            const int jb = j0 + g * E;
            if (jb >= jc1 || !warp_ok) break;
            const int nvalid = min(E, jc1 - jb);
            // t0: the largest converged-at count any finished unit had published when the PREVIOUS
            // unit of this warp started (read one unit ahead so the L2 latency is off the critical
            // path).  It never exceeds the global count u, so running this unit at least that far
            // cannot overshoot.
            const int t0 = t0_next;
            if (MODE == MODE_FIRST) t0_next = *kstar_live;
            double m[E], c0[E], w[E], c1[E], c2[E], sc[E], r[E], sexp[E], f[E];
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int c = g * E + e;
                m[e] = sm.ms[c][tid];
                c0[e] = sm.cs[c][tid];
                if (by_col) { c1[e] = sm.c1s[c]; c2[e] = sm.c2s[c]; }
                else { c1[e] = c1_row; c2[e] = c2_row; }
                const double c0m1 = __dadd_rn(c0[e], -1.0);
                w[e] = __dmul_rn(c0m1, c1[e]);                                     // x + const3
                sc[e] = (FORMULA == FORMULA_A10) ? __dmul_rn(l0_i, sm.f0s[c]) : 1.0;   // :83,:97 | S = 1
                if (MODE == MODE_FIRST) {
                    double cap = c0m1;                                             // :16
                    if constexpr (FORMULA == FORMULA_A10) cap = sm.bs[c][tid];     // :78
                    m[e] = (cap < m[e]) ? cap : m[e];                              // Rfast::Pmin = std::min(M, cap): a NaN in M stays
                }
            }
            int k = 0, kb, kk = 0;
            if (MODE == MODE_TOPUP) {
                kk = __shfl_sync(0xffffffffu, kmine, g);
                k = (kk >= target) ? target : kk;     // already there (or frozen): one evaluation regenerates its state
                if (kk < target) ++my_topups;
            }
            const int flags = newton_unit<E, FORMULA, MODE, FORMULA == FORMULA_A10, FREEZE>(
                m, c0, w, c1, c2, sc, tol_lane, k, kb, (MODE == MODE_FIRST) ? t0 : target, a.maxiter, a.delta, exptab_s, r,
                sexp, f, my_evals);
            if (flags & 6) { any_exhausted = 1; kb = a.maxiter; }
            if (MODE == MODE_TOPUP && !(flags & 3)) any_fail = 1;      // stopped at K* but not below tol
            if (!(flags & 1)) {                       // not converged: a NaN in f shows up here (it never converges)
                bool lane_nan = false;
#pragma unroll
                for (int e = 0; e < E; ++e) lane_nan = lane_nan || (row_ok && e < nvalid && (f[e] != f[e]));
                if (__any_sync(0xffffffffu, lane_nan)) any_nan = 1;
            }
            // ---- finalise: M in place, V image, this row's partial sums ---------------------------
            double vprod = 1.0;
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int c = g * E + e;
                sm.ms[c][tid] = m[e];
                if (FORMULA == FORMULA_A1) {
                    const double v = __dmul_rn(__dadd_rn(c2[e], c2[e]), r[e]);          // :36 Sigma2/temp (stale if exhausted)
                    double ex = sexp[e];                                                // == exp(M+V/2): S == 1 and the loop broke
                    if (flags & 2) ex = exp_fast(__fma_rn(0.5, v, m[e]), exptab_s);     // exhausted: M moved after temp
                    sm.cs[c][tid] = row_ok ? v : 0.0;
                    if (e < nvalid) {
                        if (a.V_out && row_ok) st_stream(a.V_out + (size_t)i + ns * (size_t)(jb + e), v);
                        acc_exp = __dadd_rn(acc_exp, ex);
                        acc_rowv = __dadd_rn(acc_rowv, v);
                        vprod = __dmul_rn(vprod, v);
                    }
                }
            }
            // sum_e log(V_e)/2 = log(prod_e V_e)/2: one log per lane instead of E (V in (0, sigma2]; a product
            // of E <= 4 such values stays far from underflow); the constant is added once per tile
            if (FORMULA == FORMULA_A1) acc_log = fma(0.5, log_fast(vprod, logtab_s), acc_log);          // R/ebpmf_log.R:318
            int knew = (flags & 8) ? kFrozen : k;
            if (MODE == MODE_TOPUP && kk == kFrozen) knew = kFrozen;
            if (lane == g) kmine = knew;
            wkmin = min(wkmin, knew);
            if (MODE == MODE_FIRST) {
                wkb = max(wkb, kb);
                // K* = max over units of the count at which the unit stopped (its first converged k >= t0, which
                // is <= u), or of the first converged k for a frozen unit: every live unit then has k_unit <= K*
                const int kpub = (flags & 8) ? kb : ((flags & 4) ? a.maxiter : k);
                if (lane == 0 && kpub > t0) publish_kstar(a.status, a.epoch, kpub);
            }
        }
        if (k_lane) kup[lane] = (unsigned short)kmine;
        if (MODE == MODE_FIRST && tid == 0) absorb_peer(a.status, a.epoch, peer_word);
        // ---- the slab leaves: final M by bulk async stores, column sums of V from the V image ------
        if (bulk) fence_async_smem();
        __syncthreads();
        if (bulk) {
            if (warp == 0 && lane < ncv)
                bulk_store(a.M_out + (size_t)r0 + ns * (size_t)(j0 + lane), smem_addr(&sm.ms[lane][0]), (unsigned)(rows_valid * 8));
        } else if (row_ok) {
            for (int c = 0; c < ncv; ++c) st_stream(a.M_out + (size_t)i + ns * (size_t)(j0 + c), sm.ms[c][tid]);
        }
        if (FORMULA == FORMULA_A1) {
            // a4: sum(Y * M) (R/ebpmf_log.R:316) over the slab's non-zeros against the final M in shared memory
            const int total = sm.ybase[kTileCols];
            for (int idx = tid; idx < min(total, kYList); idx += kThreads) {
                const int pos = sm.ylp[idx];
                acc_sym = fma(sm.ylv[idx], sm.ms[pos >> 8][pos & 255], acc_sym);
            }
            if (total > kYList) {                  // more non-zeros than the list holds: re-walk the rest
                const int *rp0 = a.rbptr + (size_t)rb * p + j0, *rp1 = rp0 + p;
                for (int c = warp; c < ncv; c += kWarps) {
                    const int lo = rp0[c], hi = rp1[c], base = sm.ybase[c] - lo;
                    for (int q = lo + lane; q < hi; q += 32)
                        if (base + q >= kYList) acc_sym = fma(__ldg(a.yx + q), sm.ms[c][a.rowidx[q] - r0], acc_sym);
                }
            }
        }
        if (FORMULA == FORMULA_A1) {
            // 16 threads per column, each 16 rows (stride 16), then a fixed shuffle tree: deterministic
            const int col = tid >> 4, part = tid & 15;
            double s = 0.0;
#pragma unroll
            for (int q = 0; q < kRowsPerCta / 16; ++q) s += sm.cs[col][part + 16 * q];
            s += __shfl_xor_sync(0xffffffffu, s, 8);
            s += __shfl_xor_sync(0xffffffffu, s, 4);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            if (part == 0 && col < ncv) a.colV[(size_t)rb * p + j0 + col] = s;
        }
    }

    // the last slab's bulk stores must have left shared memory before the CTA exits
    if (bulk && warp == 0 && lane < kTileCols) bulk_wait_read();
    // ---- tile epilogue: sums over the tile, order bookkeeping, status flags ------------------------
    if (FORMULA == FORMULA_A1) {
        if (!row_ok) { acc_exp = 0.0; acc_log = 0.0; }
        else acc_log = fma(kSlogvConst, (double)(jc1 - jc0), acc_log);
        if (a.rowpart && row_ok) a.rowpart[(size_t)span * ns + i] = acc_rowv;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        acc_exp += __shfl_xor_sync(0xffffffffu, acc_exp, o);
        acc_log += __shfl_xor_sync(0xffffffffu, acc_log, o);
        acc_sym += __shfl_xor_sync(0xffffffffu, acc_sym, o);
    }
    __syncthreads();
    if (lane == 0) { sm.wsum[0][warp] = acc_exp; sm.wsum[1][warp] = acc_log; sm.wsum[2][warp] = acc_sym; sm.wk[0][warp] = wkb; sm.wk[1][warp] = wkmin; }
    __syncthreads();
    if (tid == 0) {
        double s0 = 0.0, s1 = 0.0, s2 = 0.0;
        int hb = 0, km = kFrozen;
        for (int q = 0; q < kWarps; ++q) { s0 += sm.wsum[0][q]; s1 += sm.wsum[1][q]; s2 += sm.wsum[2][q]; hb = max(hb, sm.wk[0][q]); km = min(km, sm.wk[1][q]); }
        if (FORMULA == FORMULA_A1) { a.tileS[(size_t)tile * 3] = s0; a.tileS[(size_t)tile * 3 + 1] = s1; a.tileS[(size_t)tile * 3 + 2] = s2; }
        if (MODE == MODE_FIRST) a.tile_hard[tile] = (unsigned short)hb;
        a.tile_kmin[tile] = (unsigned short)km;
        if (MODE == MODE_TOPUP) {
            atomicAdd(&a.status->topup_tiles, 1);
            if (a.chunk_dirty) a.chunk_dirty[a.span_rank[span]] = 1;
        }
    }
    if (lane == 0) {
        if (any_exhausted) a.status->exhausted = 1;
        if (any_nan) a.status->nan = 1;
        if (any_fail) a.status->verify_fail = 1;
        if (my_topups) atomicAdd(&a.status->topup_units, (unsigned long long)my_topups);
        if (my_evals) atomicAdd(&a.status->unit_iters, (unsigned long long)my_evals);
    }
}

// column span of (one of) the hardest tile(s) of the step that just ran: where the next step should start
__global__ void __launch_bounds__(1024) hardest_span_kernel(const unsigned short *tile_hard, int ntiles, int nspans, int *out)
{
    __shared__ int best[1024];
    int b = -1;
    for (int t = threadIdx.x; t < ntiles; t += 1024) {
        const int key = ((int)tile_hard[t] << 15) | (0x7FFF - min(t % nspans, 0x7FFF));     // ties: lowest span
        b = max(b, key);
    }
    best[threadIdx.x] = b;
    __syncthreads();
    for (int s = 512; s > 0; s >>= 1) {
        if (threadIdx.x < s) best[threadIdx.x] = max(best[threadIdx.x], best[threadIdx.x + s]);
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = (best[0] < 0) ? 0 : (0x7FFF - (best[0] & 0x7FFF));
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

    // raw operands of one unit (E columns of this thread's row); a padding column shadows column jb
    struct Raw { double m[E], y[E], b[E], s2[E], s[E]; };
    auto load_raw = [&](long long jb, Raw &q) {
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const long long jj = (jb + e < jc1) ? jb + e : jb;
            q.m[e] = row_ok ? ld_stream(Msrc + i + a.n * jj) : 0.0;
            q.y[e] = row_ok ? ld_stream(a.X + i + a.n * jj) : 0.0;
            q.b[e] = row_ok ? a.Beta[ii * a.B_rs + jj * a.B_cs] : 0.0;
            q.s2[e] = row_ok ? a.Sigma2[ii * a.V_rs + jj * a.V_cs] : 1.0;
            q.s[e] = row_ok ? a.S[ii * a.S_rs + jj * a.S_cs] : 1.0;
        }
    };
    // first pass: the NEXT unit's operands are requested before this unit iterates (the unit's ~7 evaluations hide
    // the HBM latency); the top-up pass skips most units, so it loads on demand
    Raw nxt;
    if (MODE == MODE_FIRST && jc0 < jc1) load_raw(jc0, nxt);
#pragma unroll 1
    for (long long jb = jc0; jb < jc1; jb += E) {
        const long long jg = jb / E;
        int k = 0;
        const int t0 = (MODE == MODE_FIRST) ? *kstar_live : 0;
        if (MODE == MODE_TOPUP) {
            k = a.kunit[jg * a.nRW + rw];
            if (k >= target) continue;
        }
        Raw cur;
        if (MODE == MODE_FIRST) {
            cur = nxt;
            if (jb + E < jc1) load_raw(jb + E, nxt);
        } else {
            load_raw(jb, cur);
        }
        double m[E], c0[E], w[E], c1[E], c2[E], sc[E], s2[E], r[E], sexp[E], f[E];
        bool ok[E];
#pragma unroll
        for (int e = 0; e < E; ++e) {
            ok[e] = row_ok && (jb + e < jc1);
            m[e] = cur.m[e]; s2[e] = cur.s2[e]; sc[e] = cur.s[e];
            c1[e] = rcp_fast(s2[e]);                                               // :10
            c2[e] = __dmul_rn(s2[e], 0.5);                                         // :11
            c0[e] = __dadd_rn(__fma_rn(s2[e], cur.y[e], cur.b[e]), 1.0);           // :9
            const double c0m1 = __dadd_rn(c0[e], -1.0);
            w[e] = __dmul_rn(c0m1, c1[e]);                                         // x + const3
            if (MODE == MODE_FIRST) m[e] = (c0m1 < m[e]) ? c0m1 : m[e];            // :16  std::min(M, const0-1): a NaN in M stays
        }
        int kb;
        unsigned evals = 0;
        const int flags = newton_unit<E, FORMULA_A1, MODE, true, FREEZE_PERIODIC>(
            m, c0, w, c1, c2, sc, row_ok ? a.tol : INFINITY, k, kb, (MODE == MODE_FIRST) ? t0 : target, a.maxiter, 0.0,
            exptab_s, r, sexp, f, evals);
        if (flags & 6) any_exhausted = 1;              // 4: parked at maxiter-1, the top-up pass performs the last update
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
            const int kpub = (flags & 4) ? a.maxiter : k;
            if (MODE == MODE_FIRST && kpub > t0) atomicMax(&a.status->kstar, kpub);
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

// one block: scalars[0] = sym, [1] = ssexp, [2] = slogv (from the tile sums [ntiles][3]), [3] = sum(V)
__global__ void __launch_bounds__(1024) reduce_scalars_kernel(const double *v_sum, long long p, const double *tilesum,
                                                              long long ntiles, double *scalars)
{
    __shared__ double sm[4][1024];
    double acc[4] = {0, 0, 0, 0};
    for (long long j = threadIdx.x; j < p; j += 1024) acc[3] += v_sum[j];
    for (long long g = threadIdx.x; g < ntiles; g += 1024) { acc[1] += tilesum[g * 3]; acc[2] += tilesum[g * 3 + 1]; acc[0] += tilesum[g * 3 + 2]; }
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

// R/ebpmf_log.R:425 gates the by_row update on `rowMaxs(fitted(fit_flash))`, and the package imports rowMaxs from
// Rfast (NAMESPACE: importFrom(Rfast,rowMaxs); colMaxs comes from matrixStats).  Rfast::rowMaxs(x, value = FALSE)
// returns the POSITION of each row's maximum (1-based, first one on ties), not its value -- so what the reference
// compares with the threshold is a column index.  Reproduced as is.
__global__ void __launch_bounds__(256) fitted_rowmax_kernel(const double *EL, const double *EF, long long n,
                                                            long long p, int K, double *out)
{
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    double mx = -INFINITY;
    long long jbest = 0;
    for (long long j = 0; j < p; ++j) {
        double b = 0.0;
        for (int k = 0; k < K; ++k) b = fma(EL[i + n * k], EF[j + p * k], b);
        if (b > mx) { mx = b; jbest = j; }
    }
    out[i] = (double)(jbest + 1);
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
// Every sum() is taken at ITS OWN length, as R does: the sigma2-only terms over length(sigma2), the two
// mixed terms over the longer of their operands with the shorter one recycled.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) obj_terms_kernel(const double *sv, long long lsv, const double *sigma2, long long ls2,
                                                         const double *R2, long long lr2, double ss, double a0, double b0,
                                                         double *terms)
{
    __shared__ double sm[5][1024];
    double acc[5] = {0, 0, 0, 0, 0};
    const double two_pi = 6.283185307179586;
    for (long long q = threadIdx.x; q < ls2; q += 1024) {
        const double s2 = sigma2[q];
        acc[0] += ss * log(two_pi * s2) / 2;
        acc[3] += (a0 + 1) * log(s2);
        acc[4] += b0 / s2;
    }
    const long long l1 = max(lsv, ls2), l2 = max(lr2, ls2);
    for (long long q = threadIdx.x; q < l1; q += 1024) acc[1] += sv[q % lsv] / 2 / sigma2[q % ls2];
    for (long long q = threadIdx.x; q < l2; q += 1024) acc[2] += R2[q % lr2] / 2 / sigma2[q % ls2];
    for (int q = 0; q < 5; ++q) sm[q][threadIdx.x] = acc[q];
    __syncthreads();
    for (int s = 512; s > 0; s >>= 1) {
        if (threadIdx.x < s)
            for (int q = 0; q < 5; ++q) sm[q][threadIdx.x] += sm[q][threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x < 5) terms[threadIdx.x] = sm[threadIdx.x][0];
}

// ---------------------------------------------------------------------------------------
// N4: calc_split_PMF_obj_flashier  inst/code/Poisson_MF_splitting.R:373-391 (archived driver): the un-fused ELBO
// from dense Y, S, M, V.  One streaming pass (32 B per entry): thread = row, CTA = 256 rows x a column span;
// per-thread sum of  Y*M - S*exp(M+V/2) + log(2*pi*V)/2 + 0.5  (:389), sums of V per (row block, column) and per
// (span, row) for sv (:377-388).  Partials are combined by the fixed-order reduction kernels.
// ---------------------------------------------------------------------------------------
struct SplitObjArgs {
    long long n, p;
    const double *Y, *M, *V;
    const double *S; long long S_rs, S_cs;
    double *colV;        // [nRB][p]
    double *rowpart;     // [nspans][n]
    double *termpart;    // [nRB * nspans]
    int cols_per_cta;
};

__global__ void __launch_bounds__(kThreads) split_obj_kernel(const SplitObjArgs a)
{
    __shared__ double red[kWarps];
    __shared__ double vs[kTileCols][kRowsPerCta + 1];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long rb = blockIdx.y, i = rb * kRowsPerCta + tid;
    const bool row_ok = i < a.n;
    const long long jc0 = (long long)blockIdx.x * a.cols_per_cta, jc1 = min(a.p, jc0 + (long long)a.cols_per_cta);
    const double two_pi = 6.283185307179586;
    double term = 0.0, rowv = 0.0;
    for (long long j0 = jc0; j0 < jc1; j0 += kTileCols) {
        const int ncv = (int)min((long long)kTileCols, jc1 - j0);
        __syncthreads();
#pragma unroll 4
        for (int c = 0; c < kTileCols; ++c) {
            double v = 0.0;
            if (row_ok && c < ncv) {
                const long long e = i + a.n * (j0 + c);
                const double y = ld_stream(a.Y + e), m = ld_stream(a.M + e);
                v = ld_stream(a.V + e);
                const double s = a.S[i * a.S_rs + (j0 + c) * a.S_cs];
                term += ((y * m - s * exp(m + v / 2)) + log(two_pi * v) / 2) + 0.5;
                rowv += v;
            }
            vs[c][tid] = v;
        }
        __syncthreads();
        const int col = tid >> 4, part = tid & 15;          // 16 threads per column, fixed shuffle tree
        double sacc = 0.0;
#pragma unroll
        for (int q = 0; q < kRowsPerCta / 16; ++q) sacc += vs[col][part + 16 * q];
        sacc += __shfl_xor_sync(0xffffffffu, sacc, 8);
        sacc += __shfl_xor_sync(0xffffffffu, sacc, 4);
        sacc += __shfl_xor_sync(0xffffffffu, sacc, 2);
        sacc += __shfl_xor_sync(0xffffffffu, sacc, 1);
        if (part == 0 && col < ncv) a.colV[rb * a.p + j0 + col] = sacc;
    }
    if (row_ok) a.rowpart[(long long)blockIdx.x * a.n + i] = rowv;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) term += __shfl_xor_sync(0xffffffffu, term, o);
    __syncthreads();
    if (lane == 0) red[warp] = term;
    __syncthreads();
    if (tid == 0) {
        double s = 0.0;
        for (int q = 0; q < kWarps; ++q) s += red[q];
        a.termpart[rb * gridDim.x + blockIdx.x] = s;
    }
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
