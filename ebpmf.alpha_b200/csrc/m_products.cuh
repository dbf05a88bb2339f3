// N1: what flashier needs FROM the latent M, computed where M lives.
//
// flashier treats res$M as its data matrix (fit_flash$flash_fit$Y = res$M, R/ebpmf_log.R:270,309;
// flash_init(fit_flash$flash_fit$Y, ...), R/ebpmf_log_flash_update.R:35) and only ever touches it through
// products with the current loadings / factors and through the residual sums of squares R2
// (R/ebpmf_log_flash_update.R:55-64).  With these on the device M never has to cross PCIe:
//   MF  (n x K) = M %*% EF        MtL (p x K) = t(M) %*% EL
//   r2_col[j] = sum_i (M_ij - sum_k EL_ik EF_jk)^2 ,  r2_row[i] likewise over j,
// the latter from sum M^2 - 2 <M, EL EF'> + <EL EF', EL EF'>, i.e. from the products themselves and the two
// K x K Gram matrices (how flashier forms R2 as well): no third pass over M.
//
// ONE streaming kernel reads M exactly once (8 B per entry) and forms both products on the FP64 TENSOR path
// (mma.sync.m8n8k4.f64 = DMMA).  Measured on B200 (scripts/probes/dmma_probe.cu, profiles/r02_probe_dmma.txt):
// DMMA and DFMA share the FP64 pipe and the same peak (36.7 vs 34.4 TFLOP/s), but an accumulating DFMA
// (three distinct register sources) issues in 3 cycles instead of 2, so a DFMA GEMM tops out at ~2/3 of that
// peak while DMMA can reach it, and it frees issue slots for the loads.  4 K flop per entry make the pass
// FP64-bound beyond K ~ 12 (HBM-bound below): K = 22 costs >= 9 ms on the 125 000 x 30 000 slab, HBM alone 4.6 ms.
//   CTA = NW warps = 32 NW rows x a span of columns, walked in slabs of 16 columns; warp w owns rows [32w, 32w+32).
//   M slab: 16 column segments by bulk async copies, double-buffered (the next slab lands while this one is
//   multiplied); the slab's EF rows ride on the same mbarrier.
//   M %*% EF : A = M (8 rows x 4 columns per DMMA), B = EF rows of the slab (shared memory), accumulators
//              (32 rows x KP) stay in registers across the whole span;
//   t(M)%*%EL: A = t(M) (the same shared image read transposed), B = this warp's EL rows, held in registers
//              for the whole span; the slab's 16 x KP result is summed over the warps through shared memory
//              (fixed order) and written as this row block's partial, one slab behind the multiplication.
// Partials are summed by the small kernels below in a fixed order (run-to-run reproducible).
#pragma once
#include "vga_kernels.cuh"

namespace ebpmf {

struct ProdArgs {
    long long n, p;
    int K;                       // factor columns (<= KP)
    const double *M;             // n x p column-major, resident
    const double *EL;            // n x K column-major
    const double *EFt;           // p x (KP + 4) row-major, zero padded
    double *MFpart;              // [nsplit][KP][n]
    double *rowsq;               // [nsplit][n]
    double *MtLpart;             // [row blocks][p][KP]
    double *colsq;               // [row blocks][p]
    int nsplit, cols_per_split;  // column spans (multiples of 16 columns)
    int bulk;
};

template <int KP, int NW>
struct ProdSmem {
    // column stride rows + 4: the fragment loads of both products (8 rows x 4 columns, and 4 rows x 8 columns)
    // hit 16 distinct bank pairs per half warp; same argument for the EF rows (KP + 4)
    double ms[2][kTileCols][32 * NW + 4];         // M slab, double-buffered
    double efs[2][kTileCols][KP + 4];             // EF rows of the slab
    double red[2][NW][kTileCols][KP];             // per-warp t(M) EL of the slab, double-buffered
    unsigned long long mbar[2];
};

template <int KP, int NW>
__global__ void __launch_bounds__(32 * NW, 1) m_products_kernel(const ProdArgs a)
{
    constexpr int NT = KP / 8;                    // 8-wide tiles of factor columns
    constexpr int ROWS = 32 * NW, NTH = 32 * NW;
    constexpr int TPC = 16;                       // threads per column for the column sums of squares (threads 0..255)
    static_assert(NW >= 8, "the column sums of squares use 16 x 16 threads");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    ProdSmem<KP, NW> &sm = *reinterpret_cast<ProdSmem<KP, NW> *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int l4 = lane & 3, l8 = lane >> 2;
    const int n = (int)a.n, p = (int)a.p;
    const int rb = blockIdx.y, r0 = rb * ROWS;
    const int rows_valid = min(ROWS, n - r0);
    const int jc0 = blockIdx.x * a.cols_per_split;
    const int jc1 = min(p, jc0 + a.cols_per_split);
    const size_t ns = (size_t)n;
    const bool bulk = a.bulk != 0;
    if (jc0 >= jc1) return;
    const unsigned mbar_s[2] = {smem_addr(&sm.mbar[0]), smem_addr(&sm.mbar[1])};
    if (tid == 0) {
        mbar_init(mbar_s[0], 1); mbar_init(mbar_s[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_async_smem();
    }
    // rows the copies never write (ragged last row block): zero once, both buffers
    if (rows_valid < ROWS)
        for (int idx = tid; idx < 2 * kTileCols * ROWS; idx += NTH) {
            const int r = idx % ROWS, c = (idx / ROWS) % kTileCols, b = idx / (ROWS * kTileCols);
            if (r >= rows_valid) sm.ms[b][c][r] = 0.0;
        }
    // this warp's EL rows as B fragments of t(M) %*% EL: el[ks][nt] = EL[r0 + 32w + 4ks + l4][8nt + l8]
    double el[8][NT];
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
        const int i = r0 + 32 * warp + 4 * ks + l4;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            const int k = 8 * nt + l8;
            el[ks][nt] = (i < n && k < a.K) ? __ldg(a.EL + (size_t)i + ns * (size_t)k) : 0.0;
        }
    }
    double accF[4][NT][2];                         // M %*% EF: rows 32w + 8rt + l8, factor columns 8nt + 2*l4 + {0,1}
#pragma unroll
    for (int rt = 0; rt < 4; ++rt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) { accF[rt][nt][0] = 0.0; accF[rt][nt][1] = 0.0; }
    double rsq = 0.0;                              // sum over the span of M[r0 + tid, j]^2
    const int sc = (tid >> 4) & 15, sp = tid & 15;   // column / row phase for the column sums of squares

    auto issue = [&](int buf, int j0) {            // bring the slab [j0, j0+16) and its EF rows into buffer `buf`
        const int ncv = min(kTileCols, jc1 - j0);
        if (ncv < kTileCols) {                     // columns past the span: zero (no copy writes them)
            for (int idx = tid; idx < (kTileCols - ncv) * ROWS; idx += NTH) sm.ms[buf][ncv + idx / ROWS][idx % ROWS] = 0.0;
            for (int idx = tid; idx < (kTileCols - ncv) * KP; idx += NTH) sm.efs[buf][ncv + idx / KP][idx % KP] = 0.0;
        }
        if (bulk) {
            // one bulk copy per M column, spread over the warps, plus ONE copy for the slab's EF rows (EFt rows are
            // padded to KP + 4 doubles in global memory exactly as in shared memory)
            if (tid == 0) mbar_expect_tx(mbar_s[buf], (unsigned)(ncv * rows_valid * 8 + ncv * (KP + 4) * 8));
            constexpr int CPW = (kTileCols + NW - 1) / NW;      // columns per warp
            if (lane < CPW) {
                const int c = warp * CPW + lane;
                if (c < ncv)
                    bulk_load(smem_addr(&sm.ms[buf][c][0]), a.M + (size_t)r0 + ns * (size_t)(j0 + c), (unsigned)(rows_valid * 8), mbar_s[buf]);
            }
            if (warp == NW - 1 && lane == CPW)
                bulk_load(smem_addr(&sm.efs[buf][0][0]), a.EFt + (size_t)j0 * (KP + 4), (unsigned)(ncv * (KP + 4) * 8), mbar_s[buf]);
        } else {
            for (int idx = tid; idx < ncv * ROWS; idx += NTH) {
                const int c = idx / ROWS, r = idx % ROWS;
                if (r < rows_valid) sm.ms[buf][c][r] = ld_stream(a.M + (size_t)(r0 + r) + ns * (size_t)(j0 + c));
            }
            for (int idx = tid; idx < ncv * (KP + 4); idx += NTH) (&sm.efs[buf][0][0])[idx] = __ldg(a.EFt + (size_t)j0 * (KP + 4) + idx);
        }
    };
    // sum the warps' t(M) EL of a finished slab (fixed order) into this row block's partial
    auto flush = [&](int buf, int j0) {
        const int ncv = min(kTileCols, jc1 - j0);
        for (int idx = tid; idx < kTileCols * KP; idx += NTH) {
            const int c = idx / KP, k = idx - c * KP;
            if (c < ncv) {
                double s = 0.0;
#pragma unroll
                for (int w = 0; w < NW; ++w) s += sm.red[buf][w][c][k];
                a.MtLpart[((size_t)rb * p + j0 + c) * KP + k] = s;
            }
        }
    };

    __syncthreads();                               // mbarriers initialised, ragged rows zeroed
    issue(0, jc0);
    unsigned phase[2] = {0u, 0u};
    int buf = 0;
    for (int j0 = jc0; j0 < jc1; j0 += kTileCols, buf ^= 1) {
        const int ncv = min(kTileCols, jc1 - j0);
        if (bulk) { mbar_wait(mbar_s[buf], phase[buf]); phase[buf] ^= 1u; }
        // ONE barrier per slab: every warp is done with slab s-1 (its M image and EF rows may be overwritten, its
        // red[] is complete), and -- without bulk copies -- this slab's staged data is visible
        __syncthreads();
        if (j0 + kTileCols < jc1) issue(buf ^ 1, j0 + kTileCols);
        const double (*ms)[ROWS + 4] = sm.ms[buf];
        // ---- M %*% EF ---------------------------------------------------------------------------------
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            double bf[NT];
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) bf[nt] = sm.efs[buf][4 * ks + l4][8 * nt + l8];
#pragma unroll
            for (int rt = 0; rt < 4; ++rt) {
                const double af = ms[4 * ks + l4][32 * warp + 8 * rt + l8];
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) dmma884(accF[rt][nt][0], accF[rt][nt][1], af, bf[nt]);
            }
        }
        // the previous slab's t(M) EL leaves while the FP64 pipe works through the DMMAs issued above
        if (j0 > jc0) flush(buf ^ 1, j0 - kTileCols);
        // ---- t(M) %*% EL over this warp's 32 rows -------------------------------------------------------
        double accT[2][NT][2];
#pragma unroll
        for (int ct = 0; ct < 2; ++ct)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) { accT[ct][nt][0] = 0.0; accT[ct][nt][1] = 0.0; }
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
#pragma unroll
            for (int ct = 0; ct < 2; ++ct) {
                const double af = ms[8 * ct + l8][32 * warp + 4 * ks + l4];
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) dmma884(accT[ct][nt][0], accT[ct][nt][1], af, el[ks][nt]);
            }
        }
#pragma unroll
        for (int ct = 0; ct < 2; ++ct)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
                *reinterpret_cast<double2 *>(&sm.red[buf][warp][8 * ct + l8][8 * nt + 2 * l4]) = make_double2(accT[ct][nt][0], accT[ct][nt][1]);
        // ---- sums of squares: this row over the slab's columns; each column over the tile's rows ----------
        double csq = 0.0;
#pragma unroll
        for (int c = 0; c < kTileCols; ++c) { const double x = ms[c][tid]; rsq = fma(x, x, rsq); }
        if (tid < 256) {
#pragma unroll
            for (int q = 0; q < ROWS / TPC; ++q) { const double x = ms[sc][sp + TPC * q]; csq = fma(x, x, csq); }
        }
#pragma unroll
        for (int o = TPC / 2; o > 0; o >>= 1) csq += __shfl_xor_sync(0xffffffffu, csq, o);
        if (tid < 256 && sp == 0 && sc < ncv) a.colsq[(size_t)rb * p + j0 + sc] = csq;
    }
    __syncthreads();
    flush(buf ^ 1, jc0 + ((jc1 - jc0 - 1) / kTileCols) * kTileCols);       // the last slab
    // ---- this span's M %*% EF partial and row sums of squares ---------------------------------------------
    double *out = a.MFpart + (size_t)blockIdx.x * KP * ns;
#pragma unroll
    for (int rt = 0; rt < 4; ++rt) {
        const int i = r0 + 32 * warp + 8 * rt + l8;
        if (i < n) {
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                out[(size_t)(8 * nt + 2 * l4) * ns + i] = accF[rt][nt][0];
                out[(size_t)(8 * nt + 2 * l4 + 1) * ns + i] = accF[rt][nt][1];
            }
        }
    }
    if (r0 + tid < n) a.rowsq[(size_t)blockIdx.x * ns + r0 + tid] = rsq;
}

// out[t] = sum_s part[s*stride + t], t < len  (fixed order)
__global__ void sum_splits_kernel(const double *part, long long stride, long long len, int nsplit, double *out)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= len) return;
    double s = 0.0;
    for (int q = 0; q < nsplit; ++q) s += part[(size_t)q * stride + t];
    out[t] = s;
}

// MtL (p x Kc, column-major, leading dimension p) from the row-major partials [nrange][p][KP]: one thread per
// (j, k) of the partial layout, so the nrange reads of a warp are contiguous; fixed summation order
__global__ void mtl_gather_kernel(const double *part, int nrange, long long p, int KP, int Kc, double *out)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= p * KP) return;
    const long long j = t / KP; const int k = (int)(t - j * KP);
    double s = 0.0;
    for (int q = 0; q < nrange; ++q) s += part[(size_t)q * (size_t)p * KP + t];
    if (k < Kc) out[j + p * k] = s;
}

// G[k][l] = sum_i A[i + ld*k] * A[i + ld*l]  (K x K Gram matrix of an ld x K column-major matrix), one block per (k,l)
__global__ void __launch_bounds__(256) gram_kernel(const double *A, long long rows, int K, double *G)
{
    __shared__ double red[256];
    const int k = blockIdx.x / K, l = blockIdx.x % K;
    double s = 0.0;
    for (long long i = threadIdx.x; i < rows; i += 256) s = fma(A[i + rows * k], A[i + rows * l], s);
    red[threadIdx.x] = s;
    __syncthreads();
    for (int h = 128; h > 0; h >>= 1) {
        if (threadIdx.x < h) red[threadIdx.x] += red[threadIdx.x + h];
        __syncthreads();
    }
    if (threadIdx.x == 0) G[blockIdx.x] = red[0];
}

// r2[t] = sq[t] - 2 * sum_k A[t,k] * P[t,k] + sum_{k,l} A[t,k] G[k][l] A[t,l]
// A: len x K column-major (the factor on THIS side), P: the product on this side (ldp-strided, see below), G: Gram
// matrix of the OTHER factor.  p_rowmajor: P[t*ldp + k] (MtL layout) else P[t + ldp*k] (MF layout).
__global__ void r2_from_products_kernel(const double *sq, const double *A, const double *P, long long len, int K,
                                        long long ldp, int p_rowmajor, const double *G, double *r2)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= len) return;
    double cross = 0.0, quad = 0.0;
    for (int k = 0; k < K; ++k) {
        const double ak = A[t + len * k];
        cross = fma(ak, p_rowmajor ? P[t * ldp + k] : P[t + ldp * k], cross);
        double g = 0.0;
        for (int l = 0; l < K; ++l) g = fma(G[k * K + l], A[t + len * l], g);
        quad = fma(ak, g, quad);
    }
    r2[t] = (sq[t] - 2.0 * cross) + quad;
}

}  // namespace ebpmf
