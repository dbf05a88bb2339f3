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
// Two streaming kernels, one per product, each reads M exactly once (8 B per entry; the bound is HBM for
// small K and the FP64 pipe beyond K ~ 24: 1 DFMA per entry per factor column):
//   m_ef_kernel   thread = row: coalesced column walks, K accumulators per thread in registers, EF rows of the
//                 current 16-column slab broadcast from shared memory (double-buffered, one barrier per slab);
//   mt_el_kernel  CTA = 16 columns x a range of rows; per 256-row tile the M slab (16 x 2 KB column segments)
//                 and the EL tile (K x 2 KB) arrive by bulk async copies on one mbarrier; every thread owns
//                 2 columns x K/8 factor columns of the result for a quarter of the tile's rows, accumulators
//                 live in registers across the whole row range.
// Both write small per-split partials that are summed in a fixed order (run-to-run reproducible).
#pragma once
#include "vga_kernels.cuh"

namespace ebpmf {

struct ProdArgs {
    long long n, p;
    int K;                       // factor columns (<= KP)
    const double *M;             // n x p column-major, resident
    const double *EL;            // n x K column-major
    const double *EFt;           // p x KP row-major, zero padded
    double *MFpart;              // [nsplit][KP][n]
    double *rowsq;               // [nsplit][n]
    double *MtLpart;             // [nrange][p][KP]
    double *colsq;               // [nrange][p]
    int nsplit, cols_per_split;  // m_ef_kernel: column ranges (multiples of 16 columns)
    int nrange, rows_per_range;  // mt_el_kernel: row ranges (multiples of 256 rows)
    int bulk;
};

// ---- MF = M %*% EF, row sums of M^2 ---------------------------------------------------------
template <int KP>
__global__ void __launch_bounds__(kThreads, 2) m_ef_kernel(const ProdArgs a)
{
    __shared__ __align__(16) double efs[2][kTileCols][KP];
    const int tid = threadIdx.x;
    const int n = (int)a.n, p = (int)a.p;
    const int i = blockIdx.y * kRowsPerCta + tid;
    const bool row_ok = i < n;
    const int jc0 = blockIdx.x * a.cols_per_split;
    const int jc1 = min(p, jc0 + a.cols_per_split);
    const size_t ns = (size_t)n;
    double acc[KP];
#pragma unroll
    for (int k = 0; k < KP; ++k) acc[k] = 0.0;
    double sq = 0.0;
    auto stage = [&](int buf, int j0) {
        for (int idx = tid; idx < kTileCols * KP; idx += kThreads) {
            const int c = idx / KP, k = idx - c * KP;
            efs[buf][c][k] = (j0 + c < jc1) ? __ldg(a.EFt + (size_t)(j0 + c) * KP + k) : 0.0;
        }
    };
    if (jc0 < jc1) stage(0, jc0);
    __syncthreads();
    int buf = 0;
    for (int j0 = jc0; j0 < jc1; j0 += kTileCols, buf ^= 1) {
        double m[kTileCols];
        const double *mp = a.M + (size_t)i + ns * (size_t)j0;
#pragma unroll
        for (int c = 0; c < kTileCols; ++c) m[c] = (row_ok && j0 + c < jc1) ? ld_stream(mp + c * ns) : 0.0;
        if (j0 + kTileCols < jc1) stage(buf ^ 1, j0 + kTileCols);          // next slab's EF rows, other buffer
#pragma unroll
        for (int c = 0; c < kTileCols; ++c) {
            sq = fma(m[c], m[c], sq);
#pragma unroll
            for (int k = 0; k < KP; ++k) acc[k] = fma(m[c], efs[buf][c][k], acc[k]);
        }
        __syncthreads();
    }
    if (row_ok) {
        double *out = a.MFpart + (size_t)blockIdx.x * KP * ns + i;
#pragma unroll
        for (int k = 0; k < KP; ++k) out[(size_t)k * ns] = acc[k];
        a.rowsq[(size_t)blockIdx.x * ns + i] = sq;
    }
}

// ---- MtL = t(M) %*% EL, column sums of M^2 -----------------------------------------------------
constexpr int kProdStride = kRowsPerCta + 2;      // column stride (doubles) of the padded smem images: 16-byte
                                                  // aligned rows, and 8 consecutive columns hit 8 distinct bank pairs
template <int KP>
struct MtElSmem {
    double ms[kTileCols][kProdStride];
    double els[KP][kProdStride];      // after the row loop its storage is reused for the cross-group reduction
    unsigned long long mbar;
};

template <int KP>
__global__ void __launch_bounds__(kThreads, 2) mt_el_kernel(const ProdArgs a)
{
    constexpr int KT = KP / 8;                    // factor columns per thread: kt, kt+8, ...
    extern __shared__ __align__(128) unsigned char smem_raw[];
    MtElSmem<KP> &sm = *reinterpret_cast<MtElSmem<KP> *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = (int)a.n, p = (int)a.p;
    const int j0 = blockIdx.x * kTileCols;
    const int ncv = min(kTileCols, p - j0);
    const int ra = blockIdx.y * a.rows_per_range;
    const int rb = min(n, ra + a.rows_per_range);
    const size_t ns = (size_t)n;
    const bool bulk = a.bulk != 0;
    const unsigned mbar_s = smem_addr(&sm.mbar);
    if (tid == 0) {
        mbar_init(mbar_s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_async_smem();
    }
    // thread roles: rg = quarter of the tile's rows, cp = column pair, kt = first factor column
    const int rg = tid >> 6, u = tid & 63, cp = u >> 3, kt = u & 7;
    const int sc = tid >> 4, sp = tid & 15;        // column / row phase for the sums of squares
    double acc[2][KT];
#pragma unroll
    for (int q = 0; q < KT; ++q) { acc[0][q] = 0.0; acc[1][q] = 0.0; }
    double sq = 0.0;
    unsigned phase = 0;
    // what the copies never write: columns past p, factor columns past K (zeroed once)
    for (int idx = tid; idx < (kTileCols + KP) * kRowsPerCta; idx += kThreads) {
        const int c = idx >> 8, r = idx & 255;
        if (c < kTileCols) { if (c >= ncv) sm.ms[c][r] = 0.0; }
        else if (c - kTileCols >= a.K) sm.els[c - kTileCols][r] = 0.0;
    }
    for (int r0 = ra; r0 < rb; r0 += kRowsPerCta) {
        const int rows_valid = min(kRowsPerCta, rb - r0);
        __syncthreads();                             // previous tile consumed (and the mbarrier initialised)
        if (bulk) {
            if (warp == 0) {
                if (lane == 0) mbar_expect_tx(mbar_s, (unsigned)((ncv + a.K) * rows_valid * 8));
                __syncwarp();
                if (lane < ncv) bulk_load(smem_addr(&sm.ms[lane][0]), a.M + (size_t)r0 + ns * (size_t)(j0 + lane), (unsigned)(rows_valid * 8), mbar_s);
                for (int k = lane; k < a.K; k += 32)
                    bulk_load(smem_addr(&sm.els[k][0]), a.EL + (size_t)r0 + ns * (size_t)k, (unsigned)(rows_valid * 8), mbar_s);
            }
            if (rows_valid < kRowsPerCta) {          // ragged last tile: rows the copies do not write
                for (int idx = tid; idx < (kTileCols + KP) * kRowsPerCta; idx += kThreads) {
                    const int c = idx >> 8, r = idx & 255;
                    if (r >= rows_valid) { if (c < kTileCols) sm.ms[c][r] = 0.0; else sm.els[c - kTileCols][r] = 0.0; }
                }
            }
            mbar_wait(mbar_s, phase);
            phase ^= 1u;
        } else {
            for (int idx = tid; idx < (kTileCols + KP) * kRowsPerCta; idx += kThreads) {
                const int c = idx >> 8, r = idx & 255;
                if (c < kTileCols) sm.ms[c][r] = (c < ncv && r < rows_valid) ? ld_stream(a.M + (size_t)(r0 + r) + ns * (size_t)(j0 + c)) : 0.0;
                else sm.els[c - kTileCols][r] = (c - kTileCols < a.K && r < rows_valid) ? __ldg(a.EL + (size_t)(r0 + r) + ns * (size_t)(c - kTileCols)) : 0.0;
            }
        }
        __syncthreads();
        const double *m0 = &sm.ms[2 * cp][64 * rg], *m1 = &sm.ms[2 * cp + 1][64 * rg];
        const double *e0 = &sm.els[kt][64 * rg];
#pragma unroll 8
        for (int r = 0; r < 64; ++r) {
            const double x0 = m0[r], x1 = m1[r];
#pragma unroll
            for (int q = 0; q < KT; ++q) {
                const double e = e0[(size_t)q * 8 * kProdStride + r];
                acc[0][q] = fma(x0, e, acc[0][q]);
                acc[1][q] = fma(x1, e, acc[1][q]);
            }
        }
#pragma unroll
        for (int q = 0; q < kRowsPerCta / 16; ++q) { const double x = sm.ms[sc][sp + 16 * q]; sq = fma(x, x, sq); }
    }
    // combine the four row groups (fixed order), write this range's partial
    __syncthreads();
    static_assert(sizeof(double) * 4 * kTileCols * (KP + 1) <= sizeof(double) * KP * kProdStride, "reduction scratch fits the EL tile");
    struct Red { double red[4][kTileCols][KP + 1]; };
    Red &smr = *reinterpret_cast<Red *>(&sm.els[0][0]);
#pragma unroll
    for (int q = 0; q < KT; ++q) { smr.red[rg][2 * cp][kt + 8 * q] = acc[0][q]; smr.red[rg][2 * cp + 1][kt + 8 * q] = acc[1][q]; }
    sq += __shfl_xor_sync(0xffffffffu, sq, 8);
    sq += __shfl_xor_sync(0xffffffffu, sq, 4);
    sq += __shfl_xor_sync(0xffffffffu, sq, 2);
    sq += __shfl_xor_sync(0xffffffffu, sq, 1);
    __syncthreads();
    for (int idx = tid; idx < kTileCols * KP; idx += kThreads) {
        const int c = idx / KP, k = idx - c * KP;
        if (c < ncv)
            a.MtLpart[((size_t)blockIdx.y * p + j0 + c) * KP + k] = ((smr.red[0][c][k] + smr.red[1][c][k]) + smr.red[2][c][k]) + smr.red[3][c][k];
    }
    if (sp == 0 && sc < ncv) a.colsq[(size_t)blockIdx.y * p + j0 + sc] = sq;
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

// MtL (p x Kc, column-major, leading dimension p) from the row-major partials [nrange][p][KP]
__global__ void mtl_gather_kernel(const double *part, int nrange, long long p, int KP, int Kc, double *out)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= p * Kc) return;
    const long long j = t % p; const int k = (int)(t / p);
    double s = 0.0;
    for (int q = 0; q < nrange; ++q) s += part[((size_t)q * p + j) * KP + k];
    out[t] = s;
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
