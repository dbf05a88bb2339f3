// On-device synthetic inputs: the DGP of sim_data_log (R/simulate_poisson_matrix.R:9-61)
// with a counter-based RNG keyed on (seed, stream, global row, column), so a row shard
// generates exactly its rows of the global matrix whatever the number of GPUs.
// Benchmark support only ("next" row N4) -- not part of the reference API.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ebpmf {
namespace sim {

enum : uint64_t { ST_L = 1, ST_F = 2, ST_FB = 3, ST_L0 = 4, ST_F0 = 5, ST_Y = 6, ST_Y2 = 7, ST_NL = 8, ST_NF = 9, ST_S2 = 10 };

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__host__ __device__ __forceinline__ uint64_t key(uint64_t seed, uint64_t stream, uint64_t a, uint64_t b) {
    return mix64(mix64(mix64(seed ^ (stream * 0xD6E8FEB86659FD93ull)) + a) ^ (b * 0xA24BAED4963EE407ull));
}

// uniform in (0,1)
__host__ __device__ __forceinline__ double u01(uint64_t h) { return ((double)(h >> 11) + 0.5) * (1.0 / 9007199254740992.0); }

__device__ __forceinline__ double normal(uint64_t seed, uint64_t stream, uint64_t a, uint64_t b) {
    const uint64_t h = key(seed, stream, a, b);
    const double u1 = u01(h), u2 = u01(mix64(h));
    return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
}

// Poisson(lambda) by CDF inversion (lambda < 64) or a rounded normal (lambda >= 64)
__device__ __forceinline__ double poisson(double lambda, uint64_t seed, uint64_t i, uint64_t j) {
    const uint64_t h = key(seed, ST_Y, i, j);
    const double u = u01(h);
    if (lambda < 64.0) {
        double pk = exp(-lambda), cdf = pk;
        int k = 0;
        while (u > cdf && k < 400) { ++k; pk *= lambda / k; cdf += pk; }
        return (double)k;
    }
    const double u2 = u01(mix64(h ^ 0x5851F42D4C957F2Dull)), u3 = u01(mix64(h + 0x14057B7EF767814Full));
    const double z = sqrt(-2.0 * log(u2)) * cospi(2.0 * u3);
    return fmax(0.0, floor(lambda + sqrt(lambda) * z + 0.5));
}

struct GenArgs {
    long long n, p, row0;
    int K;                   // true factors; EL/EF get K+2 columns
    uint64_t seed;
    double log_offset, bg_lo, bg_hi, pi0, noise;
    double *Lt, *Ft;         // n x K, p x K  (true loadings / factors)
    double *logl0, *logf0;   // n, p
    double *EL, *EF;         // n x (K+2), p x (K+2) column-major
    double *sigma2;          // p
};

// rows: t < n ; columns: n <= t < n+p
__global__ void gen_factors_kernel(const GenArgs a)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int KT = a.K + 2;
    if (t < a.n) {
        const long long i = t, gi = a.row0 + t;
        const double l0 = a.bg_lo + (a.bg_hi - a.bg_lo) * u01(key(a.seed, ST_L0, gi, 0));
        a.logl0[i] = log(l0);
        a.EL[i] = log(l0) + a.log_offset;
        a.EL[i + a.n] = 1.0;
        for (int k = 0; k < a.K; ++k) {
            const double l = normal(a.seed, ST_L, gi, k);
            a.Lt[i + a.n * k] = l;
            a.EL[i + a.n * (k + 2)] = l + a.noise * normal(a.seed, ST_NL, gi, k);
        }
        (void)KT;
    } else if (t < a.n + a.p) {
        const long long j = t - a.n;
        const double f0 = a.bg_lo + (a.bg_hi - a.bg_lo) * u01(key(a.seed, ST_F0, j, 0));
        a.logf0[j] = log(f0);
        a.EF[j] = 1.0;
        a.EF[j + a.p] = log(f0);
        for (int k = 0; k < a.K; ++k) {
            const double keep = (u01(key(a.seed, ST_FB, j, k)) < (1.0 - a.pi0)) ? 1.0 : 0.0;
            const double f = normal(a.seed, ST_F, j, k) * keep;
            a.Ft[j + a.p * k] = f;
            a.EF[j + a.p * (k + 2)] = f + a.noise * normal(a.seed, ST_NF, j, k);
        }
        a.sigma2[j] = 0.05 + 0.95 * u01(key(a.seed, ST_S2, j, 0));
    }
}

__device__ __forceinline__ double gen_y(const GenArgs &a, long long i, long long j) {
    double s = a.logl0[i] + a.logf0[j] + a.log_offset;
    for (int k = 0; k < a.K; ++k) s = fma(a.Lt[i + a.n * k], a.Ft[j + a.p * k], s);
    return poisson(exp(s), a.seed, (uint64_t)(a.row0 + i), (uint64_t)j);
}

// one warp per column: M0 and the per-column non-zero count
__global__ void gen_count_kernel(const GenArgs a, int warm_start, double *M0, int *counts)
{
    const long long j = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (j >= a.p) return;
    const int KT = a.K + 2;
    int c = 0;
    for (long long i = lane; i < a.n; i += 32) {
        const double y = gen_y(a, i, j);
        c += (y != 0.0);
        double m0;
        if (warm_start) {
            m0 = 0.0;
            for (int k = 0; k < KT; ++k) m0 = fma(a.EL[i + a.n * k], a.EF[j + a.p * k], m0);
        } else {
            m0 = log(y + 0.5);
        }
        M0[i + a.n * j] = m0;
    }
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0) counts[j] = c;
}

__global__ void gen_fill_kernel(const GenArgs a, const int *colptr, int *rowidx, double *x)
{
    const long long j = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (j >= a.p) return;
    int pos = colptr[j];
    for (long long i0 = 0; i0 < a.n; i0 += 32) {
        const long long i = i0 + lane;
        const double v = (i < a.n) ? gen_y(a, i, j) : 0.0;
        const unsigned mask = __ballot_sync(0xffffffffu, v != 0.0);
        if (v != 0.0) {
            const int off = __popc(mask & ((1u << lane) - 1u));
            rowidx[pos + off] = (int)i;
            x[pos + off] = v;
        }
        pos += __popc(mask);
    }
}

}  // namespace sim
}  // namespace ebpmf
