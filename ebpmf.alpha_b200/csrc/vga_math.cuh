// fp64 device math for the VGA Newton kernels (sm_100a).
//
// The hot path is bound by the FP64 pipe, so the number of DP-pipe
// instructions per Newton iteration is what matters.  R's `1/x`, `exp`, `log`
// (R/vga_solver_mat.R:22-31, R/ebpmf_log.R:317-318) are replaced by
//   rcp : MUFU.RCP64H seed (rcp.approx.ftz.f64, rel. err <= 2^-20) + one cubic Newton step
//         r(1+e+e^2): 3 DFMA, result within ~1.5 ulp for normal, finite x;
//   exp : n = rint(x*64/ln2); r = x - n*ln2/64 (Cody-Waite, 2 DFMA); 2^(j/64) from a
//         64-entry table in shared memory; degree-5 polynomial: ~10 DP instructions, rel. err <= 2.3e-16;
//   log : libdevice log() (only once per entry, in the epilogue).
// Parity budget is 1e-8 relative against the oracle (BASELINE.json); these are ~1e-15.
// EBPMF_EXACT_MATH=1 switches to IEEE division and libdevice exp() (used by tests to
// bound the effect of the fast paths).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#ifndef EBPMF_EXACT_MATH
#define EBPMF_EXACT_MATH 0
#endif

namespace ebpmf {

__device__ __forceinline__ double rcp_fast(double x) {
#if EBPMF_EXACT_MATH
    return 1.0 / x;
#else
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    e = fma(e, e, e);
    return fma(r, e, r);
#endif
}

// 2^(j/64), j = 0..63, filled by exp_table_init() into shared memory (64 doubles).
#define EBPMF_EXP_TABLE_SIZE 64

__device__ __forceinline__ void exp_table_init(double *tab, int tid, int nthreads) {
    for (int j = tid; j < EBPMF_EXP_TABLE_SIZE; j += nthreads)
        tab[j] = exp2((double)j * (1.0 / EBPMF_EXP_TABLE_SIZE));
}

// libdevice exp for arguments outside the fast range (overflow, underflow, NaN, Inf); kept out of
// line so the Newton loop body stays small.
__device__ __noinline__ double exp_slow(double x) { return exp(x); }

// Constants of exp_core in the constant bank: DFMA takes them as c[bank][off] operands, so they
// cost neither registers nor the UMOV/IMAD.MOV rematerialisation that immediates need.
__constant__ double kExpC[6] = {
    92.332482616893656768,        // 64/ln2
    -0x1.62e42fee00000p-7,        // -(ln2/64 rounded to 32 bits): n*L_HI is exact
    -0x1.a39ef35793c76p-39,       // -(ln2/64 - L_HI)
    1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0};

// true when |x| >= 700 or x is NaN / Inf: decided on the high word with integer instructions
// so the test costs no FP64-pipe slot (0x4085E000_00000000 == 700.0).
__device__ __forceinline__ bool exp_out_of_range(double x) {
    return (__double2hiint(x) & 0x7fffffff) >= 0x4085E000;
}

// exp(x) for |x| < 700: n = rint(x*64/ln2), r = x - n*ln2/64, 2^(j/64) from the table,
// exp(r) - 1 = r + r^2/2 + ... + r^5/120 (the dropped r^6/720 is <= 3.5e-17 at |r| <= ln2/128).
// Measured on the host with the same fma sequence: max rel. error 2.3e-16 on [-700,700].
__device__ __forceinline__ double exp_core(double x, const double *tab) {
    const double MAGIC = 6755399441055744.0;           // 1.5 * 2^52
    const double t = fma(x, kExpC[0], MAGIC);
    const int n = __double2loint(t);
    const double nf = t - MAGIC;
    double r = fma(nf, kExpC[1], x);
    r = fma(nf, kExpC[2], r);
    double q = fma(r, kExpC[3], kExpC[4]);
    q = fma(q, r, kExpC[5]);
    q = fma(q, r, 0.5);
    q = q * r;
    q = fma(q, r, r);                                  // q = exp(r) - 1
    const double tj = tab[n & (EBPMF_EXP_TABLE_SIZE - 1)];
    const double v = fma(tj, q, tj);                   // 2^(j/64) * exp(r)
    // scale by 2^(n>>6): |x| < 700 keeps the result a normal double
    const int hi = __double2hiint(v) + ((n >> 6) << 20);
    return __hiloint2double(hi, __double2loint(v));
}

__device__ __forceinline__ double exp_fast(double x, const double *tab) {
#if EBPMF_EXACT_MATH
    (void)tab;
    return exp(x);
#else
    if (exp_out_of_range(x)) return exp_slow(x);
    return exp_core(x, tab);
#endif
}

// E exponentials with ONE range branch for the lot (the Newton loop's form)
template <int E>
__device__ __forceinline__ void exp_fast_n(const double (&x)[E], double (&out)[E], const double *tab) {
#if EBPMF_EXACT_MATH
#pragma unroll
    for (int e = 0; e < E; ++e) out[e] = exp(x[e]);
#else
    bool oor = false;
#pragma unroll
    for (int e = 0; e < E; ++e) oor = oor || exp_out_of_range(x[e]);
    if (oor) {
#pragma unroll
        for (int e = 0; e < E; ++e) out[e] = exp_slow(x[e]);
    } else {
#pragma unroll
        for (int e = 0; e < E; ++e) out[e] = exp_core(x[e], tab);
    }
#endif
}

__device__ __forceinline__ double log_fast(double x) {
    return log(x);
}

}  // namespace ebpmf
