// fp64 device math for the VGA Newton kernels (sm_100a).
//
// The hot path is bound by the FP64 pipe, so the number of DP-pipe
// instructions per Newton iteration is what matters.  R's `1/x`, `exp`, `log`
// (R/vga_solver_mat.R:22-31, R/ebpmf_log.R:317-318) are replaced by
//   rcp : MUFU.RCP64H seed (rcp.approx.ftz.f64, rel. err <= 2^-20) + one cubic Newton step
//         r(1+e+e^2): 3 DFMA, result within ~1.5 ulp for normal, finite x;
//   exp : n = rint(x*256/ln2); r = x - n*ln2/256 (Cody-Waite, 2 DFMA); 2^(j/256) from a
//         256-entry table in shared memory; degree-4 polynomial: 9 DP instructions, rel. err <= 2.4e-16;
//   log : 128-entry (1/m_i, -log(1/m_i)) table + degree-6 log1p polynomial, ~10 DP instructions.
// Parity budget is 1e-8 relative against the oracle (BASELINE.json); these are ~1e-15.
// Every operation is an explicit __fma_rn/__dmul_rn/__dadd_rn so that nvcc's mul+add contraction
// cannot differ between the two kernels that continue one iteration sequence: a unit gives
// bit-identical iterates whichever pass executes them (run-to-run reproducible results).
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
    double e = __fma_rn(-x, r, 1.0);
    e = __fma_rn(e, e, e);
    return __fma_rn(r, e, r);
#endif
}

// 2^(j/256), j = 0..255, copied by exp_table_init() into shared memory (2 KB).
#define EBPMF_EXP_TABLE_SIZE 256

// shared-space byte address of a __shared__ object
__device__ __forceinline__ unsigned smem_addr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

// The table values come from ONE place (the host, ebpmf_api.cu: math_tables()) and are copied into
// shared memory by every kernel, so all kernels iterate with bit-identical tables.
struct MathTables {
    double exp2_tab[EBPMF_EXP_TABLE_SIZE];   // 2^(j/256)
    double2 log_tab[128];         // (c_i, -log(c_i)), c_i = float(1/(1+(i+0.5)/128))
};

__device__ __forceinline__ void exp_table_init(double *tab, const MathTables *g, int tid, int nthreads) {
    for (int j = tid; j < EBPMF_EXP_TABLE_SIZE; j += nthreads) tab[j] = g->exp2_tab[j];
}

// libdevice exp for arguments outside the fast range (overflow, underflow, NaN, Inf); kept out of
// line so the Newton loop body stays small.
__device__ __noinline__ double exp_slow(double x) { return exp(x); }

// Constants of exp_core in the constant bank: DFMA takes them as c[bank][off] operands, so they
// cost neither registers nor the UMOV/IMAD.MOV rematerialisation that immediates need.
__constant__ double kExpC[5] = {
    369.3299304675746,            // 256/ln2
    -0x1.62e42fe800000p-9,        // -(ln2/256 rounded to 30 bits): n*L_HI is exact for |n| < 2^18
    -0x1.e8e7bcd5e4f1ep-39,       // -(ln2/256 - L_HI)
    1.0 / 24.0, 1.0 / 6.0};

// true when |x| >= 700 or x is NaN / Inf: decided on the high word with integer instructions
// so the test costs no FP64-pipe slot (0x4085E000_00000000 == 700.0).
__device__ __forceinline__ bool exp_out_of_range(double x) {
    return (__double2hiint(x) & 0x7fffffff) >= 0x4085E000;
}

// exp(x) for |x| < 700: n = rint(x*256/ln2), r = x - n*ln2/256, 2^(j/256) from the table,
// exp(r) - 1 = r + r^2/2 + r^3/6 + r^4/24 (the dropped r^5/120 is <= 3.8e-17 at |r| <= ln2/512).
// 9 FP64 instructions.  Measured on the host with the same fma sequence: max rel. error 2.4e-16
// on [-700,700].
// tab_s: shared-space byte address of the 256-entry table (from smem_addr()); a plain LDS, no
// generic->shared window arithmetic in the loop.
__device__ __forceinline__ double exp_core(double x, unsigned tab_s) {
    const double MAGIC = 6755399441055744.0;           // 1.5 * 2^52
    const double t = __fma_rn(x, kExpC[0], MAGIC);
    const int n = __double2loint(t);
    const double nf = __dadd_rn(t, -MAGIC);
    double r = __fma_rn(nf, kExpC[1], x);
    r = __fma_rn(nf, kExpC[2], r);
    double q = __fma_rn(r, kExpC[3], kExpC[4]);
    q = __fma_rn(q, r, 0.5);
    q = __dmul_rn(q, r);
    q = __fma_rn(q, r, r);                             // q = exp(r) - 1
    double tj;
    asm("ld.shared.f64 %0, [%1];" : "=d"(tj) : "r"(tab_s + ((unsigned)(n & (EBPMF_EXP_TABLE_SIZE - 1)) << 3)));
    const double v = __fma_rn(tj, q, tj);              // 2^(j/256) * exp(r)
    // scale by 2^(n>>8): |x| < 700 keeps the result a normal double
    const int hi = __double2hiint(v) + ((n & ~(EBPMF_EXP_TABLE_SIZE - 1)) << 12);   // + (n>>8)<<20
    return __hiloint2double(hi, __double2loint(v));
}

__device__ __forceinline__ double exp_fast(double x, unsigned tab_s) {
#if EBPMF_EXACT_MATH
    (void)tab_s;
    return exp(x);
#else
    if (exp_out_of_range(x)) return exp_slow(x);
    return exp_core(x, tab_s);
#endif
}

// E exponentials with ONE range branch for the lot (the Newton loop's form)
template <int E>
__device__ __forceinline__ void exp_fast_n(const double (&x)[E], double (&out)[E], unsigned tab_s) {
#if EBPMF_EXACT_MATH
#pragma unroll
    for (int e = 0; e < E; ++e) out[e] = exp(x[e]);
#else
    bool oor = false;
#pragma unroll
    for (int e = 0; e < E; ++e) oor = oor || exp_out_of_range(x[e]);
    if (oor) {          // rare; each entry still gets the value it would get on its own
#pragma unroll
        for (int e = 0; e < E; ++e) out[e] = exp_out_of_range(x[e]) ? exp_slow(x[e]) : exp_core(x[e], tab_s);
    } else {
#pragma unroll
        for (int e = 0; e < E; ++e) out[e] = exp_core(x[e], tab_s);
    }
#endif
}

// ---------------------------------------------------------------------------------------
// log(x): x = 2^e * m, m in [1,2); 128-entry table of c_i ~ 1/m_i (float-rounded, so
// z = m*c_i - 1 is exact in one fma, |z| <= 0.004) and l_i = -log(c_i);
// log(x) = e*ln2 + l_i + log1p(z), log1p by a degree-6 polynomial (next term z^7/7 <= 2.3e-18).
// ~10 FP64-pipe instructions; measured on the host with the same fma sequence: error
// <= 2.6e-16 * max(|log x|, 1).  Zero / negative / denormal / Inf / NaN go to libdevice.
// ---------------------------------------------------------------------------------------
#define EBPMF_LOG_TABLE_SIZE 128

__device__ __forceinline__ void log_table_init(double2 *tab, const MathTables *g, int tid, int nthreads) {
    for (int i = tid; i < EBPMF_LOG_TABLE_SIZE; i += nthreads) tab[i] = g->log_tab[i];
}

__device__ __noinline__ double log_slow(double x) { return log(x); }

__constant__ double kLogC[5] = {-1.0 / 6.0, 0.2, -0.25, 1.0 / 3.0, 0.6931471805599453};

__device__ __forceinline__ double log_fast(double x, unsigned tab_s) {
#if EBPMF_EXACT_MATH
    (void)tab_s;
    return log(x);
#else
    const int hi = __double2hiint(x);
    if ((unsigned)(hi - 0x00100000) >= 0x7fe00000u) return log_slow(x);
    const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(x));
    double2 cl;
    asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(cl.x), "=d"(cl.y) : "r"(tab_s + (((unsigned)hi >> 9) & ((EBPMF_LOG_TABLE_SIZE - 1) << 4))));
    const double z = fma(m, cl.x, -1.0);
    double p = fma(z, kLogC[0], kLogC[1]);
    p = fma(p, z, kLogC[2]);
    p = fma(p, z, kLogC[3]);
    p = fma(p, z, -0.5);
    p = p * z;
    const double res = fma(p, z, z);
    // (double)(e) without a conversion instruction: 2^52 + 2^31 magic
    const double ed = __hiloint2double(0x43300000, ((hi >> 20) - 1023) ^ 0x80000000) - 4503601774854144.0;
    return fma(ed, kLogC[4], cl.y) + res;
#endif
}

}  // namespace ebpmf
