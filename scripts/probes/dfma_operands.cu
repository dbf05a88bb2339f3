// Micro-probe: does a DFMA with three distinct REGISTER operands issue at the same rate as one
// whose multiplier/addend come from the constant bank?  (register-file bank conflicts)
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(double *out, const double *in, long long *cyc, int iters) {
    const int t = threadIdx.x;
    double a[8], b[8], c[8];
    for (int i = 0; i < 8; ++i) { a[i] = in[t + i]; b[i] = in[t + 8 + i]; c[i] = in[t + 16 + i]; }
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) a[i] = fma(a[i], 1.0000001, 1e-9);          // 1 register source
            if (MODE == 1) a[i] = fma(a[i], b[i], 1e-9);               // 2 register sources
            if (MODE == 2) a[i] = fma(a[i], b[i], c[i]);               // 3 register sources
            if (MODE == 3) a[i] = fma(b[i], c[i], a[i]);               // 3 register sources, accumulate
            if (MODE == 4) a[i] = a[i] + b[i];                         // DADD 2 regs
            if (MODE == 5) a[i] = a[i] * b[i];                         // DMUL 2 regs
            if (MODE == 6) a[i] = fma(a[i], a[i], a[i]);               // same register thrice
            if (MODE == 7) a[i] = fma(a[i], b[i], b[i]);               // 2 distinct
        }
    }
    long long t1 = clock64();
    double s = 0; for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + t] = s;
    if (t == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int MODE> void run(const char *what) {
    double *out, *in; long long *cyc, h;
    cudaMalloc(&out, 8 * 148 * 256); cudaMalloc(&in, 8 * 512); cudaMalloc(&cyc, 8);
    cudaMemset(in, 0, 8 * 512);
    const int iters = 20000;
    k<MODE><<<148, 256>>>(out, in, cyc, iters); cudaDeviceSynchronize();
    k<MODE><<<148, 256>>>(out, in, cyc, iters);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    // 2 warps per SMSP x 8 chains: pipe saturated; cycles per warp-instruction per SMSP:
    printf("%-44s %.2f cycles / warp-instr / SMSP\n", what, (double)h / iters / (8 * 2));
    cudaFree(out); cudaFree(in); cudaFree(cyc);
}
int main() {
    run<0>("DFMA a = fma(a, const, const)");
    run<1>("DFMA a = fma(a, b, const)");
    run<2>("DFMA a = fma(a, b, c)");
    run<3>("DFMA a = fma(b, c, a)");
    run<6>("DFMA a = fma(a, a, a)");
    run<7>("DFMA a = fma(a, b, b)");
    run<4>("DADD a = a + b");
    run<5>("DMUL a = a * b");
    return 0;
}
