// Micro-probe: do MUFU.RCP64H (XU pipe) and DFMA (FP64 pipe) overlap, and what does a DFMA-heavy
// stream lose when integer / LDS instructions are interleaved?  8 warps per SM sub-partition.
#include <cstdio>
#include <cuda_runtime.h>
template <int NF, int NM, int NI>
__global__ void k(double *out, const double *in, long long *cyc, int iters) {
    __shared__ double tab[64];
    const int t = threadIdx.x;
    if (t < 64) tab[t] = in[t];
    __syncthreads();
    double a[4]; int x[4];
    for (int i = 0; i < 4; ++i) { a[i] = in[t + i] + 1.5; x[i] = t + i; }
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
#pragma unroll
            for (int q = 0; q < NF; ++q) a[i] = fma(a[i], 1.0000001, 1e-9);
#pragma unroll
            for (int q = 0; q < NM; ++q) { double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a[i])); a[i] = r; }
#pragma unroll
            for (int q = 0; q < NI; ++q) { x[i] = (x[i] * 3 + 7) ^ (x[i] >> 3); }
        }
    }
    long long t1 = clock64();
    double s = 0; for (int i = 0; i < 4; ++i) s += a[i] + x[i];
    out[blockIdx.x * blockDim.x + t] = s;
    if (t == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int NF, int NM, int NI> void run(const char *what) {
    double *out, *in; long long *cyc, h;
    cudaMalloc(&out, 8 * 148 * 1024); cudaMalloc(&in, 8 * 2048); cudaMalloc(&cyc, 8);
    cudaMemset(in, 0, 8 * 2048);
    const int iters = 4000;
    k<NF, NM, NI><<<148, 1024>>>(out, in, cyc, iters); cudaDeviceSynchronize();
    k<NF, NM, NI><<<148, 1024>>>(out, in, cyc, iters);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    // 8 warps per SMSP x 4 chains
    printf("%-50s %.1f SMSP-cycles per (chain step) ; per warp-step\n", what, (double)h / iters / (4 * 8));
    cudaFree(out); cudaFree(in); cudaFree(cyc);
}
int main() {
    run<12, 0, 0>("12 DFMA");
    run<0, 1, 0>("1 MUFU.RCP64H");
    run<12, 1, 0>("12 DFMA + 1 MUFU.RCP64H   (sum 32, max 24)");
    run<12, 2, 0>("12 DFMA + 2 MUFU.RCP64H   (sum 40, max 24)");
    run<12, 0, 12>("12 DFMA + 12x3 int ops");
    run<12, 1, 12>("12 DFMA + 1 MUFU + 12x3 int ops");
    return 0;
}
