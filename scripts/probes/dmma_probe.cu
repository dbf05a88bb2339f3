// Micro-probe: FP64 tensor-core MMA (mma.sync.m8n8k4.f64 = DMMA) on B200 -- throughput alone, and what it
// costs next to DFMA and integer work (does it share the FP64 pipe?  does it hold the issue port?).
// 8 warps per SM sub-partition, ND independent accumulator tiles per warp.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// DISTINCT A / B fragments per DMMA (what a real GEMM tile loop has): MODE 0 = 8 distinct A and B registers,
// 1 = one A shared by groups of 3 DMMAs (the M %*% EF loop: af reused over nt), B distinct
template <int ND, int MODE>
__global__ void kd(double *out, const double *in, long long *cyc, int iters) {
    const int t = threadIdx.x;
    double c[12][2], a[8], b[8];
    for (int i = 0; i < 8; ++i) { a[i] = in[(t + i) & 63] + 1.0000001 + i; b[i] = in[(t + 7 + i) & 63] + 0.9999999 - i; }
    for (int i = 0; i < 12; ++i) { c[i][0] = in[i] + i; c[i][1] = in[i + 8] - i; }
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int q = 0; q < ND; ++q) dmma(c[q % 12][0], c[q % 12][1], MODE == 0 ? a[q & 7] : a[(q / 3) & 7], b[q & 7]);
    }
    long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < 12; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + t] = s;
    if (t == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int ND, int MODE> void rund(const char *what, int threads) {
    double *out, *in; long long *cyc, h;
    cudaMalloc(&out, 8 * 148 * 1024); cudaMalloc(&in, 8 * 2048); cudaMalloc(&cyc, 8);
    cudaMemset(in, 0, 8 * 2048);
    const int iters = 2000;
    kd<ND, MODE><<<148, threads>>>(out, in, cyc, iters); cudaDeviceSynchronize();
    kd<ND, MODE><<<148, threads>>>(out, in, cyc, iters); cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-44s %7.2f SMSP-cycles per DMMA (%d warps per SMSP)\n", what, (double)h / iters / ND / (threads / 128.0), threads / 128);
    cudaFree(out); cudaFree(in); cudaFree(cyc);
}

template <int ND, int NF, int NI>
__global__ void k(double *out, const double *in, long long *cyc, int iters) {
    const int t = threadIdx.x;
    double c[8][2], f[4]; int x[4];
    const double a = in[t & 63] + 1.0000001, b = in[(t + 7) & 63] + 0.9999999;
    for (int i = 0; i < 8; ++i) { c[i][0] = in[i] + i; c[i][1] = in[i + 8] - i; }
    for (int i = 0; i < 4; ++i) { f[i] = in[t & 31] + 1.5 + i; x[i] = t + i; }
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int q = 0; q < ND; ++q) dmma(c[q & 7][0], c[q & 7][1], a, b);
#pragma unroll
        for (int q = 0; q < NF; ++q) f[q & 3] = fma(f[q & 3], 1.0000001, 1e-9);
#pragma unroll
        for (int q = 0; q < NI; ++q) x[q & 3] = (x[q & 3] * 3 + 7) ^ (x[q & 3] >> 3);
    }
    long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    for (int i = 0; i < 4; ++i) s += f[i] + x[i];
    out[blockIdx.x * blockDim.x + t] = s;
    if (t == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int ND, int NF, int NI> void run(const char *what) {
    double *out, *in; long long *cyc, h;
    cudaMalloc(&out, 8 * 148 * 1024); cudaMalloc(&in, 8 * 2048); cudaMalloc(&cyc, 8);
    cudaMemset(in, 0, 8 * 2048);
    const int iters = 2000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<ND, NF, NI><<<148, 1024>>>(out, in, cyc, iters); cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<ND, NF, NI><<<148, 1024>>>(out, in, cyc, iters);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double per_warp_iter = (double)h / iters / 8.0;          // SMSP cycles per warp-iteration (8 warps per SMSP)
    const double flops = (512.0 * ND + 64.0 * NF) * iters * 32.0 * 148.0;   // warp-level: DMMA 8x8x4x2, DFMA 32x2; 32 warps per SM
    printf("%-44s %7.1f SMSP-cycles per warp-iteration  %6.2f TFLOP/s fp64\n", what, per_warp_iter, flops / (ms * 1e-3) / 1e12);
    cudaFree(out); cudaFree(in); cudaFree(cyc);
}

int main() {
    run<8, 0, 0>("8 DMMA");
    run<0, 16, 0>("16 DFMA");
    run<8, 16, 0>("8 DMMA + 16 DFMA");
    run<8, 0, 32>("8 DMMA + 32 int");
    run<0, 16, 32>("16 DFMA + 32 int");
    run<8, 16, 32>("8 DMMA + 16 DFMA + 32 int");
    run<1, 0, 0>("1 DMMA (latency-bound chain per tile)");
    run<2, 16, 16>("2 DMMA + 16 DFMA + 16 int");
    rund<24, 0>("24 DMMA, distinct A,B regs", 1024);
    rund<24, 0>("24 DMMA, distinct A,B regs", 256);
    rund<24, 1>("24 DMMA, A shared by 3, B distinct", 256);
    rund<24, 0>("24 DMMA, distinct A,B regs", 128);
    return 0;
}
