// Micro-probe: dependent-issue latency and throughput of DFMA / MUFU.RCP64H on B200 as a function of
// independent chains per warp (ILP) and warps per SM sub-partition.  Build: nvcc -arch=sm_100a -O3.
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double *out, long long *cyc, int iters, double seed) {
    double a[ILP];
    for (int i = 0; i < ILP; ++i) a[i] = seed + i;
    const double b = 1.0000001, c = 1e-9;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) a[i] = fma(a[i], b, c);
    }
    long long t1 = clock64();
    double s = 0; for (int i = 0; i < ILP; ++i) s += a[i];
    if (s == 123.456) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int ILP>
__global__ void krcp(double *out, long long *cyc, int iters, double seed) {
    double a[ILP];
    for (int i = 0; i < ILP; ++i) a[i] = seed + i;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) { double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a[i])); a[i] = r; }
    }
    long long t1 = clock64();
    double s = 0; for (int i = 0; i < ILP; ++i) s += a[i];
    if (s == 123.456) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int ILP> void run(int threads, const char *what, bool rcp) {
    double *out; long long *cyc, h;
    cudaMalloc(&out, 8); cudaMalloc(&cyc, 8);
    const int iters = 20000;
    if (rcp) krcp<ILP><<<148, threads>>>(out, cyc, iters, 1.5); else k<ILP><<<148, threads>>>(out, cyc, iters, 1.5);
    cudaDeviceSynchronize();
    if (rcp) krcp<ILP><<<148, threads>>>(out, cyc, iters, 1.5); else k<ILP><<<148, threads>>>(out, cyc, iters, 1.5);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double per_iter = (double)h / iters;             // cycles per loop iteration of one warp
    int warps_per_smsp = threads / 32 / 4; if (warps_per_smsp < 1) warps_per_smsp = 1;
    printf("%s ILP=%d warps/SMSP=%d : %.2f cyc per dependent op, SMSP issue interval %.2f cyc/warp-instr\n", what, ILP,
           threads >= 128 ? threads / 128 : 1, per_iter, per_iter / (ILP * (threads >= 128 ? threads / 128 : 1)));
    cudaFree(out); cudaFree(cyc);
}
int main() {
    for (int th : {32, 128, 256, 512, 1024}) { run<1>(th, "DFMA", false); run<2>(th, "DFMA", false); run<4>(th, "DFMA", false); run<8>(th, "DFMA", false); }
    for (int th : {32, 128, 512}) { run<1>(th, "RCP64H", true); run<2>(th, "RCP64H", true); run<4>(th, "RCP64H", true); }
    return 0;
}
