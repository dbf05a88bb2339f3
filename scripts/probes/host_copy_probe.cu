// Host-side copy probe behind the design of the pinned ring in ebpmf_api.cu (HostPipe):
// what a D2H / H2D of a large matrix costs into pinned vs pageable host memory, how fast T host
// threads move a pinned slot into a pageable destination, and what the two overlapped achieve.
// Output: profiles/r02_probe_host_copy.txt
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include <atomic>
#include <unistd.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

static void par_memcpy(char *dst, const char *src, size_t bytes, int T) {
    std::vector<std::thread> th;
    const size_t per = (bytes / T + 4095) & ~(size_t)4095;
    for (int t = 0; t < T; ++t) {
        const size_t lo = std::min(bytes, per * t), hi = std::min(bytes, per * (t + 1));
        th.emplace_back([=] { if (hi > lo) memcpy(dst + lo, src + lo, hi - lo); });
    }
    for (auto &t : th) t.join();
}

int main(int argc, char **argv) {
    const size_t GB = 1ull << 30;
    const size_t bytes = (argc > 1 ? atof(argv[1]) : 4.0) * GB;
    printf("host cores online: %ld, page size %ld, buffer %.1f GiB\n", sysconf(_SC_NPROCESSORS_ONLN), sysconf(_SC_PAGESIZE), bytes / (double)GB);
    char *d = nullptr, *pin = nullptr;
    CK(cudaMalloc(&d, bytes));
    CK(cudaMemset(d, 1, bytes));
    double t = now();
    CK(cudaMallocHost(&pin, bytes));
    printf("cudaMallocHost %.1f GiB: %.3f s\n", bytes / (double)GB, now() - t);
    char *pg = (char *)malloc(bytes);
    t = now();
    memset(pg, 0, bytes);
    printf("first touch (memset, 1 thread) of pageable: %.2f GB/s\n", bytes / 1e9 / (now() - t));
    cudaStream_t s; CK(cudaStreamCreate(&s));
    for (int rep = 0; rep < 2; ++rep) {
        t = now(); CK(cudaMemcpyAsync(pin, d, bytes, cudaMemcpyDeviceToHost, s)); CK(cudaStreamSynchronize(s));
        printf("D2H -> pinned:   %.2f GB/s\n", bytes / 1e9 / (now() - t));
        t = now(); CK(cudaMemcpyAsync(pg, d, bytes, cudaMemcpyDeviceToHost, s)); CK(cudaStreamSynchronize(s));
        printf("D2H -> pageable: %.2f GB/s\n", bytes / 1e9 / (now() - t));
        t = now(); CK(cudaMemcpyAsync(d, pin, bytes, cudaMemcpyHostToDevice, s)); CK(cudaStreamSynchronize(s));
        printf("H2D <- pinned:   %.2f GB/s\n", bytes / 1e9 / (now() - t));
        t = now(); CK(cudaMemcpyAsync(d, pg, bytes, cudaMemcpyHostToDevice, s)); CK(cudaStreamSynchronize(s));
        printf("H2D <- pageable: %.2f GB/s\n", bytes / 1e9 / (now() - t));
    }
    for (int T : {1, 2, 4, 8, 12, 16, 24, 32}) {
        t = now(); par_memcpy(pg, pin, bytes, T);
        printf("memcpy pinned -> pageable (warm), %2d threads: %.2f GB/s\n", T, bytes / 1e9 / (now() - t));
    }
    for (int T : {1, 4, 8, 16}) {
        char *fresh = (char *)malloc(bytes);
        t = now(); par_memcpy(fresh, pin, bytes, T);
        printf("memcpy pinned -> FRESH pageable (page faults), %2d threads: %.2f GB/s\n", T, bytes / 1e9 / (now() - t));
        free(fresh);
    }
    t = now();
    cudaError_t e = cudaHostRegister(pg, bytes, cudaHostRegisterDefault);
    printf("cudaHostRegister %.1f GiB: %.3f s (%s)\n", bytes / (double)GB, now() - t, cudaGetErrorString(e));
    if (e == cudaSuccess) { t = now(); cudaHostUnregister(pg); printf("cudaHostUnregister: %.3f s\n", now() - t); }
    // the pipeline: D2H in slot-sized pieces into a pinned ring, T threads drain the ring into pageable memory
    for (size_t slot_mb : {8, 32, 128}) {
        for (int T : {2, 4, 8, 16}) {
            const size_t slot = slot_mb << 20;
            const int NS = 16;
            const size_t npieces = bytes / slot;
            std::vector<cudaEvent_t> ev(NS);
            for (auto &x : ev) CK(cudaEventCreateWithFlags(&x, cudaEventDisableTiming));
            std::vector<std::atomic<long>> done(NS);        // piece index drained from slot s (+1), 0 = free
            for (auto &x : done) x = 0;
            std::atomic<long> issued{0}, next{0};
            std::vector<long> slot_piece(NS, -1);
            t = now();
            std::vector<std::thread> th;
            std::atomic<long> drained{0};
            for (int w = 0; w < T; ++w)
                th.emplace_back([&] {
                    cudaSetDevice(0);
                    for (;;) {
                        const long k = next.fetch_add(1);
                        if (k >= (long)npieces) break;
                        while (issued.load(std::memory_order_acquire) <= k) std::this_thread::yield();
                        const int sl = (int)(k % NS);
                        cudaEventSynchronize(ev[sl]);
                        memcpy(pg + (size_t)k * slot, pin + (size_t)sl * slot, slot);
                        done[sl].store(k + 1, std::memory_order_release);
                        drained.fetch_add(1);
                    }
                });
            for (long k = 0; k < (long)npieces; ++k) {
                const int sl = (int)(k % NS);
                if (k >= NS) while (done[sl].load(std::memory_order_acquire) != k - NS + 1) std::this_thread::yield();
                cudaMemcpyAsync(pin + (size_t)sl * slot, d + (size_t)k * slot, slot, cudaMemcpyDeviceToHost, s);
                cudaEventRecord(ev[sl], s);
                issued.store(k + 1, std::memory_order_release);
            }
            for (auto &x : th) x.join();
            printf("pipelined D2H -> ring(%d x %zu MB) -> pageable, %2d threads: %.2f GB/s\n", NS, slot_mb, T,
                   npieces * slot / 1e9 / (now() - t));
            for (auto &x : ev) cudaEventDestroy(x);
        }
    }
    return 0;
}
