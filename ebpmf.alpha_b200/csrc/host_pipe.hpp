// HostPipe: moves a column-major fp64 matrix between the device and CALLER-OWNED host memory that may
// be pageable (what R's Rf_allocMatrix / REAL(x) hands the .Call shim), at close to pinned-copy speed
// and asynchronously to the compute stream.
//
// Measured on the GPU box (profiles/r02_probe_host_copy.txt): cudaMemcpy D2H into pageable memory runs
// at 21 GB/s (H2D from pageable 11 GB/s) against 57 GB/s for pinned memory; pinning the caller's buffer
// (cudaHostRegister) costs 0.09 s per GiB -- more than the copy itself.  So: a small ring of pinned
// slots, the copy engine fills (drains) slot after slot on a dedicated stream, and a few host threads
// memcpy each slot into (out of) the caller's matrix while the next slots are in flight (53 GB/s with
// 8 threads).  A caller buffer that is already pinned goes straight through, no ring.
//
// Transfers are JOBS (whole column ranges) queued to one dispatcher thread per direction; a D2H job may
// name an event on another stream to wait for (the kernel that produces the columns), an H2D job an
// event to record when its last piece has been issued (what the consuming kernel waits for).
#pragma once
#include <cuda_runtime.h>
#include <sys/mman.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace ebpmf {

class HostPipe {
   public:
    struct Job {
        bool to_host = true;
        const double *src = nullptr;     // D2H: device, packed (leading dimension == rows)   H2D: host, leading dimension ld
        double *dst = nullptr;           // D2H: host, leading dimension ld                    H2D: device, packed
        size_t rows = 0, cols = 0, ld = 0;
        cudaEvent_t after = nullptr;     // D2H: wait for this event (recorded before the job was queued)
        cudaEvent_t done = nullptr;      // H2D: recorded on the copy stream once every piece has been issued
        std::atomic<long> pieces_left{0};
        std::atomic<bool> all_pieces_queued{false};
        bool finished = false;           // D2H: data is in dst          H2D: `done` has been recorded
    };
    using JobRef = std::shared_ptr<Job>;

    HostPipe() = default;
    HostPipe(const HostPipe &) = delete;
    ~HostPipe() { shutdown(); }

    bool ready() const { return started_; }

    // slot_bytes x nslots pinned per direction; nworkers memcpy threads shared by both directions
    int init(int device, size_t slot_bytes, int nslots, int nworkers, std::string &err) {
        if (started_) return 0;
        device_ = device;
        slot_bytes_ = slot_bytes;
        cudaError_t e = cudaSetDevice(device);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&out_stream_, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&in_stream_, cudaStreamNonBlocking);
        for (int d = 0; d < 2 && e == cudaSuccess; ++d) {
            e = cudaMallocHost((void **)&ring_[d], slot_bytes * (size_t)nslots);
            slots_[d].resize((size_t)nslots);
            for (int s = 0; s < nslots && e == cudaSuccess; ++s) {
                slots_[d][(size_t)s].ptr = ring_[d] + (size_t)s * slot_bytes;
                e = cudaEventCreateWithFlags(&slots_[d][(size_t)s].ev, cudaEventDisableTiming);
            }
        }
        if (e != cudaSuccess) {
            err = std::string("HostPipe: ") + cudaGetErrorString(e);
            (void)cudaGetLastError();
            release();
            return 1;
        }
        stop_ = false;
        for (int w = 0; w < nworkers; ++w) workers_.emplace_back([this] { worker_loop(); });
        disp_[0] = std::thread([this] { dispatch_loop(0); });
        disp_[1] = std::thread([this] { dispatch_loop(1); });
        started_ = true;
        return 0;
    }

    void shutdown() {
        if (!started_) { release(); return; }
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
        }
        cv_jobs_.notify_all(); cv_work_.notify_all(); cv_slots_.notify_all(); cv_done_.notify_all();
        for (auto &t : workers_) t.join();
        workers_.clear();
        for (auto &t : disp_) if (t.joinable()) t.join();
        started_ = false;
        release();
    }

    // columns of a packed device matrix -> host (leading dimension ld >= rows); returns at once
    JobRef d2h(const double *src_dev, double *dst, size_t rows, size_t cols, size_t ld, cudaEvent_t after) {
        auto j = std::make_shared<Job>();
        j->to_host = true; j->src = src_dev; j->dst = dst; j->rows = rows; j->cols = cols; j->ld = ld; j->after = after;
        enqueue(j, 0);
        return j;
    }
    // host (leading dimension ld) -> packed device matrix; `done` is recorded when the last piece is issued
    JobRef h2d(double *dst_dev, const double *src, size_t rows, size_t cols, size_t ld, cudaEvent_t done) {
        auto j = std::make_shared<Job>();
        j->to_host = false; j->src = src; j->dst = dst_dev; j->rows = rows; j->cols = cols; j->ld = ld; j->done = done;
        enqueue(j, 1);
        return j;
    }
    // D2H: the data is in the caller's memory.  H2D: `done` has been recorded (a stream may now wait on it).
    int wait(const JobRef &j, std::string &err) {
        std::unique_lock<std::mutex> lk(mu_);
        cv_done_.wait(lk, [&] { return j->finished || failed_; });
        if (failed_) { err = err_; return 1; }
        return 0;
    }
    bool failed() const { return failed_; }
    cudaStream_t out_stream() const { return out_stream_; }

   private:
    struct Slot {
        char *ptr = nullptr;
        cudaEvent_t ev = nullptr;
        int state = 0;                    // 0 free, 1 in use, 2 (H2D) copy issued: free once ev has completed
    };
    struct Piece {
        JobRef job;
        int dir = 0, slot = -1;           // slot < 0: direct (pinned caller memory), nothing for a worker to do
        size_t c0 = 0, nc = 0, r0 = 0, nr = 0;
    };

    void release() {
        for (int d = 0; d < 2; ++d) {
            for (auto &s : slots_[d]) if (s.ev) cudaEventDestroy(s.ev);
            slots_[d].clear();
            if (ring_[d]) cudaFreeHost(ring_[d]);
            ring_[d] = nullptr;
        }
        if (out_stream_) cudaStreamDestroy(out_stream_);
        if (in_stream_) cudaStreamDestroy(in_stream_);
        out_stream_ = in_stream_ = nullptr;
    }

    void fail(cudaError_t e, const char *what) {
        std::lock_guard<std::mutex> lk(mu_);
        if (!failed_) { failed_ = true; err_ = std::string("HostPipe ") + what + ": " + cudaGetErrorString(e); }
        cv_done_.notify_all(); cv_slots_.notify_all();
    }

    void enqueue(const JobRef &j, int dir) {
        {
            std::lock_guard<std::mutex> lk(mu_);
            jobs_[dir].push_back(j);
        }
        cv_jobs_.notify_all();
    }

    static bool is_pinned(const void *p) {
        cudaPointerAttributes at;
        const cudaError_t e = cudaPointerGetAttributes(&at, p);
        if (e != cudaSuccess) { (void)cudaGetLastError(); return false; }
        return at.type == cudaMemoryTypeHost;
    }

    // take the next slot of the ring (FIFO order) once it is free
    int acquire(int dir, size_t &cursor) {
        const int s = (int)(cursor % slots_[dir].size());
        ++cursor;
        std::unique_lock<std::mutex> lk(mu_);
        cv_slots_.wait(lk, [&] { return slots_[dir][(size_t)s].state != 1 || stop_ || failed_; });
        if (stop_ || failed_) return -1;
        if (slots_[dir][(size_t)s].state == 2) {           // H2D copy out of this slot was issued: wait for the engine
            lk.unlock();
            cudaEventSynchronize(slots_[dir][(size_t)s].ev);
            lk.lock();
        }
        slots_[dir][(size_t)s].state = 1;
        return s;
    }

    void finish_piece(const Piece &pc) {
        Job &j = *pc.job;
        if (j.pieces_left.fetch_sub(1) == 1 && j.all_pieces_queued.load()) complete(pc.job);
    }
    void complete(const JobRef &j) {
        if (!j->to_host && j->done) {
            const cudaError_t e = cudaEventRecord(j->done, in_stream_);
            if (e != cudaSuccess) fail(e, "cudaEventRecord");
        }
        {
            std::lock_guard<std::mutex> lk(mu_);
            j->finished = true;
        }
        cv_done_.notify_all();
    }

    void dispatch_loop(int dir) {
        cudaSetDevice(device_);
        size_t cursor = 0;
        for (;;) {
            JobRef j;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_jobs_.wait(lk, [&] { return stop_ || !jobs_[dir].empty(); });
                if (stop_) return;
                j = jobs_[dir].front();
                jobs_[dir].pop_front();
            }
            if (failed_) { complete(j); continue; }
            const bool to_host = j->to_host;
            char *host = to_host ? (char *)j->dst : (char *)j->src;
            const bool direct = is_pinned(host);
            cudaStream_t st = to_host ? out_stream_ : in_stream_;
            if (to_host && j->after) {
                const cudaError_t e = cudaStreamWaitEvent(out_stream_, j->after, 0);
                if (e != cudaSuccess) fail(e, "cudaStreamWaitEvent");
            }
            if (to_host && !direct && j->rows * j->cols * 8 >= (size_t)(64u << 20)) {
                // a freshly allocated destination (R allocates the result per call) is all page faults: ask for huge pages
                const uintptr_t lo = ((uintptr_t)host + 4095) & ~(uintptr_t)4095;
                const uintptr_t hi = ((uintptr_t)host + ((j->cols - 1) * j->ld + j->rows) * 8) & ~(uintptr_t)4095;
                if (hi > lo) (void)madvise((void *)lo, hi - lo, MADV_HUGEPAGE);
            }
            // pieces: whole columns when a column fits a slot, else row ranges of one column
            const size_t col_bytes = j->rows * 8;
            const size_t cols_per = direct ? std::max<size_t>(1, (size_t)(256u << 20) / std::max<size_t>(col_bytes, 1))
                                           : std::max<size_t>(1, slot_bytes_ / std::max<size_t>(col_bytes, 1));
            const size_t rows_per = (direct || col_bytes <= slot_bytes_) ? j->rows : slot_bytes_ / 8;
            j->pieces_left.store(1);                          // the dispatcher's own reference, dropped after the loop
            for (size_t c0 = 0; c0 < j->cols && !failed_ && !stop_; c0 += cols_per) {
                const size_t nc = std::min(cols_per, j->cols - c0);
                for (size_t r0 = 0; r0 < j->rows; r0 += rows_per) {
                    const size_t nr = std::min(rows_per, j->rows - r0);
                    Piece pc;
                    pc.job = j; pc.dir = dir; pc.c0 = c0; pc.nc = nc; pc.r0 = r0; pc.nr = nr;
                    if (direct) {
                        cudaError_t e;
                        if (to_host)
                            e = cudaMemcpy2DAsync(j->dst + c0 * j->ld + r0, j->ld * 8, j->src + c0 * j->rows + r0, j->rows * 8,
                                                  nr * 8, nc, cudaMemcpyDeviceToHost, st);
                        else
                            e = cudaMemcpy2DAsync(j->dst + c0 * j->rows + r0, j->rows * 8, j->src + c0 * j->ld + r0, j->ld * 8,
                                                  nr * 8, nc, cudaMemcpyHostToDevice, st);
                        if (e != cudaSuccess) fail(e, "cudaMemcpy2DAsync");
                        continue;
                    }
                    pc.slot = acquire(dir, cursor);
                    if (pc.slot < 0) break;
                    j->pieces_left.fetch_add(1);
                    if (to_host) {
                        Slot &s = slots_[0][(size_t)pc.slot];
                        cudaError_t e = (nr == j->rows)
                            ? cudaMemcpyAsync(s.ptr, j->src + c0 * j->rows, nr * nc * 8, cudaMemcpyDeviceToHost, st)
                            : cudaMemcpy2DAsync(s.ptr, nr * 8, j->src + c0 * j->rows + r0, j->rows * 8, nr * 8, nc,
                                                cudaMemcpyDeviceToHost, st);
                        if (e == cudaSuccess) e = cudaEventRecord(s.ev, st);
                        if (e != cudaSuccess) fail(e, "D2H into the ring");
                    }
                    {
                        std::lock_guard<std::mutex> lk(mu_);
                        work_.push_back(pc);
                    }
                    cv_work_.notify_one();
                }
            }
            if (direct && !failed_) {
                // nothing went through the workers: D2H completes when the stream has drained; H2D only needs `done`
                if (to_host) {
                    const cudaError_t e = cudaStreamSynchronize(st);
                    if (e != cudaSuccess) fail(e, "cudaStreamSynchronize");
                }
            }
            j->all_pieces_queued.store(true);
            if (j->pieces_left.fetch_sub(1) == 1) complete(j);
        }
    }

    void worker_loop() {
        cudaSetDevice(device_);
        for (;;) {
            Piece pc;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_work_.wait(lk, [&] { return stop_ || !work_.empty(); });
                if (stop_ && work_.empty()) return;
                pc = work_.front();
                work_.pop_front();
            }
            Job &j = *pc.job;
            Slot &s = slots_[pc.dir][(size_t)pc.slot];
            if (j.to_host) {
                const cudaError_t e = cudaEventSynchronize(s.ev);
                if (e != cudaSuccess) fail(e, "cudaEventSynchronize");
                if (pc.nr == j.rows && j.ld == j.rows) {
                    memcpy(j.dst + pc.c0 * j.ld, s.ptr, pc.nr * pc.nc * 8);
                } else {
                    for (size_t c = 0; c < pc.nc; ++c)
                        memcpy(j.dst + (pc.c0 + c) * j.ld + pc.r0, s.ptr + c * pc.nr * 8, pc.nr * 8);
                }
                {
                    std::lock_guard<std::mutex> lk(mu_);
                    s.state = 0;
                }
                cv_slots_.notify_all();
            } else {
                if (pc.nr == j.rows && j.ld == j.rows) {
                    memcpy(s.ptr, j.src + pc.c0 * j.ld, pc.nr * pc.nc * 8);
                } else {
                    for (size_t c = 0; c < pc.nc; ++c)
                        memcpy(s.ptr + c * pc.nr * 8, j.src + (pc.c0 + c) * j.ld + pc.r0, pc.nr * 8);
                }
                cudaError_t e = (pc.nr == j.rows)
                    ? cudaMemcpyAsync(j.dst + pc.c0 * j.rows, s.ptr, pc.nr * pc.nc * 8, cudaMemcpyHostToDevice, in_stream_)
                    : cudaMemcpy2DAsync(j.dst + pc.c0 * j.rows + pc.r0, j.rows * 8, s.ptr, pc.nr * 8, pc.nr * 8, pc.nc,
                                        cudaMemcpyHostToDevice, in_stream_);
                if (e == cudaSuccess) e = cudaEventRecord(s.ev, in_stream_);
                if (e != cudaSuccess) fail(e, "H2D out of the ring");
                {
                    std::lock_guard<std::mutex> lk(mu_);
                    s.state = 2;
                }
                cv_slots_.notify_all();
            }
            finish_piece(pc);
        }
    }

    int device_ = 0;
    size_t slot_bytes_ = 0;
    bool started_ = false;
    cudaStream_t out_stream_ = nullptr, in_stream_ = nullptr;
    char *ring_[2] = {nullptr, nullptr};
    std::vector<Slot> slots_[2];
    std::mutex mu_;
    std::condition_variable cv_jobs_, cv_work_, cv_slots_, cv_done_;
    std::deque<JobRef> jobs_[2];
    std::deque<Piece> work_;
    std::vector<std::thread> workers_;
    std::thread disp_[2];
    bool stop_ = true;
    std::atomic<bool> failed_{false};
    std::string err_;
};

// HostPool: warm host blocks for RESULT matrices.  R allocates a fresh n x p matrix for every result
// (Rf_allocMatrix); at 30 GB that is 7 million first-touch page faults per call and halves the copy-out
// rate (profiles/r02_probe_host_copy.txt: 17-29 GB/s into fresh pages against 53 GB/s into touched ones).
// R lets a package supply the memory of a vector (allocVector3 with an R_allocator_t); the .Call glue hands
// it blocks from this pool and the pool takes them back when R's garbage collector frees the matrix, so from
// the second outer iteration on every result lands in pages that are already mapped and page-locked (blocks of
// 64 MB and more are cudaHostRegister'ed once, ~0.09 s per GiB; EBPMF_B200_POOL_PIN=0 keeps them pageable): the
// copy engine then writes them directly, no staging memcpy -- which matters most when several GPUs share the
// host's memory bandwidth (2 GPUs: 33.9 GB/s per GPU page-locked against 22.4 GB/s through the ring).
class HostPool {
   public:
    static HostPool &instance() { static HostPool p; return p; }

    void *alloc(size_t bytes) {
        const size_t kHuge = (size_t)2 << 20;
        bytes = (std::max<size_t>(bytes, 1) + kHuge - 1) / kHuge * kHuge;      // whole huge pages
        {
            std::lock_guard<std::mutex> lk(mu_);
            size_t best = free_.size();
            for (size_t q = 0; q < free_.size(); ++q)
                if (free_[q].bytes >= bytes && free_[q].bytes - bytes <= free_[q].bytes / 4 &&
                    (best == free_.size() || free_[q].bytes < free_[best].bytes)) best = q;
            if (best != free_.size()) {
                Block b = free_[best];
                free_.erase(free_.begin() + (long)best);
                kept_bytes_ -= b.bytes;
                live_.push_back(b);
                ++reused_;
                return b.ptr;
            }
        }
        const size_t len = bytes;
        void *ptr = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        if (ptr == MAP_FAILED) return nullptr;
        (void)madvise(ptr, len, MADV_HUGEPAGE);
        Block b;
        b.ptr = ptr; b.bytes = len; b.pinned = false;
        const char *pin = getenv("EBPMF_B200_POOL_PIN");         // default on; "0" keeps the blocks pageable
        if (!(pin && atoi(pin) == 0) && len >= ((size_t)64 << 20)) {
            populate(ptr, len);
            b.pinned = (cudaHostRegister(ptr, len, cudaHostRegisterPortable) == cudaSuccess);
            if (!b.pinned) (void)cudaGetLastError();
        }
        std::lock_guard<std::mutex> lk(mu_);
        live_.push_back(b);
        ++fresh_;
        return ptr;
    }

    // back to the pool (kept warm up to the keep limit); returns false when ptr is not one of ours
    bool release(void *ptr) {
        Block b;
        {
            std::lock_guard<std::mutex> lk(mu_);
            size_t q = 0;
            for (; q < live_.size(); ++q) if (live_[q].ptr == ptr) break;
            if (q == live_.size()) return false;
            b = live_[q];
            live_.erase(live_.begin() + (long)q);
            if (kept_bytes_ + b.bytes <= keep_limit()) {
                free_.push_back(b);
                kept_bytes_ += b.bytes;
                return true;
            }
        }
        unmap(b);
        return true;
    }

    void trim() {
        std::vector<Block> drop;
        {
            std::lock_guard<std::mutex> lk(mu_);
            drop.swap(free_);
            kept_bytes_ = 0;
        }
        for (auto &b : drop) unmap(b);
    }

    void stats(double *out4) {
        std::lock_guard<std::mutex> lk(mu_);
        out4[0] = (double)fresh_; out4[1] = (double)reused_; out4[2] = (double)kept_bytes_; out4[3] = (double)live_.size();
    }

   private:
    struct Block { void *ptr = nullptr; size_t bytes = 0; bool pinned = false; };
    static size_t keep_limit() {
        const char *v = getenv("EBPMF_B200_POOL_KEEP_GB");
        const double gb = v ? atof(v) : 96.0;
        return (size_t)(std::max(0.0, gb) * 1073741824.0);
    }
    static void populate(void *ptr, size_t len) {      // touch every page from a few threads (first-touch zeroing)
        const int nt = 8;
        std::vector<std::thread> th;
        const size_t per = (len / nt + 4095) & ~(size_t)4095;
        for (int t = 0; t < nt; ++t)
            th.emplace_back([=] {
                const size_t lo = (size_t)t * per, hi = std::min(len, lo + per);
                for (size_t o = lo; o < hi; o += 4096) ((volatile char *)ptr)[o] = 0;
            });
        for (auto &t : th) t.join();
    }
    static void unmap(const Block &b) {
        if (b.pinned) (void)cudaHostUnregister(b.ptr);
        (void)munmap(b.ptr, b.bytes);
    }
    std::mutex mu_;
    std::vector<Block> live_, free_;
    size_t kept_bytes_ = 0;
    unsigned long long fresh_ = 0, reused_ = 0;
};

}  // namespace ebpmf
