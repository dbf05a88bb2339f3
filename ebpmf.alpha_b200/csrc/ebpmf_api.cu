// C-ABI implementation of include/ebpmf_b200.h: context, device buffers, pass orchestration
// of the global-stop Newton solve, reductions and the NCCL combine.  No PyTorch, no CPU
// fallback: every entry point fails with EBPMF_ERR_CUDA when no sm_100 GPU is usable.
#include "../../include/ebpmf_b200.h"
#include "vga_kernels.cuh"
#include "simgen.cuh"
#include "host_pipe.hpp"
#include "m_products.cuh"

#include <cub/device/device_radix_sort.cuh>

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <unistd.h>
#include <nccl.h>      // types only; the library is dlopen'ed (nccl_dyn below)

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

using namespace ebpmf;

// -----------------------------------------------------------------------------------------
// NCCL through dlopen: single-GPU users need no NCCL at all
// -----------------------------------------------------------------------------------------
namespace {
struct NcclDyn {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool load(std::string &err) {
        if (handle) return true;
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *nm : names) {
            handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (handle) break;
        }
        if (!handle) { err = std::string("cannot dlopen libnccl.so.2: ") + dlerror(); return false; }
#define EBPMF_SYM(field, name)                                                         \
        *(void **)(&field) = dlsym(handle, name);                                      \
        if (!field) { err = std::string("NCCL symbol missing: ") + name; handle = nullptr; return false; }
        EBPMF_SYM(GetUniqueId, "ncclGetUniqueId");
        EBPMF_SYM(CommInitRank, "ncclCommInitRank");
        EBPMF_SYM(CommDestroy, "ncclCommDestroy");
        EBPMF_SYM(AllReduce, "ncclAllReduce");
        EBPMF_SYM(AllGather, "ncclAllGather");
        EBPMF_SYM(GetErrorString, "ncclGetErrorString");
#undef EBPMF_SYM
        return true;
    }
};
NcclDyn g_nccl;
thread_local std::string g_last_error;

struct DevBuf {
    void *ptr = nullptr;
    size_t cap = 0;
};
}  // namespace

struct ebpmf_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_p1 = nullptr, ev_p2 = nullptr, ev_p3 = nullptr, ev_setup = nullptr;
    std::vector<cudaEvent_t> ev_chunk, ev_in;    // host-pipelined step: pass 1 of chunk c done / M_in of chunk c on the device
    std::string err;

    // shapes
    long long n = 0, p = 0, nnz = 0;
    long long nRB = 0, nRW = 0;
    int K = 0, KTP = 0, ldk = 0;
    int var_type = -1;
    int freeze = 1;            // 1 (default): periodic fast-forward, bit-identical to iterating on; 0: plain iteration
    double delta = 0.0;        // 0: exact update count (default); > 0: delta mode (ebpmf_ctx_set_option)
    double debug_verify_scale = 1.0;   // test hook: pass 2 verifies against tol*scale (exercises the retry loop)
    int use_history = 1;       // order the first pass by the previous step's per-tile counts
    int speculate = 1;         // host-pipelined step: copy a chunk out as soon as its first pass is done
    int pipe_chunks = 0;       // host-pipelined step: number of column chunks (0: chosen from the matrix size)
    bool can_rewind = false;
    int cur_nspans = 0, cur_ntiles = 0, cur_cols = 0;   // tiling of the step in flight (fused_args)
    unsigned long long diag_topups = 0, diag_iters = 0;
    int diag_topup_tiles = 0;
    double last_kernel_ms = 0.0;                  // device time of the last call that has no ebpmf_solve_info (fixed_iter)
    unsigned long long launches = 0;              // kernels launched by this context (bench.py's gpu_launches)
    double diag_spec_bytes = 0.0, diag_recopy_bytes = 0.0;   // last host-pipelined step: speculative / repeated D2H bytes
    bool have_y = false, have_m = false, have_f = false, have_s2 = false, have_v = false, have_vsum = false;

    // order of the first pass: per-tile counts of the previous step on this shape
    bool have_hist = false;
    long long hist_n = 0, hist_p = 0;
    int hist_cols = 0;
    int hard_span = 0;         // column span of (one of) the hardest tile(s) of the previous step
    double spec_waste = 0.0;   // share of the previous pipelined step's speculative copies that had to be repeated

    // resident buffers
    DevBuf colptr, rowidx, yx, rbptr;
    DevBuf Mcur, Malt, V;
    DevBuf EL, EF, EFt, sigma2, l0, f0;
    DevBuf sc1, sc2;   // 1/sigma2, sigma2/2
    DevBuf kunit, colV, tileS, rowpart;
    DevBuf tile_hard, tile_kmin;
    DevBuf red;        // [4 scalars: sym, ssexp, slogv, sumV][v_sum ...]
    DevBuf colsums;    // colsym[p] | v_sum scratch[p]
    DevBuf status;     // Status
    DevBuf tables;     // MathTables
    DevBuf order_keys, order_keys2, order_ids, perm, cub_tmp, span_rank, chunk_dirty, small_out;
    DevBuf misc;       // small scratch (R2, fitmax, obj terms, lfactorial partials)
    DevBuf scratch[6]; // dense-signature operands
    DevBuf prod;       // ebpmf_ctx_m_products scratch

    Status *h_status = nullptr;   // pinned
    double *h_small = nullptr;    // pinned, 64 doubles
    int *h_ints = nullptr;        // pinned, 4096 ints: hard span, chunk dirty flags

    std::unique_ptr<HostPipe> pipe;   // pinned ring + copy threads, created on the first host-buffer step

    // NCCL
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    // the other ranks' Status words mapped into this device's address space (NVLink peer memory)
    DevBuf peer_ptrs;               // Status* [npeers]
    int npeers = 0;
    std::vector<void *> ipc_opened; // pointers to close at destroy
    unsigned epoch = 0;             // solve counter, identical on every rank
};

namespace {

int fail(ebpmf_ctx *ctx, int code, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf;
    g_last_error = buf;
    return code;
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            return fail(ctx, EBPMF_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), \
                        __FILE__, __LINE__);                                                       \
    } while (0)

#define NC(call)                                                                                   \
    do {                                                                                           \
        ncclResult_t r__ = (call);                                                                 \
        if (r__ != ncclSuccess)                                                                    \
            return fail(ctx, EBPMF_ERR_NCCL, "%s failed: %s", #call, g_nccl.GetErrorString(r__));  \
    } while (0)

#define CHECK(expr)                         \
    do {                                    \
        int rc__ = (expr);                  \
        if (rc__ != EBPMF_OK) return rc__;  \
    } while (0)

int reserve(ebpmf_ctx *ctx, DevBuf &b, size_t bytes) {
    if (bytes <= b.cap) return EBPMF_OK;
    if (b.ptr) { cudaFree(b.ptr); b.ptr = nullptr; b.cap = 0; }
    cudaError_t e = cudaMalloc(&b.ptr, bytes);
    if (e != cudaSuccess) {
        b.ptr = nullptr;
        (void)cudaGetLastError();
        return fail(ctx, e == cudaErrorMemoryAllocation ? EBPMF_ERR_NOMEM : EBPMF_ERR_CUDA,
                    "cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
    }
    b.cap = bytes;
    return EBPMF_OK;
}

void release(DevBuf &b) {
    if (b.ptr) cudaFree(b.ptr);
    b.ptr = nullptr;
    b.cap = 0;
}

template <typename T> T *as(DevBuf &b) { return reinterpret_cast<T *>(b.ptr); }

// register tile of factor columns: 8/16/24/32; more than 32 columns are walked in chunks of 32
int ktp_for(int K) {
    if (K <= 8) return 8;
    if (K <= 16) return 16;
    if (K <= 24) return 24;
    if (K <= 4096) return 32;
    return -1;
}

// row length of the transposed, zero-padded EF (EFt): KTP, or the next multiple of 32 when K > 32
int ldk_for(int K) { return K <= 32 ? ktp_for(K) : ((K + 31) / 32) * 32; }

#ifndef EBPMF_E
#define EBPMF_E 2
#endif
constexpr int kE = EBPMF_E;   // columns per unit in the fused / dense kernels

int cols_per_cta_for(long long p, long long nRB) {
    // enough CTAs for ~16 waves of 148 SMs x 2 resident CTAs, spans of 16..256 columns
    const long long want = 148LL * 2 * 16;
    long long spans = std::max(1LL, (want + nRB - 1) / nRB);
    long long c = (p + spans - 1) / spans;
    c = ((c + kTileCols - 1) / kTileCols) * kTileCols;
    c = std::min(256LL, std::max((long long)kTileCols, c));
    return (int)c;
}

// How one step walks the matrix: tiles = (row block, column span); the spans are grouped into column
// CHUNKS that are launched (and, in the host-buffer step, copied) one after the other.  Chunk 0 is the
// single span that held the hardest tile of the previous step, so that K* reaches the global count
// before anything else starts.
struct StepPlan {
    int cols_per_cta = 0, nspans = 0, ntiles = 0;
    int nchunks = 1;
    std::vector<int> span0, span1;        // chunk c covers the spans [span0[c], span1[c])
    std::vector<int> tile_begin, tile_count;   // its positions in the launch order
    std::vector<int> span_rank;           // span -> chunk
    bool ordered = false;                 // ctx->perm holds the order (else natural order, one chunk)
};

StepPlan make_plan(ebpmf_ctx *ctx, int want_chunks) {
    StepPlan pl;
    pl.cols_per_cta = cols_per_cta_for(ctx->p, ctx->nRB);
    pl.nspans = (int)((ctx->p + pl.cols_per_cta - 1) / pl.cols_per_cta);
    pl.ntiles = pl.nspans * (int)ctx->nRB;
    const bool hist = ctx->use_history && ctx->have_hist && ctx->hist_n == ctx->n && ctx->hist_p == ctx->p &&
                      ctx->hist_cols == pl.cols_per_cta;
    want_chunks = std::max(1, std::min(want_chunks, pl.nspans));
    if (want_chunks == 1) {
        pl.span0 = {0}; pl.span1 = {pl.nspans};
    } else {
        const int hs = hist ? std::min(std::max(ctx->hard_span, 0), pl.nspans - 1) : 0;
        const int per = (pl.nspans - 1 + (want_chunks - 1) - 1) / (want_chunks - 1);      // spans per ordinary chunk
        pl.span0.push_back(hs); pl.span1.push_back(hs + 1);
        for (int s = 0; s < hs; s += per) { pl.span0.push_back(s); pl.span1.push_back(std::min(s + per, hs)); }
        for (int s = hs + 1; s < pl.nspans; s += per) { pl.span0.push_back(s); pl.span1.push_back(std::min(s + per, pl.nspans)); }
    }
    pl.nchunks = (int)pl.span0.size();
    pl.span_rank.assign((size_t)pl.nspans, 0);
    int pos = 0;
    for (int c = 0; c < pl.nchunks; ++c) {
        for (int sp = pl.span0[(size_t)c]; sp < pl.span1[(size_t)c]; ++sp) pl.span_rank[(size_t)sp] = c;
        pl.tile_begin.push_back(pos);
        pl.tile_count.push_back((pl.span1[(size_t)c] - pl.span0[(size_t)c]) * (int)ctx->nRB);
        pos += pl.tile_count.back();
    }
    pl.ordered = hist || pl.nchunks > 1;
    return pl;
}

__global__ void order_keys_rank_kernel(const unsigned short *tile_hard, const int *span_rank, int ntiles, int nspans,
                                       int use_hist, unsigned *keys, int *ids) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntiles) return;
    const unsigned hard = use_hist ? (unsigned)tile_hard[t] : 0u;
    keys[t] = ((unsigned)span_rank[t % nspans] << 16) | (0xFFFFu - hard);
    ids[t] = t;
}

// ctx->perm = tile ids in launch order: by chunk, inside a chunk by decreasing hardness of the previous step
// (stable radix sort: equal keys keep the natural order)
int build_order(ebpmf_ctx *ctx, const StepPlan &pl) {
    if (!pl.ordered) return EBPMF_OK;
    const int nt = pl.ntiles;
    if (pl.nchunks > 0xFFFF) return fail(ctx, EBPMF_ERR_ARG, "too many column chunks");
    CHECK(reserve(ctx, ctx->order_keys, sizeof(unsigned) * (size_t)nt));
    CHECK(reserve(ctx, ctx->order_keys2, sizeof(unsigned) * (size_t)nt));
    CHECK(reserve(ctx, ctx->order_ids, sizeof(int) * (size_t)nt));
    CHECK(reserve(ctx, ctx->perm, sizeof(int) * (size_t)nt));
    CHECK(reserve(ctx, ctx->span_rank, sizeof(int) * (size_t)pl.nspans));
    CHECK(reserve(ctx, ctx->chunk_dirty, sizeof(int) * (size_t)std::max(pl.nchunks, 1)));
    CU(cudaMemcpyAsync(ctx->span_rank.ptr, pl.span_rank.data(), sizeof(int) * (size_t)pl.nspans, cudaMemcpyHostToDevice,
                       ctx->stream));
    const bool hist = ctx->use_history && ctx->have_hist && ctx->hist_n == ctx->n && ctx->hist_p == ctx->p &&
                      ctx->hist_cols == pl.cols_per_cta;
    order_keys_rank_kernel<<<(nt + 255) / 256, 256, 0, ctx->stream>>>(as<unsigned short>(ctx->tile_hard), as<int>(ctx->span_rank),
                                                                      nt, pl.nspans, hist ? 1 : 0, as<unsigned>(ctx->order_keys),
                                                                      as<int>(ctx->order_ids));
    CU(cudaGetLastError());
    size_t tmp_bytes = 0;
    CU(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, as<unsigned>(ctx->order_keys), as<unsigned>(ctx->order_keys2),
                                       as<int>(ctx->order_ids), as<int>(ctx->perm), nt, 0, 32, ctx->stream));
    CHECK(reserve(ctx, ctx->cub_tmp, tmp_bytes));
    CU(cub::DeviceRadixSort::SortPairs(ctx->cub_tmp.ptr, tmp_bytes, as<unsigned>(ctx->order_keys), as<unsigned>(ctx->order_keys2),
                                       as<int>(ctx->order_ids), as<int>(ctx->perm), nt, 0, 32, ctx->stream));
    ctx->launches += 4;
    return EBPMF_OK;
}

template <int KTP, int FORMULA, int MODE, int FREEZE>
int launch_fused_inst(ebpmf_ctx *ctx, const FusedArgs &a, int count) {
    auto kern = vga_fused_kernel<KTP, kE, FORMULA, MODE, FREEZE>;
    const int smem = (int)sizeof(FusedSmem<KTP, FORMULA>);
    static bool configured[64] = {false};   // per device
    if (ctx->device < 64 ? !configured[ctx->device] : true) {
        CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        if (ctx->device < 64) configured[ctx->device] = true;
    }
    kern<<<dim3((unsigned)count), dim3(kThreads), smem, ctx->stream>>>(a);
    CU(cudaGetLastError());
    ++ctx->launches;
    return EBPMF_OK;
}

template <int KTP, int FORMULA, int MODE>
int launch_fused_ktp(ebpmf_ctx *ctx, const FusedArgs &a, int count) {
    if (a.delta > 0) return launch_fused_inst<KTP, FORMULA, MODE, FREEZE_DELTA>(ctx, a, count);
    if (ctx->freeze) return launch_fused_inst<KTP, FORMULA, MODE, FREEZE_PERIODIC>(ctx, a, count);
    return launch_fused_inst<KTP, FORMULA, MODE, FREEZE_NONE>(ctx, a, count);
}

// one launch of the fused kernel over `count` tiles starting at position `tile_begin` of the order
template <int FORMULA, int MODE>
int launch_fused(ebpmf_ctx *ctx, FusedArgs a, int tile_begin, int count) {
    if (count <= 0) return EBPMF_OK;
    a.tile_begin = tile_begin;
    switch (ctx->KTP) {
        case 8: return launch_fused_ktp<8, FORMULA, MODE>(ctx, a, count);
        case 16: return launch_fused_ktp<16, FORMULA, MODE>(ctx, a, count);
        case 24: return launch_fused_ktp<24, FORMULA, MODE>(ctx, a, count);
        case 32: return launch_fused_ktp<32, FORMULA, MODE>(ctx, a, count);
    }
    return fail(ctx, EBPMF_ERR_ARG, "K = %d factor columns not supported", ctx->K);
}

int stride_for(int kind, long long n, long long *rs, long long *cs) {
    switch (kind) {
        case EBPMF_OP_SCALAR: *rs = 0; *cs = 0; return EBPMF_OK;
        case EBPMF_OP_ROWVEC: *rs = 1; *cs = 0; return EBPMF_OK;
        case EBPMF_OP_COLVEC: *rs = 0; *cs = 1; return EBPMF_OK;
        case EBPMF_OP_MATRIX: *rs = 1; *cs = n; return EBPMF_OK;
    }
    return EBPMF_ERR_ARG;
}

size_t operand_len(int kind, long long n, long long p) {
    switch (kind) {
        case EBPMF_OP_SCALAR: return 1;
        case EBPMF_OP_ROWVEC: return (size_t)n;
        case EBPMF_OP_COLVEC: return (size_t)p;
        default: return (size_t)n * (size_t)p;
    }
}

// Combine the Status words over ranks (MAX) -- no-op without a communicator.
int allreduce_status(ebpmf_ctx *ctx) {
    if (!ctx->comm) return EBPMF_OK;
    NC(g_nccl.AllReduce(ctx->status.ptr, ctx->status.ptr, 4, ncclInt32, ncclMax, ctx->comm, ctx->stream));
    return EBPMF_OK;
}

// Pull the Status words to the host (synchronises the compute stream).
int fetch_status(ebpmf_ctx *ctx) {
    CU(cudaMemcpyAsync(ctx->h_status, ctx->status.ptr, sizeof(Status), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return EBPMF_OK;
}

int set_target(ebpmf_ctx *ctx, int target) {
    Status s;
    memset(&s, 0, sizeof s);
    s.kstar = target;
    *ctx->h_status = s;
    CU(cudaMemcpyAsync(ctx->status.ptr, ctx->h_status, kStatusResetBytes, cudaMemcpyHostToDevice, ctx->stream));
    return EBPMF_OK;
}

int ensure_common(ebpmf_ctx *ctx) {
    CU(cudaSetDevice(ctx->device));
    return EBPMF_OK;
}

// reductions of the per-tile partials into red = [sym, ssexp, slogv, sumV][v_sum...]; also the column
// span of the hardest tile, where the NEXT step's first pass will start
int run_reductions(ebpmf_ctx *ctx) {
    const long long n = ctx->n, p = ctx->p;
    double *red = as<double>(ctx->red);
    double *colv = (ctx->var_type == EBPMF_VAR_BY_COL) ? red + 4 : as<double>(ctx->colsums);
    reduce_over_rows_kernel<1><<<(unsigned)((p + 31) / 32), dim3(32, 8), 0, ctx->stream>>>(
        as<double>(ctx->colV), ctx->nRB, p, colv);
    CU(cudaGetLastError());
    reduce_scalars_kernel<<<1, 1024, 0, ctx->stream>>>(colv, p, as<double>(ctx->tileS), ctx->cur_ntiles, red);
    CU(cudaGetLastError());
    ctx->launches += 2;
    if (ctx->var_type == EBPMF_VAR_BY_ROW) {
        reduce_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(as<double>(ctx->rowpart), n, ctx->cur_nspans,
                                                                                 red + 4);
        CU(cudaGetLastError());
        ++ctx->launches;
    }
    hardest_span_kernel<<<1, 1024, 0, ctx->stream>>>(as<unsigned short>(ctx->tile_hard), ctx->cur_ntiles, ctx->cur_nspans,
                                                     as<int>(ctx->small_out));
    CU(cudaGetLastError());
    ++ctx->launches;
    CU(cudaMemcpyAsync(ctx->h_ints, ctx->small_out.ptr, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    if (ctx->comm) {
        const size_t count = 4 + (ctx->var_type == EBPMF_VAR_BY_COL ? (size_t)p : 0);
        NC(g_nccl.AllReduce(red, red, count, ncclFloat64, ncclSum, ctx->comm, ctx->stream));
    }
    return EBPMF_OK;
}

// The pass orchestration shared by the fused step (A1), low_memory (A10) and the dense solver.
// first / topup are callables that enqueue one pass on ctx->stream.
template <typename First, typename Topup, typename After>
int run_global_stop(ebpmf_ctx *ctx, int maxiter, First first, Topup topup, After after, ebpmf_solve_info *info) {
    CU(cudaMemsetAsync(ctx->status.ptr, 0, kStatusResetBytes, ctx->stream));     // `pub` survives (epoch-tagged)
    ++ctx->epoch;
    CU(cudaEventRecord(ctx->ev0, ctx->stream));
    CHECK(first());
    CU(cudaEventRecord(ctx->ev_p1, ctx->stream));
    CHECK(allreduce_status(ctx));
    CHECK(topup());
    int passes = 2;
    for (;;) {
        CHECK(allreduce_status(ctx));
        CU(cudaEventRecord(ctx->ev_p2, ctx->stream));
        CHECK(after());                      // reductions (speculative: redone if the verify failed)
        CU(cudaEventRecord(ctx->ev1, ctx->stream));
        CHECK(fetch_status(ctx));
        const Status s = *ctx->h_status;
        if (s.nan)
            return fail(ctx, EBPMF_ERR_NAN, "missing value where TRUE/FALSE needed (NaN in f at the stop test)");
        if (s.verify_fail && s.kstar < maxiter) {
            CHECK(set_target(ctx, s.kstar + 1));
            CHECK(topup());
            ++passes;
            continue;
        }
        ctx->diag_topups = s.topup_units;
        ctx->diag_iters = s.unit_iters;
        ctx->diag_topup_tiles = s.topup_tiles;
        if (info) {
            info->n_updates = s.kstar;
            info->converged = (s.kstar < maxiter) ? 1 : 0;
            info->passes = passes;
            info->reserved = 0;
            float ms = 0.f;
            CU(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
            info->kernel_ms = ms;
            CU(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev_p1));
            info->pass1_ms = ms;
            CU(cudaEventElapsedTime(&ms, ctx->ev_p1, ctx->ev_p2));
            info->pass2_ms = ms;
            CU(cudaEventElapsedTime(&ms, ctx->ev_p2, ctx->ev1));
            info->reduce_ms = ms;
        }
        return EBPMF_OK;
    }
}

FusedArgs fused_args(ebpmf_ctx *ctx, const StepPlan &pl, int maxiter, double tol, bool want_v) {
    FusedArgs a;
    memset(&a, 0, sizeof a);
    a.n = ctx->n; a.p = ctx->p; a.K = ctx->K; a.var_type = ctx->var_type;
    a.M_in = as<double>(ctx->Mcur); a.M_out = as<double>(ctx->Malt);
    a.V_out = want_v ? as<double>(ctx->V) : nullptr;
    a.EL = as<double>(ctx->EL); a.EFt = as<double>(ctx->EFt); a.ldk = ctx->ldk;
    a.sc1 = as<double>(ctx->sc1); a.sc2 = as<double>(ctx->sc2);
    a.l0 = as<double>(ctx->l0); a.f0 = as<double>(ctx->f0);
    a.rowidx = as<int>(ctx->rowidx); a.yx = as<double>(ctx->yx); a.rbptr = as<int>(ctx->rbptr);
    a.kunit = as<unsigned short>(ctx->kunit); a.nRW = ctx->nRW; a.nJG = (ctx->p + kE - 1) / kE;
    a.colV = as<double>(ctx->colV); a.tileS = as<double>(ctx->tileS);
    a.rowpart = (ctx->var_type == EBPMF_VAR_BY_ROW) ? as<double>(ctx->rowpart) : nullptr;
    a.tile_hard = as<unsigned short>(ctx->tile_hard); a.tile_kmin = as<unsigned short>(ctx->tile_kmin);
    a.perm = pl.ordered ? as<int>(ctx->perm) : nullptr;
    a.tile_begin = 0;
    a.nspans = pl.nspans; a.cols_per_cta = pl.cols_per_cta;
    a.chunk_dirty = (pl.nchunks > 1) ? as<int>(ctx->chunk_dirty) : nullptr;
    a.span_rank = (pl.nchunks > 1) ? as<int>(ctx->span_rank) : nullptr;
    a.bulk = (ctx->n % 2 == 0 && !getenv("EBPMF_B200_NO_BULK")) ? 1 : 0;
    a.status = as<Status>(ctx->status);
    a.peers = reinterpret_cast<const Status *const *>(ctx->peer_ptrs.ptr);
    a.npeers = ctx->npeers;
    a.epoch = ctx->epoch + 1;      // run_global_stop increments the counter before the first launch
    a.tables = as<MathTables>(ctx->tables);
    a.maxiter = maxiter; a.tol = tol; a.delta = ctx->delta; a.tol_verify = tol * ctx->debug_verify_scale;
    ctx->cur_nspans = pl.nspans; ctx->cur_ntiles = pl.ntiles; ctx->cur_cols = pl.cols_per_cta;
    return a;
}

// one O(nnz) pass over the dgCMatrix slots before they are trusted by the kernels (the row indices
// address shared memory and the offset table is built by binary search)
int validate_csc(ebpmf_ctx *ctx, long long n, long long p, const int32_t *colptr, const int32_t *rowidx) {
    if (colptr[0] != 0) return fail(ctx, EBPMF_ERR_ARG, "bad dgCMatrix @p slot: p[0] != 0");
    for (long long j = 0; j < p; ++j) {
        const int lo = colptr[j], hi = colptr[j + 1];
        if (hi < lo) return fail(ctx, EBPMF_ERR_ARG, "bad dgCMatrix @p slot: not non-decreasing at column %lld", j);
        int prev = -1;
        for (int q = lo; q < hi; ++q) {
            const int r = rowidx[q];
            if (r <= prev || r >= n)
                return fail(ctx, EBPMF_ERR_ARG, "bad dgCMatrix @i slot: column %lld holds row index %d (rows must be increasing, < %lld)", j, r, n);
            prev = r;
        }
    }
    return EBPMF_OK;
}

int upload_csc(ebpmf_ctx *ctx, long long n, long long p, const int32_t *colptr, const int32_t *rowidx,
               const double *x) {
    CHECK(validate_csc(ctx, n, p, colptr, rowidx));
    const long long nnz = colptr[p];
    CHECK(reserve(ctx, ctx->colptr, sizeof(int) * (size_t)(p + 1)));
    CHECK(reserve(ctx, ctx->rowidx, sizeof(int) * (size_t)std::max(1LL, nnz)));
    CHECK(reserve(ctx, ctx->yx, sizeof(double) * (size_t)std::max(1LL, nnz)));
    CU(cudaMemcpyAsync(ctx->colptr.ptr, colptr, sizeof(int) * (size_t)(p + 1), cudaMemcpyHostToDevice, ctx->stream));
    if (nnz > 0) {
        CU(cudaMemcpyAsync(ctx->rowidx.ptr, rowidx, sizeof(int) * (size_t)nnz, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(ctx->yx.ptr, x, sizeof(double) * (size_t)nnz, cudaMemcpyHostToDevice, ctx->stream));
    }
    ctx->nnz = nnz;
    return EBPMF_OK;
}

// shape-dependent scratch of the fused step (grow-only), and nRB / nRW
int reserve_shape_buffers(ebpmf_ctx *ctx, long long n, long long p) {
    if (n != ctx->n || p != ctx->p) ctx->have_hist = false;
    ctx->n = n; ctx->p = p;
    ctx->nRB = (n + kRowsPerCta - 1) / kRowsPerCta;
    ctx->nRW = (n + 31) / 32;
    const long long nJG = (p + kE - 1) / kE;
    const int cols = cols_per_cta_for(p, ctx->nRB);
    const long long nspans = (p + cols - 1) / cols;
    const long long ntiles = nspans * ctx->nRB;
    if (ntiles > 0x7fffffffLL) return fail(ctx, EBPMF_ERR_ARG, "matrix too large for one shard");
    CHECK(reserve(ctx, ctx->kunit, sizeof(unsigned short) * (size_t)nJG * (size_t)ctx->nRW));
    CHECK(reserve(ctx, ctx->status, sizeof(Status)));
    CHECK(reserve(ctx, ctx->colV, sizeof(double) * (size_t)ctx->nRB * (size_t)p));
    CHECK(reserve(ctx, ctx->tileS, sizeof(double) * 3 * (size_t)ntiles));
    CHECK(reserve(ctx, ctx->tile_hard, sizeof(unsigned short) * (size_t)ntiles));
    CHECK(reserve(ctx, ctx->tile_kmin, sizeof(unsigned short) * (size_t)ntiles));
    CHECK(reserve(ctx, ctx->small_out, sizeof(int) * 16));
    CHECK(reserve(ctx, ctx->colsums, sizeof(double) * 2 * (size_t)p));
    CHECK(reserve(ctx, ctx->sc1, sizeof(double) * (size_t)std::max(n, p)));
    CHECK(reserve(ctx, ctx->sc2, sizeof(double) * (size_t)std::max(n, p)));
    CHECK(reserve(ctx, ctx->red, sizeof(double) * (4 + (size_t)std::max(n, p))));
    CHECK(reserve(ctx, ctx->misc, sizeof(double) * (size_t)(3 * std::max(n, p) + 4096)));
    return EBPMF_OK;
}

// after colptr/rowidx/yx are resident: shapes, row-block offset table, partial buffers
int finish_set_y(ebpmf_ctx *ctx, long long n, long long p) {
    CHECK(reserve_shape_buffers(ctx, n, p));
    const long long cells = (ctx->nRB + 1) * p;
    CHECK(reserve(ctx, ctx->rbptr, sizeof(int) * (size_t)cells));
    build_rbptr_kernel<<<(unsigned)((cells + 255) / 256), 256, 0, ctx->stream>>>(
        as<int>(ctx->colptr), as<int>(ctx->rowidx), p, ctx->nRB, as<int>(ctx->rbptr));
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->have_y = true;
    ctx->have_hist = false;
    ctx->have_m = ctx->have_f = ctx->have_s2 = ctx->have_v = ctx->have_vsum = false;
    return EBPMF_OK;
}

int dense_to_csc(ebpmf_ctx *ctx, long long n, long long p, const double *dY /* device */) {
    // counts -> colptr -> fill; nnz must fit the int32 @p slot of a dgCMatrix
    CHECK(reserve(ctx, ctx->colptr, sizeof(int) * (size_t)(p + 1)));
    CHECK(reserve(ctx, ctx->misc, sizeof(double) * (size_t)(3 * std::max(n, p) + 4096)));
    int *counts = as<int>(ctx->misc);
    const unsigned blocks = (unsigned)((p * 32 + 255) / 256);
    count_nnz_kernel<<<blocks, 256, 0, ctx->stream>>>(dY, n, p, counts);
    CU(cudaGetLastError());
    scan_counts_kernel<<<1, 1024, 0, ctx->stream>>>(counts, p, as<int>(ctx->colptr));
    CU(cudaGetLastError());
    int nnz = 0;
    CU(cudaMemcpyAsync(&nnz, as<int>(ctx->colptr) + p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (nnz < 0) return fail(ctx, EBPMF_ERR_ARG, "more than 2^31-1 non-zeros in one shard");
    CHECK(reserve(ctx, ctx->rowidx, sizeof(int) * (size_t)std::max(1, nnz)));
    CHECK(reserve(ctx, ctx->yx, sizeof(double) * (size_t)std::max(1, nnz)));
    fill_csc_kernel<<<blocks, 256, 0, ctx->stream>>>(dY, n, p, as<int>(ctx->colptr), as<int>(ctx->rowidx),
                                                     as<double>(ctx->yx));
    CU(cudaGetLastError());
    ctx->nnz = nnz;
    return EBPMF_OK;
}

int sum_lfactorial_dev(ebpmf_ctx *ctx, const double *dx, long long len, double *out) {
    const int blocks = (int)std::min<long long>(2048, std::max<long long>(1, (len + 255) / 256));
    CHECK(reserve(ctx, ctx->misc, sizeof(double) * 4096));
    double *partials = as<double>(ctx->misc);
    lfactorial_partial_kernel<<<blocks, 256, 0, ctx->stream>>>(dx, len, partials);
    CU(cudaGetLastError());
    sum_partials_kernel<<<1, 1024, 0, ctx->stream>>>(partials, blocks, partials + 2048);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(ctx->h_small, partials + 2048, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    *out = ctx->h_small[0];
    return EBPMF_OK;
}


// host matrix (rows x cols, leading dimension ld >= rows) <-> packed device matrix (ld == rows)
int copy_in_ld(ebpmf_ctx *ctx, void *dst, const double *src, long long rows, long long cols, long long ld) {
    if (ld == rows) {
        CU(cudaMemcpyAsync(dst, src, sizeof(double) * (size_t)rows * (size_t)cols, cudaMemcpyHostToDevice, ctx->stream));
    } else {
        CU(cudaMemcpy2DAsync(dst, sizeof(double) * (size_t)rows, src, sizeof(double) * (size_t)ld,
                             sizeof(double) * (size_t)rows, (size_t)cols, cudaMemcpyHostToDevice, ctx->stream));
    }
    return EBPMF_OK;
}

int copy_out_ld(ebpmf_ctx *ctx, double *dst, const void *src, long long rows, long long cols, long long ld) {
    if (ld == rows) {
        CU(cudaMemcpyAsync(dst, src, sizeof(double) * (size_t)rows * (size_t)cols, cudaMemcpyDeviceToHost, ctx->stream));
    } else {
        CU(cudaMemcpy2DAsync(dst, sizeof(double) * (size_t)ld, src, sizeof(double) * (size_t)rows,
                             sizeof(double) * (size_t)rows, (size_t)cols, cudaMemcpyDeviceToHost, ctx->stream));
    }
    return EBPMF_OK;
}

// pinned ring + copy threads for caller-owned (possibly pageable) host matrices; created on first use
int ensure_pipe(ebpmf_ctx *ctx) {
    if (ctx->pipe && ctx->pipe->ready()) return EBPMF_OK;
    auto env_int = [](const char *name, long dflt, long lo, long hi) {
        const char *v = getenv(name);
        if (!v) return dflt;
        const long x = atol(v);
        return std::min(hi, std::max(lo, x));
    };
    long cores = sysconf(_SC_NPROCESSORS_ONLN);
    if (cores < 1) cores = 1;
    const long slot_mb = env_int("EBPMF_B200_PIPE_SLOT_MB", 32, 1, 1024);
    const long nslots = env_int("EBPMF_B200_PIPE_SLOTS", 12, 2, 256);
    const long nthreads = env_int("EBPMF_B200_PIPE_THREADS", std::min(8L, std::max(2L, cores / 2)), 1, 64);
    ctx->pipe.reset(new (std::nothrow) HostPipe());
    if (!ctx->pipe) return fail(ctx, EBPMF_ERR_NOMEM, "out of host memory");
    std::string err;
    if (ctx->pipe->init(ctx->device, (size_t)slot_mb << 20, (int)nslots, (int)nthreads, err)) {
        ctx->pipe.reset();
        return fail(ctx, EBPMF_ERR_CUDA, "%s", err.c_str());
    }
    return EBPMF_OK;
}

int pipe_wait(ebpmf_ctx *ctx, const HostPipe::JobRef &j) {
    std::string err;
    if (ctx->pipe->wait(j, err)) return fail(ctx, EBPMF_ERR_CUDA, "%s", err.c_str());
    return EBPMF_OK;
}

constexpr size_t kPipeMinBytes = (size_t)32 << 20;   // below this a plain cudaMemcpy is as good

// whole-matrix host -> device (packed) on the compute stream's timeline; returns when the data is on the device
int matrix_in(ebpmf_ctx *ctx, void *dst, const double *src, long long rows, long long cols, long long ld) {
    const size_t bytes = sizeof(double) * (size_t)rows * (size_t)cols;
    if (bytes < kPipeMinBytes) {
        CHECK(copy_in_ld(ctx, dst, src, rows, cols, ld));
        CU(cudaStreamSynchronize(ctx->stream));
        return EBPMF_OK;
    }
    CHECK(ensure_pipe(ctx));
    CU(cudaStreamSynchronize(ctx->stream));             // nothing on the compute stream still uses dst
    auto j = ctx->pipe->h2d((double *)dst, src, (size_t)rows, (size_t)cols, (size_t)ld, ctx->ev_setup);
    CHECK(pipe_wait(ctx, j));
    CU(cudaEventSynchronize(ctx->ev_setup));
    return EBPMF_OK;
}

// whole-matrix device (packed) -> host; the compute stream must have produced src already (callers synchronise)
int matrix_out(ebpmf_ctx *ctx, double *dst, const void *src, long long rows, long long cols, long long ld) {
    const size_t bytes = sizeof(double) * (size_t)rows * (size_t)cols;
    if (bytes < kPipeMinBytes) {
        CHECK(copy_out_ld(ctx, dst, src, rows, cols, ld));
        CU(cudaStreamSynchronize(ctx->stream));
        return EBPMF_OK;
    }
    CHECK(ensure_pipe(ctx));
    CU(cudaStreamSynchronize(ctx->stream));
    auto j = ctx->pipe->d2h((const double *)src, dst, (size_t)rows, (size_t)cols, (size_t)ld, nullptr);
    return pipe_wait(ctx, j);
}

// several packed n x p host matrices (NULL entries skipped) -> device, queued together so that the staging threads
// and the copy engine stay busy across matrix boundaries; returns when all of them are on the device
int matrices_in(ebpmf_ctx *ctx, void *const *dst, const double *const *src, int count, long long rows, long long cols) {
    const size_t bytes = sizeof(double) * (size_t)rows * (size_t)cols;
    if (bytes < kPipeMinBytes) {
        for (int q = 0; q < count; ++q)
            if (src[q]) CU(cudaMemcpyAsync(dst[q], src[q], bytes, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        return EBPMF_OK;
    }
    CHECK(ensure_pipe(ctx));
    CU(cudaStreamSynchronize(ctx->stream));
    std::vector<HostPipe::JobRef> jobs;
    for (int q = 0; q < count; ++q)
        if (src[q]) jobs.push_back(ctx->pipe->h2d((double *)dst[q], src[q], (size_t)rows, (size_t)cols, (size_t)rows, ctx->ev_setup));
    int rc = EBPMF_OK;
    for (auto &j : jobs) { const int r = pipe_wait(ctx, j); if (rc == EBPMF_OK) rc = r; }
    CHECK(rc);
    CU(cudaEventSynchronize(ctx->ev_setup));           // recorded after the LAST job's last piece (one in-order copy stream)
    return EBPMF_OK;
}

// up to two packed n x p device matrices -> host (the second pair may be NULL); the compute stream is synchronised first
int matrices_out(ebpmf_ctx *ctx, double *dst0, const void *src0, double *dst1, const void *src1, long long rows, long long cols) {
    const size_t bytes = sizeof(double) * (size_t)rows * (size_t)cols;
    if (bytes < kPipeMinBytes) {
        CU(cudaMemcpyAsync(dst0, src0, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        if (dst1) CU(cudaMemcpyAsync(dst1, src1, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        return EBPMF_OK;
    }
    CHECK(ensure_pipe(ctx));
    CU(cudaStreamSynchronize(ctx->stream));
    auto j0 = ctx->pipe->d2h((const double *)src0, dst0, (size_t)rows, (size_t)cols, (size_t)rows, nullptr);
    HostPipe::JobRef j1;
    if (dst1) j1 = ctx->pipe->d2h((const double *)src1, dst1, (size_t)rows, (size_t)cols, (size_t)rows, nullptr);
    int rc = pipe_wait(ctx, j0);
    if (j1) { const int r = pipe_wait(ctx, j1); if (rc == EBPMF_OK) rc = r; }
    return rc;
}

int set_m_ld(ebpmf_ctx *ctx, const double *M, long long ld) {
    if (!ctx || !M) return fail(ctx, EBPMF_ERR_ARG, "set_m: bad arguments");
    if (!ctx->have_y) return fail(ctx, EBPMF_ERR_STATE, "set_m: Y not set");
    CHECK(ensure_common(ctx));
    const size_t bytes = sizeof(double) * (size_t)ctx->n * (size_t)ctx->p;
    CHECK(reserve(ctx, ctx->Mcur, bytes));
    CHECK(reserve(ctx, ctx->Malt, bytes));
    CHECK(matrix_in(ctx, ctx->Mcur.ptr, M, ctx->n, ctx->p, ld));
    ctx->have_m = true; ctx->can_rewind = false;
    return EBPMF_OK;
}

int set_factors_ld(ebpmf_ctx *ctx, int64_t K, const double *EL, long long ldl, const double *EF, bool sync = true) {
    if (!ctx || !EL || !EF || K < 1) return fail(ctx, EBPMF_ERR_ARG, "set_factors: bad arguments");
    if (!ctx->have_y) return fail(ctx, EBPMF_ERR_STATE, "set_factors: Y not set");
    const int ktp = ktp_for((int)K);
    if (ktp < 0) return fail(ctx, EBPMF_ERR_ARG, "K = %lld factor columns not supported (max 4096)", (long long)K);
    const int ldk = ldk_for((int)K);
    CHECK(ensure_common(ctx));
    const long long n = ctx->n, p = ctx->p;
    CHECK(reserve(ctx, ctx->EL, sizeof(double) * (size_t)n * (size_t)K));
    CHECK(reserve(ctx, ctx->EF, sizeof(double) * (size_t)p * (size_t)K));
    CHECK(reserve(ctx, ctx->EFt, sizeof(double) * (size_t)p * (size_t)ldk));
    CHECK(copy_in_ld(ctx, ctx->EL.ptr, EL, n, K, ldl));
    CU(cudaMemcpyAsync(ctx->EF.ptr, EF, sizeof(double) * (size_t)p * (size_t)K, cudaMemcpyHostToDevice, ctx->stream));
    transpose_ef_kernel<<<(unsigned)((p * ldk + 255) / 256), 256, 0, ctx->stream>>>(as<double>(ctx->EF), p, (int)K, ldk,
                                                                                   as<double>(ctx->EFt));
    CU(cudaGetLastError());
    ++ctx->launches;
    if (sync) CU(cudaStreamSynchronize(ctx->stream));
    ctx->K = (int)K; ctx->KTP = ktp; ctx->ldk = ldk; ctx->have_f = true;
    return EBPMF_OK;
}

int get_matrix_ld(ebpmf_ctx *ctx, bool want_v, double *out, long long ld) {
    if (!ctx || !out) return fail(ctx, EBPMF_ERR_ARG, "get: bad arguments");
    if (want_v ? !ctx->have_v : !ctx->have_m) return fail(ctx, EBPMF_ERR_STATE, want_v ? "get_v: the last step did not materialise V (want_v = 0)" : "get_m: M not set");
    CHECK(ensure_common(ctx));
    return matrix_out(ctx, out, want_v ? ctx->V.ptr : ctx->Mcur.ptr, ctx->n, ctx->p, ld);
}


// Map the other ranks' Status words into this device's address space so the pass-1 kernels can read
// their published counts over NVLink.  Same process (ebpmf_group): raw pointers + peer access;
// other processes (torchrun): CUDA IPC handles, exchanged with one ncclAllGather.  Best effort: any
// failure leaves npeers = 0 and the ranks only meet in the all-reduce between the passes.
struct PeerInfo {
    long long pid;
    int device, valid;
    unsigned long long raw;
    cudaIpcMemHandle_t handle;
};

int setup_peers(ebpmf_ctx *ctx) {
    ctx->npeers = 0;
    if (!ctx->comm || ctx->world < 2 || getenv("EBPMF_B200_NO_PEER")) return EBPMF_OK;
    const int W = ctx->world;
    PeerInfo mine;
    memset(&mine, 0, sizeof mine);
    mine.pid = (long long)getpid(); mine.device = ctx->device; mine.raw = (unsigned long long)ctx->status.ptr;
    mine.valid = (cudaIpcGetMemHandle(&mine.handle, ctx->status.ptr) == cudaSuccess) ? 1 : 0;
    (void)cudaGetLastError();
    CHECK(reserve(ctx, ctx->scratch[0], sizeof(PeerInfo) * (size_t)(W + 1)));
    PeerInfo *d = as<PeerInfo>(ctx->scratch[0]);
    CU(cudaMemcpyAsync(d + W, &mine, sizeof mine, cudaMemcpyHostToDevice, ctx->stream));
    NC(g_nccl.AllGather(d + W, d, sizeof(PeerInfo), ncclChar, ctx->comm, ctx->stream));
    std::vector<PeerInfo> all((size_t)W);
    CU(cudaMemcpyAsync(all.data(), d, sizeof(PeerInfo) * (size_t)W, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    std::vector<void *> ptrs;
    for (int r = 0; r < W; ++r) {
        if (r == ctx->rank) continue;
        void *p = nullptr;
        if (all[(size_t)r].pid == mine.pid) {
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, ctx->device, all[(size_t)r].device) == cudaSuccess && can) {
                cudaError_t e = cudaDeviceEnablePeerAccess(all[(size_t)r].device, 0);
                if (e == cudaSuccess || e == cudaErrorPeerAccessAlreadyEnabled) p = (void *)all[(size_t)r].raw;
            }
        } else if (all[(size_t)r].valid) {
            if (cudaIpcOpenMemHandle(&p, all[(size_t)r].handle, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess)
                ctx->ipc_opened.push_back(p);
            else
                p = nullptr;
        }
        (void)cudaGetLastError();
        if (p) ptrs.push_back(p);
    }
    if (ptrs.empty()) return EBPMF_OK;
    CHECK(reserve(ctx, ctx->peer_ptrs, sizeof(void *) * ptrs.size()));
    CU(cudaMemcpyAsync(ctx->peer_ptrs.ptr, ptrs.data(), sizeof(void *) * ptrs.size(), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->npeers = (int)ptrs.size();
    return EBPMF_OK;
}

}  // namespace

// =========================================================================================
// exported C ABI
// =========================================================================================
extern "C" {

// ---- warm host blocks for result matrices (HostPool, host_pipe.hpp) ------------------------------
void *ebpmf_host_alloc(size_t bytes) { return HostPool::instance().alloc(bytes); }
int ebpmf_host_free(void *ptr) { return HostPool::instance().release(ptr) ? EBPMF_OK : EBPMF_ERR_ARG; }
void ebpmf_host_pool_trim(void) { HostPool::instance().trim(); }
void ebpmf_host_pool_stats(double *out4) { if (out4) HostPool::instance().stats(out4); }

const char *ebpmf_version(void) { return "ebpmf_b200 0.1.0 (sm_100a)"; }

int ebpmf_device_count(int *count) {
    ebpmf_ctx *ctx = nullptr;
    if (!count) return fail(ctx, EBPMF_ERR_ARG, "count is NULL");
    int c = 0;
    cudaError_t e = cudaGetDeviceCount(&c);
    if (e != cudaSuccess) { *count = 0; (void)cudaGetLastError(); return fail(ctx, EBPMF_ERR_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e)); }
    *count = c;
    return EBPMF_OK;
}

const char *ebpmf_last_error(const ebpmf_ctx *ctx) { return ctx ? ctx->err.c_str() : g_last_error.c_str(); }

int ebpmf_ctx_create(int device, ebpmf_ctx **out) {
    ebpmf_ctx *ctx = nullptr;
    if (!out) return fail(ctx, EBPMF_ERR_ARG, "out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        (void)cudaGetLastError();
        return fail(ctx, EBPMF_ERR_CUDA, "no usable CUDA device (%s); this library has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    }
    if (device < 0 || device >= count) return fail(ctx, EBPMF_ERR_ARG, "device %d out of range [0,%d)", device, count);
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(ctx, EBPMF_ERR_CUDA, "device %d is sm_%d%d; the kernels are built for sm_100a only", device,
                    prop.major, prop.minor);
    ctx = new (std::nothrow) ebpmf_ctx();
    if (!ctx) return fail(nullptr, EBPMF_ERR_NOMEM, "out of host memory");
    ctx->device = device;
    if (const char *dl = getenv("EBPMF_B200_DELTA")) {          // same range check as ebpmf_ctx_set_option
        const double d = atof(dl);
        if (!(d >= 0.0) || d > 1e-10) {
            delete ctx;
            return fail(nullptr, EBPMF_ERR_ARG, "EBPMF_B200_DELTA must be in [0, 1e-10]");
        }
        ctx->delta = d;
    }
    if (const char *fz = getenv("EBPMF_B200_FREEZE")) ctx->freeze = atoi(fz) ? 1 : 0;
    if (const char *hs = getenv("EBPMF_B200_HISTORY")) ctx->use_history = atoi(hs) ? 1 : 0;
    if (const char *sp = getenv("EBPMF_B200_SPECULATE")) ctx->speculate = atoi(sp) ? 1 : 0;
    cudaError_t ce;
#define CREATE_STEP(call)                                                                                        \
    ce = (call);                                                                                                 \
    if (ce != cudaSuccess) {                                                                                     \
        fail(nullptr, EBPMF_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(ce));                           \
        delete ctx;                                                                                              \
        return EBPMF_ERR_CUDA;                                                                                   \
    }
    CREATE_STEP(cudaSetDevice(device));
    CREATE_STEP(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    CREATE_STEP(cudaEventCreate(&ctx->ev0));
    CREATE_STEP(cudaEventCreate(&ctx->ev1));
    CREATE_STEP(cudaEventCreate(&ctx->ev_p1));
    CREATE_STEP(cudaEventCreate(&ctx->ev_p2));
    CREATE_STEP(cudaEventCreate(&ctx->ev_p3));
    CREATE_STEP(cudaEventCreateWithFlags(&ctx->ev_setup, cudaEventDisableTiming));
    CREATE_STEP(cudaMallocHost((void **)&ctx->h_status, sizeof(Status)));
    CREATE_STEP(cudaMallocHost((void **)&ctx->h_small, sizeof(double) * 64));
    CREATE_STEP(cudaMallocHost((void **)&ctx->h_ints, sizeof(int) * 4096));
    CREATE_STEP(cudaMalloc(&ctx->tables.ptr, sizeof(MathTables)));
    ctx->tables.cap = sizeof(MathTables);
    CREATE_STEP(cudaMalloc(&ctx->status.ptr, sizeof(Status)));
    ctx->status.cap = sizeof(Status);
    CREATE_STEP(cudaMemset(ctx->status.ptr, 0, sizeof(Status)));
    {
        static MathTables host_tables;          // built once on the host: every kernel uses the same bits
        static bool built = false;
        if (!built) {
            for (int j = 0; j < EBPMF_EXP_TABLE_SIZE; ++j) host_tables.exp2_tab[j] = std::exp2((double)j / EBPMF_EXP_TABLE_SIZE);
            for (int i = 0; i < 128; ++i) {
                const float c = (float)(1.0 / (1.0 + ((double)i + 0.5) / 128.0));
                host_tables.log_tab[i].x = (double)c;
                host_tables.log_tab[i].y = -std::log((double)c);
            }
            built = true;
        }
        CREATE_STEP(cudaMemcpy(ctx->tables.ptr, &host_tables, sizeof(MathTables), cudaMemcpyHostToDevice));
    }
#undef CREATE_STEP
    *out = ctx;
    return EBPMF_OK;
}

int ebpmf_ctx_destroy(ebpmf_ctx *ctx) {
    if (!ctx) return EBPMF_OK;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (void *q : ctx->ipc_opened) cudaIpcCloseMemHandle(q);
    if (ctx->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(ctx->comm);
    ctx->pipe.reset();                     // joins the copy threads, frees the pinned rings
    DevBuf *bufs[] = {&ctx->colptr, &ctx->rowidx, &ctx->yx, &ctx->rbptr, &ctx->Mcur, &ctx->Malt, &ctx->V,
                      &ctx->EL, &ctx->EF, &ctx->EFt, &ctx->sigma2, &ctx->l0, &ctx->f0, &ctx->kunit,
                      &ctx->colV, &ctx->tileS, &ctx->rowpart, &ctx->tile_hard, &ctx->tile_kmin, &ctx->red, &ctx->colsums,
                      &ctx->status, &ctx->misc, &ctx->sc1, &ctx->sc2, &ctx->tables, &ctx->peer_ptrs,
                      &ctx->order_keys, &ctx->order_keys2, &ctx->order_ids, &ctx->perm, &ctx->cub_tmp, &ctx->span_rank,
                      &ctx->chunk_dirty, &ctx->small_out, &ctx->prod};
    for (DevBuf *b : bufs) release(*b);
    for (DevBuf &b : ctx->scratch) release(b);
    if (ctx->h_status) cudaFreeHost(ctx->h_status);
    if (ctx->h_small) cudaFreeHost(ctx->h_small);
    if (ctx->h_ints) cudaFreeHost(ctx->h_ints);
    cudaEvent_t evs[] = {ctx->ev0, ctx->ev1, ctx->ev_p1, ctx->ev_p2, ctx->ev_p3, ctx->ev_setup};
    for (cudaEvent_t e : evs) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : ctx->ev_chunk) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : ctx->ev_in) if (e) cudaEventDestroy(e);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return EBPMF_OK;
}

int ebpmf_nccl_unique_id(void *id128) {
    ebpmf_ctx *ctx = nullptr;
    if (!id128) return fail(ctx, EBPMF_ERR_ARG, "id128 is NULL");
    std::string err;
    if (!g_nccl.load(err)) return fail(ctx, EBPMF_ERR_NCCL, "%s", err.c_str());
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    NC(g_nccl.GetUniqueId(&id));
    memcpy(id128, &id, sizeof id);
    return EBPMF_OK;
}

int ebpmf_ctx_comm_init(ebpmf_ctx *ctx, int rank, int world, const void *id128) {
    if (!ctx || !id128 || world < 1 || rank < 0 || rank >= world) return fail(ctx, EBPMF_ERR_ARG, "bad comm arguments");
    CHECK(ensure_common(ctx));
    std::string err;
    if (!g_nccl.load(err)) return fail(ctx, EBPMF_ERR_NCCL, "%s", err.c_str());
    ncclUniqueId id;
    memcpy(&id, id128, sizeof id);
    NC(g_nccl.CommInitRank(&ctx->comm, world, id, rank));
    ctx->rank = rank; ctx->world = world;
    return setup_peers(ctx);
}

int ebpmf_ctx_set_y_csc(ebpmf_ctx *ctx, int64_t n, int64_t p, const int32_t *colptr, const int32_t *rowidx,
                        const double *x) {
    if (!ctx || n < 1 || p < 1 || !colptr) return fail(ctx, EBPMF_ERR_ARG, "set_y_csc: bad arguments");
    if (colptr[p] > 0 && (!rowidx || !x)) return fail(ctx, EBPMF_ERR_ARG, "set_y_csc: @i / @x missing");
    CHECK(ensure_common(ctx));
    CHECK(upload_csc(ctx, n, p, colptr, rowidx, x));
    return finish_set_y(ctx, n, p);
}

int ebpmf_ctx_set_y_dense(ebpmf_ctx *ctx, int64_t n, int64_t p, const double *Y) {
    if (!ctx || n < 1 || p < 1 || !Y) return fail(ctx, EBPMF_ERR_ARG, "set_y_dense: bad arguments");
    CHECK(ensure_common(ctx));
    const size_t bytes = sizeof(double) * (size_t)n * (size_t)p;
    CHECK(reserve(ctx, ctx->Malt, bytes));
    CU(cudaMemcpyAsync(ctx->Malt.ptr, Y, bytes, cudaMemcpyHostToDevice, ctx->stream));
    CHECK(dense_to_csc(ctx, n, p, as<double>(ctx->Malt)));
    return finish_set_y(ctx, n, p);
}

int ebpmf_ctx_sum_lfactorial(ebpmf_ctx *ctx, double *out) {
    if (!ctx || !out) return fail(ctx, EBPMF_ERR_ARG, "sum_lfactorial: bad arguments");
    if (!ctx->have_y) return fail(ctx, EBPMF_ERR_STATE, "sum_lfactorial: Y not set");
    CHECK(ensure_common(ctx));
    return sum_lfactorial_dev(ctx, as<double>(ctx->yx), ctx->nnz, out);
}

int ebpmf_sum_lfactorial(ebpmf_ctx *ctx, int64_t len, const double *x, double *out) {
    if (!ctx || len < 0 || !out || (len > 0 && !x)) return fail(ctx, EBPMF_ERR_ARG, "sum_lfactorial: bad arguments");
    CHECK(ensure_common(ctx));
    if (len == 0) { *out = 0.0; return EBPMF_OK; }
    CHECK(reserve(ctx, ctx->scratch[0], sizeof(double) * (size_t)len));
    CU(cudaMemcpyAsync(ctx->scratch[0].ptr, x, sizeof(double) * (size_t)len, cudaMemcpyHostToDevice, ctx->stream));
    return sum_lfactorial_dev(ctx, as<double>(ctx->scratch[0]), len, out);
}

int ebpmf_ctx_set_m(ebpmf_ctx *ctx, const double *M) { return set_m_ld(ctx, M, ctx ? ctx->n : 0); }

int ebpmf_ctx_set_factors(ebpmf_ctx *ctx, int64_t K, const double *EL, const double *EF) {
    return set_factors_ld(ctx, K, EL, ctx ? ctx->n : 0, EF);
}

}  // extern "C"

namespace {

int set_sigma2_impl(ebpmf_ctx *ctx, int var_type, const double *sigma2, bool sync) {
    if (!ctx || !sigma2) return fail(ctx, EBPMF_ERR_ARG, "set_sigma2: bad arguments");
    if (var_type < 0 || var_type > 2) return fail(ctx, EBPMF_ERR_VAR_TYPE, "Non-supported var type");
    if (!ctx->have_y) return fail(ctx, EBPMF_ERR_STATE, "set_sigma2: Y not set");
    CHECK(ensure_common(ctx));
    const size_t len = var_type == EBPMF_VAR_BY_COL ? (size_t)ctx->p : var_type == EBPMF_VAR_BY_ROW ? (size_t)ctx->n : 1;
    CHECK(reserve(ctx, ctx->sigma2, sizeof(double) * std::max((size_t)ctx->p, (size_t)ctx->n)));
    CU(cudaMemcpyAsync(ctx->sigma2.ptr, sigma2, sizeof(double) * len, cudaMemcpyHostToDevice, ctx->stream));
    if (var_type == EBPMF_VAR_BY_ROW) {
        const int cols = cols_per_cta_for(ctx->p, ctx->nRB);
        const long long nspans = (ctx->p + cols - 1) / cols;
        CHECK(reserve(ctx, ctx->rowpart, sizeof(double) * (size_t)nspans * (size_t)ctx->n));
    }
    sigma_consts_kernel<<<(unsigned)((len + 255) / 256), 256, 0, ctx->stream>>>(as<double>(ctx->sigma2), (long long)len,
                                                                                as<double>(ctx->sc1), as<double>(ctx->sc2));
    CU(cudaGetLastError());
    ++ctx->launches;
    if (sync) CU(cudaStreamSynchronize(ctx->stream));
    ctx->var_type = var_type; ctx->have_s2 = true;
    return EBPMF_OK;
}

// Multi-GPU: every rank says whether its preparation (allocations, uploads, argument checks) succeeded
// BEFORE the first collective of the solve, so that one failing rank makes all ranks return an error
// instead of leaving the others blocked in ncclAllReduce.
int ranks_agree(ebpmf_ctx *ctx, int rc) {
    if (!ctx->comm) return rc;
    int *flag = as<int>(ctx->small_out) + 8;
    ctx->h_ints[8] = (rc == EBPMF_OK) ? 1 : 0;
    if (cudaMemcpyAsync(flag, ctx->h_ints + 8, sizeof(int), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess ||
        g_nccl.AllReduce(flag, flag, 1, ncclInt32, ncclMin, ctx->comm, ctx->stream) != ncclSuccess ||
        cudaMemcpyAsync(ctx->h_ints + 8, flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
        cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
        (void)cudaGetLastError();
        return rc != EBPMF_OK ? rc : fail(ctx, EBPMF_ERR_NCCL, "the ranks could not agree on the state of the solve");
    }
    if (rc != EBPMF_OK) return rc;
    if (ctx->h_ints[8] != 1) return fail(ctx, EBPMF_ERR_STATE, "another rank failed while preparing this solve");
    return EBPMF_OK;
}

// everything a fused A1 step needs before its first collective
int prepare_step(ebpmf_ctx *ctx, const char *who, int maxiter, bool want_v, int want_chunks, StepPlan &pl) {
    if (maxiter < 1 || maxiter > 65534) return fail(ctx, EBPMF_ERR_ARG, "%s: maxiter must be in [1, 65534]", who);
    if (!ctx->have_y || !ctx->have_m || !ctx->have_f || !ctx->have_s2)
        return fail(ctx, EBPMF_ERR_STATE, "%s: Y, M, factors and sigma2 must be set first", who);
    CHECK(ensure_common(ctx));
    if (want_v) CHECK(reserve(ctx, ctx->V, sizeof(double) * (size_t)ctx->n * (size_t)ctx->p));
    pl = make_plan(ctx, want_chunks);
    CHECK(build_order(ctx, pl));
    return EBPMF_OK;
}

// bookkeeping after a successful fused step: sums to the host, the resident M is now res$M, order history
int finish_step(ebpmf_ctx *ctx, const StepPlan &pl, bool want_v, ebpmf_solve_info *info) {
    CU(cudaMemcpyAsync(ctx->h_small, ctx->red.ptr, sizeof(double) * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (info) { info->sym = ctx->h_small[0]; info->ssexp = ctx->h_small[1]; info->slogv = ctx->h_small[2]; }
    std::swap(ctx->Mcur, ctx->Malt);      // the resident M is now res$M
    ctx->have_v = want_v; ctx->have_vsum = true; ctx->can_rewind = true;
    ctx->have_hist = true; ctx->hist_n = ctx->n; ctx->hist_p = ctx->p; ctx->hist_cols = pl.cols_per_cta;
    ctx->hard_span = ctx->h_ints[0];
    return EBPMF_OK;
}

int vga_step_impl(ebpmf_ctx *ctx, int maxiter, double tol, int want_v, ebpmf_solve_info *info) {
    if (!ctx) return fail(ctx, EBPMF_ERR_ARG, "vga_step: ctx is NULL");
    StepPlan pl;
    int rc = prepare_step(ctx, "vga_step", maxiter, want_v != 0, 1, pl);
    CHECK(ranks_agree(ctx, rc));
    const FusedArgs a = fused_args(ctx, pl, maxiter, tol, want_v != 0);
    ctx->have_v = false; ctx->have_vsum = false;
    rc = run_global_stop(
        ctx, maxiter, [&] { return launch_fused<FORMULA_A1, MODE_FIRST>(ctx, a, 0, pl.ntiles); },
        [&] { return launch_fused<FORMULA_A1, MODE_TOPUP>(ctx, a, 0, pl.ntiles); }, [&] { return run_reductions(ctx); }, info);
    if (rc != EBPMF_OK) return rc;
    return finish_step(ctx, pl, want_v != 0, info);
}

int get_v_sum_impl(ebpmf_ctx *ctx, double *v_sum) {
    if (!ctx || !v_sum) return fail(ctx, EBPMF_ERR_ARG, "get_v_sum: bad arguments");
    if (!ctx->have_vsum) return fail(ctx, EBPMF_ERR_STATE, "get_v_sum: no VGA step has run");
    CHECK(ensure_common(ctx));
    const double *red = as<double>(ctx->red);
    if (ctx->var_type == EBPMF_VAR_CONSTANT) {
        CU(cudaMemcpyAsync(v_sum, red + 3, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    } else {
        const size_t len = ctx->var_type == EBPMF_VAR_BY_COL ? (size_t)ctx->p : (size_t)ctx->n;
        CU(cudaMemcpyAsync(v_sum, red + 4, sizeof(double) * len, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CU(cudaStreamSynchronize(ctx->stream));
    return EBPMF_OK;
}

// The whole step with CALLER-OWNED host buffers (pageable is fine), pipelined by column chunks:
//   M_in (optional: NULL = the resident M is the input) goes up chunk by chunk through the pinned ring while the
//   first pass already runs on the chunks that have arrived; every chunk's result is copied out as soon as its
//   first pass is done (speculatively: with the tiles taken hardest-first, only the first chunk still changes
//   in the top-up pass) and the chunks the top-up pass touched are copied again.  ld_* are leading dimensions
//   of the host matrices (>= n; a row shard of a larger matrix has ld = its row count).
int step_host_impl(ebpmf_ctx *ctx, const double *M_in, long long ld_in, int64_t K, const double *EL, long long ldl,
                   const double *EF, int var_type, const double *sigma2, int maxiter, double tol, double *M_out,
                   long long ld_out, double *V_out, double *v_sum, ebpmf_solve_info *info) {
    if (!ctx || !M_out) return fail(ctx, EBPMF_ERR_ARG, "vga_step_host: bad arguments");
    StepPlan pl;
    const long long n = ctx->n, p = ctx->p;
    const size_t bytes = sizeof(double) * (size_t)n * (size_t)p;
    const bool piped = bytes >= kPipeMinBytes;
    int rc = EBPMF_OK;
    do {
        if (!ctx->have_y) { rc = fail(ctx, EBPMF_ERR_STATE, "vga_step_host: Y not set"); break; }
        if (!M_in && !ctx->have_m) { rc = fail(ctx, EBPMF_ERR_STATE, "vga_step_resident_host: upload M once with ebpmf_ctx_set_m"); break; }
        if ((rc = set_factors_ld(ctx, K, EL, ldl, EF, false)) != EBPMF_OK) break;
        if ((rc = set_sigma2_impl(ctx, var_type, sigma2, false)) != EBPMF_OK) break;
        if (M_in) {
            if ((rc = reserve(ctx, ctx->Mcur, bytes)) != EBPMF_OK) break;
            if ((rc = reserve(ctx, ctx->Malt, bytes)) != EBPMF_OK) break;
            if (!piped) { if ((rc = set_m_ld(ctx, M_in, ld_in)) != EBPMF_OK) break; }
            ctx->have_m = true; ctx->can_rewind = false;
        }
        if (piped && (rc = ensure_pipe(ctx)) != EBPMF_OK) break;
        int want_chunks = 1;
        if (piped) {
            want_chunks = ctx->pipe_chunks > 0 ? ctx->pipe_chunks
                                               : (int)std::min<size_t>(32, std::max<size_t>(4, bytes / ((size_t)1280 << 20)));
        }
        rc = prepare_step(ctx, "vga_step_host", maxiter, V_out != nullptr, want_chunks, pl);
        if (rc == EBPMF_OK) {
            while ((int)ctx->ev_chunk.size() < pl.nchunks) {
                cudaEvent_t e1 = nullptr, e2 = nullptr;
                if (cudaEventCreateWithFlags(&e1, cudaEventDisableTiming) != cudaSuccess ||
                    cudaEventCreateWithFlags(&e2, cudaEventDisableTiming) != cudaSuccess) {
                    rc = fail(ctx, EBPMF_ERR_CUDA, "cudaEventCreate failed");
                    break;
                }
                ctx->ev_chunk.push_back(e1); ctx->ev_in.push_back(e2);
            }
        }
    } while (0);
    CHECK(ranks_agree(ctx, rc));

    const FusedArgs a = fused_args(ctx, pl, maxiter, tol, V_out != nullptr);
    ctx->have_v = false; ctx->have_vsum = false;
    auto col0 = [&](int c) { return (long long)pl.span0[(size_t)c] * pl.cols_per_cta; };
    auto col1 = [&](int c) { return std::min(p, (long long)pl.span1[(size_t)c] * pl.cols_per_cta); };
    const bool hist = pl.ordered && ctx->use_history && ctx->have_hist;     // make_plan checked the shape
    const bool spec = piped && ctx->speculate && hist && pl.nchunks > 1 && ctx->spec_waste < 0.5;
    std::vector<HostPipe::JobRef> in_jobs, out_jobs((size_t)pl.nchunks);
    if (pl.nchunks > 1) CU(cudaMemsetAsync(ctx->chunk_dirty.ptr, 0, sizeof(int) * (size_t)pl.nchunks, ctx->stream));
    if (piped && M_in) {
        CU(cudaStreamSynchronize(ctx->stream));         // nothing queued earlier may still read the buffer M arrives in
        for (int c = 0; c < pl.nchunks; ++c)
            in_jobs.push_back(ctx->pipe->h2d(as<double>(ctx->Mcur) + (size_t)col0(c) * (size_t)n, M_in + (size_t)col0(c) * (size_t)ld_in,
                                             (size_t)n, (size_t)(col1(c) - col0(c)), (size_t)ld_in, ctx->ev_in[(size_t)c]));
    }
    double spec_bytes = 0.0, recopy_bytes = 0.0;
    ebpmf_solve_info local;
    memset(&local, 0, sizeof local);
    rc = run_global_stop(
        ctx, maxiter,
        [&] {
            for (int c = 0; c < pl.nchunks; ++c) {
                if (piped && M_in) {
                    CHECK(pipe_wait(ctx, in_jobs[(size_t)c]));                   // its last piece has been issued
                    CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_in[(size_t)c], 0));
                }
                CHECK((launch_fused<FORMULA_A1, MODE_FIRST>(ctx, a, pl.tile_begin[(size_t)c], pl.tile_count[(size_t)c])));
                if (spec) {
                    CU(cudaEventRecord(ctx->ev_chunk[(size_t)c], ctx->stream));
                    out_jobs[(size_t)c] = ctx->pipe->d2h(as<double>(ctx->Malt) + (size_t)col0(c) * (size_t)n,
                                                         M_out + (size_t)col0(c) * (size_t)ld_out, (size_t)n,
                                                         (size_t)(col1(c) - col0(c)), (size_t)ld_out, ctx->ev_chunk[(size_t)c]);
                    spec_bytes += 8.0 * (double)n * (double)(col1(c) - col0(c));
                }
            }
            return (int)EBPMF_OK;
        },
        [&] { return launch_fused<FORMULA_A1, MODE_TOPUP>(ctx, a, 0, pl.ntiles); },
        [&] {
            CHECK(run_reductions(ctx));
            if (pl.nchunks > 1)
                CU(cudaMemcpyAsync(ctx->h_ints + 16, ctx->chunk_dirty.ptr, sizeof(int) * (size_t)pl.nchunks, cudaMemcpyDeviceToHost,
                                   ctx->stream));
            return (int)EBPMF_OK;
        },
        &local);
    if (rc != EBPMF_OK) {
        if (piped) {                                    // let every queued transfer finish before the caller's buffers go away
            for (auto &j : in_jobs) if (j) (void)pipe_wait(ctx, j);
            for (auto &j : out_jobs) if (j) (void)pipe_wait(ctx, j);
        }
        return rc;
    }
    CHECK(finish_step(ctx, pl, V_out != nullptr, &local));        // synchronises the compute stream; Mcur is the result
    if (info) *info = local;
    if (!piped) {
        CHECK(copy_out_ld(ctx, M_out, ctx->Mcur.ptr, n, p, ld_out));
        if (V_out) CHECK(copy_out_ld(ctx, V_out, ctx->V.ptr, n, p, ld_out));
        CU(cudaStreamSynchronize(ctx->stream));
    } else {
        // chunks whose speculative copy is stale (the top-up pass rewrote them, or a K*+1 retry ran), or all of them
        for (int c = 0; c < pl.nchunks; ++c) {
            const bool stale = !spec || local.passes > 2 || (pl.nchunks > 1 && ctx->h_ints[16 + c] != 0);
            if (!stale) continue;
            if (out_jobs[(size_t)c]) {
                CHECK(pipe_wait(ctx, out_jobs[(size_t)c]));       // never two writers on the same host columns
                recopy_bytes += 8.0 * (double)n * (double)(col1(c) - col0(c));
            }
            out_jobs[(size_t)c] = ctx->pipe->d2h(as<double>(ctx->Mcur) + (size_t)col0(c) * (size_t)n,
                                                 M_out + (size_t)col0(c) * (size_t)ld_out, (size_t)n, (size_t)(col1(c) - col0(c)),
                                                 (size_t)ld_out, nullptr);
        }
        HostPipe::JobRef vjob;
        if (V_out) vjob = ctx->pipe->d2h(as<double>(ctx->V), V_out, (size_t)n, (size_t)p, (size_t)ld_out, nullptr);
        for (auto &j : out_jobs) if (j) CHECK(pipe_wait(ctx, j));
        if (vjob) CHECK(pipe_wait(ctx, vjob));
        ctx->spec_waste = spec_bytes > 0 ? recopy_bytes / spec_bytes : 0.0;
    }
    ctx->diag_spec_bytes = spec_bytes; ctx->diag_recopy_bytes = recopy_bytes;
    if (v_sum) CHECK(get_v_sum_impl(ctx, v_sum));
    return EBPMF_OK;
}

}  // namespace

extern "C" {

int ebpmf_ctx_set_sigma2(ebpmf_ctx *ctx, int var_type, const double *sigma2) {
    return set_sigma2_impl(ctx, var_type, sigma2, true);
}

int ebpmf_ctx_vga_step(ebpmf_ctx *ctx, int maxiter, double tol, int want_v, ebpmf_solve_info *info) {
    return vga_step_impl(ctx, maxiter, tol, want_v, info);
}

int ebpmf_ctx_rewind_m(ebpmf_ctx *ctx) {
    if (!ctx) return fail(ctx, EBPMF_ERR_ARG, "rewind_m: ctx is NULL");
    if (!ctx->can_rewind) return fail(ctx, EBPMF_ERR_STATE, "rewind_m: no step to undo");
    std::swap(ctx->Mcur, ctx->Malt);
    ctx->can_rewind = false;
    return EBPMF_OK;
}

int ebpmf_ctx_get_m(ebpmf_ctx *ctx, double *M) { return get_matrix_ld(ctx, false, M, ctx ? ctx->n : 0); }

int ebpmf_ctx_get_v(ebpmf_ctx *ctx, double *V) { return get_matrix_ld(ctx, true, V, ctx ? ctx->n : 0); }

int ebpmf_ctx_get_v_sum(ebpmf_ctx *ctx, double *v_sum) { return get_v_sum_impl(ctx, v_sum); }

int ebpmf_ctx_update_sigma2(ebpmf_ctx *ctx, const double *R2, double a0, double b0, double cap_var_mean_ratio,
                            int64_t n_total, int64_t p_total, double *sigma2_out) {
    if (!ctx || !R2 || !sigma2_out) return fail(ctx, EBPMF_ERR_ARG, "update_sigma2: bad arguments");
    if (!ctx->have_vsum || !ctx->have_s2) return fail(ctx, EBPMF_ERR_STATE, "update_sigma2: no VGA step has run");
    CHECK(ensure_common(ctx));
    const int vt = ctx->var_type;
    const long long len = vt == EBPMF_VAR_BY_COL ? ctx->p : vt == EBPMF_VAR_BY_ROW ? ctx->n : 1;
    const double half_m = vt == EBPMF_VAR_BY_COL ? (double)n_total / 2
                        : vt == EBPMF_VAR_BY_ROW ? (double)p_total / 2
                                                 : (double)n_total * (double)p_total / 2;       // :417,:428,:438
    double *dR2 = as<double>(ctx->misc);
    double *fitmax = dR2 + std::max(ctx->n, ctx->p);
    CU(cudaMemcpyAsync(dR2, R2, sizeof(double) * (size_t)len, cudaMemcpyHostToDevice, ctx->stream));
    if (cap_var_mean_ratio > 0) {
        if (!ctx->have_f) return fail(ctx, EBPMF_ERR_STATE, "update_sigma2: cap gate needs the factors");
        if (vt == EBPMF_VAR_BY_ROW) {
            fitted_rowmax_kernel<<<(unsigned)((ctx->n + 255) / 256), 256, 0, ctx->stream>>>(
                as<double>(ctx->EL), as<double>(ctx->EF), ctx->n, ctx->p, ctx->K, fitmax);
            CU(cudaGetLastError());
        } else {
            double *colmax = (vt == EBPMF_VAR_BY_COL) ? fitmax : fitmax + 1;
            fitted_colmax_kernel<<<(unsigned)ctx->p, 256, 0, ctx->stream>>>(as<double>(ctx->EL), as<double>(ctx->EF),
                                                                           ctx->n, ctx->p, ctx->K, colmax);
            CU(cudaGetLastError());
            if (vt == EBPMF_VAR_CONSTANT) {
                max_reduce_kernel<<<1, 1024, 0, ctx->stream>>>(colmax, ctx->p, fitmax);
                CU(cudaGetLastError());
            }
            if (ctx->comm) {   // rows are sharded: the column / global max spans all ranks
                NC(g_nccl.AllReduce(fitmax, fitmax, (size_t)len, ncclFloat64, ncclMax, ctx->comm, ctx->stream));
            }
        }
    }
    const double *vsum = as<double>(ctx->red) + (vt == EBPMF_VAR_CONSTANT ? 3 : 4);
    update_sigma2_kernel<<<(unsigned)((len + 255) / 256), 256, 0, ctx->stream>>>(
        vsum, dR2, len, half_m, a0, b0, cap_var_mean_ratio, fitmax, as<double>(ctx->sigma2));
    CU(cudaGetLastError());
    sigma_consts_kernel<<<(unsigned)((len + 255) / 256), 256, 0, ctx->stream>>>(as<double>(ctx->sigma2), len,
                                                                               as<double>(ctx->sc1), as<double>(ctx->sc2));
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(sigma2_out, ctx->sigma2.ptr, sizeof(double) * (size_t)len, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return EBPMF_OK;
}

int ebpmf_update_sigma2(ebpmf_ctx *ctx, int var_type, int64_t n, int64_t p, const double *sigma2, const double *v_sum,
                        const double *R2, double a0, double b0, double cap_var_mean_ratio, int64_t K,
                        const double *EL, const double *EF, double *sigma2_out) {
    if (!ctx || !sigma2 || !v_sum || !R2 || !sigma2_out || n < 1 || p < 1)
        return fail(ctx, EBPMF_ERR_ARG, "update_sigma2: bad arguments");
    if (var_type < 0 || var_type > 2) return fail(ctx, EBPMF_ERR_VAR_TYPE, "Non-supported var type");
    if (cap_var_mean_ratio > 0 && (!EL || !EF || K < 1))
        return fail(ctx, EBPMF_ERR_ARG, "update_sigma2: the cap gate needs EL and EF");
    CHECK(ensure_common(ctx));
    const long long len = var_type == EBPMF_VAR_BY_COL ? p : var_type == EBPMF_VAR_BY_ROW ? n : 1;
    const double half_m = var_type == EBPMF_VAR_BY_COL ? (double)n / 2
                        : var_type == EBPMF_VAR_BY_ROW ? (double)p / 2 : (double)n * (double)p / 2;
    const size_t big = (size_t)std::max(n, p);
    CHECK(reserve(ctx, ctx->scratch[0], sizeof(double) * (4 * big + 8)));
    double *dS = as<double>(ctx->scratch[0]), *dV = dS + big, *dR = dV + big, *fitmax = dR + big;
    CU(cudaMemcpyAsync(dS, sigma2, sizeof(double) * (size_t)len, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(dV, v_sum, sizeof(double) * (size_t)len, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(dR, R2, sizeof(double) * (size_t)len, cudaMemcpyHostToDevice, ctx->stream));
    if (cap_var_mean_ratio > 0) {
        CHECK(reserve(ctx, ctx->scratch[1], sizeof(double) * ((size_t)n + (size_t)p) * (size_t)K));
        double *dEL = as<double>(ctx->scratch[1]), *dEF = dEL + (size_t)n * K;
        CU(cudaMemcpyAsync(dEL, EL, sizeof(double) * (size_t)n * K, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(dEF, EF, sizeof(double) * (size_t)p * K, cudaMemcpyHostToDevice, ctx->stream));
        if (var_type == EBPMF_VAR_BY_ROW) {
            fitted_rowmax_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(dEL, dEF, n, p, (int)K, fitmax);
        } else {
            double *colmax = (var_type == EBPMF_VAR_BY_COL) ? fitmax : fitmax + 1;
            fitted_colmax_kernel<<<(unsigned)p, 256, 0, ctx->stream>>>(dEL, dEF, n, p, (int)K, colmax);
            if (var_type == EBPMF_VAR_CONSTANT) max_reduce_kernel<<<1, 1024, 0, ctx->stream>>>(colmax, p, fitmax);
        }
        CU(cudaGetLastError());
    }
    update_sigma2_kernel<<<(unsigned)((len + 255) / 256), 256, 0, ctx->stream>>>(dV, dR, len, half_m, a0, b0,
                                                                                 cap_var_mean_ratio, fitmax, dS);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(sigma2_out, dS, sizeof(double) * (size_t)len, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return EBPMF_OK;
}

int ebpmf_calc_obj(ebpmf_ctx *ctx, int64_t n, int64_t p, double sym, double ssexp, double slogv, int64_t len_sv,
                   const double *sv, int64_t len_sigma2, const double *sigma2, int64_t len_r2, const double *R2,
                   double KL_LF, double const_, double ss, double a0, double b0, double *obj) {
    if (!ctx || !sv || !sigma2 || !R2 || !obj || len_sv < 1 || len_sigma2 < 1 || len_r2 < 1)
        return fail(ctx, EBPMF_ERR_ARG, "calc_obj: bad arguments");
    CHECK(ensure_common(ctx));
    CHECK(reserve(ctx, ctx->scratch[0], sizeof(double) * ((size_t)len_sv + (size_t)len_sigma2 + (size_t)len_r2 + 8)));
    double *d = as<double>(ctx->scratch[0]);
    double *dS = d + len_sv, *dR = dS + len_sigma2, *dT = dR + len_r2;
    CU(cudaMemcpyAsync(d, sv, sizeof(double) * (size_t)len_sv, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(dS, sigma2, sizeof(double) * (size_t)len_sigma2, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(dR, R2, sizeof(double) * (size_t)len_r2, cudaMemcpyHostToDevice, ctx->stream));
    obj_terms_kernel<<<1, 1024, 0, ctx->stream>>>(d, len_sv, dS, len_sigma2, dR, len_r2, ss, a0, b0, dT);
    CU(cudaGetLastError());
    ++ctx->launches;
    CU(cudaMemcpyAsync(ctx->h_small, dT, sizeof(double) * 5, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    const double *t = ctx->h_small;
    // R/ebpmf_log.R:407, left to right
    *obj = sym - ssexp + slogv + 0.5 * (double)n * (double)p - t[0] - t[1] - t[2] + const_ + KL_LF - t[3] - t[4];
    return EBPMF_OK;
}

// calc_split_PMF_obj_flashier(Y,S,sigma2,M,V,fit_flash,KL_LF,const,var_type)  inst/code/Poisson_MF_splitting.R:373-391
int ebpmf_calc_split_obj(ebpmf_ctx *ctx, int64_t n, int64_t p, const double *Y, const double *S, int S_kind,
                         const double *M, const double *V, int var_type, int64_t len_sigma2, const double *sigma2,
                         int64_t len_r2, const double *R2, double KL_LF, double const_, double *obj) {
    if (!ctx || !Y || !S || !M || !V || !sigma2 || !R2 || !obj || n < 1 || p < 1 || len_sigma2 < 1 || len_r2 < 1)
        return fail(ctx, EBPMF_ERR_ARG, "calc_split_obj: bad arguments");
    if (var_type < 0 || var_type > 2) return fail(ctx, EBPMF_ERR_VAR_TYPE, "Non-supported var type");
    CHECK(ensure_common(ctx));
    SplitObjArgs a;
    memset(&a, 0, sizeof a);
    if (stride_for(S_kind, n, &a.S_rs, &a.S_cs)) return fail(ctx, EBPMF_ERR_ARG, "calc_split_obj: bad operand kind");
    const size_t NP = (size_t)n * (size_t)p, bytes = sizeof(double) * NP;
    const long long nRB = (n + kRowsPerCta - 1) / kRowsPerCta;
    const int cols = cols_per_cta_for(p, nRB);
    const int nspans = (int)((p + cols - 1) / cols);
    const size_t lS = operand_len(S_kind, n, p);
    for (int b = 0; b < 3; ++b) CHECK(reserve(ctx, ctx->scratch[b], bytes));
    const long long lsv = var_type == EBPMF_VAR_BY_COL ? p : var_type == EBPMF_VAR_BY_ROW ? n : 1;
    // small scratch: S (when not a matrix), partials, sv, sigma2, R2, terms
    const size_t small = (lS == NP ? 0 : lS) + (size_t)nRB * p + (size_t)nspans * n + (size_t)nRB * nspans + (size_t)std::max(n, p) + 8 +
                         (size_t)len_sigma2 + (size_t)len_r2 + 64;
    CHECK(reserve(ctx, ctx->scratch[4], sizeof(double) * small));
    if (lS == NP) CHECK(reserve(ctx, ctx->scratch[3], bytes));
    double *q = as<double>(ctx->scratch[4]);
    double *dS = (lS == NP) ? as<double>(ctx->scratch[3]) : q; if (lS != NP) q += lS;
    double *colV = q; q += (size_t)nRB * p;
    double *rowpart = q; q += (size_t)nspans * n;
    double *termpart = q; q += (size_t)nRB * nspans;
    double *sv = q; q += (size_t)std::max(n, p);
    double *sv_total = q; q += 8;
    double *dsig = q; q += len_sigma2;
    double *dR2 = q; q += len_r2;
    double *terms = q;
    {
        void *dsts[4] = {ctx->scratch[0].ptr, ctx->scratch[1].ptr, ctx->scratch[2].ptr, dS};
        const double *srcs[4] = {Y, M, V, lS == NP ? S : nullptr};
        CHECK(matrices_in(ctx, dsts, srcs, 4, n, p));
    }
    if (lS != NP) CU(cudaMemcpyAsync(dS, S, sizeof(double) * lS, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(dsig, sigma2, sizeof(double) * (size_t)len_sigma2, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(dR2, R2, sizeof(double) * (size_t)len_r2, cudaMemcpyHostToDevice, ctx->stream));
    a.n = n; a.p = p; a.Y = as<double>(ctx->scratch[0]); a.M = as<double>(ctx->scratch[1]); a.V = as<double>(ctx->scratch[2]);
    a.S = dS; a.colV = colV; a.rowpart = rowpart; a.termpart = termpart; a.cols_per_cta = cols;
    CU(cudaEventRecord(ctx->ev0, ctx->stream));
    split_obj_kernel<<<dim3((unsigned)nspans, (unsigned)nRB), kThreads, 0, ctx->stream>>>(a);
    CU(cudaGetLastError());
    CU(cudaEventRecord(ctx->ev1, ctx->stream));
    // sv (:377-388): column sums, row sums, or the total of V
    double ss = 0.0;
    if (var_type == EBPMF_VAR_BY_ROW) {
        reduce_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(rowpart, n, nspans, sv);
        ss = (double)p;
    } else {
        reduce_over_rows_kernel<1><<<(unsigned)((p + 31) / 32), dim3(32, 8), 0, ctx->stream>>>(colV, nRB, p, sv);
        ss = (double)n;
        if (var_type == EBPMF_VAR_CONSTANT) {
            sum_partials_kernel<<<1, 1024, 0, ctx->stream>>>(sv, (int)p, sv_total);
            ss = (double)n * (double)p;
        }
    }
    CU(cudaGetLastError());
    // sum of the entry terms -> terms[8]; the sigma2 terms of :389 through obj_terms_kernel (a0 = -1, b0 = 0: its two prior terms vanish)
    sum_partials_kernel<<<1, 1024, 0, ctx->stream>>>(termpart, (int)(nRB * nspans), terms + 8);
    const double *svp = (var_type == EBPMF_VAR_CONSTANT) ? sv_total : sv;
    obj_terms_kernel<<<1, 1024, 0, ctx->stream>>>(svp, lsv, dsig, len_sigma2, dR2, len_r2, ss, -1.0, 0.0, terms);
    CU(cudaGetLastError());
    ctx->launches += 4;
    CU(cudaMemcpyAsync(ctx->h_small, terms, sizeof(double) * 9, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ctx->last_kernel_ms = ms;
    const double *tq = ctx->h_small;
    *obj = tq[8] - tq[0] - tq[1] - tq[2] + const_ + KL_LF;      // :389, left to right
    return EBPMF_OK;
}

}  // extern "C"
namespace {
// the host matrices of a *_step_host call must have the shape of the resident Y (R: "non-conformable arrays")
int check_shape(ebpmf_ctx *ctx, const char *who, int64_t n, int64_t p) {
    if (!ctx->have_y) return fail(ctx, EBPMF_ERR_STATE, "%s: Y not set", who);
    if (n != ctx->n || p != ctx->p)
        return fail(ctx, EBPMF_ERR_ARG, "%s: non-conformable arrays: M is %lld x %lld but the resident Y is %lld x %lld", who,
                    (long long)n, (long long)p, ctx->n, ctx->p);
    return EBPMF_OK;
}
}  // namespace
extern "C" {

int ebpmf_vga_step_host(ebpmf_ctx *ctx, int64_t n, int64_t p, const double *M_in, int64_t K, const double *EL,
                        const double *EF, int var_type, const double *sigma2, int maxiter, double tol, double *M_out,
                        double *V_out, double *v_sum, ebpmf_solve_info *info) {
    if (!ctx || !M_in || !M_out) return fail(ctx, EBPMF_ERR_ARG, "vga_step_host: bad arguments");
    CHECK(check_shape(ctx, "vga_step_host", n, p));
    return step_host_impl(ctx, M_in, n, K, EL, n, EF, var_type, sigma2, maxiter, tol, M_out, n, V_out, v_sum, info);
}

// The per-iteration call of ebpmf_log's outer loop once M lives on the device: R feeds res$M back
// unchanged (fit_flash$flash_fit$Y = res$M, R/ebpmf_log.R:309; flash_init(fit_flash$flash_fit$Y, ..),
// R/ebpmf_log_flash_update.R:35), so the step's input M IS the resident M left by the previous step
// and only the factors and sigma2 cross PCIe on the way in.
int ebpmf_vga_step_resident_host(ebpmf_ctx *ctx, int64_t n, int64_t p, int64_t K, const double *EL, const double *EF,
                                 int var_type, const double *sigma2, int maxiter, double tol, double *M_out,
                                 double *V_out, double *v_sum, ebpmf_solve_info *info) {
    if (!ctx || !M_out) return fail(ctx, EBPMF_ERR_ARG, "vga_step_resident_host: bad arguments");
    CHECK(check_shape(ctx, "vga_step_resident_host", n, p));
    return step_host_impl(ctx, nullptr, 0, K, EL, n, EF, var_type, sigma2, maxiter, tol, M_out, n, V_out, v_sum, info);
}

// ---- N1: products of the resident M with the factors, and the residual sums of squares -----------
}  // extern "C"
namespace {
// warps (32-row groups) per CTA of the products kernel: one CTA per SM for KP >= 16 (159+ registers), so 12 warps
// hide more latency than 8; KP = 8 fits two CTAs of 8 warps (measured on the bench slab: K = 22: 14.6 vs 16.2 ms
// kernel time with 12 vs 8 warps, K = 8: 6.8 vs 7.3 ms with 8 vs 12)
#ifndef EBPMF_PROD_NW
#define EBPMF_PROD_NW 12
#endif
inline int prod_rows_for(int KP) { return KP <= 8 ? 256 : 32 * EBPMF_PROD_NW; }

template <int KP, int NW>
int launch_products(ebpmf_ctx *ctx, const ProdArgs &a) {
    const long long nRBp = (a.n + 32 * NW - 1) / (32 * NW);
    const int smem = (int)sizeof(ProdSmem<KP, NW>);
    static bool configured[64] = {false};
    if (ctx->device < 64 ? !configured[ctx->device] : true) {
        CU(cudaFuncSetAttribute(m_products_kernel<KP, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        if (ctx->device < 64) configured[ctx->device] = true;
    }
    m_products_kernel<KP, NW><<<dim3((unsigned)a.nsplit, (unsigned)nRBp), 32 * NW, smem, ctx->stream>>>(a);
    CU(cudaGetLastError());
    ++ctx->launches;
    return EBPMF_OK;
}
}  // namespace
extern "C" {

int ebpmf_ctx_m_products(ebpmf_ctx *ctx, int64_t Kw, const double *EL_w, const double *EF_w, double *MF, double *MtL,
                         double *r2_col, double *r2_row, double *kernel_ms) {
    if (!ctx || !EL_w || !EF_w || !MF || !MtL || Kw < 1 || Kw > 256)
        return fail(ctx, EBPMF_ERR_ARG, "m_products: bad arguments (1 <= Kw <= 256)");
    if (!ctx->have_y || !ctx->have_m) return fail(ctx, EBPMF_ERR_STATE, "m_products: no resident M");
    CHECK(ensure_common(ctx));
    const long long n = ctx->n, p = ctx->p;
    const int K = (int)Kw;
    const int KPmax = ktp_for(std::min(K, 32));
    const long long nRB = (n + 255) / 256;          // the most row blocks any chunk's kernel uses (8 warps per CTA)
    // column spans: enough CTAs (one per SM at a time) for a dozen waves, few enough that the M %*% EF partials stay small
    const int nsplit = (int)std::max(1LL, std::min<long long>((p + kTileCols - 1) / kTileCols, (148LL * 12 + nRB - 1) / nRB));
    int cols_per_split = (int)((p + nsplit - 1) / nsplit);
    cols_per_split = ((cols_per_split + kTileCols - 1) / kTileCols) * kTileCols;
    // device scratch, carved out of one buffer
    size_t off = 0;
    auto carve = [&](size_t doubles) { const size_t o = off; off += (doubles + 1) & ~(size_t)1; return o; };
    const size_t oEL = carve((size_t)n * K), oEF = carve((size_t)p * K), oEFt = carve((size_t)p * (KPmax + 4));
    const size_t oMFp = carve((size_t)nsplit * KPmax * n), oRsq = carve((size_t)nsplit * n);
    const size_t oMtLp = carve((size_t)nRB * p * KPmax), oCsq = carve((size_t)nRB * p);
    const size_t oMF = carve((size_t)n * K), oMtL = carve((size_t)p * K), oGL = carve((size_t)K * K), oGF = carve((size_t)K * K);
    const size_t oSqR = carve((size_t)n), oSqC = carve((size_t)p), oR2r = carve((size_t)n), oR2c = carve((size_t)p);
    int rc = reserve(ctx, ctx->prod, sizeof(double) * off);
    CHECK(ranks_agree(ctx, rc));
    double *base = as<double>(ctx->prod);
    CU(cudaMemcpyAsync(base + oEL, EL_w, sizeof(double) * (size_t)n * K, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(base + oEF, EF_w, sizeof(double) * (size_t)p * K, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaEventRecord(ctx->ev0, ctx->stream));
    float ms_a = 0.f, ms_b = 0.f;
    for (int k0 = 0; k0 < K; k0 += 32) {             // factor columns in chunks of at most 32 (one pass over M each)
        const int Kc = std::min(32, K - k0);
        const int KP = ktp_for(Kc);
        transpose_ef_kernel<<<(unsigned)((p * (KP + 4) + 255) / 256), 256, 0, ctx->stream>>>(base + oEF + (size_t)p * k0, p, Kc, KP + 4, base + oEFt);
        CU(cudaGetLastError());
        ProdArgs a;
        memset(&a, 0, sizeof a);
        a.n = n; a.p = p; a.K = Kc; a.M = as<double>(ctx->Mcur); a.EL = base + oEL + (size_t)n * k0; a.EFt = base + oEFt;
        a.MFpart = base + oMFp; a.rowsq = base + oRsq; a.MtLpart = base + oMtLp; a.colsq = base + oCsq;
        a.nsplit = nsplit; a.cols_per_split = cols_per_split;
        a.bulk = (n % 2 == 0 && !getenv("EBPMF_B200_NO_BULK")) ? 1 : 0;
        const int prows = prod_rows_for(KP);
        const int nrange = (int)((n + prows - 1) / prows);       // t(M) %*% EL: one partial per row block of this chunk's kernel
        CU(cudaEventRecord(ctx->ev_p1, ctx->stream));
        switch (KP) {
            case 8: CHECK((launch_products<8, 8>(ctx, a))); break;
            case 16: CHECK((launch_products<16, EBPMF_PROD_NW>(ctx, a))); break;
            case 24: CHECK((launch_products<24, EBPMF_PROD_NW>(ctx, a))); break;
            default: CHECK((launch_products<32, EBPMF_PROD_NW>(ctx, a))); break;
        }
        CU(cudaEventRecord(ctx->ev_p2, ctx->stream));
        sum_splits_kernel<<<(unsigned)(((long long)Kc * n + 255) / 256), 256, 0, ctx->stream>>>(base + oMFp, (long long)KP * n, (long long)Kc * n, nsplit,
                                                                                             base + oMF + (size_t)n * k0);
        mtl_gather_kernel<<<(unsigned)((p * KP + 255) / 256), 256, 0, ctx->stream>>>(base + oMtLp, nrange, p, KP, Kc, base + oMtL + (size_t)p * k0);
        CU(cudaGetLastError());
        if (k0 == 0) {                               // sums of squares of M: any chunk's pass gives them
            sum_splits_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(base + oRsq, n, n, nsplit, base + oSqR);
            sum_splits_kernel<<<(unsigned)((p + 255) / 256), 256, 0, ctx->stream>>>(base + oCsq, p, p, nrange, base + oSqC);
            CU(cudaGetLastError());
        }
        CU(cudaEventRecord(ctx->ev_p3, ctx->stream));
        ctx->launches += 3 + (k0 == 0 ? 2 : 0);
        if (kernel_ms) {                             // per-kernel times of the chunk (events are reused: read them now)
            CU(cudaEventSynchronize(ctx->ev_p3));
            float t = 0.f;
            CU(cudaEventElapsedTime(&t, ctx->ev_p1, ctx->ev_p2)); ms_a += t;
            CU(cudaEventElapsedTime(&t, ctx->ev_p2, ctx->ev_p3)); ms_b += t;
        }
    }
    if (r2_col || r2_row) {
        gram_kernel<<<(unsigned)(K * K), 256, 0, ctx->stream>>>(base + oEL, n, K, base + oGL);
        gram_kernel<<<(unsigned)(K * K), 256, 0, ctx->stream>>>(base + oEF, p, K, base + oGF);
        CU(cudaGetLastError());
        ctx->launches += 2;
    }
    if (ctx->comm) {                                 // rows are sharded: t(M) EL, the column sums and EL'EL span all ranks
        NC(g_nccl.AllReduce(base + oMtL, base + oMtL, (size_t)p * K, ncclFloat64, ncclSum, ctx->comm, ctx->stream));
        NC(g_nccl.AllReduce(base + oSqC, base + oSqC, (size_t)p, ncclFloat64, ncclSum, ctx->comm, ctx->stream));
        if (r2_col) NC(g_nccl.AllReduce(base + oGL, base + oGL, (size_t)K * K, ncclFloat64, ncclSum, ctx->comm, ctx->stream));
    }
    if (r2_col) {
        r2_from_products_kernel<<<(unsigned)((p + 255) / 256), 256, 0, ctx->stream>>>(base + oSqC, base + oEF, base + oMtL, p, K, p, 0,
                                                                                     base + oGL, base + oR2c);
        CU(cudaGetLastError());
        ++ctx->launches;
    }
    if (r2_row) {
        r2_from_products_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(base + oSqR, base + oEL, base + oMF, n, K, n, 0,
                                                                                     base + oGF, base + oR2r);
        CU(cudaGetLastError());
        ++ctx->launches;
    }
    CU(cudaEventRecord(ctx->ev1, ctx->stream));
    CU(cudaMemcpyAsync(MF, base + oMF, sizeof(double) * (size_t)n * K, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(MtL, base + oMtL, sizeof(double) * (size_t)p * K, cudaMemcpyDeviceToHost, ctx->stream));
    if (r2_col) CU(cudaMemcpyAsync(r2_col, base + oR2c, sizeof(double) * (size_t)p, cudaMemcpyDeviceToHost, ctx->stream));
    if (r2_row) CU(cudaMemcpyAsync(r2_row, base + oR2r, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (kernel_ms) {
        float t = 0.f;
        CU(cudaEventElapsedTime(&t, ctx->ev0, ctx->ev1));
        kernel_ms[0] = t; kernel_ms[1] = ms_a; kernel_ms[2] = ms_b;
    }
    return EBPMF_OK;
}

// ---- drop-in dense signatures ------------------------------------------------------------

int ebpmf_vga_newton(ebpmf_ctx *ctx, int64_t n, int64_t p, const double *M, const double *X,
                     const int32_t *X_colptr, const int32_t *X_rowidx, const double *X_x, const double *S, int S_kind,
                     const double *Beta, int Beta_kind, const double *Sigma2, int Sigma2_kind, int maxiter, double tol,
                     double *M_out, double *V_out, ebpmf_solve_info *info) {
    if (!ctx || !M || !S || !Beta || !Sigma2 || !M_out || n < 1 || p < 1)
        return fail(ctx, EBPMF_ERR_ARG, "vga_newton: bad arguments");
    if (!X && !(X_colptr && (X_colptr[p] == 0 || (X_rowidx && X_x))))
        return fail(ctx, EBPMF_ERR_ARG, "vga_newton: X must be dense or dgCMatrix slots");
    if (maxiter < 1 || maxiter > 65534) return fail(ctx, EBPMF_ERR_ARG, "vga_newton: maxiter must be in [1, 65534]");
    CHECK(ensure_common(ctx));
    DenseArgs a;
    memset(&a, 0, sizeof a);
    if (stride_for(S_kind, n, &a.S_rs, &a.S_cs) || stride_for(Beta_kind, n, &a.B_rs, &a.B_cs) ||
        stride_for(Sigma2_kind, n, &a.V_rs, &a.V_cs))
        return fail(ctx, EBPMF_ERR_ARG, "vga_newton: bad operand kind");
    const size_t NP = (size_t)n * (size_t)p, bytes = sizeof(double) * NP;
    // scratch: 0 M_in, 1 M_out, 2 X, 3 V, 4 Beta|Sigma2|S packed
    CHECK(reserve(ctx, ctx->scratch[0], bytes));
    CHECK(reserve(ctx, ctx->scratch[1], bytes));
    CHECK(reserve(ctx, ctx->scratch[2], bytes));
    if (V_out) CHECK(reserve(ctx, ctx->scratch[3], bytes));
    const size_t lS = operand_len(S_kind, n, p), lB = operand_len(Beta_kind, n, p), lV = operand_len(Sigma2_kind, n, p);
    CHECK(reserve(ctx, ctx->scratch[4], sizeof(double) * (lS + lB + lV)));
    double *dS = as<double>(ctx->scratch[4]), *dB = dS + lS, *dV = dB + lB;
    {   // n x p operands go through the pinned ring (pageable R memory at close to pinned speed), the rest directly
        void *dsts[5] = {ctx->scratch[0].ptr, ctx->scratch[2].ptr, dS, dB, dV};
        const double *srcs[5] = {M, X, lS == NP ? S : nullptr, lB == NP ? Beta : nullptr, lV == NP ? Sigma2 : nullptr};
        CHECK(matrices_in(ctx, dsts, srcs, 5, n, p));
    }
    if (lS != NP) CU(cudaMemcpyAsync(dS, S, sizeof(double) * lS, cudaMemcpyHostToDevice, ctx->stream));
    if (lB != NP) CU(cudaMemcpyAsync(dB, Beta, sizeof(double) * lB, cudaMemcpyHostToDevice, ctx->stream));
    if (lV != NP) CU(cudaMemcpyAsync(dV, Sigma2, sizeof(double) * lV, cudaMemcpyHostToDevice, ctx->stream));
    if (!X) {
        const long long nnz = X_colptr[p];
        CHECK(reserve(ctx, ctx->scratch[5], (sizeof(int) + sizeof(double)) * (size_t)std::max(1LL, nnz) + sizeof(int) * (size_t)(p + 1) + 16));
        double *dx = as<double>(ctx->scratch[5]);
        int *di = reinterpret_cast<int *>(dx + std::max(1LL, nnz));
        int *dp = di + std::max(1LL, nnz);
        CU(cudaMemcpyAsync(dp, X_colptr, sizeof(int) * (size_t)(p + 1), cudaMemcpyHostToDevice, ctx->stream));
        if (nnz > 0) {
            CU(cudaMemcpyAsync(di, X_rowidx, sizeof(int) * (size_t)nnz, cudaMemcpyHostToDevice, ctx->stream));
            CU(cudaMemcpyAsync(dx, X_x, sizeof(double) * (size_t)nnz, cudaMemcpyHostToDevice, ctx->stream));
        }
        CU(cudaMemsetAsync(ctx->scratch[2].ptr, 0, bytes, ctx->stream));
        scatter_csc_kernel<<<(unsigned)((p * 32 + 255) / 256), 256, 0, ctx->stream>>>(dp, di, dx, n, p,
                                                                                      as<double>(ctx->scratch[2]));
        CU(cudaGetLastError());
    }
    const long long nRB = (n + kRowsPerCta - 1) / kRowsPerCta, nRW = (n + 31) / 32, nJG = (p + kE - 1) / kE;
    CHECK(reserve(ctx, ctx->kunit, sizeof(unsigned short) * (size_t)nJG * (size_t)nRW));
    CHECK(reserve(ctx, ctx->status, sizeof(Status)));
    a.n = n; a.p = p;
    a.M_in = as<double>(ctx->scratch[0]); a.M_out = as<double>(ctx->scratch[1]);
    a.V_out = V_out ? as<double>(ctx->scratch[3]) : nullptr;
    a.X = as<double>(ctx->scratch[2]);
    a.S = dS; a.Beta = dB; a.Sigma2 = dV;
    a.kunit = as<unsigned short>(ctx->kunit); a.nRW = nRW;
    a.status = as<Status>(ctx->status);
    a.tables = as<MathTables>(ctx->tables);
    a.maxiter = maxiter; a.tol = tol;
    a.cols_per_cta = cols_per_cta_for(p, nRB);
    dim3 grid((unsigned)((p + a.cols_per_cta - 1) / a.cols_per_cta), (unsigned)nRB);
    int rc = run_global_stop(
        ctx, maxiter,
        [&] { vga_dense_kernel<kE, MODE_FIRST><<<grid, kThreads, 0, ctx->stream>>>(a); CU(cudaGetLastError()); return EBPMF_OK; },
        [&] { vga_dense_kernel<kE, MODE_TOPUP><<<grid, kThreads, 0, ctx->stream>>>(a); CU(cudaGetLastError()); return EBPMF_OK; },
        [&] { return EBPMF_OK; }, info);
    if (rc != EBPMF_OK) return rc;
    if (info) { info->sym = info->ssexp = info->slogv = 0.0; }
    CHECK(matrices_out(ctx, M_out, ctx->scratch[1].ptr, V_out, V_out ? ctx->scratch[3].ptr : nullptr, n, p));
    return EBPMF_OK;
}

int ebpmf_vga_fixed_iter(ebpmf_ctx *ctx, int64_t n, int64_t p, const double *M, const double *V, const double *Y,
                         const double *S, int S_kind, const double *Beta, int Beta_kind, const double *Sigma2,
                         int Sigma2_kind, int maxiter, double *M_out, double *V_out) {
    if (!ctx || !M || !Y || !S || !Beta || !Sigma2 || !M_out || !V_out || n < 1 || p < 1)
        return fail(ctx, EBPMF_ERR_ARG, "vga_fixed_iter: bad arguments");
    if (maxiter < 1) return fail(ctx, EBPMF_ERR_ARG, "vga_fixed_iter: maxiter must be >= 1");
    CHECK(ensure_common(ctx));
    FixedArgs a;
    memset(&a, 0, sizeof a);
    if (stride_for(S_kind, n, &a.S_rs, &a.S_cs) || stride_for(Beta_kind, n, &a.B_rs, &a.B_cs) ||
        stride_for(Sigma2_kind, n, &a.V_rs, &a.V_cs))
        return fail(ctx, EBPMF_ERR_ARG, "vga_fixed_iter: bad operand kind");
    const size_t NP = (size_t)n * (size_t)p, bytes = sizeof(double) * NP;
    for (int b = 0; b < 4; ++b) CHECK(reserve(ctx, ctx->scratch[b], bytes));
    if (V) CHECK(reserve(ctx, ctx->scratch[5], bytes));
    const size_t lS = operand_len(S_kind, n, p), lB = operand_len(Beta_kind, n, p), lV = operand_len(Sigma2_kind, n, p);
    CHECK(reserve(ctx, ctx->scratch[4], sizeof(double) * (lS + lB + lV)));
    double *dS = as<double>(ctx->scratch[4]), *dB = dS + lS, *dV = dB + lB;
    {
        void *dsts[6] = {ctx->scratch[0].ptr, ctx->scratch[2].ptr, V ? ctx->scratch[5].ptr : nullptr, dS, dB, dV};
        const double *srcs[6] = {M, Y, V, lS == NP ? S : nullptr, lB == NP ? Beta : nullptr, lV == NP ? Sigma2 : nullptr};
        CHECK(matrices_in(ctx, dsts, srcs, 6, n, p));
    }
    if (lS != NP) CU(cudaMemcpyAsync(dS, S, sizeof(double) * lS, cudaMemcpyHostToDevice, ctx->stream));
    if (lB != NP) CU(cudaMemcpyAsync(dB, Beta, sizeof(double) * lB, cudaMemcpyHostToDevice, ctx->stream));
    if (lV != NP) CU(cudaMemcpyAsync(dV, Sigma2, sizeof(double) * lV, cudaMemcpyHostToDevice, ctx->stream));
    a.n = n; a.p = p;
    a.M_in = as<double>(ctx->scratch[0]); a.V_in = V ? as<double>(ctx->scratch[5]) : nullptr;
    a.Y = as<double>(ctx->scratch[2]);
    a.S = dS; a.Beta = dB; a.Sigma2 = dV;
    a.M_out = as<double>(ctx->scratch[1]); a.V_out = as<double>(ctx->scratch[3]);
    a.tables = as<MathTables>(ctx->tables);
    a.maxiter = maxiter;
    const long long nRB = (n + kRowsPerCta - 1) / kRowsPerCta;
    dim3 grid((unsigned)std::min<long long>(p, 4096), (unsigned)nRB);
    CU(cudaEventRecord(ctx->ev0, ctx->stream));
    vga_fixed_iter_kernel<<<grid, kThreads, 0, ctx->stream>>>(a);
    CU(cudaGetLastError());
    CU(cudaEventRecord(ctx->ev1, ctx->stream));
    ++ctx->launches;
    CU(cudaStreamSynchronize(ctx->stream));
    CHECK(matrices_out(ctx, M_out, ctx->scratch[1].ptr, V_out, ctx->scratch[3].ptr, n, p));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ctx->last_kernel_ms = ms;
    return EBPMF_OK;
}

int ebpmf_vga_low_memory(ebpmf_ctx *ctx, int64_t n, int64_t p, int64_t K, const double *M, const double *Y,
                         const int32_t *Y_colptr, const int32_t *Y_rowidx, const double *Y_x, const double *l0,
                         const double *f0, const double *EL, const double *EF, const double *sigma2, int var_type,
                         int maxiter, double tol, double *M_out, ebpmf_solve_info *info) {
    if (!ctx || !M || !l0 || !f0 || !EL || !EF || !sigma2 || !M_out || n < 1 || p < 1 || K < 1)
        return fail(ctx, EBPMF_ERR_ARG, "vga_low_memory: bad arguments");
    if (maxiter < 1 || maxiter > 65534) return fail(ctx, EBPMF_ERR_ARG, "vga_low_memory: maxiter must be in [1, 65534]");
    // R falls through both `if(var_type==...)` blocks for an unknown var_type and returns the
    // clamped M (:78,:108); the clamp needs the kernels too, so run zero iterations of by_col
    // semantics is NOT the same thing -- report the unsupported type instead of guessing.
    if (var_type < 0 || var_type > 2) return fail(ctx, EBPMF_ERR_VAR_TYPE, "Non-supported var type");
    // This entry point takes over the context's resident state (Y, M, factors, sigma2) and does NOT give it
    // back: whatever happens, the resident state is invalidated on return, so that a later resident step on
    // the same context fails with EBPMF_ERR_STATE instead of silently running on this call's matrices.
    struct Invalidate {
        ebpmf_ctx *c;
        ~Invalidate() { c->have_y = c->have_m = c->have_f = c->have_s2 = c->have_v = c->have_vsum = c->have_hist = false; c->can_rewind = false; }
    } invalidate{ctx};
    if (Y) CHECK(ebpmf_ctx_set_y_dense(ctx, n, p, Y));
    else CHECK(ebpmf_ctx_set_y_csc(ctx, n, p, Y_colptr, Y_rowidx, Y_x));
    CHECK(ebpmf_ctx_set_factors(ctx, K, EL, EF));
    CHECK(ebpmf_ctx_set_sigma2(ctx, var_type, sigma2));
    CHECK(ebpmf_ctx_set_m(ctx, M));
    CHECK(reserve(ctx, ctx->l0, sizeof(double) * (size_t)n));
    CHECK(reserve(ctx, ctx->f0, sizeof(double) * (size_t)p));
    CU(cudaMemcpyAsync(ctx->l0.ptr, l0, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->f0.ptr, f0, sizeof(double) * (size_t)p, cudaMemcpyHostToDevice, ctx->stream));
    StepPlan pl = make_plan(ctx, 1);
    CHECK(build_order(ctx, pl));
    const FusedArgs a = fused_args(ctx, pl, maxiter, tol, false);
    int rc = run_global_stop(
        ctx, maxiter, [&] { return launch_fused<FORMULA_A10, MODE_FIRST>(ctx, a, 0, pl.ntiles); },
        [&] { return launch_fused<FORMULA_A10, MODE_TOPUP>(ctx, a, 0, pl.ntiles); }, [&] { return (int)EBPMF_OK; }, info);
    if (rc != EBPMF_OK) return rc;
    if (info) { info->sym = info->ssexp = info->slogv = 0.0; }
    std::swap(ctx->Mcur, ctx->Malt);
    return ebpmf_ctx_get_m(ctx, M_out);
}

int ebpmf_ctx_get_shape(ebpmf_ctx *ctx, int64_t *n, int64_t *p, int64_t *K, int64_t *nnz) {
    if (!ctx) return fail(ctx, EBPMF_ERR_ARG, "get_shape: ctx is NULL");
    if (!ctx->have_y) return fail(ctx, EBPMF_ERR_STATE, "get_shape: Y not set");
    if (n) *n = ctx->n;
    if (p) *p = ctx->p;
    if (K) *K = ctx->have_f ? ctx->K : 0;
    if (nnz) *nnz = ctx->nnz;
    return EBPMF_OK;
}

int ebpmf_ctx_get_y_csc(ebpmf_ctx *ctx, int32_t *colptr, int32_t *rowidx, double *x) {
    if (!ctx || !colptr) return fail(ctx, EBPMF_ERR_ARG, "get_y_csc: bad arguments");
    if (!ctx->have_y) return fail(ctx, EBPMF_ERR_STATE, "get_y_csc: Y not set");
    CHECK(ensure_common(ctx));
    CU(cudaMemcpyAsync(colptr, ctx->colptr.ptr, sizeof(int) * (size_t)(ctx->p + 1), cudaMemcpyDeviceToHost, ctx->stream));
    if (ctx->nnz > 0 && rowidx && x) {
        CU(cudaMemcpyAsync(rowidx, ctx->rowidx.ptr, sizeof(int) * (size_t)ctx->nnz, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaMemcpyAsync(x, ctx->yx.ptr, sizeof(double) * (size_t)ctx->nnz, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CU(cudaStreamSynchronize(ctx->stream));
    return EBPMF_OK;
}

int ebpmf_ctx_get_factors(ebpmf_ctx *ctx, double *EL, double *EF) {
    if (!ctx || !EL || !EF) return fail(ctx, EBPMF_ERR_ARG, "get_factors: bad arguments");
    if (!ctx->have_f) return fail(ctx, EBPMF_ERR_STATE, "get_factors: factors not set");
    CHECK(ensure_common(ctx));
    CU(cudaMemcpyAsync(EL, ctx->EL.ptr, sizeof(double) * (size_t)ctx->n * (size_t)ctx->K, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(EF, ctx->EF.ptr, sizeof(double) * (size_t)ctx->p * (size_t)ctx->K, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return EBPMF_OK;
}

int ebpmf_ctx_get_sigma2(ebpmf_ctx *ctx, double *sigma2) {
    if (!ctx || !sigma2) return fail(ctx, EBPMF_ERR_ARG, "get_sigma2: bad arguments");
    if (!ctx->have_s2) return fail(ctx, EBPMF_ERR_STATE, "get_sigma2: sigma2 not set");
    CHECK(ensure_common(ctx));
    const size_t len = ctx->var_type == EBPMF_VAR_BY_COL ? (size_t)ctx->p : ctx->var_type == EBPMF_VAR_BY_ROW ? (size_t)ctx->n : 1;
    CU(cudaMemcpyAsync(sigma2, ctx->sigma2.ptr, sizeof(double) * len, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return EBPMF_OK;
}

int ebpmf_ctx_get_block(ebpmf_ctx *ctx, int which, int64_t row0, int64_t nrows, int64_t col0, int64_t ncols, double *out) {
    if (!ctx || !out || row0 < 0 || nrows < 1 || col0 < 0 || ncols < 1 || which < 0 || which > 2)
        return fail(ctx, EBPMF_ERR_ARG, "get_block: bad arguments");
    if (!ctx->have_m || row0 + nrows > ctx->n || col0 + ncols > ctx->p)
        return fail(ctx, EBPMF_ERR_STATE, "get_block: M not set or block out of range");
    if (which == 1 && !ctx->have_v) return fail(ctx, EBPMF_ERR_STATE, "get_block: V was not materialised");
    if (which == 2 && !ctx->can_rewind) return fail(ctx, EBPMF_ERR_STATE, "get_block: no step to look behind");
    CHECK(ensure_common(ctx));
    // a step never writes its input: after it the result is in Mcur and the input still in Malt
    const double *base = which == 1 ? as<double>(ctx->V) : which == 2 ? as<double>(ctx->Malt) : as<double>(ctx->Mcur);
    const double *src = base + row0 + (size_t)ctx->n * (size_t)col0;
    CU(cudaMemcpy2DAsync(out, sizeof(double) * (size_t)nrows, src, sizeof(double) * (size_t)ctx->n,
                         sizeof(double) * (size_t)nrows, (size_t)ncols, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return EBPMF_OK;
}

int ebpmf_ctx_get_rows(ebpmf_ctx *ctx, int which, int64_t row0, int64_t nrows, double *out) {
    if (!ctx || !ctx->have_m) return fail(ctx, EBPMF_ERR_STATE, "get_rows: M not set or rows out of range");
    return ebpmf_ctx_get_block(ctx, which, row0, nrows, 0, ctx->p, out);
}

int ebpmf_ctx_sim_data_log(ebpmf_ctx *ctx, int64_t n, int64_t p, int64_t K, int64_t row0, uint64_t seed,
                           double log_offset, double bg_lo, double bg_hi, double pi0, double noise, int warm_start,
                           int64_t *nnz_out) {
    if (!ctx || n < 1 || p < 1 || K < 1 || row0 < 0 || !(bg_lo > 0) || !(bg_hi >= bg_lo))
        return fail(ctx, EBPMF_ERR_ARG, "sim_data_log: bad arguments");
    const int KT = (int)K + 2;
    const int ktp = ktp_for(KT);
    if (ktp < 0 || KT > 32) return fail(ctx, EBPMF_ERR_ARG, "sim_data_log: K + 2 = %d factor columns not supported (max 32)", KT);
    CHECK(ensure_common(ctx));
    const size_t NP = (size_t)n * (size_t)p;
    CHECK(reserve(ctx, ctx->Mcur, sizeof(double) * NP));
    CHECK(reserve(ctx, ctx->Malt, sizeof(double) * NP));
    CHECK(reserve(ctx, ctx->EL, sizeof(double) * (size_t)n * KT));
    CHECK(reserve(ctx, ctx->EF, sizeof(double) * (size_t)p * KT));
    CHECK(reserve(ctx, ctx->EFt, sizeof(double) * (size_t)p * ktp));
    CHECK(reserve(ctx, ctx->sigma2, sizeof(double) * (size_t)std::max(n, p)));
    CHECK(reserve(ctx, ctx->scratch[0], sizeof(double) * ((size_t)n * K + (size_t)p * K + (size_t)n + (size_t)p)));
    CHECK(reserve(ctx, ctx->colptr, sizeof(int) * (size_t)(p + 1)));
    CHECK(reserve(ctx, ctx->misc, sizeof(double) * (size_t)(3 * std::max(n, p) + 4096)));
    sim::GenArgs g;
    g.n = n; g.p = p; g.row0 = row0; g.K = (int)K; g.seed = seed;
    g.log_offset = log_offset; g.bg_lo = bg_lo; g.bg_hi = bg_hi; g.pi0 = pi0; g.noise = noise;
    g.Lt = as<double>(ctx->scratch[0]); g.Ft = g.Lt + (size_t)n * K; g.logl0 = g.Ft + (size_t)p * K; g.logf0 = g.logl0 + n;
    g.EL = as<double>(ctx->EL); g.EF = as<double>(ctx->EF); g.sigma2 = as<double>(ctx->sigma2);
    sim::gen_factors_kernel<<<(unsigned)((n + p + 255) / 256), 256, 0, ctx->stream>>>(g);
    CU(cudaGetLastError());
    int *counts = as<int>(ctx->misc);
    const unsigned blocks = (unsigned)((p * 32 + 255) / 256);
    sim::gen_count_kernel<<<blocks, 256, 0, ctx->stream>>>(g, warm_start, as<double>(ctx->Mcur), counts);
    CU(cudaGetLastError());
    scan_counts_kernel<<<1, 1024, 0, ctx->stream>>>(counts, p, as<int>(ctx->colptr));
    CU(cudaGetLastError());
    int nnz = 0;
    CU(cudaMemcpyAsync(&nnz, as<int>(ctx->colptr) + p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (nnz < 0) return fail(ctx, EBPMF_ERR_ARG, "sim_data_log: more than 2^31-1 non-zeros in one shard");
    CHECK(reserve(ctx, ctx->rowidx, sizeof(int) * (size_t)std::max(1, nnz)));
    CHECK(reserve(ctx, ctx->yx, sizeof(double) * (size_t)std::max(1, nnz)));
    sim::gen_fill_kernel<<<blocks, 256, 0, ctx->stream>>>(g, as<int>(ctx->colptr), as<int>(ctx->rowidx), as<double>(ctx->yx));
    CU(cudaGetLastError());
    ctx->nnz = nnz;
    CHECK(finish_set_y(ctx, n, p));
    transpose_ef_kernel<<<(unsigned)((p * ktp + 255) / 256), 256, 0, ctx->stream>>>(as<double>(ctx->EF), p, KT, ktp,
                                                                                   as<double>(ctx->EFt));
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(ctx->stream));
    sigma_consts_kernel<<<(unsigned)((p + 255) / 256), 256, 0, ctx->stream>>>(as<double>(ctx->sigma2), p,
                                                                             as<double>(ctx->sc1), as<double>(ctx->sc2));
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->K = KT; ctx->KTP = ktp; ctx->ldk = ktp; ctx->var_type = EBPMF_VAR_BY_COL;
    ctx->have_m = ctx->have_f = ctx->have_s2 = true;
    if (nnz_out) *nnz_out = nnz;
    return EBPMF_OK;
}

int ebpmf_ctx_set_option(ebpmf_ctx *ctx, const char *name, double value) {
    if (!ctx || !name) return fail(ctx, EBPMF_ERR_ARG, "set_option: bad arguments");
    if (!strcmp(name, "delta")) {
        if (!(value >= 0.0) || value > 1e-10) return fail(ctx, EBPMF_ERR_ARG, "set_option: delta must be in [0, 1e-10]");
        ctx->delta = value;
        return EBPMF_OK;
    }
    if (!strcmp(name, "debug_verify_scale")) {
        if (!(value > 0.0) || value > 1.0) return fail(ctx, EBPMF_ERR_ARG, "set_option: debug_verify_scale must be in (0, 1]");
        ctx->debug_verify_scale = value;
        return EBPMF_OK;
    }
    if (!strcmp(name, "freeze") || !strcmp(name, "history") || !strcmp(name, "speculate")) {
        if (value != 0.0 && value != 1.0) return fail(ctx, EBPMF_ERR_ARG, "set_option: %s is 0 or 1", name);
        (name[0] == 'f' ? ctx->freeze : name[0] == 'h' ? ctx->use_history : ctx->speculate) = (int)value;
        return EBPMF_OK;
    }
    if (!strcmp(name, "pipe_chunks")) {
        if (!(value >= 0.0) || value > 4096.0) return fail(ctx, EBPMF_ERR_ARG, "set_option: pipe_chunks must be in [0, 4096]");
        ctx->pipe_chunks = (int)value;
        return EBPMF_OK;
    }
    if (!strcmp(name, "forget_history")) {
        ctx->have_hist = false; ctx->spec_waste = 0.0;
        return EBPMF_OK;
    }
    return fail(ctx, EBPMF_ERR_ARG, "set_option: unknown option '%s'", name);
}

int ebpmf_ctx_peer_count(ebpmf_ctx *ctx, int *npeers) {
    if (!ctx || !npeers) return fail(ctx, EBPMF_ERR_ARG, "peer_count: bad arguments");
    *npeers = ctx->npeers;
    return EBPMF_OK;
}

int ebpmf_ctx_last_diag(ebpmf_ctx *ctx, uint64_t *topup_units, uint64_t *unit_evals) {
    if (!ctx) return fail(ctx, EBPMF_ERR_ARG, "last_diag: ctx is NULL");
    if (topup_units) *topup_units = ctx->diag_topups;
    if (unit_evals) *unit_evals = ctx->diag_iters;
    return EBPMF_OK;
}

int ebpmf_ctx_counters(ebpmf_ctx *ctx, double *out, int count) {
    if (!ctx || !out || count < 1) return fail(ctx, EBPMF_ERR_ARG, "counters: bad arguments");
    const double v[9] = {(double)ctx->diag_topups, (double)ctx->diag_iters, (double)ctx->diag_topup_tiles,
                         (double)ctx->launches, ctx->diag_spec_bytes, ctx->diag_recopy_bytes, (double)ctx->cur_ntiles,
                         ctx->have_hist ? 1.0 : 0.0, ctx->last_kernel_ms};
    for (int q = 0; q < count && q < 9; ++q) out[q] = v[q];
    return EBPMF_OK;
}

int ebpmf_probe_fp64_tflops(ebpmf_ctx *ctx, double *tflops) {
    if (!ctx || !tflops) return fail(ctx, EBPMF_ERR_ARG, "probe: bad arguments");
    CHECK(ensure_common(ctx));
    CHECK(reserve(ctx, ctx->misc, sizeof(double) * 4096));
    const int iters = 20000, blocks = 148 * 8;
    fp64_probe_kernel<<<blocks, 256, 0, ctx->stream>>>(as<double>(ctx->misc), 1000, 1.0);   // warm-up
    CU(cudaEventRecord(ctx->ev0, ctx->stream));
    fp64_probe_kernel<<<blocks, 256, 0, ctx->stream>>>(as<double>(ctx->misc), iters, 1.0);
    CU(cudaEventRecord(ctx->ev1, ctx->stream));
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(ctx->stream));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    const double flops = 2.0 * 8.0 * (double)iters * 256.0 * (double)blocks;
    *tflops = flops / (ms * 1e-3) / 1e12;
    return EBPMF_OK;
}


// ---- single-process multi-GPU group (what an R session uses: R is one process) -------------
}  // extern "C"

struct ebpmf_group {
    std::vector<ebpmf_ctx *> ctx;
    std::vector<long long> row0, nrows;
    long long n = 0, p = 0;
    bool have_m = false;
    std::string err;
};

namespace {
int gfail(ebpmf_group *g, int code, const std::string &msg) {
    if (g) g->err = msg;
    g_last_error = msg;
    return code;
}

// one host thread per GPU; every thread runs fn(rank) against its own context.  The per-context
// calls are synchronous, so the threads are what overlaps the GPUs (and they meet in the NCCL
// collectives inside ebpmf_ctx_vga_step).
template <typename F> int run_ranks(ebpmf_group *g, F fn) {
    const int N = (int)g->ctx.size();
    std::vector<int> rc(N, EBPMF_OK);
    std::vector<std::thread> th;
    for (int r = 1; r < N; ++r) th.emplace_back([&, r] { rc[r] = fn(r); });
    rc[0] = fn(0);
    for (auto &t : th) t.join();
    for (int r = 0; r < N; ++r)
        if (rc[r] != EBPMF_OK) return gfail(g, rc[r], std::string("rank ") + std::to_string(r) + ": " + g->ctx[r]->err);
    return EBPMF_OK;
}
}  // namespace

extern "C" {

int ebpmf_group_create(int ndev, const int *devices, ebpmf_group **out) {
    if (!out || ndev < 1) return gfail(nullptr, EBPMF_ERR_ARG, "group_create: bad arguments");
    *out = nullptr;
    ebpmf_group *g = new (std::nothrow) ebpmf_group();
    if (!g) return gfail(nullptr, EBPMF_ERR_NOMEM, "out of host memory");
    for (int r = 0; r < ndev; ++r) {
        ebpmf_ctx *c = nullptr;
        int rc = ebpmf_ctx_create(devices ? devices[r] : r, &c);
        if (rc != EBPMF_OK) {
            for (ebpmf_ctx *q : g->ctx) ebpmf_ctx_destroy(q);
            delete g;
            return rc;
        }
        g->ctx.push_back(c);
    }
    if (ndev > 1) {
        unsigned char id[128];
        int rc = ebpmf_nccl_unique_id(id);
        if (rc == EBPMF_OK) rc = run_ranks(g, [&](int r) { return ebpmf_ctx_comm_init(g->ctx[r], r, ndev, id); });
        if (rc != EBPMF_OK) {
            const std::string msg = g->err.empty() ? g_last_error : g->err;
            for (ebpmf_ctx *q : g->ctx) ebpmf_ctx_destroy(q);
            delete g;
            return gfail(nullptr, rc, msg);
        }
    }
    *out = g;
    return EBPMF_OK;
}

int ebpmf_group_destroy(ebpmf_group *g) {
    if (!g) return EBPMF_OK;
    for (ebpmf_ctx *c : g->ctx) ebpmf_ctx_destroy(c);
    delete g;
    return EBPMF_OK;
}

const char *ebpmf_group_last_error(const ebpmf_group *g) { return g ? g->err.c_str() : g_last_error.c_str(); }

int ebpmf_group_size(const ebpmf_group *g) { return g ? (int)g->ctx.size() : 0; }

// Rows are dealt to the GPUs in 256-row blocks (the kernels' row-block size), as evenly as possible.
int ebpmf_group_set_y_csc(ebpmf_group *g, int64_t n, int64_t p, const int32_t *colptr, const int32_t *rowidx,
                          const double *x) {
    if (!g || n < 1 || p < 1 || !colptr) return gfail(g, EBPMF_ERR_ARG, "group_set_y_csc: bad arguments");
    const int N = (int)g->ctx.size();
    if (colptr[p] > 0 && (!rowidx || !x)) return gfail(g, EBPMF_ERR_ARG, "group_set_y_csc: @i / @x missing");
    if (validate_csc(g->ctx[0], n, p, colptr, rowidx) != EBPMF_OK) return gfail(g, EBPMF_ERR_ARG, g->ctx[0]->err);
    const long long nblk = (n + kRowsPerCta - 1) / kRowsPerCta;
    if (nblk < N) return gfail(g, EBPMF_ERR_ARG, "group_set_y_csc: fewer 256-row blocks than GPUs; use a smaller group");
    g->row0.assign(N, 0); g->nrows.assign(N, 0);
    long long r0 = 0;
    for (int r = 0; r < N; ++r) {
        long long rows = (nblk / N + (r < nblk % N ? 1 : 0)) * kRowsPerCta;
        rows = std::max(0LL, std::min(rows, (long long)n - r0));
        g->row0[r] = r0; g->nrows[r] = rows; r0 += rows;
    }
    g->n = n; g->p = p; g->have_m = false;
    return run_ranks(g, [&](int r) {
        // the row slice Y[row0:row0+nrows, ] as its own dgCMatrix (what `Y[b,]` does in R/ebpmf_log.R:199)
        const long long lo = g->row0[r], hi = lo + g->nrows[r];
        std::vector<int32_t> cp((size_t)p + 1, 0);
        for (long long j = 0; j < p; ++j) {
            int c = 0;
            for (int q = colptr[j]; q < colptr[j + 1]; ++q) c += (rowidx[q] >= lo && rowidx[q] < hi);
            cp[(size_t)j + 1] = cp[(size_t)j] + c;
        }
        std::vector<int32_t> ri((size_t)std::max(1, cp[(size_t)p]));
        std::vector<double> xx((size_t)std::max(1, cp[(size_t)p]));
        for (long long j = 0; j < p; ++j) {
            int w = cp[(size_t)j];
            for (int q = colptr[j]; q < colptr[j + 1]; ++q)
                if (rowidx[q] >= lo && rowidx[q] < hi) { ri[(size_t)w] = (int32_t)(rowidx[q] - lo); xx[(size_t)w] = x[q]; ++w; }
        }
        return ebpmf_ctx_set_y_csc(g->ctx[r], g->nrows[r], p, cp.data(), ri.data(), xx.data());
    });
}

int ebpmf_group_set_option(ebpmf_group *g, const char *name, double value) {
    if (!g) return gfail(g, EBPMF_ERR_ARG, "group_set_option: group is NULL");
    for (ebpmf_ctx *c : g->ctx) {
        int rc = ebpmf_ctx_set_option(c, name, value);
        if (rc != EBPMF_OK) return gfail(g, rc, c->err);
    }
    return EBPMF_OK;
}

// The whole VGA step on the FULL host matrices; each GPU works on its row shard (strided copies out
// of / into the column-major host arrays), the shards meet in the NCCL combine.  v_sum has length
// p ('by_col'), n ('by_row', written shard by shard) or 1.
int ebpmf_group_vga_step_host(ebpmf_group *g, int64_t n_in, int64_t p_in, const double *M_in, int64_t K, const double *EL,
                              const double *EF, int var_type, const double *sigma2, int maxiter, double tol,
                              double *M_out, double *V_out, double *v_sum, ebpmf_solve_info *info) {
    if (!g || !M_out || !EL || !EF || !sigma2) return gfail(g, EBPMF_ERR_ARG, "group_vga_step_host: bad arguments");
    if (g->n < 1) return gfail(g, EBPMF_ERR_STATE, "group_vga_step_host: Y not set");
    if (n_in != g->n || p_in != g->p) return gfail(g, EBPMF_ERR_ARG, "group_vga_step_host: non-conformable arrays (M does not have the shape of the resident Y)");
    if (!M_in && !g->have_m) return gfail(g, EBPMF_ERR_STATE, "group_vga_step_host: M_in is NULL and no M is resident (ebpmf_group_set_m)");
    const int N = (int)g->ctx.size();
    std::vector<ebpmf_solve_info> infos((size_t)N);
    const long long n = g->n;
    int rc = run_ranks(g, [&](int r) {
        ebpmf_ctx *c = g->ctx[r];
        const long long lo = g->row0[r];
        double *vs = nullptr;
        if (v_sum) vs = (var_type == EBPMF_VAR_BY_ROW) ? v_sum + lo : (r == 0 ? v_sum : nullptr);
        return step_host_impl(c, M_in ? M_in + lo : nullptr, n, K, EL + lo, n, EF, var_type,
                              var_type == EBPMF_VAR_BY_ROW ? sigma2 + lo : sigma2, maxiter, tol, M_out + lo, n,
                              V_out ? V_out + lo : nullptr, vs, &infos[(size_t)r]);
    });
    if (rc != EBPMF_OK) { g->have_m = false; return rc; }
    g->have_m = true;                       // every shard now holds its rows of res$M
    if (info) {
        *info = infos[0];                       // sums and the update count are global after the combine
        for (int r = 1; r < N; ++r) info->kernel_ms = std::max(info->kernel_ms, infos[(size_t)r].kernel_ms);
    }
    return EBPMF_OK;
}

// M_init -> the shards, once per ebpmf_log() call; afterwards ebpmf_group_vga_step_host(.., M_in = NULL, ..)
int ebpmf_group_set_m(ebpmf_group *g, int64_t n_in, int64_t p_in, const double *M) {
    if (!g || !M) return gfail(g, EBPMF_ERR_ARG, "group_set_m: bad arguments");
    if (g->n < 1) return gfail(g, EBPMF_ERR_STATE, "group_set_m: Y not set");
    if (n_in != g->n || p_in != g->p) return gfail(g, EBPMF_ERR_ARG, "group_set_m: non-conformable arrays");
    const long long n = g->n;
    int rc = run_ranks(g, [&](int r) { return set_m_ld(g->ctx[r], M + g->row0[r], n); });
    g->have_m = (rc == EBPMF_OK);
    return rc;
}

// sum(lfactorial(Y@x)) over all shards
int ebpmf_group_sum_lfactorial(ebpmf_group *g, double *out) {
    if (!g || !out) return gfail(g, EBPMF_ERR_ARG, "group_sum_lfactorial: bad arguments");
    std::vector<double> part(g->ctx.size(), 0.0);
    int rc = run_ranks(g, [&](int r) { return ebpmf_ctx_sum_lfactorial(g->ctx[r], &part[(size_t)r]); });
    if (rc != EBPMF_OK) return rc;
    double s = 0.0;
    for (double v : part) s += v;
    *out = s;
    return EBPMF_OK;
}

}  // extern "C"

// ---- batch_size-faithful, host-streamed step (R/ebpmf_log.R:195-203, :254-300) ---------------
namespace {
// the Y part of a context: swapped into a worker context for the duration of one batch
struct YState {
    DevBuf colptr, rowidx, yx, rbptr;
    long long n = 0, nnz = 0;
};

void swap_y(ebpmf_ctx *c, YState &y) {
    std::swap(c->colptr, y.colptr); std::swap(c->rowidx, y.rowidx);
    std::swap(c->yx, y.yx); std::swap(c->rbptr, y.rbptr);
    std::swap(c->n, y.n); std::swap(c->nnz, y.nnz);
}
}  // namespace

struct ebpmf_batched {
    std::vector<ebpmf_ctx *> slot;      // worker contexts, possibly on several devices; batch b runs on slot b % S
    std::vector<YState> ys;             // one per batch, resident
    std::vector<long long> row0;        // nbatch + 1
    long long n = 0, p = 0;
    double lfact = 0.0;                 // sum(lfactorial(Y@x)) of the whole matrix
    std::string err;
};

namespace {
int bfail(ebpmf_batched *h, int code, const std::string &msg) {
    if (h) h->err = msg;
    g_last_error = msg;
    return code;
}

void release_batches(ebpmf_batched *h) {
    const size_t S = h->slot.size();
    for (size_t b = 0; b < h->ys.size(); ++b) {
        YState &y = h->ys[b];
        cudaSetDevice(h->slot[b % S]->device);       // the slab lives on the device of the slot that runs it
        release(y.colptr); release(y.rowidx); release(y.yx); release(y.rbptr);
    }
    h->ys.clear();
}

// worker context c takes batch y: Y pointers, shapes, shape-dependent scratch
int attach_batch(ebpmf_ctx *c, YState &y, long long p) {
    swap_y(c, y);
    CHECK(ensure_common(c));
    CHECK(reserve_shape_buffers(c, c->n, p));
    c->have_y = true;
    c->have_m = c->have_f = c->have_s2 = c->have_v = c->have_vsum = false;
    return EBPMF_OK;
}

void detach_batch(ebpmf_ctx *c, YState &y) {
    swap_y(c, y);
    c->have_y = false;
}
}  // namespace

extern "C" {

// Batches are independent (their stop tests are separate), so several GPUs need no collective: slot s
// lives on devices[s % ndev] and batch b always runs on slot b % S, where its Y slab is resident.
int ebpmf_batched_create_multi(int ndev, const int *devices, int nslots_per_device, ebpmf_batched **out) {
    if (!out || ndev < 1 || ndev > 64 || nslots_per_device < 1 || nslots_per_device > 8)
        return bfail(nullptr, EBPMF_ERR_ARG, "batched_create: need 1..64 devices and 1..8 slots per device");
    *out = nullptr;
    ebpmf_batched *h = new (std::nothrow) ebpmf_batched();
    if (!h) return bfail(nullptr, EBPMF_ERR_NOMEM, "out of host memory");
    for (int s = 0; s < ndev * nslots_per_device; ++s) {
        ebpmf_ctx *c = nullptr;
        int rc = ebpmf_ctx_create(devices ? devices[s % ndev] : s % ndev, &c);
        if (rc != EBPMF_OK) {
            for (ebpmf_ctx *q : h->slot) ebpmf_ctx_destroy(q);
            delete h;
            return rc;
        }
        h->slot.push_back(c);
    }
    *out = h;
    return EBPMF_OK;
}

int ebpmf_batched_create(int device, int nslots, ebpmf_batched **out) {
    return ebpmf_batched_create_multi(1, &device, nslots, out);
}

int ebpmf_batched_destroy(ebpmf_batched *h) {
    if (!h) return EBPMF_OK;
    release_batches(h);
    for (ebpmf_ctx *c : h->slot) ebpmf_ctx_destroy(c);
    delete h;
    return EBPMF_OK;
}

const char *ebpmf_batched_last_error(const ebpmf_batched *h) { return h ? h->err.c_str() : g_last_error.c_str(); }

int ebpmf_batched_set_option(ebpmf_batched *h, const char *name, double value) {
    if (!h) return bfail(h, EBPMF_ERR_ARG, "batched_set_option: handle is NULL");
    for (ebpmf_ctx *c : h->slot) {
        int rc = ebpmf_ctx_set_option(c, name, value);
        if (rc != EBPMF_OK) return bfail(h, rc, c->err);
    }
    return EBPMF_OK;
}

// Y[b,] for every batch b (what `lapply(batches, function(b) Y[b,])` does, R/ebpmf_log.R:199): two passes
// over the non-zeros, then one upload per batch; each slab becomes its own resident CSC with its own
// row-block offset table.
int ebpmf_batched_set_y_csc(ebpmf_batched *h, int64_t n, int64_t p, const int32_t *colptr, const int32_t *rowidx,
                            const double *x, int64_t nbatch, const int64_t *batch_row0) {
    if (!h || n < 1 || p < 1 || !colptr || nbatch < 1 || !batch_row0)
        return bfail(h, EBPMF_ERR_ARG, "batched_set_y_csc: bad arguments");
    if (batch_row0[0] != 0 || batch_row0[nbatch] != n) return bfail(h, EBPMF_ERR_ARG, "batched_set_y_csc: batch_row0 must run from 0 to n");
    for (int64_t b = 0; b < nbatch; ++b)
        if (batch_row0[b + 1] <= batch_row0[b]) return bfail(h, EBPMF_ERR_ARG, "batched_set_y_csc: empty or unordered batch");
    const long long nnz = colptr[p];
    if (nnz > 0 && (!rowidx || !x)) return bfail(h, EBPMF_ERR_ARG, "batched_set_y_csc: bad arguments");
    release_batches(h);
    h->n = n; h->p = p;
    h->row0.assign(batch_row0, batch_row0 + nbatch + 1);
    std::vector<int32_t> batch_of((size_t)n);
    for (int64_t b = 0; b < nbatch; ++b)
        for (long long i = batch_row0[b]; i < batch_row0[b + 1]; ++i) batch_of[(size_t)i] = (int32_t)b;
    std::vector<std::vector<int32_t>> cp((size_t)nbatch, std::vector<int32_t>((size_t)p + 1, 0));
    for (long long j = 0; j < p; ++j)
        for (int q = colptr[j]; q < colptr[j + 1]; ++q) {
            if (rowidx[q] < 0 || rowidx[q] >= n) return bfail(h, EBPMF_ERR_ARG, "batched_set_y_csc: row index out of range");
            ++cp[(size_t)batch_of[(size_t)rowidx[q]]][(size_t)j + 1];
        }
    std::vector<std::vector<int32_t>> ri((size_t)nbatch);
    std::vector<std::vector<double>> xx((size_t)nbatch);
    for (int64_t b = 0; b < nbatch; ++b) {
        for (long long j = 0; j < p; ++j) cp[(size_t)b][(size_t)j + 1] += cp[(size_t)b][(size_t)j];
        ri[(size_t)b].resize((size_t)std::max(1, cp[(size_t)b][(size_t)p]));
        xx[(size_t)b].resize((size_t)std::max(1, cp[(size_t)b][(size_t)p]));
    }
    {
        std::vector<int32_t> fill((size_t)nbatch);
        for (long long j = 0; j < p; ++j) {
            for (int64_t b = 0; b < nbatch; ++b) fill[(size_t)b] = cp[(size_t)b][(size_t)j];
            for (int q = colptr[j]; q < colptr[j + 1]; ++q) {
                const int32_t b = batch_of[(size_t)rowidx[q]];
                const int32_t w = fill[(size_t)b]++;
                ri[(size_t)b][(size_t)w] = (int32_t)(rowidx[q] - batch_row0[b]);
                xx[(size_t)b][(size_t)w] = x[q];
            }
        }
    }
    ebpmf_ctx *c0 = h->slot[0];
    h->ys.resize((size_t)nbatch);
    for (int64_t b = 0; b < nbatch; ++b) {
        ebpmf_ctx *cb = h->slot[(size_t)b % h->slot.size()];       // upload on the device that will run the batch
        const long long nb = batch_row0[b + 1] - batch_row0[b];
        int rc = ebpmf_ctx_set_y_csc(cb, nb, p, cp[(size_t)b].data(), ri[(size_t)b].data(), xx[(size_t)b].data());
        if (rc != EBPMF_OK) return bfail(h, rc, cb->err);
        swap_y(cb, h->ys[(size_t)b]);           // the slab now belongs to the batch store
        cb->have_y = false;
        std::vector<int32_t>().swap(ri[(size_t)b]);
        std::vector<double>().swap(xx[(size_t)b]);
    }
    h->lfact = 0.0;
    if (nnz > 0) {
        // needs a context with scratch: borrow batch 0
        int rc = attach_batch(c0, h->ys[0], p);
        if (rc == EBPMF_OK) rc = ebpmf_sum_lfactorial(c0, nnz, x, &h->lfact);
        detach_batch(c0, h->ys[0]);
        if (rc != EBPMF_OK) return bfail(h, rc, c0->err);
    }
    return EBPMF_OK;
}

int ebpmf_batched_sum_lfactorial(ebpmf_batched *h, double *out) {
    if (!h || !out) return bfail(h, EBPMF_ERR_ARG, "batched_sum_lfactorial: bad arguments");
    if (h->n < 1) return bfail(h, EBPMF_ERR_STATE, "batched_sum_lfactorial: Y not set");
    *out = h->lfact;
    return EBPMF_OK;
}

int ebpmf_batched_vga_step_host(ebpmf_batched *h, int64_t n_in, int64_t p_in, const double *M_in, int64_t K,
                                const double *EL, const double *EF, int var_type, const double *sigma2, int maxiter,
                                double tol, double *M_out, double *V_out, double *v_sum, ebpmf_solve_info *info,
                                int32_t *n_updates, int32_t *converged) {
    if (!h || !M_in || !M_out || !EL || !EF || !sigma2) return bfail(h, EBPMF_ERR_ARG, "batched_vga_step_host: bad arguments");
    if (h->n < 1 || h->ys.empty()) return bfail(h, EBPMF_ERR_STATE, "batched_vga_step_host: Y not set");
    if (n_in != h->n || p_in != h->p) return bfail(h, EBPMF_ERR_ARG, "batched_vga_step_host: non-conformable arrays (M does not have the shape of the resident Y)");
    const int S = (int)h->slot.size();
    const long long n = h->n, p = h->p;
    const int nb = (int)h->ys.size();
    std::vector<ebpmf_solve_info> infos((size_t)nb);
    std::vector<double> vcol;                                  // per-batch colSums(V_b) / sum(V_b), added in batch order
    const size_t vlen = (var_type == EBPMF_VAR_BY_COL) ? (size_t)p : 1;
    if (v_sum && var_type != EBPMF_VAR_BY_ROW) vcol.assign((size_t)nb * vlen, 0.0);
    std::vector<int> rc((size_t)S, EBPMF_OK);
    std::vector<int> failed_batch((size_t)S, -1);
    auto work = [&](int s) -> int {
        ebpmf_ctx *c = h->slot[(size_t)s];
        for (int b = s; b < nb; b += S) {
            const long long lo = h->row0[(size_t)b];
            YState &y = h->ys[(size_t)b];
            int r = attach_batch(c, y, p);
            if (r == EBPMF_OK) r = set_factors_ld(c, K, EL + lo, n, EF);                                   // EF[[1]][b,], :266
            if (r == EBPMF_OK) r = ebpmf_ctx_set_sigma2(c, var_type, var_type == EBPMF_VAR_BY_ROW ? sigma2 + lo : sigma2);   // :267
            if (r == EBPMF_OK) r = set_m_ld(c, M_in + lo, n);                                              // flash_fit$Y[b,], :263
            if (r == EBPMF_OK) r = ebpmf_ctx_vga_step(c, maxiter, tol, V_out != nullptr, &infos[(size_t)b]);
            if (r == EBPMF_OK) r = get_matrix_ld(c, false, M_out + lo, n);                                 // :270
            if (r == EBPMF_OK && V_out) r = get_matrix_ld(c, true, V_out + lo, n);
            if (r == EBPMF_OK && v_sum)
                r = ebpmf_ctx_get_v_sum(c, var_type == EBPMF_VAR_BY_ROW ? v_sum + lo : vcol.data() + (size_t)b * vlen);
            detach_batch(c, y);
            if (r != EBPMF_OK) { failed_batch[(size_t)s] = b; return r; }
        }
        return EBPMF_OK;
    };
    {
        std::vector<std::thread> th;
        for (int s = 1; s < S; ++s) th.emplace_back([&, s] { rc[(size_t)s] = work(s); });
        rc[0] = work(0);
        for (auto &t : th) t.join();
    }
    for (int s = 0; s < S; ++s)
        if (rc[(size_t)s] != EBPMF_OK)
            return bfail(h, rc[(size_t)s], std::string("batch ") + std::to_string(failed_batch[(size_t)s]) + ": " + h->slot[(size_t)s]->err);
    if (v_sum && var_type != EBPMF_VAR_BY_ROW) {
        for (size_t q = 0; q < vlen; ++q) v_sum[q] = 0.0;                                                  // v_sum = 0, :255
        for (int b = 0; b < nb; ++b)
            for (size_t q = 0; q < vlen; ++q) v_sum[q] += vcol[(size_t)b * vlen + q];                      // :284, :286
    }
    if (info) {
        memset(info, 0, sizeof *info);
        info->converged = 1;
        for (int b = 0; b < nb; ++b) {
            const ebpmf_solve_info &ib = infos[(size_t)b];
            info->sym += ib.sym; info->ssexp += ib.ssexp; info->slogv += ib.slogv;                         // :274-276
            info->n_updates = std::max(info->n_updates, ib.n_updates);
            info->converged = info->converged && ib.converged;
            info->passes = std::max(info->passes, ib.passes);
            info->kernel_ms += ib.kernel_ms; info->pass1_ms += ib.pass1_ms;
            info->pass2_ms += ib.pass2_ms; info->reduce_ms += ib.reduce_ms;
        }
    }
    for (int b = 0; b < nb; ++b) {
        if (n_updates) n_updates[b] = infos[(size_t)b].n_updates;
        if (converged) converged[b] = infos[(size_t)b].converged;
    }
    return EBPMF_OK;
}

}  // extern "C"
