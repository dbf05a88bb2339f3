// C-ABI implementation of include/ebpmf_b200.h: context, device buffers, pass orchestration
// of the global-stop Newton solve, reductions and the NCCL combine.  No PyTorch, no CPU
// fallback: every entry point fails with EBPMF_ERR_CUDA when no sm_100 GPU is usable.
#include "../../include/ebpmf_b200.h"
#include "vga_kernels.cuh"
#include "simgen.cuh"

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
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_copy = nullptr, ev_p1 = nullptr, ev_p2 = nullptr;
    std::string err;

    // shapes
    long long n = 0, p = 0, nnz = 0;
    long long nRB = 0, nRW = 0;
    int K = 0, KTP = 0, ldk = 0;
    int var_type = -1;
    int pipeline = 0;          // 0: fused first pass, 1: split (const0 kernel + streaming Newton), 2: ordered split,
                               // 3: scored (FP32 scoring pass, then the fused first pass in score order)
    double delta = 0.0;        // 0: exact update count (default); > 0: delta mode (ebpmf_ctx_set_option)
    unsigned scored_min_ctas = 2 * 148 * 3;   // scored pipeline: below two waves of CTAs the order cannot matter
    double debug_verify_scale = 1.0;   // test hook: pass 2 verifies against tol*scale (exercises the retry loop)
    bool can_rewind = false;
    unsigned long long diag_topups = 0, diag_iters = 0;
    bool have_y = false, have_m = false, have_f = false, have_s2 = false, have_v = false, have_vsum = false;

    // resident buffers
    DevBuf colptr, rowidx, yx, rbptr;
    DevBuf Mcur, Malt, V;
    DevBuf EL, EF, EFt, sigma2, l0, f0;
    DevBuf C0;         // const0 of the last first pass (n x p)
    DevBuf sc1, sc2;   // 1/sigma2, sigma2/2
    DevBuf kunit, colV, unitS, rowpart;
    DevBuf red;        // [4 scalars: sym, ssexp, slogv, sumV][v_sum ...]
    DevBuf colsums;    // colsym[p] | v_sum scratch[p] | unitsum[nJG][2]
    DevBuf status;     // Status
    DevBuf tables;     // MathTables
    DevBuf score, score_sorted, cta_ids, perm, cub_tmp;   // ordered pipeline (pipeline == 2)
    DevBuf misc;       // small scratch (R2, fitmax, obj terms, lfactorial partials)
    DevBuf scratch[6]; // dense-signature operands

    Status *h_status = nullptr;   // pinned
    double *h_small = nullptr;    // pinned, 64 doubles

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

template <int KTP, int FORMULA, bool DELTA, bool PERM = false>
int launch_first_ktp2(ebpmf_ctx *ctx, const FusedArgs &a, dim3 grid) {
    auto kern = vga_fused_kernel<KTP, kE, FORMULA, DELTA, false, PERM>;
    const int smem = (int)sizeof(FusedSmem<KTP>);
    static bool configured[8] = {false, false, false, false, false, false, false, false};   // per device
    if (ctx->device < 8 ? !configured[ctx->device] : true) {
        CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        if (ctx->device < 8) configured[ctx->device] = true;
    }
    kern<<<grid, dim3(kThreads), smem, ctx->stream>>>(a);
    CU(cudaGetLastError());
    return EBPMF_OK;
}

template <int KTP, int FORMULA>
int launch_first_ktp(ebpmf_ctx *ctx, const FusedArgs &a, dim3 grid) {
    if (FORMULA == FORMULA_A1 && a.perm)      // scored pipeline: tiles in score order
        return a.delta > 0 ? launch_first_ktp2<KTP, FORMULA_A1, true, true>(ctx, a, grid)
                           : launch_first_ktp2<KTP, FORMULA_A1, false, true>(ctx, a, grid);
    return a.delta > 0 ? launch_first_ktp2<KTP, FORMULA, true>(ctx, a, grid)
                       : launch_first_ktp2<KTP, FORMULA, false>(ctx, a, grid);
}

template <int KTP, int FORMULA>
int launch_prep_ktp(ebpmf_ctx *ctx, const FusedArgs &a, dim3 grid) {
    auto kern = vga_prep_kernel<KTP, FORMULA>;
    const int smem = (int)sizeof(PrepSmem<KTP>);
    kern<<<grid, dim3(kThreads), smem, ctx->stream>>>(a);
    CU(cudaGetLastError());
    return EBPMF_OK;
}

template <int FORMULA>
int launch_prep(ebpmf_ctx *ctx, const FusedArgs &a, dim3 grid) {
    switch (ctx->KTP) {
        case 8: return launch_prep_ktp<8, FORMULA>(ctx, a, grid);
        case 16: return launch_prep_ktp<16, FORMULA>(ctx, a, grid);
        case 24: return launch_prep_ktp<24, FORMULA>(ctx, a, grid);
        case 32: return launch_prep_ktp<32, FORMULA>(ctx, a, grid);
    }
    return fail(ctx, EBPMF_ERR_ARG, "K = %d factor columns not supported", ctx->K);
}

template <int KTP>
int launch_iter0_ktp(ebpmf_ctx *ctx, const FusedArgs &a, dim3 grid) {
    auto kern = vga_fused_kernel<KTP, kE, FORMULA_A1, false, true>;
    const int smem = (int)sizeof(FusedSmem<KTP>);
    static bool configured[8] = {false, false, false, false, false, false, false, false};   // per device
    if (ctx->device < 8 ? !configured[ctx->device] : true) {
        CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        if (ctx->device < 8) configured[ctx->device] = true;
    }
    kern<<<grid, dim3(kThreads), smem, ctx->stream>>>(a);
    CU(cudaGetLastError());
    return EBPMF_OK;
}

__global__ void iota_kernel(int *ids, int count) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < count) ids[t] = t;
}

int reserve_order_buffers(ebpmf_ctx *ctx, int ncta) {
    CHECK(reserve(ctx, ctx->score, sizeof(float) * (size_t)ncta));
    CHECK(reserve(ctx, ctx->score_sorted, sizeof(float) * (size_t)ncta));
    CHECK(reserve(ctx, ctx->cta_ids, sizeof(int) * (size_t)ncta));
    CHECK(reserve(ctx, ctx->perm, sizeof(int) * (size_t)ncta));
    return EBPMF_OK;
}

// ctx->perm = CTA ids by decreasing ctx->score (cub radix sort, ~60 us for 57 k CTAs)
int sort_ctas_by_score(ebpmf_ctx *ctx, int ncta) {
    iota_kernel<<<(ncta + 255) / 256, 256, 0, ctx->stream>>>(as<int>(ctx->cta_ids), ncta);
    CU(cudaGetLastError());
    size_t tmp_bytes = 0;
    CU(cub::DeviceRadixSort::SortPairsDescending(nullptr, tmp_bytes, as<float>(ctx->score), as<float>(ctx->score_sorted),
                                                 as<int>(ctx->cta_ids), as<int>(ctx->perm), ncta, 0, 32, ctx->stream));
    CHECK(reserve(ctx, ctx->cub_tmp, tmp_bytes));
    CU(cub::DeviceRadixSort::SortPairsDescending(ctx->cub_tmp.ptr, tmp_bytes, as<float>(ctx->score),
                                                 as<float>(ctx->score_sorted), as<int>(ctx->cta_ids), as<int>(ctx->perm),
                                                 ncta, 0, 32, ctx->stream));
    return EBPMF_OK;
}

template <int KTP>
int launch_score_ktp(ebpmf_ctx *ctx, const FusedArgs &a, dim3 grid) {
    const int smem = (int)sizeof(ScoreSmem<KTP>);
    static bool configured[8] = {false, false, false, false, false, false, false, false};   // per device
    if (ctx->device < 8 ? !configured[ctx->device] : true) {
        CU(cudaFuncSetAttribute(vga_score_kernel<KTP, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        CU(cudaFuncSetAttribute(vga_score_kernel<KTP, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        if (ctx->device < 8) configured[ctx->device] = true;
    }
    if (a.var_type == EBPMF_VAR_BY_COL) vga_score_kernel<KTP, true><<<grid, dim3(kThreads), smem, ctx->stream>>>(a);
    else vga_score_kernel<KTP, false><<<grid, dim3(kThreads), smem, ctx->stream>>>(a);
    CU(cudaGetLastError());
    return EBPMF_OK;
}

// Scored pipeline (pipeline == 3, FORMULA_A1): an FP32 scoring pass ranks the tiles by their largest
// first Newton step, then the SAME fused kernel takes them in that order.  The slowest entries are
// met first, K* is published early and (almost) every unit runs to the global count inside pass 1,
// in registers; the top-up pass finds next to nothing to do.  Results are bit-identical to
// pipeline 0 (a unit's iterates do not depend on the order the units are taken in).
int launch_score_and_sort(ebpmf_ctx *ctx, FusedArgs &a, dim3 grid) {
    const int ncta = (int)(grid.x * grid.y);
    CHECK(reserve_order_buffers(ctx, ncta));
    a.score = as<float>(ctx->score);
    a.perm = nullptr;
    switch (ctx->KTP) {
        case 8: CHECK((launch_score_ktp<8>(ctx, a, grid))); break;
        case 16: CHECK((launch_score_ktp<16>(ctx, a, grid))); break;
        case 24: CHECK((launch_score_ktp<24>(ctx, a, grid))); break;
        case 32: CHECK((launch_score_ktp<32>(ctx, a, grid))); break;
        default: return fail(ctx, EBPMF_ERR_ARG, "K = %d factor columns not supported", ctx->K);
    }
    CHECK(sort_ctas_by_score(ctx, ncta));
    a.perm = as<int>(ctx->perm);
    return EBPMF_OK;
}

// Ordered pipeline (pipeline == 2, FORMULA_A1, maxiter >= 2): stage A (the fused kernel's ITER0_ONLY
// form) performs Newton iteration 0 and scores every CTA by its largest first step; the streaming kernel
// then walks the CTAs by decreasing score, so the slowest entries -- the ones that set the global update count --
// are met first and almost no unit finishes below K* (the top-up pass becomes nearly empty).
template <bool DELTA>
int launch_first_ordered(ebpmf_ctx *ctx, FusedArgs a, dim3 grid) {
    const int ncta = (int)(grid.x * grid.y);
    CHECK(reserve_order_buffers(ctx, ncta));
    a.score = as<float>(ctx->score);
    a.perm = nullptr;
    a.nspans = (int)grid.x;
    switch (ctx->KTP) {   // stage A: tile staging + Beta + Newton iteration 0 (the fused kernel's ITER0_ONLY form)
        case 8: CHECK((launch_iter0_ktp<8>(ctx, a, grid))); break;
        case 16: CHECK((launch_iter0_ktp<16>(ctx, a, grid))); break;
        case 24: CHECK((launch_iter0_ktp<24>(ctx, a, grid))); break;
        case 32: CHECK((launch_iter0_ktp<32>(ctx, a, grid))); break;
        default: return fail(ctx, EBPMF_ERR_ARG, "K = %d factor columns not supported", ctx->K);
    }
    if (ctx->comm)    // "some entry is not converged at k = 0" is a property of the WHOLE matrix
        NC(g_nccl.AllReduce(&as<Status>(ctx->status)->notconv0, &as<Status>(ctx->status)->notconv0, 1, ncclInt32,
                            ncclMax, ctx->comm, ctx->stream));
    CHECK(sort_ctas_by_score(ctx, ncta));
    a.perm = as<int>(ctx->perm);
    vga_stream_kernel<kE, FORMULA_A1, MODE_FIRST, DELTA><<<dim3((unsigned)ncta), dim3(kThreads), 0, ctx->stream>>>(a);
    CU(cudaGetLastError());
    return EBPMF_OK;
}

// pass 1 of the fused path (K1 / K4): either ONE fused kernel (tile staging + Beta rebuild +
// register-resident Newton), or the split pipeline: const0 kernel, then streaming Newton.
template <int FORMULA>
int launch_first(ebpmf_ctx *ctx, const FusedArgs &a) {
    const long long ncols = a.col1 - a.col0;
    if (ncols <= 0) return EBPMF_OK;
    dim3 grid((unsigned)((ncols + a.cols_per_cta - 1) / a.cols_per_cta), (unsigned)ctx->nRB);
    if (ctx->pipeline == 2 && FORMULA == FORMULA_A1 && a.maxiter >= 2 && a.col0 == 0 && a.col1 == ctx->p)
        return a.delta > 0 ? launch_first_ordered<true>(ctx, a, grid) : launch_first_ordered<false>(ctx, a, grid);
    if (ctx->pipeline == 1 || ctx->pipeline == 2) {
        CHECK((launch_prep<FORMULA>(ctx, a, grid)));
        if (a.delta > 0) vga_stream_kernel<kE, FORMULA, MODE_FIRST, true><<<grid, dim3(kThreads), 0, ctx->stream>>>(a);
        else vga_stream_kernel<kE, FORMULA, MODE_FIRST, false><<<grid, dim3(kThreads), 0, ctx->stream>>>(a);
        CU(cudaGetLastError());
        return EBPMF_OK;
    }
    FusedArgs b = a;
    b.perm = nullptr;
    if (ctx->pipeline == 3 && FORMULA == FORMULA_A1 && a.maxiter >= 2 && ctx->K <= 32 && a.cols_per_cta <= kMaxSpanCols &&
        grid.x * grid.y >= ctx->scored_min_ctas)
        CHECK(launch_score_and_sort(ctx, b, grid));
    switch (ctx->KTP) {
        case 8: return launch_first_ktp<8, FORMULA>(ctx, b, grid);
        case 16: return launch_first_ktp<16, FORMULA>(ctx, b, grid);
        case 24: return launch_first_ktp<24, FORMULA>(ctx, b, grid);
        case 32: return launch_first_ktp<32, FORMULA>(ctx, b, grid);
    }
    return fail(ctx, EBPMF_ERR_ARG, "K = %d factor columns not supported", ctx->K);
}

// pass 2 (streaming top-up to the global count)
template <int FORMULA>
int launch_topup(ebpmf_ctx *ctx, const FusedArgs &a) {
    const long long ncols = a.col1 - a.col0;
    if (ncols <= 0) return EBPMF_OK;
    dim3 grid((unsigned)((ncols + a.cols_per_cta - 1) / a.cols_per_cta), (unsigned)ctx->nRB);
    if (a.delta > 0) vga_stream_kernel<kE, FORMULA, MODE_TOPUP, true><<<grid, dim3(kThreads), 0, ctx->stream>>>(a);
    else vga_stream_kernel<kE, FORMULA, MODE_TOPUP, false><<<grid, dim3(kThreads), 0, ctx->stream>>>(a);
    CU(cudaGetLastError());
    return EBPMF_OK;
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

int alloc_solver_state(ebpmf_ctx *ctx, long long n, long long p) {
    ctx->nRB = (n + kRowsPerCta - 1) / kRowsPerCta;
    ctx->nRW = (n + 31) / 32;
    const long long nJG = (p + kE - 1) / kE;
    CHECK(reserve(ctx, ctx->kunit, sizeof(unsigned short) * (size_t)nJG * (size_t)ctx->nRW));
    CHECK(reserve(ctx, ctx->status, sizeof(Status)));
    return EBPMF_OK;
}

// reductions of the per-unit partials into red = [sym, ssexp, slogv, sumV][v_sum...]
// (the final M is in Malt at this point: sum(Y*M) runs over the CSC non-zeros only)
int run_reductions(ebpmf_ctx *ctx) {
    const long long n = ctx->n, p = ctx->p;
    const long long nJG = (p + kE - 1) / kE;
    double *red = as<double>(ctx->red);
    double *colsym = as<double>(ctx->colsums);
    double *colv = (ctx->var_type == EBPMF_VAR_BY_COL) ? red + 4 : colsym + p;
    double *unitsum = colsym + 2 * p;
    reduce_over_rows_kernel<1><<<(unsigned)((p + 31) / 32), dim3(32, 8), 0, ctx->stream>>>(
        as<double>(ctx->colV), ctx->nRW, p, colv);
    CU(cudaGetLastError());
    reduce_over_rows_kernel<2><<<(unsigned)((nJG + 31) / 32), dim3(32, 8), 0, ctx->stream>>>(
        as<double>(ctx->unitS), ctx->nRW, nJG, unitsum);
    CU(cudaGetLastError());
    sym_nnz_kernel<<<(unsigned)((p * 32 + 255) / 256), 256, 0, ctx->stream>>>(
        as<int>(ctx->colptr), as<int>(ctx->rowidx), as<double>(ctx->yx), as<double>(ctx->Malt), n, p, colsym);
    CU(cudaGetLastError());
    reduce_scalars_kernel<<<1, 1024, 0, ctx->stream>>>(colsym, colv, p, unitsum, nJG, red);
    CU(cudaGetLastError());
    if (ctx->var_type == EBPMF_VAR_BY_ROW) {
        reduce_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(as<double>(ctx->rowpart), n, nJG,
                                                                                 red + 4);
        CU(cudaGetLastError());
    }
    if (ctx->comm) {
        const size_t count = 4 + (ctx->var_type == EBPMF_VAR_BY_COL ? (size_t)p : 0);
        NC(g_nccl.AllReduce(red, red, count, ncclFloat64, ncclSum, ctx->comm, ctx->stream));
    }
    return EBPMF_OK;
}

// The pass orchestration shared by the fused step (A1) and low_memory (A10).
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
        if (info) {
            info->n_updates = s.kstar;
            info->converged = (s.kstar < maxiter) ? 1 : 0;
            info->passes = passes;
            info->reserved = 0;
            ctx->diag_topups = s.topup_units;
            ctx->diag_iters = s.unit_iters;
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

FusedArgs fused_args(ebpmf_ctx *ctx, int maxiter, double tol, bool want_v) {
    FusedArgs a;
    memset(&a, 0, sizeof a);
    a.n = ctx->n; a.p = ctx->p; a.K = ctx->K; a.var_type = ctx->var_type;
    a.M_in = as<double>(ctx->Mcur); a.M_out = as<double>(ctx->Malt);
    a.V_out = want_v ? as<double>(ctx->V) : nullptr;
    a.C0 = as<double>(ctx->C0);
    a.EL = as<double>(ctx->EL); a.EFt = as<double>(ctx->EFt); a.ldk = ctx->ldk; a.sigma2 = as<double>(ctx->sigma2);
    a.sc1 = as<double>(ctx->sc1); a.sc2 = as<double>(ctx->sc2);
    a.l0 = as<double>(ctx->l0); a.f0 = as<double>(ctx->f0);
    a.rowidx = as<int>(ctx->rowidx); a.yx = as<double>(ctx->yx); a.rbptr = as<int>(ctx->rbptr);
    a.kunit = as<unsigned short>(ctx->kunit); a.nRW = ctx->nRW; a.nJG = (ctx->p + kE - 1) / kE;
    a.colV = as<double>(ctx->colV); a.unitS = as<double>(ctx->unitS);
    a.rowpart = (ctx->var_type == EBPMF_VAR_BY_ROW) ? as<double>(ctx->rowpart) : nullptr;
    a.status = as<Status>(ctx->status);
    a.peers = reinterpret_cast<const Status *const *>(ctx->peer_ptrs.ptr);
    a.npeers = ctx->npeers;
    a.epoch = ctx->epoch + 1;      // run_global_stop increments the counter before the first launch
    a.tables = as<MathTables>(ctx->tables);
    a.maxiter = maxiter; a.tol = tol; a.delta = ctx->delta; a.tol_verify = tol * ctx->debug_verify_scale;
    a.col0 = 0; a.col1 = ctx->p;
    a.cols_per_cta = cols_per_cta_for(ctx->p, ctx->nRB);
    return a;
}

int upload_csc(ebpmf_ctx *ctx, long long n, long long p, const int32_t *colptr, const int32_t *rowidx,
               const double *x) {
    const long long nnz = colptr[p];
    if (nnz < 0 || colptr[0] != 0) return fail(ctx, EBPMF_ERR_ARG, "bad dgCMatrix @p slot");
    CHECK(reserve(ctx, ctx->colptr, sizeof(int) * (size_t)(p + 1)));
    CHECK(reserve(ctx, ctx->rowidx, sizeof(int) * (size_t)std::max(1LL, nnz)));
    CHECK(reserve(ctx, ctx->yx, sizeof(double) * (size_t)std::max(1LL, nnz)));
    CU(cudaMemcpyAsync(ctx->colptr.ptr, colptr, sizeof(int) * (size_t)(p + 1), cudaMemcpyHostToDevice, ctx->stream));
    if (nnz > 0) {
        CU(cudaMemcpyAsync(ctx->rowidx.ptr, rowidx, sizeof(int) * (size_t)nnz, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(ctx->yx.ptr, x, sizeof(double) * (size_t)nnz, cudaMemcpyHostToDevice, ctx->stream));
    }
    ctx->nnz = nnz;
    (void)n;
    return EBPMF_OK;
}

// after colptr/rowidx/yx are resident: shapes, row-block offset table, partial buffers
// shape-dependent scratch of the fused step (grow-only), and nRB / nRW
int reserve_shape_buffers(ebpmf_ctx *ctx, long long n, long long p) {
    ctx->n = n; ctx->p = p;
    CHECK(alloc_solver_state(ctx, n, p));
    const long long nJG = (p + kE - 1) / kE;
    CHECK(reserve(ctx, ctx->colV, sizeof(double) * (size_t)ctx->nRW * (size_t)p));
    CHECK(reserve(ctx, ctx->unitS, sizeof(double) * 2 * (size_t)ctx->nRW * (size_t)nJG));
    CHECK(reserve(ctx, ctx->colsums, sizeof(double) * (2 * (size_t)p + 2 * (size_t)nJG)));
    CHECK(reserve(ctx, ctx->sc1, sizeof(double) * (size_t)std::max(n, p)));
    CHECK(reserve(ctx, ctx->sc2, sizeof(double) * (size_t)std::max(n, p)));
    CHECK(reserve(ctx, ctx->red, sizeof(double) * (4 + (size_t)std::max(n, p))));
    CHECK(reserve(ctx, ctx->misc, sizeof(double) * (size_t)(3 * std::max(n, p) + 4096)));
    return EBPMF_OK;
}

int finish_set_y(ebpmf_ctx *ctx, long long n, long long p) {
    ctx->n = n; ctx->p = p;
    CHECK(alloc_solver_state(ctx, n, p));
    const long long cells = (ctx->nRB + 1) * p;
    CHECK(reserve(ctx, ctx->rbptr, sizeof(int) * (size_t)cells));
    build_rbptr_kernel<<<(unsigned)((cells + 255) / 256), 256, 0, ctx->stream>>>(
        as<int>(ctx->colptr), as<int>(ctx->rowidx), p, ctx->nRB, as<int>(ctx->rbptr));
    CU(cudaGetLastError());
    const long long nJG = (p + kE - 1) / kE;
    CHECK(reserve(ctx, ctx->colV, sizeof(double) * (size_t)ctx->nRW * (size_t)p));
    CHECK(reserve(ctx, ctx->unitS, sizeof(double) * 2 * (size_t)ctx->nRW * (size_t)nJG));
    CHECK(reserve(ctx, ctx->colsums, sizeof(double) * (2 * (size_t)p + 2 * (size_t)nJG)));
    CHECK(reserve(ctx, ctx->sc1, sizeof(double) * (size_t)std::max(n, p)));
    CHECK(reserve(ctx, ctx->sc2, sizeof(double) * (size_t)std::max(n, p)));
    CHECK(reserve(ctx, ctx->red, sizeof(double) * (4 + (size_t)std::max(n, p))));
    CHECK(reserve(ctx, ctx->misc, sizeof(double) * (size_t)(3 * std::max(n, p) + 4096)));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->have_y = true;
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

int set_m_ld(ebpmf_ctx *ctx, const double *M, long long ld) {
    if (!ctx || !M) return fail(ctx, EBPMF_ERR_ARG, "set_m: bad arguments");
    if (!ctx->have_y) return fail(ctx, EBPMF_ERR_STATE, "set_m: Y not set");
    CHECK(ensure_common(ctx));
    const size_t bytes = sizeof(double) * (size_t)ctx->n * (size_t)ctx->p;
    CHECK(reserve(ctx, ctx->Mcur, bytes));
    CHECK(reserve(ctx, ctx->Malt, bytes));
    CHECK(copy_in_ld(ctx, ctx->Mcur.ptr, M, ctx->n, ctx->p, ld));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->have_m = true; ctx->can_rewind = false;
    return EBPMF_OK;
}

int set_factors_ld(ebpmf_ctx *ctx, int64_t K, const double *EL, long long ldl, const double *EF) {
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
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->K = (int)K; ctx->KTP = ktp; ctx->ldk = ldk; ctx->have_f = true;
    return EBPMF_OK;
}

int get_matrix_ld(ebpmf_ctx *ctx, bool want_v, double *out, long long ld) {
    if (!ctx || !out) return fail(ctx, EBPMF_ERR_ARG, "get: bad arguments");
    if (want_v ? !ctx->have_v : !ctx->have_m) return fail(ctx, EBPMF_ERR_STATE, want_v ? "get_v: the last step did not materialise V (want_v = 0)" : "get_m: M not set");
    CHECK(ensure_common(ctx));
    CHECK(copy_out_ld(ctx, out, want_v ? ctx->V.ptr : ctx->Mcur.ptr, ctx->n, ctx->p, ld));
    CU(cudaStreamSynchronize(ctx->stream));
    return EBPMF_OK;
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
    if (const char *pl = getenv("EBPMF_B200_PIPELINE"))
        ctx->pipeline = !strcmp(pl, "scored") ? 3 : !strcmp(pl, "ordered") ? 2 : !strcmp(pl, "split") ? 1 : 0;
    if (const char *dl = getenv("EBPMF_B200_DELTA")) ctx->delta = atof(dl);
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
    CREATE_STEP(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    CREATE_STEP(cudaEventCreate(&ctx->ev0));
    CREATE_STEP(cudaEventCreate(&ctx->ev1));
    CREATE_STEP(cudaEventCreate(&ctx->ev_p1));
    CREATE_STEP(cudaEventCreate(&ctx->ev_p2));
    CREATE_STEP(cudaEventCreateWithFlags(&ctx->ev_copy, cudaEventDisableTiming));
    CREATE_STEP(cudaMallocHost((void **)&ctx->h_status, sizeof(Status)));
    CREATE_STEP(cudaMallocHost((void **)&ctx->h_small, sizeof(double) * 64));
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
    DevBuf *bufs[] = {&ctx->colptr, &ctx->rowidx, &ctx->yx, &ctx->rbptr, &ctx->Mcur, &ctx->Malt, &ctx->V,
                      &ctx->EL, &ctx->EF, &ctx->EFt, &ctx->sigma2, &ctx->l0, &ctx->f0, &ctx->kunit,
                      &ctx->colV, &ctx->unitS, &ctx->rowpart, &ctx->red, &ctx->colsums, &ctx->status, &ctx->misc,
                      &ctx->C0, &ctx->sc1, &ctx->sc2, &ctx->tables, &ctx->peer_ptrs,
                      &ctx->score, &ctx->score_sorted, &ctx->cta_ids, &ctx->perm, &ctx->cub_tmp};
    for (DevBuf *b : bufs) release(*b);
    for (DevBuf &b : ctx->scratch) release(b);
    if (ctx->h_status) cudaFreeHost(ctx->h_status);
    if (ctx->h_small) cudaFreeHost(ctx->h_small);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->ev_p1) cudaEventDestroy(ctx->ev_p1);
    if (ctx->ev_p2) cudaEventDestroy(ctx->ev_p2);
    if (ctx->ev_copy) cudaEventDestroy(ctx->ev_copy);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
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

int ebpmf_ctx_set_sigma2(ebpmf_ctx *ctx, int var_type, const double *sigma2) {
    if (!ctx || !sigma2) return fail(ctx, EBPMF_ERR_ARG, "set_sigma2: bad arguments");
    if (var_type < 0 || var_type > 2) return fail(ctx, EBPMF_ERR_VAR_TYPE, "Non-supported var type");
    if (!ctx->have_y) return fail(ctx, EBPMF_ERR_STATE, "set_sigma2: Y not set");
    CHECK(ensure_common(ctx));
    const size_t len = var_type == EBPMF_VAR_BY_COL ? (size_t)ctx->p : var_type == EBPMF_VAR_BY_ROW ? (size_t)ctx->n : 1;
    CHECK(reserve(ctx, ctx->sigma2, sizeof(double) * std::max((size_t)ctx->p, (size_t)ctx->n)));
    CU(cudaMemcpyAsync(ctx->sigma2.ptr, sigma2, sizeof(double) * len, cudaMemcpyHostToDevice, ctx->stream));
    if (var_type == EBPMF_VAR_BY_ROW) {
        const long long nJG = (ctx->p + kE - 1) / kE;
        CHECK(reserve(ctx, ctx->rowpart, sizeof(double) * (size_t)nJG * (size_t)ctx->n));
    }
    sigma_consts_kernel<<<(unsigned)((len + 255) / 256), 256, 0, ctx->stream>>>(as<double>(ctx->sigma2), (long long)len,
                                                                                as<double>(ctx->sc1), as<double>(ctx->sc2));
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->var_type = var_type; ctx->have_s2 = true;
    return EBPMF_OK;
}

int ebpmf_ctx_vga_step(ebpmf_ctx *ctx, int maxiter, double tol, int want_v, ebpmf_solve_info *info) {
    if (!ctx) return fail(ctx, EBPMF_ERR_ARG, "vga_step: ctx is NULL");
    if (maxiter < 1 || maxiter > 65535) return fail(ctx, EBPMF_ERR_ARG, "vga_step: maxiter must be in [1, 65535]");
    if (!ctx->have_y || !ctx->have_m || !ctx->have_f || !ctx->have_s2)
        return fail(ctx, EBPMF_ERR_STATE, "vga_step: Y, M, factors and sigma2 must be set first");
    CHECK(ensure_common(ctx));
    if (want_v) CHECK(reserve(ctx, ctx->V, sizeof(double) * (size_t)ctx->n * (size_t)ctx->p));
    CHECK(reserve(ctx, ctx->C0, sizeof(double) * (size_t)ctx->n * (size_t)ctx->p));
    const FusedArgs a = fused_args(ctx, maxiter, tol, want_v != 0);
    ctx->have_v = false; ctx->have_vsum = false;
    int rc = run_global_stop(
        ctx, maxiter, [&] { return launch_first<FORMULA_A1>(ctx, a); },
        [&] { return launch_topup<FORMULA_A1>(ctx, a); }, [&] { return run_reductions(ctx); }, info);
    if (rc != EBPMF_OK) return rc;
    CU(cudaMemcpyAsync(ctx->h_small, ctx->red.ptr, sizeof(double) * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (info) { info->sym = ctx->h_small[0]; info->ssexp = ctx->h_small[1]; info->slogv = ctx->h_small[2]; }
    std::swap(ctx->Mcur, ctx->Malt);      // the resident M is now res$M
    ctx->have_v = want_v != 0; ctx->have_vsum = true; ctx->can_rewind = true;
    return EBPMF_OK;
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

int ebpmf_ctx_get_v_sum(ebpmf_ctx *ctx, double *v_sum) {
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

int ebpmf_calc_obj(ebpmf_ctx *ctx, int64_t n, int64_t p, double sym, double ssexp, double slogv, int64_t len,
                   const double *sv, const double *sigma2, const double *R2, double KL_LF, double const_, double ss,
                   double a0, double b0, double *obj) {
    if (!ctx || !sv || !sigma2 || !R2 || !obj || len < 1) return fail(ctx, EBPMF_ERR_ARG, "calc_obj: bad arguments");
    CHECK(ensure_common(ctx));
    CHECK(reserve(ctx, ctx->scratch[0], sizeof(double) * (3 * (size_t)len + 8)));
    double *d = as<double>(ctx->scratch[0]);
    CU(cudaMemcpyAsync(d, sv, sizeof(double) * (size_t)len, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(d + len, sigma2, sizeof(double) * (size_t)len, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(d + 2 * len, R2, sizeof(double) * (size_t)len, cudaMemcpyHostToDevice, ctx->stream));
    obj_terms_kernel<<<1, 1024, 0, ctx->stream>>>(d, d + len, d + 2 * len, len, ss, a0, b0, d + 3 * len);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(ctx->h_small, d + 3 * len, sizeof(double) * 5, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    const double *t = ctx->h_small;
    // R/ebpmf_log.R:407, left to right
    *obj = sym - ssexp + slogv + 0.5 * (double)n * (double)p - t[0] - t[1] - t[2] + const_ + KL_LF - t[3] - t[4];
    return EBPMF_OK;
}

int ebpmf_vga_step_host(ebpmf_ctx *ctx, const double *M_in, int64_t K, const double *EL, const double *EF,
                        int var_type, const double *sigma2, int maxiter, double tol, double *M_out, double *V_out,
                        double *v_sum, ebpmf_solve_info *info) {
    if (!ctx || !M_in || !M_out) return fail(ctx, EBPMF_ERR_ARG, "vga_step_host: bad arguments");
    CHECK(ebpmf_ctx_set_factors(ctx, K, EL, EF));
    CHECK(ebpmf_ctx_set_sigma2(ctx, var_type, sigma2));
    CHECK(ebpmf_ctx_set_m(ctx, M_in));
    CHECK(ebpmf_ctx_vga_step(ctx, maxiter, tol, V_out != nullptr, info));
    CHECK(ebpmf_ctx_get_m(ctx, M_out));
    if (V_out) CHECK(ebpmf_ctx_get_v(ctx, V_out));
    if (v_sum) CHECK(ebpmf_ctx_get_v_sum(ctx, v_sum));
    return EBPMF_OK;
}

// The per-iteration call of ebpmf_log's outer loop once M lives on the device: R feeds res$M back
// unchanged (fit_flash$flash_fit$Y = res$M, R/ebpmf_log.R:309; flash_init(fit_flash$flash_fit$Y, ..),
// R/ebpmf_log_flash_update.R:35), so the step's input M IS the resident M left by the previous step
// and only the factors and sigma2 cross PCIe on the way in.
int ebpmf_vga_step_resident_host(ebpmf_ctx *ctx, int64_t K, const double *EL, const double *EF, int var_type,
                                 const double *sigma2, int maxiter, double tol, double *M_out, double *V_out,
                                 double *v_sum, ebpmf_solve_info *info) {
    if (!ctx || !M_out) return fail(ctx, EBPMF_ERR_ARG, "vga_step_resident_host: bad arguments");
    if (!ctx->have_m) return fail(ctx, EBPMF_ERR_STATE, "vga_step_resident_host: upload M once with ebpmf_ctx_set_m");
    CHECK(ebpmf_ctx_set_factors(ctx, K, EL, EF));
    CHECK(ebpmf_ctx_set_sigma2(ctx, var_type, sigma2));
    CHECK(ebpmf_ctx_vga_step(ctx, maxiter, tol, V_out != nullptr, info));
    CHECK(ebpmf_ctx_get_m(ctx, M_out));
    if (V_out) CHECK(ebpmf_ctx_get_v(ctx, V_out));
    if (v_sum) CHECK(ebpmf_ctx_get_v_sum(ctx, v_sum));
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
    if (maxiter < 1 || maxiter > 65535) return fail(ctx, EBPMF_ERR_ARG, "vga_newton: maxiter must be in [1, 65535]");
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
    CU(cudaMemcpyAsync(ctx->scratch[0].ptr, M, bytes, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(dS, S, sizeof(double) * lS, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(dB, Beta, sizeof(double) * lB, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(dV, Sigma2, sizeof(double) * lV, cudaMemcpyHostToDevice, ctx->stream));
    if (X) {
        CU(cudaMemcpyAsync(ctx->scratch[2].ptr, X, bytes, cudaMemcpyHostToDevice, ctx->stream));
    } else {
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
    CU(cudaMemcpyAsync(M_out, ctx->scratch[1].ptr, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (V_out) CU(cudaMemcpyAsync(V_out, ctx->scratch[3].ptr, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
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
    CU(cudaMemcpyAsync(ctx->scratch[0].ptr, M, bytes, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->scratch[2].ptr, Y, bytes, cudaMemcpyHostToDevice, ctx->stream));
    if (V) CU(cudaMemcpyAsync(ctx->scratch[5].ptr, V, bytes, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(dS, S, sizeof(double) * lS, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(dB, Beta, sizeof(double) * lB, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(dV, Sigma2, sizeof(double) * lV, cudaMemcpyHostToDevice, ctx->stream));
    a.n = n; a.p = p;
    a.M_in = as<double>(ctx->scratch[0]); a.V_in = V ? as<double>(ctx->scratch[5]) : nullptr;
    a.Y = as<double>(ctx->scratch[2]);
    a.S = dS; a.Beta = dB; a.Sigma2 = dV;
    a.M_out = as<double>(ctx->scratch[1]); a.V_out = as<double>(ctx->scratch[3]);
    a.tables = as<MathTables>(ctx->tables);
    a.maxiter = maxiter;
    const long long nRB = (n + kRowsPerCta - 1) / kRowsPerCta;
    dim3 grid((unsigned)std::min<long long>(p, 4096), (unsigned)nRB);
    vga_fixed_iter_kernel<<<grid, kThreads, 0, ctx->stream>>>(a);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(M_out, ctx->scratch[1].ptr, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(V_out, ctx->scratch[3].ptr, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return EBPMF_OK;
}

int ebpmf_vga_low_memory(ebpmf_ctx *ctx, int64_t n, int64_t p, int64_t K, const double *M, const double *Y,
                         const int32_t *Y_colptr, const int32_t *Y_rowidx, const double *Y_x, const double *l0,
                         const double *f0, const double *EL, const double *EF, const double *sigma2, int var_type,
                         int maxiter, double tol, double *M_out, ebpmf_solve_info *info) {
    if (!ctx || !M || !l0 || !f0 || !EL || !EF || !sigma2 || !M_out || n < 1 || p < 1 || K < 1)
        return fail(ctx, EBPMF_ERR_ARG, "vga_low_memory: bad arguments");
    if (maxiter < 1 || maxiter > 65535) return fail(ctx, EBPMF_ERR_ARG, "vga_low_memory: maxiter must be in [1, 65535]");
    // R falls through both `if(var_type==...)` blocks for an unknown var_type and returns the
    // clamped M (:78,:108); the clamp needs the kernels too, so run zero iterations of by_col
    // semantics is NOT the same thing -- report the unsupported type instead of guessing.
    if (var_type < 0 || var_type > 2) return fail(ctx, EBPMF_ERR_VAR_TYPE, "Non-supported var type");
    // this entry point owns the context's resident state for the duration of the call
    if (Y) CHECK(ebpmf_ctx_set_y_dense(ctx, n, p, Y));
    else CHECK(ebpmf_ctx_set_y_csc(ctx, n, p, Y_colptr, Y_rowidx, Y_x));
    CHECK(ebpmf_ctx_set_factors(ctx, K, EL, EF));
    CHECK(ebpmf_ctx_set_sigma2(ctx, var_type, sigma2));
    CHECK(ebpmf_ctx_set_m(ctx, M));
    CHECK(reserve(ctx, ctx->l0, sizeof(double) * (size_t)n));
    CHECK(reserve(ctx, ctx->f0, sizeof(double) * (size_t)p));
    CU(cudaMemcpyAsync(ctx->l0.ptr, l0, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->f0.ptr, f0, sizeof(double) * (size_t)p, cudaMemcpyHostToDevice, ctx->stream));
    CHECK(reserve(ctx, ctx->C0, sizeof(double) * (size_t)n * (size_t)p));
    const FusedArgs a = fused_args(ctx, maxiter, tol, false);
    int rc = run_global_stop(
        ctx, maxiter, [&] { return launch_first<FORMULA_A10>(ctx, a); },
        [&] { return launch_topup<FORMULA_A10>(ctx, a); }, [&] { return EBPMF_OK; }, info);
    if (rc != EBPMF_OK) return rc;
    if (info) { info->sym = info->ssexp = info->slogv = 0.0; }
    std::swap(ctx->Mcur, ctx->Malt);
    ctx->have_vsum = false; ctx->have_v = false;
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

int ebpmf_ctx_get_rows(ebpmf_ctx *ctx, int which, int64_t row0, int64_t nrows, double *out) {
    if (!ctx || !out || row0 < 0 || nrows < 1) return fail(ctx, EBPMF_ERR_ARG, "get_rows: bad arguments");
    if (!ctx->have_m || row0 + nrows > ctx->n) return fail(ctx, EBPMF_ERR_STATE, "get_rows: M not set or rows out of range");
    if (which != 0 && !ctx->have_v) return fail(ctx, EBPMF_ERR_STATE, "get_rows: V was not materialised");
    CHECK(ensure_common(ctx));
    const double *src = (which ? as<double>(ctx->V) : as<double>(ctx->Mcur)) + row0;
    CU(cudaMemcpy2DAsync(out, sizeof(double) * (size_t)nrows, src, sizeof(double) * (size_t)ctx->n,
                         sizeof(double) * (size_t)nrows, (size_t)ctx->p, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return EBPMF_OK;
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
    if (!strcmp(name, "pipeline")) {
        if (value != 0.0 && value != 1.0 && value != 2.0 && value != 3.0)
            return fail(ctx, EBPMF_ERR_ARG, "set_option: pipeline is 0, 1, 2 or 3");
        ctx->pipeline = (int)value;
        return EBPMF_OK;
    }
    if (!strcmp(name, "scored_min_ctas")) {
        if (!(value >= 0.0) || value > 1e9) return fail(ctx, EBPMF_ERR_ARG, "set_option: scored_min_ctas must be in [0, 1e9]");
        ctx->scored_min_ctas = (unsigned)value;
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
    const long long nblk = (n + kRowsPerCta - 1) / kRowsPerCta;
    if (nblk < N) return gfail(g, EBPMF_ERR_ARG, "group_set_y_csc: fewer 256-row blocks than GPUs; use a smaller group");
    g->row0.assign(N, 0); g->nrows.assign(N, 0);
    long long r0 = 0;
    for (int r = 0; r < N; ++r) {
        long long rows = (nblk / N + (r < nblk % N ? 1 : 0)) * kRowsPerCta;
        rows = std::max(0LL, std::min(rows, (long long)n - r0));
        g->row0[r] = r0; g->nrows[r] = rows; r0 += rows;
    }
    g->n = n; g->p = p;
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
int ebpmf_group_vga_step_host(ebpmf_group *g, const double *M_in, int64_t K, const double *EL, const double *EF,
                              int var_type, const double *sigma2, int maxiter, double tol, double *M_out,
                              double *V_out, double *v_sum, ebpmf_solve_info *info) {
    if (!g || !M_in || !M_out || !EL || !EF || !sigma2) return gfail(g, EBPMF_ERR_ARG, "group_vga_step_host: bad arguments");
    if (g->n < 1) return gfail(g, EBPMF_ERR_STATE, "group_vga_step_host: Y not set");
    const int N = (int)g->ctx.size();
    std::vector<ebpmf_solve_info> infos((size_t)N);
    const long long n = g->n;
    int rc = run_ranks(g, [&](int r) {
        ebpmf_ctx *c = g->ctx[r];
        const long long lo = g->row0[r];
        CHECK(set_factors_ld(c, K, EL + lo, n, EF));
        CHECK(ebpmf_ctx_set_sigma2(c, var_type, var_type == EBPMF_VAR_BY_ROW ? sigma2 + lo : sigma2));
        CHECK(set_m_ld(c, M_in + lo, n));
        CHECK(ebpmf_ctx_vga_step(c, maxiter, tol, V_out != nullptr, &infos[(size_t)r]));
        CHECK(get_matrix_ld(c, false, M_out + lo, n));
        if (V_out) CHECK(get_matrix_ld(c, true, V_out + lo, n));
        if (v_sum) {
            if (var_type == EBPMF_VAR_BY_ROW) CHECK(ebpmf_ctx_get_v_sum(c, v_sum + lo));
            else if (r == 0) CHECK(ebpmf_ctx_get_v_sum(c, v_sum));
        }
        return (int)EBPMF_OK;
    });
    if (rc != EBPMF_OK) return rc;
    if (info) {
        *info = infos[0];                       // sums and the update count are global after the combine
        for (int r = 1; r < N; ++r) info->kernel_ms = std::max(info->kernel_ms, infos[(size_t)r].kernel_ms);
    }
    return EBPMF_OK;
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

int ebpmf_batched_vga_step_host(ebpmf_batched *h, const double *M_in, int64_t K, const double *EL, const double *EF,
                                int var_type, const double *sigma2, int maxiter, double tol, double *M_out,
                                double *V_out, double *v_sum, ebpmf_solve_info *info, int32_t *n_updates,
                                int32_t *converged) {
    if (!h || !M_in || !M_out || !EL || !EF || !sigma2) return bfail(h, EBPMF_ERR_ARG, "batched_vga_step_host: bad arguments");
    if (h->n < 1 || h->ys.empty()) return bfail(h, EBPMF_ERR_STATE, "batched_vga_step_host: Y not set");
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
