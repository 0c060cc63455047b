// fdb_internal.h -- context / array objects behind the opaque handles of fdb200.h
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdarg.h>
#include <stdio.h>
#include <string>

#include "../../include/fdb200.h"
#include "fdb_mirror.cuh"

struct fdb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    bool async = false;
    bool nansafe = false;
    bool lane_sharing = true;   // see fdb_sweep.cu (SHARE)
    int grad_group = 1;         // chunks swept per block by the catalogue gradient kernel (fdb_set_chunk_grouping); 1 = independent sweeps
    bool use_dmma = true;       // tanh_affine: DMMA contraction (fdb_dmma.cu) vs CUDA-core kernel
    int dmma_tma_store = 1;     // 0 never, 1 when the result is mirrored to peers, 2 always -- fused DMMA epilogue: result blocks staged in shared memory and written by TMA tensor stores (fdb_set_dmma_tma_store)
    int dmma_bn = 0;            // column-tile width of the DMMA contraction: 0 = chosen per shape, or 16 / 32 / 64 (fdb_set_dmma_tile)
    int sm_count = 148;
    int64_t launches = 0;
    // reusable device scratch (block partials of reductions, staging of host entry points)
    double* scratch = nullptr;
    size_t scratch_bytes = 0;
    // pinned host staging for the *_host entry points
    unsigned* counter = nullptr;   // zero-initialised ticket for single-launch reductions
    int64_t* cursor = nullptr;     // device-resident chunk window {index, chunksize} (fdb_graph.cu)
    bool capturing = false;        // between fdb_graph_begin and fdb_graph_end
    bool saved_async = false;
    double* pinned = nullptr;
    size_t pinned_bytes = 0;
    // device staging of the *_host entry points that run drivers with their own scratch needs (fdb_multi.cu)
    double* hostws = nullptr;
    size_t hostws_bytes = 0;
    struct fdb_exchange* xch = nullptr;   // enabled peer-memory result exchange (fdb_exchange.cu), or null
    struct fdb_jit_cache* jit = nullptr;  // kernels specialised at run time for op-programs (fdb_jit.cu)
    bool use_jit = true;
    bool dmma_configured = false;         // cudaFuncSetAttribute(dmma_gemm_kernel, ...) done on THIS device (per device, not per process)
    struct fdb_multi* multi = nullptr;    // owning single-process multi-GPU group (fdb_multi.cu), or null
    // sticky failure of an asynchronous step (a peer never reached an exchange barrier): reported by the next
    // fdb_sync / synchronous call / fdb_exchange_* on this context; results produced since are invalid
    unsigned* xch_timeout_flag = nullptr; // pinned host word the barrier kernel bumps when it gives up
    bool poisoned = false;
    // plain Float64 arrays an op-program reads next to x (closed-over data of the user's f): fdb_program_bind_params
    const double* prm_ptr[FDB_MAX_PARAMS] = {nullptr, nullptr, nullptr, nullptr};
    int64_t prm_len[FDB_MAX_PARAMS] = {0, 0, 0, 0};
    int n_prm = 0;
    std::string err;
};

struct fdb_array {
    fdb_ctx* ctx = nullptr;
    double* data = nullptr;
    int64_t n = 0;     // elements
    int64_t ld = 0;    // plane stride in doubles
    int N = 0;         // chunk size (0 for level 0)
    int level = 0;     // 0 plain, 1 Dual, 2 nested Dual
    bool owned = true;
    int planes() const { return level == 0 ? 1 : (level == 1 ? N + 1 : (N + 1) * (N + 1)); }
};

// symmetric result arena of one rank (fdb_exchange.cu)
struct fdb_exchange {
    fdb_ctx* ctx = nullptr;
    char* base = nullptr;          // local arena: data | flag page
    size_t data_bytes = 0;
    int rank = 0, world = 1;
    char* peer[FDB_MAX_RANKS] = {nullptr};   // peer[r] = rank r's arena mapped here (peer[rank] = base)
    bool connected = false;
    unsigned long long epoch = 0;
    unsigned long long timeout_ns = 20ull * 1000000000ull;   // a barrier gives up after this long (fdb_exchange_set_timeout)
    bool in_process = false;       // peers linked with cudaDeviceEnablePeerAccess inside one process (fdb_exchange_connect_local)
    // NVSwitch multicast (fdb_multicast.cu): the arena is a cuMemCreate allocation shared as a POSIX file descriptor, bound
    // into a multicast object together with every peer's arena; a store to mc_base + off lands at off in EVERY rank's arena
    bool vmm = false;                       // arena (and the peer mappings) are virtual-memory-management mappings, not cudaMalloc / IPC
    unsigned long long mem_handle = 0;      // CUmemGenericAllocationHandle of the local arena
    size_t map_bytes = 0;                   // mapped size (data | flag page, rounded up to the multicast granularity)
    unsigned long long mc_handle = 0;       // CUmemGenericAllocationHandle of the multicast object (0 = none)
    char* mc_base = nullptr;                // multicast mapping of all arenas (null until fdb_exchange_multicast_bind)
    bool mc_on = true;                      // result stores go through mc_base when it is mapped (fdb_exchange_set_multicast)
    int exported_fd[2] = {-1, -1};          // descriptors handed out by fdb_exchange_share_handle / fdb_exchange_multicast_create
};

namespace fdb {

int multicast_destroy(fdb_exchange* x);     // fdb_multicast.cu: unmap / release a virtual-memory arena and its multicast object

extern thread_local std::string g_init_error;

// Every extern "C" entry point runs with its context's device current and restores the caller's device on return:
// several contexts (one per GPU) may live in one process, and a host framework (torch, CUDA.jl) may have switched devices.
struct DeviceGuard {
    int prev = -1; bool switched = false;
    explicit DeviceGuard(int dev) {
        if (dev < 0) return;
        if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) switched = (cudaSetDevice(dev) == cudaSuccess);
    }
    ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};
#define FDB_ENTER(c) fdb::DeviceGuard fdb_device_guard__(fdb::device_of(c))

inline int device_of(const fdb_ctx* c) { return c ? c->device : -1; }

inline int fail(fdb_ctx* ctx, int code, const char* fmt, ...) __attribute__((format(printf, 3, 4)));
inline int fail(fdb_ctx* ctx, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf; else g_init_error = buf;
    return code;
}

#define FDB_CUDA(ctx, expr)                                                                        \
    do {                                                                                           \
        cudaError_t e__ = (expr);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            return fdb::fail(ctx, e__ == cudaErrorMemoryAllocation ? FDB_ERR_OOM : FDB_ERR_CUDA,   \
                             "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

// after a host synchronisation: did an exchange barrier on this context give up (dead / slow peer)?  Then every result
// stored since is incomplete: the context stays poisoned and every later synchronising call reports it.
inline int check_poison(fdb_ctx* ctx, const char* what) {
    if (ctx->xch_timeout_flag && *(volatile unsigned*)ctx->xch_timeout_flag) ctx->poisoned = true;
    if (ctx->poisoned)
        return fail(ctx, FDB_ERR_CUDA, "%s: a peer never reached an exchange barrier (timeout); results since then are incomplete", what);
    return FDB_OK;
}

// after a kernel launch: check the launch, count it, and drain the stream unless async
inline int finish(fdb_ctx* ctx, int nlaunch, const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(ctx, FDB_ERR_CUDA, "%s: launch failed: %s", what, cudaGetErrorString(e));
    ctx->launches += nlaunch;
    if (!ctx->async) {
        e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) return fail(ctx, FDB_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
        return check_poison(ctx, what);
    }
    return FDB_OK;
}

int ensure_scratch(fdb_ctx* ctx, size_t bytes);
int ensure_pinned(fdb_ctx* ctx, size_t bytes);
int ensure_counter(fdb_ctx* ctx);
// fdb_exchange.cu: mirrors for a driver whose outputs are `outs` (null entries skipped); fails unless every
// output lies inside the enabled exchange's arena
int mirrors_for(fdb_ctx* ctx, fdb_mirrors* m, const fdb_array* const* outs, int nouts, const char* what);

// fdb_dmma.cu: batched tanh.(A*x .+ b) chunk sweeps on the FP64 tensor-core contraction
int tanh_affine_dmma(fdb_ctx* ctx, const fdb_array* A, int64_t lda, const fdb_array* b, const fdb_array* x, int64_t m, int N, int64_t lo,
                     int64_t hi, int64_t nchunks, int lcs, fdb_array* J, int64_t ldJ, double* yout, bool* used_dmma, const fdb_mirrors& Mr);
bool dmma_eligible(const double* A, int64_t lda, int64_t m, int64_t k);
int matvec_product_lanes(fdb_ctx* ctx, const double* A, int64_t lda, int64_t m, const fdb_array* x, int N, int64_t lo, int64_t hi, int64_t nchunks,
                         double** Y_out, int64_t* ldy_out, int* nlaunch);

// fdb_jit.cu
void jit_release(fdb_ctx* ctx);
int jit_launch(fdb_ctx* ctx, void* entry, unsigned gridx, unsigned gridy, unsigned threads, void** args, const char* what);

// cudaFuncSetAttribute is a per-device setting: remember per (kernel instantiation, device)
inline bool first_use_on_device(unsigned long long& mask, int dev) {
    const bool first = !((mask >> (dev & 63)) & 1ull);
    mask |= 1ull << (dev & 63);
    return first;
}

inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }
inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

// host mirror of the reference's chunk bookkeeping (src/gradient.jl:121-125)
struct Plan {
    int64_t xlen, remainder, lastchunksize, lastchunkindex, middle_hi, nchunks;
};
inline Plan make_plan(int64_t xlen, int64_t N) {
    Plan p;
    p.xlen = xlen;
    p.remainder = xlen % N;
    p.lastchunksize = p.remainder == 0 ? N : p.remainder;
    p.lastchunkindex = xlen - p.lastchunksize + 1;
    p.middle_hi = (xlen - p.lastchunksize) / N;
    p.nchunks = p.middle_hi + 1;
    return p;
}

// dispatch a run-time chunk size to a compile-time one (N = 1..12, src/prelude.jl:12 `@nif 12`)
#define FDB_DISPATCH_N(N_, ...)                                        \
    switch (N_) {                                                      \
        case 1: { constexpr int NN = 1; __VA_ARGS__; } break;          \
        case 2: { constexpr int NN = 2; __VA_ARGS__; } break;          \
        case 3: { constexpr int NN = 3; __VA_ARGS__; } break;          \
        case 4: { constexpr int NN = 4; __VA_ARGS__; } break;          \
        case 5: { constexpr int NN = 5; __VA_ARGS__; } break;          \
        case 6: { constexpr int NN = 6; __VA_ARGS__; } break;          \
        case 7: { constexpr int NN = 7; __VA_ARGS__; } break;          \
        case 8: { constexpr int NN = 8; __VA_ARGS__; } break;          \
        case 9: { constexpr int NN = 9; __VA_ARGS__; } break;          \
        case 10: { constexpr int NN = 10; __VA_ARGS__; } break;        \
        case 11: { constexpr int NN = 11; __VA_ARGS__; } break;        \
        case 12: { constexpr int NN = 12; __VA_ARGS__; } break;        \
        default: return fdb::fail(ctx, FDB_ERR_BADARG, "chunk size N=%d outside 1..12", (int)(N_)); \
    }
#define FDB_DISPATCH_NS(ns_, ...)                                     \
    if (ns_) { constexpr bool NS = true; __VA_ARGS__; } else { constexpr bool NS = false; __VA_ARGS__; }

}  // namespace fdb
