// fdb_exchange.cu -- multi-GPU result exchange over NVLink peer memory (one process per GPU).
//
// The chunk sweeps of src/gradient.jl:141-152 / src/jacobian.jl:199-210 are independent, so each rank
// sweeps a contiguous chunk range and owns a contiguous slice of the result.  The
// reference's caller expects the whole `result`; rather than all-gathering after the sweeps, every rank
// keeps its result in a symmetric arena that the peers map through CUDA IPC, and the driver kernels
// repeat each extraction store into every peer's arena (fdb_mirror.cuh).  What is left of the collective
// is a barrier: flag words at the end of each arena, written with release/system scope by the peers and
// polled with acquire/system scope locally, enqueued on the context's stream like any other kernel.
#include <string.h>

#include "fdb_internal.h"

namespace fdb {

static constexpr size_t FLAG_PAGE = 256;            // FDB_MAX_RANKS u64 arrival flags, then the timeout counter
static constexpr size_t TIMEOUT_OFF = 128;

struct PeerFlags { unsigned long long* flags[FDB_MAX_RANKS]; };

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t;
}

// thread r: tell rank r that this rank has arrived at `epoch`, then wait until rank r has.
__global__ void exchange_barrier_kernel(PeerFlags P, int rank, int world, unsigned long long epoch, unsigned long long timeout_ns,
                                        unsigned* host_flag) {
    const int r = threadIdx.x;
    if (r >= world || r == rank) return;
    __threadfence_system();       // everything this stream stored into peer arenas is ordered before the flag
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(P.flags[r] + rank), "l"(epoch) : "memory");
    const unsigned long long* mine = P.flags[rank] + r;
    const unsigned long long t0 = globaltimer_ns();
    for (;;) {
        unsigned long long v;
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine) : "memory");
        if (v >= epoch) break;
        if (globaltimer_ns() - t0 > timeout_ns) {      // a peer died: give up instead of hanging the GPU
            atomicAdd(reinterpret_cast<unsigned*>(reinterpret_cast<char*>(P.flags[rank]) + TIMEOUT_OFF), 1u);
            if (host_flag) { *reinterpret_cast<volatile unsigned*>(host_flag) = 1u; __threadfence_system(); }   // sticky, host-visible: check_poison
            break;
        }
        __nanosleep(200);
    }
}

// ---- store-bandwidth probe: the roofline denominator of the fused exchange -------------------------------
// Every rank pushes the first n doubles of its arena into all peers at once.  VARIANT 0: element-major, each
// thread repeats its 8-byte store for every peer in the order of `d` (what mstore does); 1: the same with
// 16-byte stores; 2: peer-major, blockIdx.y picks the peer and the block streams a contiguous range to it.
template <int VARIANT>
__global__ void __launch_bounds__(256) exchange_probe_kernel(const double* __restrict__ src, int64_t n, const fdb_mirrors M) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    if (VARIANT == 0) {
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
            const double v = src[i];
            for (int r = 0; r < M.count; r++) *reinterpret_cast<double*>(reinterpret_cast<char*>(const_cast<double*>(src + i)) + M.delta[r]) = v;
        }
    } else if (VARIANT == 1) {
        for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 2; i + 1 < n; i += stride * 2) {
            const double2 v = *reinterpret_cast<const double2*>(src + i);
            for (int r = 0; r < M.count; r++) *reinterpret_cast<double2*>(reinterpret_cast<char*>(const_cast<double*>(src + i)) + M.delta[r]) = v;
        }
    } else {
        const int r = blockIdx.y;
        char* dst = reinterpret_cast<char*>(const_cast<double*>(src)) + M.delta[r];
        for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 2; i + 1 < n; i += stride * 2)
            *reinterpret_cast<double2*>(dst + i * 8) = *reinterpret_cast<const double2*>(src + i);
    }
}

int mirrors_for(fdb_ctx* ctx, fdb_mirrors* m, const fdb_array* const* outs, int nouts, const char* what) {
    *m = fdb_mirrors();
    fdb_exchange* x = ctx->xch;
    if (!x) return FDB_OK;
    for (int i = 0; i < nouts; i++) {
        if (!outs[i]) continue;
        const char* p = reinterpret_cast<const char*>(outs[i]->data);
        if (p < x->base || p + (size_t)outs[i]->n * sizeof(double) > x->base + x->data_bytes)
            return fail(ctx, FDB_ERR_BADARG, "%s: with a peer exchange enabled every result array must be a view of the exchange arena "
                                             "(fdb_exchange_view); output %d is not", what, i);
    }
    if (x->mc_base && x->mc_on && x->world > 1) {      // NVSwitch multicast: one store, replicated into every arena (fdb_multicast.cu)
        m->mc = 1; m->count = 1; m->delta[0] = (long long)(x->mc_base - x->base);
        return FDB_OK;
    }
    // rank r visits its peers as r+1, r+2, ...: at any instant the ranks address different destinations
    for (int k = 1; k < x->world; k++) {
        const int r = (x->rank + k) % x->world;
        m->delta[m->count++] = (long long)(x->peer[r] - x->base);
    }
    return FDB_OK;
}

}  // namespace fdb

using namespace fdb;

extern "C" int fdb_exchange_create(fdb_ctx* ctx, int64_t n_doubles, int rank, int world, fdb_exchange** out) {
    FDB_ENTER(ctx);
    if (!ctx || !out || n_doubles < 0) return fail(ctx, FDB_ERR_BADARG, "fdb_exchange_create: bad argument");
    if (world < 1 || world > FDB_MAX_RANKS || rank < 0 || rank >= world)
        return fail(ctx, FDB_ERR_BADARG, "fdb_exchange_create: rank %d of %d (at most %d ranks, one NVSwitch domain)", rank, world, FDB_MAX_RANKS);
    fdb_exchange* x = new fdb_exchange();
    x->ctx = ctx; x->rank = rank; x->world = world;
    x->data_bytes = (size_t)round_up(n_doubles > 0 ? n_doubles : 1, 32) * sizeof(double);
    cudaSetDevice(ctx->device);
    cudaError_t e = cudaMalloc(&x->base, x->data_bytes + FLAG_PAGE);     // cudaMalloc, not a pool: IPC needs a plain allocation
    if (e != cudaSuccess) { delete x; return fail(ctx, e == cudaErrorMemoryAllocation ? FDB_ERR_OOM : FDB_ERR_CUDA,
                                                  "cudaMalloc(%zu bytes): %s", x->data_bytes + FLAG_PAGE, cudaGetErrorString(e)); }
    e = cudaMemset(x->base, 0, x->data_bytes + FLAG_PAGE);
    if (e != cudaSuccess) { cudaFree(x->base); delete x; return fail(ctx, FDB_ERR_CUDA, "cudaMemset: %s", cudaGetErrorString(e)); }
    if (!ctx->xch_timeout_flag) {     // pinned, device-visible word: a barrier that gives up poisons the context (fdb_internal.h check_poison)
        e = cudaHostAlloc(reinterpret_cast<void**>(&ctx->xch_timeout_flag), 64, cudaHostAllocMapped | cudaHostAllocPortable);
        if (e != cudaSuccess) { cudaFree(x->base); delete x; return fail(ctx, FDB_ERR_CUDA, "cudaHostAlloc: %s", cudaGetErrorString(e)); }
        *ctx->xch_timeout_flag = 0;
    }
    x->peer[rank] = x->base;
    x->connected = (world == 1);
    *out = x;
    return FDB_OK;
}

extern "C" int fdb_exchange_handle(fdb_exchange* x, unsigned char handle[FDB_IPC_HANDLE_BYTES]) {
    FDB_ENTER(x ? x->ctx : nullptr);
    if (!x || !handle) return FDB_ERR_BADARG;
    static_assert(sizeof(cudaIpcMemHandle_t) == FDB_IPC_HANDLE_BYTES, "IPC handle size");
    if (x->vmm) return fail(x->ctx, FDB_ERR_BADARG, "fdb_exchange_handle: a multicast arena is shared with fdb_exchange_share_handle, not through CUDA IPC");
    cudaIpcMemHandle_t h;
    FDB_CUDA(x->ctx, cudaIpcGetMemHandle(&h, x->base));
    memcpy(handle, &h, sizeof h);
    return FDB_OK;
}

extern "C" int fdb_exchange_connect(fdb_exchange* x, const unsigned char* handles) {
    FDB_ENTER(x ? x->ctx : nullptr);
    if (!x || !handles) return FDB_ERR_BADARG;
    fdb_ctx* ctx = x->ctx;
    if (x->vmm) return fail(ctx, FDB_ERR_BADARG, "fdb_exchange_connect: a multicast arena connects with fdb_exchange_connect_multicast");
    if (x->connected) return FDB_OK;
    cudaSetDevice(ctx->device);
    for (int r = 0; r < x->world; r++) {
        if (r == x->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + (size_t)r * FDB_IPC_HANDLE_BYTES, sizeof h);
        void* p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            for (int q = 0; q < r; q++) if (q != x->rank && x->peer[q]) { cudaIpcCloseMemHandle(x->peer[q]); x->peer[q] = nullptr; }
            return fail(ctx, FDB_ERR_CUDA, "cudaIpcOpenMemHandle(rank %d's arena): %s", r, cudaGetErrorString(e));
        }
        x->peer[r] = static_cast<char*>(p);
    }
    x->connected = true;
    return FDB_OK;
}

// Single-process form of fdb_exchange_connect: the `world` exchanges live in THIS process, one per device (one Julia
// session driving all GPUs).  CUDA IPC cannot open a handle in the process that exported it, and does not need to:
// with peer access enabled both ways the arenas are directly addressable (unified virtual addressing), so the
// peers' base pointers are linked as they are.
extern "C" int fdb_exchange_connect_local(fdb_exchange* const* xs, int world) {
    if (!xs || world < 1 || world > FDB_MAX_RANKS) return FDB_ERR_BADARG;
    for (int r = 0; r < world; r++)
        if (!xs[r] || xs[r]->world != world || xs[r]->rank != r) return fail(xs[0] ? xs[0]->ctx : nullptr, FDB_ERR_BADARG,
            "fdb_exchange_connect_local: xs[%d] must be rank %d of %d", r, r, world);
    for (int r = 0; r < world; r++)
        if (xs[r]->data_bytes != xs[0]->data_bytes || xs[r]->vmm) return fail(xs[0]->ctx, FDB_ERR_BADARG, "fdb_exchange_connect_local: arenas must be symmetric fdb_exchange_create arenas");
    int prev = 0; cudaGetDevice(&prev);
    for (int r = 0; r < world; r++) {
        fdb_ctx* ctx = xs[r]->ctx;
        cudaSetDevice(ctx->device);
        for (int q = 0; q < world; q++) {
            if (q == r) continue;
            const int pd = xs[q]->ctx->device;
            if (pd == ctx->device) continue;                                        // two contexts on one GPU (tests): plain pointers
            int can = 0; cudaDeviceCanAccessPeer(&can, ctx->device, pd);
            if (!can) { cudaSetDevice(prev); return fail(ctx, FDB_ERR_CUDA, "device %d cannot map device %d's memory (no peer access)", ctx->device, pd); }
            cudaError_t e = cudaDeviceEnablePeerAccess(pd, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
            if (e != cudaSuccess) { cudaSetDevice(prev); return fail(ctx, FDB_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d -> %d): %s", ctx->device, pd, cudaGetErrorString(e)); }
        }
    }
    cudaSetDevice(prev);
    for (int r = 0; r < world; r++) {
        for (int q = 0; q < world; q++) xs[r]->peer[q] = xs[q]->base;
        xs[r]->connected = true; xs[r]->in_process = true;
    }
    return FDB_OK;
}

extern "C" int fdb_exchange_set_timeout(fdb_exchange* x, double seconds) {
    if (!x || !(seconds > 0.0)) return FDB_ERR_BADARG;
    x->timeout_ns = (unsigned long long)(seconds * 1e9);
    return FDB_OK;
}

extern "C" int fdb_exchange_view(fdb_exchange* x, int64_t offset, int64_t n, int64_t ld, fdb_array** out) {
    if (!x || !out) return FDB_ERR_BADARG;
    if (offset < 0 || n < 0 || ld < n || (size_t)(offset + ld) * sizeof(double) > x->data_bytes)
        return fail(x->ctx, FDB_ERR_BADARG, "fdb_exchange_view: [%lld, %lld) outside the arena of %zu doubles", (long long)offset,
                    (long long)(offset + ld), x->data_bytes / sizeof(double));
    return fdb_array_wrap(x->ctx, reinterpret_cast<double*>(x->base) + offset, n, ld, 0, 0, out);
}

extern "C" int fdb_exchange_enable(fdb_ctx* ctx, fdb_exchange* x) {
    FDB_ENTER(ctx);
    if (!ctx) return FDB_ERR_BADARG;
    if (x && x->ctx != ctx) return fail(ctx, FDB_ERR_BADARG, "fdb_exchange_enable: exchange belongs to another context");
    if (x && !x->connected) return fail(ctx, FDB_ERR_BADARG, "fdb_exchange_enable: call fdb_exchange_connect first");
    ctx->xch = x;
    return FDB_OK;
}

extern "C" int fdb_exchange_barrier(fdb_exchange* x) {
    FDB_ENTER(x ? x->ctx : nullptr);
    if (!x) return FDB_ERR_BADARG;
    fdb_ctx* ctx = x->ctx;
    if (!x->connected) return fail(ctx, FDB_ERR_BADARG, "fdb_exchange_barrier: not connected");
    if (ctx->capturing) return fail(ctx, FDB_ERR_UNSUPPORTED, "fdb_exchange_barrier: the epoch is a launch argument; not capturable");
    if (x->world == 1) return FDB_OK;
    PeerFlags P;
    for (int r = 0; r < FDB_MAX_RANKS; r++)
        P.flags[r] = r < x->world ? reinterpret_cast<unsigned long long*>(x->peer[r] + x->data_bytes) : nullptr;
    x->epoch++;
    exchange_barrier_kernel<<<1, 32, 0, ctx->stream>>>(P, x->rank, x->world, x->epoch, x->timeout_ns, ctx->xch_timeout_flag);
    return finish(ctx, 1, "fdb_exchange_barrier");
}

extern "C" int fdb_exchange_measure_store_bw(fdb_exchange* x, int64_t n_doubles, int variant, int rotate, double* gb_per_s) {
    FDB_ENTER(x ? x->ctx : nullptr);
    if (!x || !gb_per_s) return FDB_ERR_BADARG;
    fdb_ctx* ctx = x->ctx;
    if (!x->connected || x->world < 2) return fail(ctx, FDB_ERR_BADARG, "fdb_exchange_measure_store_bw: needs a connected exchange with peers");
    if (n_doubles < 2 || (size_t)n_doubles * sizeof(double) > x->data_bytes || variant < 0 || variant > 2)
        return fail(ctx, FDB_ERR_BADARG, "fdb_exchange_measure_store_bw: bad size or variant");
    fdb_mirrors M;
    for (int k = 1; k < x->world; k++) {
        const int r = rotate ? (x->rank + k) % x->world : (k - 1 < x->rank ? k - 1 : k);
        M.delta[M.count++] = (long long)(x->peer[r] - x->base);
    }
    const bool was_async = ctx->async; ctx->async = true;
    cudaEvent_t e0, e1;
    FDB_CUDA(ctx, cudaEventCreate(&e0)); FDB_CUDA(ctx, cudaEventCreate(&e1));
    float best = 1e30f;
    const double* src = reinterpret_cast<const double*>(x->base);
    for (int rep = 0; rep < 4; rep++) {
        int rc = fdb_exchange_barrier(x);                                                    // all ranks push at the same time
        if (rc) { ctx->async = was_async; cudaEventDestroy(e0); cudaEventDestroy(e1); return rc; }
        cudaEventRecord(e0, ctx->stream);
        if (variant == 0) exchange_probe_kernel<0><<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(src, n_doubles, M);
        else if (variant == 1) exchange_probe_kernel<1><<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(src, n_doubles, M);
        else exchange_probe_kernel<2><<<dim3(ctx->sm_count * 2, M.count), 256, 0, ctx->stream>>>(src, n_doubles, M);
        cudaEventRecord(e1, ctx->stream);
        cudaError_t ce = cudaEventSynchronize(e1);
        if (ce != cudaSuccess) {
            ctx->async = was_async; cudaEventDestroy(e0); cudaEventDestroy(e1);
            return fail(ctx, FDB_ERR_CUDA, "fdb_exchange_measure_store_bw: %s", cudaGetErrorString(ce));
        }
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
        ctx->launches++;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    ctx->async = was_async;
    *gb_per_s = (double)n_doubles * 8.0 * M.count / (best * 1e-3) / 1e9;      // bytes this rank sent per second
    return FDB_OK;
}

// all-gather / broadcast for data the drivers did not mirror themselves (e.g. results of the materialised loop path):
// copy [offset, offset+count) of the local arena to the same place in every peer's arena.
extern "C" int fdb_exchange_push(fdb_exchange* x, int64_t offset, int64_t count) {
    FDB_ENTER(x ? x->ctx : nullptr);
    if (!x) return FDB_ERR_BADARG;
    fdb_ctx* ctx = x->ctx;
    if (!x->connected) return fail(ctx, FDB_ERR_BADARG, "fdb_exchange_push: not connected");
    if (offset < 0 || count < 0 || (size_t)(offset + count) * sizeof(double) > x->data_bytes)
        return fail(ctx, FDB_ERR_BADARG, "fdb_exchange_push: [%lld, %lld) outside the arena", (long long)offset, (long long)(offset + count));
    if (x->world == 1 || count == 0) return FDB_OK;
    fdb_mirrors M;
    if (x->mc_base && x->mc_on) { M.count = 1; M.delta[0] = (long long)(x->mc_base - x->base); }      // one copy, replicated by the switch
    else for (int k = 1; k < x->world; k++) M.delta[M.count++] = (long long)(x->peer[(x->rank + k) % x->world] - x->base);
    const double* src = reinterpret_cast<const double*>(x->base) + offset;
    int64_t blocks = cdiv(count, 256 * 4); if (blocks > ctx->sm_count * 8) blocks = ctx->sm_count * 8; if (blocks < 1) blocks = 1;
    exchange_probe_kernel<0><<<(unsigned)blocks, 256, 0, ctx->stream>>>(src, count, M);
    return finish(ctx, 1, "fdb_exchange_push");
}

extern "C" int fdb_exchange_timeouts(fdb_exchange* x, int64_t* count) {
    FDB_ENTER(x ? x->ctx : nullptr);
    if (!x || !count) return FDB_ERR_BADARG;
    unsigned v = 0;
    FDB_CUDA(x->ctx, cudaMemcpyAsync(&v, x->base + x->data_bytes + TIMEOUT_OFF, sizeof v, cudaMemcpyDeviceToHost, x->ctx->stream));
    FDB_CUDA(x->ctx, cudaStreamSynchronize(x->ctx->stream));
    *count = v;
    return FDB_OK;
}

extern "C" int fdb_exchange_destroy(fdb_exchange* x) {
    FDB_ENTER(x ? x->ctx : nullptr);
    if (!x) return FDB_ERR_BADARG;
    fdb_ctx* ctx = x->ctx;
    if (ctx->xch == x) ctx->xch = nullptr;
    cudaStreamSynchronize(ctx->stream);
    if (x->vmm) { multicast_destroy(x); delete x; return FDB_OK; }
    if (!x->in_process)
        for (int r = 0; r < x->world; r++)
            if (r != x->rank && x->peer[r]) cudaIpcCloseMemHandle(x->peer[r]);
    cudaFree(x->base);
    delete x;
    return FDB_OK;
}
