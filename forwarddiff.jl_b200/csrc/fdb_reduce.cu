// fdb_reduce.cu -- sum / prod and fused map-reduce evaluation of catalogue targets on materialised SoA Dual
// arrays (what `sum`, `prod` and `f(xdual)` are in the reference's chunk loop, src/gradient.jl:135,144,151).
// HBM-bound: 8(N+1) bytes per dual element read.  Two kernels: register windows (views, small arrays) and a
// TMA-pipelined stream through shared memory (large aligned arrays).
#include "fdb_internal.h"
#include "fdb_targets.cuh"
#include "fdb_plane_io.cuh"
#include "fdb_tma.cuh"

namespace fdb {

// ---- map-reduce over a materialised dual array -----------------------------------------
// pass 1: each block reduces a grid-strided set of terms and writes its K accumulators
// (AoS, K*(N+1) doubles) to scratch; pass 2: one block combines them in block order
// (deterministic) and applies F::finalize.
// A thread's register window over W = 2 + HALO consecutive elements: the pair (i, i+1) comes
// from one 16-byte load per plane when the planes are 16-byte aligned, the halo from 8-byte loads
// (the neighbouring thread's pair: an L1 hit).  Terms i and i+1 then index it with compile-time
// offsets after inlining.
template <int N, bool NS, int W> struct WindowLoader {
    Dual<double, N, NS> e[W];
    int64_t base;
    FDB_DEV const Dual<double, N, NS>& operator()(int64_t j) const { return e[j - base]; }
};
template <class F, int N, bool NS, int THREADS>
__global__ void __launch_bounds__(THREADS, 2) mapreduce_pass1(const double* __restrict__ base, int64_t ld, int64_t n,
                                                              double* __restrict__ partial, unsigned* __restrict__ counter, bool vec,
                                                              double* __restrict__ out, int64_t ldo) {
    typedef Dual<double, N, NS> D;
    constexpr int W = 2 + F::HALO;
    D acc[F::K];
    D like; like.v = 0.0;
#pragma unroll
    for (int k = 0; k < N; k++) like.p[k] = 0.0;
#pragma unroll
    for (int k = 0; k < F::K; k++) acc[k] = F::init(k, like);
    const int64_t nt = F::nterms(n);
    for (int64_t i = ((int64_t)blockIdx.x * THREADS + threadIdx.x) * 2; i < nt; i += (int64_t)gridDim.x * THREADS * 2) {
        WindowLoader<N, NS, W> x; x.base = i;
        const bool pair = i + 1 < n;
        Dual2<N, NS> p2 = load2<N, NS>(base, ld, i, pair, false, vec);
        x.e[0] = p2.a; x.e[1] = p2.b;
#pragma unroll
        for (int w = 2; w < W; w++) {
            const int64_t j = i + w < n ? i + w : n - 1;     // clamp: only read by terms that do not exist
            x.e[w].v = base[j];
#pragma unroll
            for (int k = 0; k < N; k++) x.e[w].p[k] = base[(int64_t)(k + 1) * ld + j];
        }
        F::term(x, i, n, acc);
        if (i + 1 < nt) F::term(x, i + 1, n, acc);
    }
    block_combine<F, D, THREADS>(acc);
    __shared__ bool is_last;
    if (threadIdx.x == 0) {
        double* o = partial + (int64_t)blockIdx.x * F::K * (N + 1);
#pragma unroll
        for (int k = 0; k < F::K; k++) {
            o[k * (N + 1)] = acc[k].v;
#pragma unroll
            for (int j = 0; j < N; j++) o[k * (N + 1) + 1 + j] = acc[k].p[j];
        }
        __threadfence();
        const unsigned ticket = atomicAdd(counter, 1u);
        is_last = (ticket == gridDim.x - 1);
        if (is_last) *counter = 0u;            // self-cleaning for the next launch on this stream
    }
    __syncthreads();
    if (!is_last) return;
    // ---- the last block to finish combines all block partials in block order (deterministic) ----
    __threadfence();
#pragma unroll
    for (int k = 0; k < F::K; k++) acc[k] = F::init(k, like);
    for (int b = threadIdx.x; b < (int)gridDim.x; b += THREADS) {
        const double* o = partial + (int64_t)b * F::K * (N + 1);
#pragma unroll
        for (int k = 0; k < F::K; k++) {
            D t; t.v = __ldcg(o + k * (N + 1));
#pragma unroll
            for (int j = 0; j < N; j++) t.p[j] = __ldcg(o + k * (N + 1) + 1 + j);
            acc[k] = F::combine(k, acc[k], t);
        }
    }
    __syncthreads();
    block_combine<F, D, THREADS>(acc);
    if (threadIdx.x == 0) {
        D y = F::finalize(acc, n);
        out[0] = y.v;
#pragma unroll
        for (int j = 0; j < N; j++) out[(int64_t)(j + 1) * ldo] = y.p[j];
    }
}
template <class F, int N, bool NS, int THREADS>
__global__ void __launch_bounds__(THREADS) mapreduce_pass2(const double* __restrict__ partial, int nblocks, int64_t n,
                                                           double* __restrict__ out, int64_t ldo) {
    typedef Dual<double, N, NS> D;
    D acc[F::K];
    D like; like.v = 0.0;
#pragma unroll
    for (int k = 0; k < N; k++) like.p[k] = 0.0;
#pragma unroll
    for (int k = 0; k < F::K; k++) acc[k] = F::init(k, like);
    for (int b = threadIdx.x; b < nblocks; b += THREADS) {
        const double* o = partial + (int64_t)b * F::K * (N + 1);
#pragma unroll
        for (int k = 0; k < F::K; k++) {
            D t; t.v = o[k * (N + 1)];
#pragma unroll
            for (int j = 0; j < N; j++) t.p[j] = o[k * (N + 1) + 1 + j];
            acc[k] = F::combine(k, acc[k], t);
        }
    }
    block_combine<F, D, THREADS>(acc);
    if (threadIdx.x == 0) {
        D y = F::finalize(acc, n);
        out[0] = y.v;
#pragma unroll
        for (int j = 0; j < N; j++) out[(int64_t)(j + 1) * ldo] = y.p[j];
    }
}

// ---- TMA-pipelined map-reduce -------------------------------------------------------------------
// The register-window kernel above can keep at most ~100 KB of loads in flight per SM (its loads live in
// registers) and cannot overlap a compute-heavy term with the loads it waits on.  Here a producer thread
// streams tiles of MR_TILE elements x (N+1) planes into a ring of MR_STAGES shared-memory stages with TMA
// bulk copies (one per plane, mbarrier complete_tx); four consumer warps evaluate the terms of a tile out of
// shared memory while the next tiles are in flight.  Persistent CTAs, grid = SMs x CTAs that fit.
constexpr int MR_TILE = 512, MR_STAGES = 4, MR_CONSUMERS = 256, MR_THREADS = MR_CONSUMERS + 32;
constexpr int MR_ROW = MR_TILE + 2;                     // tile row incl. a 2-element halo, 16-byte multiple

template <int N, bool NS> struct SmemWindow {           // loader over one staged tile
    const double* s; int64_t t0;
    FDB_DEV Dual<double, N, NS> operator()(int64_t j) const {
        const int e = (int)(j - t0);
        Dual<double, N, NS> d; d.v = s[e];
#pragma unroll
        for (int k = 0; k < N; k++) d.p[k] = s[(k + 1) * MR_ROW + e];
        return d;
    }
};

template <class F, int N, bool NS>
__global__ void __launch_bounds__(MR_THREADS) mapreduce_tma_kernel(const double* __restrict__ base, int64_t ld, int64_t n, int64_t ntiles,
                                                                   double* __restrict__ partial, unsigned* __restrict__ counter,
                                                                   double* __restrict__ out, int64_t ldo) {
    typedef Dual<double, N, NS> D;
    constexpr int P = N + 1;
    constexpr int STAGE_DOUBLES = P * MR_ROW;
    extern __shared__ __align__(128) unsigned char mr_smem_raw[];
    double* stages = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(mr_smem_raw) + 127) & ~uintptr_t(127));
    uint64_t* full = reinterpret_cast<uint64_t*>(stages + MR_STAGES * STAGE_DOUBLES);
    uint64_t* empty = full + MR_STAGES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < MR_STAGES; s++) { tma::mbar_init(&full[s], 1); tma::mbar_init(&empty[s], MR_CONSUMERS / 32); }
        tma::mbar_fence_init();
    }
    __syncthreads();
    const int64_t nt = F::nterms(n);
    D acc[F::K];
    D like; like.v = 0.0;
#pragma unroll
    for (int k = 0; k < N; k++) like.p[k] = 0.0;
#pragma unroll
    for (int k = 0; k < F::K; k++) acc[k] = F::init(k, like);

    if (warp == MR_CONSUMERS / 32) {
        // producer warp: lane q stages plane q of the tile (one TMA bulk copy each, issued together)
        int it = 0;
        for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, it++) {
            const int st = it % MR_STAGES;
            if (it >= MR_STAGES) tma::mbar_wait(&empty[st], ((it / MR_STAGES) - 1) & 1);
            const int64_t t0 = tile * MR_TILE;
            int64_t cnt = ld - t0; if (cnt > MR_ROW) cnt = MR_ROW;         // ld is a multiple of 32: cnt is even
            if (lane == 0) tma::mbar_expect_tx(&full[st], (uint32_t)(cnt * 8 * P));
            __syncwarp();
            double* dst = stages + st * STAGE_DOUBLES;
            if (lane < P) tma::bulk_g2s(dst + lane * MR_ROW, base + (int64_t)lane * ld + t0, (uint32_t)(cnt * 8), &full[st]);
        }
    } else {
        int it = 0;
        for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, it++) {
            const int st = it % MR_STAGES;
            tma::mbar_wait(&full[st], (it / MR_STAGES) & 1);
            const int64_t t0 = tile * MR_TILE;
            const double* s = stages + st * STAGE_DOUBLES + 2 * threadIdx.x;
            const int64_t i = t0 + 2 * threadIdx.x;
            WindowLoader<N, NS, 2 + F::HALO> x; x.base = i;       // pair via 16-byte shared loads (conflict free), halo 8-byte
            {
                const double2 v = *reinterpret_cast<const double2*>(s);
                x.e[0].v = v.x; x.e[1].v = v.y;
#pragma unroll
                for (int k = 0; k < N; k++) {
                    const double2 u = *reinterpret_cast<const double2*>(s + (k + 1) * MR_ROW);
                    x.e[0].p[k] = u.x; x.e[1].p[k] = u.y;
                }
#pragma unroll
                for (int w = 2; w < 2 + F::HALO; w++) {
                    x.e[w].v = s[w];
#pragma unroll
                    for (int k = 0; k < N; k++) x.e[w].p[k] = s[(k + 1) * MR_ROW + w];
                }
            }
            if (i < nt) F::term(x, i, n, acc);
            if (i + 1 < nt) F::term(x, i + 1, n, acc);
            __syncwarp();
            if (lane == 0) tma::mbar_arrive(&empty[st]);
        }
    }
    // producer warp carries neutral accumulators through the block reduction
    constexpr int RT = MR_THREADS;                            // all warps (producer included) take part in the block reduction
    block_combine<F, D, RT>(acc);
    __shared__ bool is_last;
    if (threadIdx.x == 0) {
        double* o = partial + (int64_t)blockIdx.x * F::K * P;
#pragma unroll
        for (int k = 0; k < F::K; k++) {
            o[k * P] = acc[k].v;
#pragma unroll
            for (int j = 0; j < N; j++) o[k * P + 1 + j] = acc[k].p[j];
        }
        __threadfence();
        const unsigned ticket = atomicAdd(counter, 1u);
        is_last = (ticket == gridDim.x - 1);
        if (is_last) *counter = 0u;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
#pragma unroll
    for (int k = 0; k < F::K; k++) acc[k] = F::init(k, like);
    for (int b = threadIdx.x; b < (int)gridDim.x; b += MR_THREADS) {
        const double* o = partial + (int64_t)b * F::K * P;
#pragma unroll
        for (int k = 0; k < F::K; k++) {
            D t; t.v = __ldcg(o + k * P);
#pragma unroll
            for (int j = 0; j < N; j++) t.p[j] = __ldcg(o + k * P + 1 + j);
            acc[k] = F::combine(k, acc[k], t);
        }
    }
    __syncthreads();
    block_combine<F, D, RT>(acc);
    if (threadIdx.x == 0) {
        D y = F::finalize(acc, n);
        out[0] = y.v;
#pragma unroll
        for (int j = 0; j < N; j++) out[(int64_t)(j + 1) * ldo] = y.p[j];
    }
}

template <class F, int N, bool NS>
static int mapreduce_tma_launch(fdb_ctx* ctx, fdb_array* out1, const fdb_array* a, int64_t ntiles) {
    const size_t smem = (size_t)MR_STAGES * (N + 1) * MR_ROW * 8 + 2 * MR_STAGES * 8 + 128;
    static unsigned long long configured = 0;      // per device: several contexts (GPUs) may live in one process
    if (first_use_on_device(configured, ctx->device)) {
        cudaError_t e = cudaFuncSetAttribute(mapreduce_tma_kernel<F, N, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail(ctx, FDB_ERR_CUDA, "cudaFuncSetAttribute(mapreduce_tma_kernel): %s", cudaGetErrorString(e));
    }
    int per_sm = (int)((220 * 1024) / smem); if (per_sm < 1) per_sm = 1; if (per_sm > 6) per_sm = 6;
    int64_t want = (int64_t)ctx->sm_count * per_sm;
    const int nblocks = (int)(ntiles < want ? ntiles : want);
    int rc = ensure_scratch(ctx, (size_t)nblocks * F::K * (N + 1) * sizeof(double)); if (rc) return rc;
    rc = ensure_counter(ctx); if (rc) return rc;
    mapreduce_tma_kernel<F, N, NS><<<nblocks, MR_THREADS, smem, ctx->stream>>>(a->data, a->ld, a->n, ntiles, ctx->scratch, ctx->counter,
                                                                             out1->data, out1->ld);
    return FDB_OK;
}

template <class F>
static int mapreduce_launch(fdb_ctx* ctx, fdb_array* out1, const fdb_array* a, const char* what) {
    constexpr int THREADS = 256;
    int64_t nt = F::nterms(a->n);
    // large, aligned, padded arrays stream through the TMA pipeline; views and small arrays use register windows
    const int64_t ntiles = cdiv(nt, MR_TILE);
    if (!ctx->nansafe && nt >= 64 * MR_TILE && ((((uintptr_t)a->data) & 15) == 0) && (a->ld % 32 == 0) &&
        (ntiles - 1) * MR_TILE + 2 <= a->ld && a->owned) {
        int rc = FDB_OK;
        FDB_DISPATCH_N(a->N, (rc = mapreduce_tma_launch<F, NN, false>(ctx, out1, a, ntiles)));
        if (rc) return rc;
        return finish(ctx, 1, what);
    }
    int64_t want = cdiv(nt > 0 ? nt : 1, THREADS * 2 * 2);       // >= 2 pairs per thread
    int nblocks = (int)(want < (int64_t)ctx->sm_count * 6 ? want : (int64_t)ctx->sm_count * 6);
    if (nblocks < 1) nblocks = 1;
    const bool vec = ((((uintptr_t)a->data) & 15) == 0) && ((a->ld & 1) == 0);
    int rc = ensure_scratch(ctx, (size_t)nblocks * F::K * (a->N + 1) * sizeof(double)); if (rc) return rc;
    rc = ensure_counter(ctx); if (rc) return rc;
    FDB_DISPATCH_N(a->N, FDB_DISPATCH_NS(ctx->nansafe,
        (mapreduce_pass1<F, NN, NS, THREADS><<<nblocks, THREADS, 0, ctx->stream>>>(a->data, a->ld, a->n, ctx->scratch, ctx->counter, vec,
                                                                                    out1->data, out1->ld))));
    return finish(ctx, 1, what);
}

}  // namespace fdb

using namespace fdb;

static int check_same(fdb_ctx* ctx, const fdb_array* out, const fdb_array* a, const char* what) {
    if (!ctx || !out || !a) return fail(ctx, FDB_ERR_BADARG, "%s: null argument", what);
    if (out->level != 1 || a->level != 1) return fail(ctx, FDB_ERR_UNSUPPORTED, "%s: operands must be level-1 Dual arrays", what);
    if (out->N != a->N) return fail(ctx, FDB_ERR_BADARG, "%s: chunk sizes differ (%d vs %d)", what, out->N, a->N);
    return FDB_OK;
}
static int check_reduce(fdb_ctx* ctx, fdb_array* out1, const fdb_array* a, const char* what) {
    int rc = check_same(ctx, out1, a, what); if (rc) return rc;
    if (out1->n != 1) return fail(ctx, FDB_ERR_SHAPE, "%s: output must have length 1", what);
    return FDB_OK;
}
extern "C" int fdb_reduce(fdb_ctx* ctx, int redop, fdb_array* out1, const fdb_array* a) {
    FDB_ENTER(ctx);
    int rc = check_reduce(ctx, out1, a, "fdb_reduce"); if (rc) return rc;
    if (redop == FDB_SUM) return mapreduce_launch<SumAll>(ctx, out1, a, "fdb_reduce(sum)");
    if (redop == FDB_PROD) return mapreduce_launch<ProdAll>(ctx, out1, a, "fdb_reduce(prod)");
    return fail(ctx, FDB_ERR_UNSUPPORTED, "fdb_reduce: unknown reduction %d", redop);
}
extern "C" int fdb_eval_scalar_fn(fdb_ctx* ctx, int fn, fdb_array* out1, const fdb_array* xdual) {
    FDB_ENTER(ctx);
    int rc = check_reduce(ctx, out1, xdual, "fdb_eval_scalar_fn"); if (rc) return rc;
    switch (fn) {
        case FDB_FN_ROSENBROCK: return mapreduce_launch<Rosenbrock>(ctx, out1, xdual, "fdb_eval_scalar_fn(rosenbrock)");
        case FDB_FN_ACKLEY: return mapreduce_launch<Ackley>(ctx, out1, xdual, "fdb_eval_scalar_fn(ackley)");
        case FDB_FN_SUMSQ: return mapreduce_launch<SumSq>(ctx, out1, xdual, "fdb_eval_scalar_fn(sumsq)");
        case FDB_FN_PROD: return mapreduce_launch<ProdAll>(ctx, out1, xdual, "fdb_eval_scalar_fn(prod)");
    }
    return fail(ctx, FDB_ERR_UNSUPPORTED, "fdb_eval_scalar_fn: unknown target %d", fn);
}
