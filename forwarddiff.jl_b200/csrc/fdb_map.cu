// fdb_map.cu -- array arithmetic on materialised SoA Dual arrays: the
// elementwise (broadcast) and reduction kernels a user `f` is made of.
//
// These replace src/dual.jl's scalar ops applied through Julia broadcast / sum /
// prod.  They are HBM-bound: every
// element is read once and written once, so the roofline is
//   8(N+1) bytes x (dual elements read + written) / measured HBM bandwidth.
// Layout choice: one thread owns two consecutive elements; each of the N+1 planes
// is read with a 16-byte load, so a warp reads 512 contiguous bytes per plane and
// every thread has N+1 independent loads in flight before the first use.
#include "fdb_internal.h"
#include "fdb_targets.cuh"
#include "fdb_tma.cuh"

namespace fdb {

// ---- plane IO: two consecutive elements per thread -----------------------------------
template <int N, bool NS> struct Dual2 { Dual<double, N, NS> a, b; };

// `bcast` arrays have a single element that every position reads
template <int N, bool NS>
FDB_DEV Dual2<N, NS> load2(const double* __restrict__ base, int64_t ld, int64_t e, bool pair, bool bcast, bool vec) {
    Dual2<N, NS> r;
    if (bcast) {
        r.a.v = base[0];
#pragma unroll
        for (int k = 0; k < N; k++) r.a.p[k] = base[(int64_t)(k + 1) * ld];
        r.b = r.a;
    } else if (pair && vec) {
        double2 t = __ldg(reinterpret_cast<const double2*>(base + e));
        r.a.v = t.x; r.b.v = t.y;
#pragma unroll
        for (int k = 0; k < N; k++) {
            double2 u = __ldg(reinterpret_cast<const double2*>(base + (int64_t)(k + 1) * ld + e));
            r.a.p[k] = u.x; r.b.p[k] = u.y;
        }
    } else {
        r.a.v = base[e];
#pragma unroll
        for (int k = 0; k < N; k++) r.a.p[k] = base[(int64_t)(k + 1) * ld + e];
        r.b = r.a;
        if (pair) {            // unaligned view: scalar loads for the second element too
            r.b.v = base[e + 1];
#pragma unroll
            for (int k = 0; k < N; k++) r.b.p[k] = base[(int64_t)(k + 1) * ld + e + 1];
        }
    }
    return r;
}
FDB_DEV void loadreal2(const double* __restrict__ base, int64_t e, bool pair, bool bcast, bool vec, double& a, double& b) {
    if (bcast) { a = b = base[0]; }
    else if (pair && vec) { double2 t = __ldg(reinterpret_cast<const double2*>(base + e)); a = t.x; b = t.y; }
    else { a = b = base[e]; if (pair) b = base[e + 1]; }
}
template <int N, bool NS>
FDB_DEV void store2(double* __restrict__ base, int64_t ld, int64_t e, bool pair, bool vec, const Dual<double, N, NS>& a, const Dual<double, N, NS>& b) {
    if (pair && vec) {
        *reinterpret_cast<double2*>(base + e) = make_double2(a.v, b.v);
#pragma unroll
        for (int k = 0; k < N; k++) *reinterpret_cast<double2*>(base + (int64_t)(k + 1) * ld + e) = make_double2(a.p[k], b.p[k]);
    } else {
        base[e] = a.v;
#pragma unroll
        for (int k = 0; k < N; k++) base[(int64_t)(k + 1) * ld + e] = a.p[k];
        if (pair) {
            base[e + 1] = b.v;
#pragma unroll
            for (int k = 0; k < N; k++) base[(int64_t)(k + 1) * ld + e + 1] = b.p[k];
        }
    }
}

template <int N, bool NS> FDB_DEV Dual<double, N, NS> apply1(int op, const Dual<double, N, NS>& x) {
    switch (op) {
        case FDB_NEG: return -x;
        case FDB_EXP: return exp_v(x);
        case FDB_LOG: return log_v(x);
        case FDB_SIN: return sin_v(x);
        case FDB_COS: return cos_v(x);
        case FDB_TANH: return tanh_v(x);
        case FDB_SQRT: return sqrt_v(x);
        case FDB_ABS: return abs_v(x);
        case FDB_ABS2: return abs2_v(x);
        case FDB_INV: return inv_v(x);
        case FDB_POW0: return pow0(x);
        case FDB_POW1: return pow1(x);
        case FDB_POW2: return pow2(x);
        case FDB_POW3: return pow3(x);
        case FDB_TAN: return tan_v(x);
        case FDB_EXP2: return exp2_v(x);
        case FDB_LOG2: return log2_v(x);
        case FDB_LOG10: return log10_v(x);
        case FDB_LOG1P: return log1p_v(x);
        case FDB_EXPM1: return expm1_v(x);
        case FDB_SINH: return sinh_v(x);
        case FDB_COSH: return cosh_v(x);
        case FDB_ASIN: return asin_v(x);
        case FDB_ACOS: return acos_v(x);
        case FDB_ATAN: return atan_v(x);
        default: return cbrt_v(x);
    }
}
template <int N, bool NS> FDB_DEV Dual<double, N, NS> apply2_dd(int op, const Dual<double, N, NS>& x, const Dual<double, N, NS>& y) {
    switch (op) {
        case FDB_ADD: return x + y;
        case FDB_SUB: return x - y;
        case FDB_MUL: return x * y;
        case FDB_DIV: return x / y;
        case FDB_POW: return pow_v(x, y);
        case FDB_MAX: return max_v(x, y);
        case FDB_MIN: return min_v(x, y);
        case FDB_ATAN2: return atan2_v(x, y);
        default: return hypot_v(x, y);
    }
}
template <int N, bool NS> FDB_DEV Dual<double, N, NS> apply2_dr(int op, const Dual<double, N, NS>& x, double s) {
    switch (op) {
        case FDB_ADD: return x + s;
        case FDB_SUB: return x - s;
        case FDB_MUL: return x * s;
        case FDB_DIV: return x / s;
        default: return pow_v(x, s);
    }
}
template <int N, bool NS> FDB_DEV Dual<double, N, NS> apply2_rd(int op, double s, const Dual<double, N, NS>& y) {
    switch (op) {
        case FDB_ADD: return s + y;
        case FDB_SUB: return s - y;
        case FDB_MUL: return s * y;
        case FDB_DIV: return s / y;
        default: return pow_v(s, y);
    }
}

template <int N, bool NS>
__global__ void __launch_bounds__(256) map1_kernel(int op, double* __restrict__ out, int64_t ldo, const double* __restrict__ a,
                                                    int64_t lda, int64_t n, bool vec) {
    int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 2;
    if (e >= n) return;
    bool pair = e + 1 < n;
    Dual2<N, NS> x = load2<N, NS>(a, lda, e, pair, false, vec);
    Dual<double, N, NS> r0 = apply1(op, x.a), r1 = r0;
    if (pair) r1 = apply1(op, x.b);
    store2<N, NS>(out, ldo, e, pair, vec, r0, r1);
}

// KIND 0: Dual (op) Dual, 1: Dual (op) Real, 2: Real (op) Dual.  The Real operand is a
// plain array (rb != null) or the scalar s.
template <int N, bool NS, int KIND>
__global__ void __launch_bounds__(256) map2_kernel(int op, double* __restrict__ out, int64_t ldo, const double* __restrict__ a,
                                                    int64_t lda, bool abcast, const double* __restrict__ b, int64_t ldb,
                                                    bool bbcast, double s, int64_t n, bool vec) {
    int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 2;
    if (e >= n) return;
    bool pair = e + 1 < n;
    Dual<double, N, NS> r0, r1;
    if (KIND == 0) {
        Dual2<N, NS> x = load2<N, NS>(a, lda, e, pair, abcast, vec);
        Dual2<N, NS> y = load2<N, NS>(b, ldb, e, pair, bbcast, vec);
        r0 = apply2_dd(op, x.a, y.a); r1 = r0;
        if (pair) r1 = apply2_dd(op, x.b, y.b);
    } else {
        Dual2<N, NS> x = load2<N, NS>(a, lda, e, pair, abcast, vec);   // `a` is always the Dual operand
        double s0 = s, s1 = s;
        if (b) loadreal2(b, e, pair, bbcast, vec, s0, s1);
        if (KIND == 1) { r0 = apply2_dr(op, x.a, s0); r1 = r0; if (pair) r1 = apply2_dr(op, x.b, s1); }
        else           { r0 = apply2_rd(op, s0, x.a); r1 = r0; if (pair) r1 = apply2_rd(op, s1, x.b); }
    }
    store2<N, NS>(out, ldo, e, pair, vec, r0, r1);
}

template <int N, bool NS>
__global__ void __launch_bounds__(256) fma_kernel(int kind, double* __restrict__ out, int64_t ldo, const double* __restrict__ a, int64_t lda,
                                                   const double* __restrict__ b, int64_t ldb, const double* __restrict__ c, int64_t ldc,
                                                   double s, int64_t n, bool vec) {
    int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 2;
    if (e >= n) return;
    bool pair = e + 1 < n;
    Dual2<N, NS> x = load2<N, NS>(a, lda, e, pair, false, vec);
    Dual<double, N, NS> r0, r1;
    if (kind == 0) {
        Dual2<N, NS> y = load2<N, NS>(b, ldb, e, pair, false, vec), z = load2<N, NS>(c, ldc, e, pair, false, vec);
        r0 = fma_xyz(x.a, y.a, z.a); r1 = fma_xyz(x.b, y.b, z.b);
    } else if (kind == 1) {
        Dual2<N, NS> y = load2<N, NS>(b, ldb, e, pair, false, vec);
        r0 = fma_xy(x.a, y.a, s); r1 = fma_xy(x.b, y.b, s);
    } else {
        Dual2<N, NS> z = load2<N, NS>(c, ldc, e, pair, false, vec);
        r0 = fma_xz(x.a, s, z.a); r1 = fma_xz(x.b, s, z.b);
    }
    store2<N, NS>(out, ldo, e, pair, vec, r0, r1);
}

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
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(mapreduce_tma_kernel<F, N, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail(ctx, FDB_ERR_CUDA, "cudaFuncSetAttribute(mapreduce_tma_kernel): %s", cudaGetErrorString(e));
        configured = true;
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

static bool vec_ok(const fdb_array* a) {
    return a == nullptr || ((((uintptr_t)a->data) & 15) == 0 && (a->ld & 1) == 0);
}
static int check_same(fdb_ctx* ctx, const fdb_array* out, const fdb_array* a, const char* what) {
    if (!ctx || !out || !a) return fail(ctx, FDB_ERR_BADARG, "%s: null argument", what);
    if (out->level != 1 || a->level != 1) return fail(ctx, FDB_ERR_UNSUPPORTED, "%s: operands must be level-1 Dual arrays", what);
    if (out->N != a->N) return fail(ctx, FDB_ERR_BADARG, "%s: chunk sizes differ (%d vs %d)", what, out->N, a->N);
    return FDB_OK;
}

extern "C" int fdb_map1(fdb_ctx* ctx, int op, fdb_array* out, const fdb_array* a) {
    int rc = check_same(ctx, out, a, "fdb_map1"); if (rc) return rc;
    if (op < FDB_NEG || op > FDB_CBRT) return fail(ctx, FDB_ERR_UNSUPPORTED, "fdb_map1: unknown op %d", op);
    if (out->n != a->n) return fail(ctx, FDB_ERR_SHAPE, "fdb_map1: DimensionMismatch");
    if (a->n == 0) return FDB_OK;
    unsigned grid = (unsigned)cdiv(cdiv(a->n, 2), 256);
    const bool vec = vec_ok(out) && vec_ok(a);
    FDB_DISPATCH_N(a->N, FDB_DISPATCH_NS(ctx->nansafe,
        (map1_kernel<NN, NS><<<grid, 256, 0, ctx->stream>>>(op, out->data, out->ld, a->data, a->ld, a->n, vec))));
    return finish(ctx, 1, "fdb_map1");
}

extern "C" int fdb_map2(fdb_ctx* ctx, int op, fdb_array* out, const fdb_array* a, const fdb_array* b) {
    if (!ctx || !out || !a || !b) return fail(ctx, FDB_ERR_BADARG, "fdb_map2: null argument");
    if (op < FDB_ADD || op > FDB_HYPOT) return fail(ctx, FDB_ERR_UNSUPPORTED, "fdb_map2: unknown op %d", op);
    if (out->level != 1 || a->level > 1 || b->level > 1 || (a->level == 0 && b->level == 0))
        return fail(ctx, FDB_ERR_UNSUPPORTED, "fdb_map2: need a level-1 output and at least one level-1 operand");
    int N = out->N;
    if ((a->level == 1 && a->N != N) || (b->level == 1 && b->N != N)) return fail(ctx, FDB_ERR_BADARG, "fdb_map2: chunk sizes differ");
    int64_t n = out->n;
    bool ab = (a->n == 1 && n != 1), bb = (b->n == 1 && n != 1);
    if ((!ab && a->n != n) || (!bb && b->n != n)) return fail(ctx, FDB_ERR_SHAPE, "fdb_map2: DimensionMismatch");
    if (n == 0) return FDB_OK;
    if ((a->level == 0 || b->level == 0) && op > FDB_POW) return fail(ctx, FDB_ERR_UNSUPPORTED, "fdb_map2: max/min need two Dual operands");
    unsigned grid = (unsigned)cdiv(cdiv(n, 2), 256);
    const bool vec = vec_ok(out) && vec_ok(a) && vec_ok(b);
    if (a->level == 1 && b->level == 1) {
        FDB_DISPATCH_N(N, FDB_DISPATCH_NS(ctx->nansafe,
            (map2_kernel<NN, NS, 0><<<grid, 256, 0, ctx->stream>>>(op, out->data, out->ld, a->data, a->ld, ab, b->data, b->ld, bb, 0.0, n, vec))));
    } else if (a->level == 1) {
        FDB_DISPATCH_N(N, FDB_DISPATCH_NS(ctx->nansafe,
            (map2_kernel<NN, NS, 1><<<grid, 256, 0, ctx->stream>>>(op, out->data, out->ld, a->data, a->ld, ab, b->data, b->ld, bb, 0.0, n, vec))));
    } else {
        FDB_DISPATCH_N(N, FDB_DISPATCH_NS(ctx->nansafe,
            (map2_kernel<NN, NS, 2><<<grid, 256, 0, ctx->stream>>>(op, out->data, out->ld, b->data, b->ld, bb, a->data, a->ld, ab, 0.0, n, vec))));
    }
    return finish(ctx, 1, "fdb_map2");
}

extern "C" int fdb_map2_scalar(fdb_ctx* ctx, int op, fdb_array* out, const fdb_array* a, double s, int scalar_first) {
    int rc = check_same(ctx, out, a, "fdb_map2_scalar"); if (rc) return rc;
    if (op < FDB_ADD || op > FDB_POW) return fail(ctx, FDB_ERR_UNSUPPORTED, "fdb_map2_scalar: unknown op %d", op);
    if (out->n != a->n) return fail(ctx, FDB_ERR_SHAPE, "fdb_map2_scalar: DimensionMismatch");
    if (a->n == 0) return FDB_OK;
    unsigned grid = (unsigned)cdiv(cdiv(a->n, 2), 256);
    const bool vec = vec_ok(out) && vec_ok(a);
    if (!scalar_first) {
        FDB_DISPATCH_N(a->N, FDB_DISPATCH_NS(ctx->nansafe,
            (map2_kernel<NN, NS, 1><<<grid, 256, 0, ctx->stream>>>(op, out->data, out->ld, a->data, a->ld, false, nullptr, 0, false, s, a->n, vec))));
    } else {
        FDB_DISPATCH_N(a->N, FDB_DISPATCH_NS(ctx->nansafe,
            (map2_kernel<NN, NS, 2><<<grid, 256, 0, ctx->stream>>>(op, out->data, out->ld, a->data, a->ld, false, nullptr, 0, false, s, a->n, vec))));
    }
    return finish(ctx, 1, "fdb_map2_scalar");
}

extern "C" int fdb_fma(fdb_ctx* ctx, int kind, fdb_array* out, const fdb_array* a, const fdb_array* b, const fdb_array* c, double s) {
    int rc = check_same(ctx, out, a, "fdb_fma"); if (rc) return rc;
    if (kind < 0 || kind > 2) return fail(ctx, FDB_ERR_BADARG, "fdb_fma: kind %d", kind);
    if ((kind != 2 && !b) || (kind != 1 && !c)) return fail(ctx, FDB_ERR_BADARG, "fdb_fma: missing operand");
    if (out->n != a->n || (kind != 2 && (b->n != a->n || b->N != a->N || b->level != 1)) ||
        (kind != 1 && (c->n != a->n || c->N != a->N || c->level != 1)))
        return fail(ctx, FDB_ERR_SHAPE, "fdb_fma: DimensionMismatch");
    if (a->n == 0) return FDB_OK;
    unsigned grid = (unsigned)cdiv(cdiv(a->n, 2), 256);
    const bool vec = vec_ok(out) && vec_ok(a) && vec_ok(b) && vec_ok(c);
    FDB_DISPATCH_N(a->N, FDB_DISPATCH_NS(ctx->nansafe,
        (fma_kernel<NN, NS><<<grid, 256, 0, ctx->stream>>>(kind, out->data, out->ld, a->data, a->ld, b ? b->data : nullptr,
                                                            b ? b->ld : 0, c ? c->data : nullptr, c ? c->ld : 0, s, a->n, vec))));
    return finish(ctx, 1, "fdb_fma");
}

static int check_reduce(fdb_ctx* ctx, fdb_array* out1, const fdb_array* a, const char* what) {
    int rc = check_same(ctx, out1, a, what); if (rc) return rc;
    if (out1->n != 1) return fail(ctx, FDB_ERR_SHAPE, "%s: output must have length 1", what);
    return FDB_OK;
}
extern "C" int fdb_reduce(fdb_ctx* ctx, int redop, fdb_array* out1, const fdb_array* a) {
    int rc = check_reduce(ctx, out1, a, "fdb_reduce"); if (rc) return rc;
    if (redop == FDB_SUM) return mapreduce_launch<SumAll>(ctx, out1, a, "fdb_reduce(sum)");
    if (redop == FDB_PROD) return mapreduce_launch<ProdAll>(ctx, out1, a, "fdb_reduce(prod)");
    return fail(ctx, FDB_ERR_UNSUPPORTED, "fdb_reduce: unknown reduction %d", redop);
}
extern "C" int fdb_eval_scalar_fn(fdb_ctx* ctx, int fn, fdb_array* out1, const fdb_array* xdual) {
    int rc = check_reduce(ctx, out1, xdual, "fdb_eval_scalar_fn"); if (rc) return rc;
    switch (fn) {
        case FDB_FN_ROSENBROCK: return mapreduce_launch<Rosenbrock>(ctx, out1, xdual, "fdb_eval_scalar_fn(rosenbrock)");
        case FDB_FN_ACKLEY: return mapreduce_launch<Ackley>(ctx, out1, xdual, "fdb_eval_scalar_fn(ackley)");
        case FDB_FN_SUMSQ: return mapreduce_launch<SumSq>(ctx, out1, xdual, "fdb_eval_scalar_fn(sumsq)");
        case FDB_FN_PROD: return mapreduce_launch<ProdAll>(ctx, out1, xdual, "fdb_eval_scalar_fn(prod)");
    }
    return fail(ctx, FDB_ERR_UNSUPPORTED, "fdb_eval_scalar_fn: unknown target %d", fn);
}
