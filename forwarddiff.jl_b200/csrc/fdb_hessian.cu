// fdb_hessian.cu -- batched nested-dual Hessian: replaces src/hessian.jl:14-71, i.e.
// jacobian(y -> gradient(f, y, gcfg), x, jcfg) with Dual{T,Dual{T,V,N},N} work arrays
// (src/config.jl:198-201), every (outer chunk, inner chunk) pair in one launch.
//
// Block (ci, co) evaluates f once on the nested duals that the reference would have
// materialised for inner chunk ci and outer chunk co, and stores
//     H[istart+k, ostart+j] = partials(partials(y, k), j)      (inner gradient.jl:74-81,
//                                                               outer jacobian.jl:109-118)
// plus, from the last outer chunk, grad[istart+k] = value(partials(y,k)) and
// value = value(value(y)) (hessian.jl:50-55).
//
// Two-level layout: a nested dual has (N+1)^2 lanes [a][b], a = inner index (0 = value),
// b = outer index.  Terms that read no seeded position of either window have identical
// lanes in each of the four quadrants {a=0,a>0} x {b=0,b>0}; they are evaluated on
// Dual{Dual{1},1} (four lanes) by all threads.  The few terms near a window are evaluated
// per inner lane k on Dual{Dual{N},1}; their contributions are summed in term order.
// Restricted to additive targets (K = 1, combine = +, finalize = identity).
#include "fdb_internal.h"
#include "fdb_mirror.cuh"
#include "fdb_targets.cuh"

namespace fdb {

// far-from-both-windows loader: Dual{Dual{1},1} with all partials zero
template <bool NS> struct FFLoader {
    typedef Dual<Dual<double, 1, NS>, 1, NS> DO;
    const double* x;
    FDB_DEV DO operator()(int64_t i) const {
        DO d; d.v.v = x[i]; d.v.p[0] = 0.0; d.p[0].v = 0.0; d.p[0].p[0] = 0.0; return d;
    }
};
// general loader for inner lane k: value = outer-seeded Dual{N}; the single inner partial lane
// carries one(Dual{N}) iff pos is slot k of the inner window (src/apiutils.jl:125-143 with
// Partials{N,Dual{T,V,N}} seeds, src/config.jl:225-226)
template <int N, bool NS> struct NearLoader {
    typedef Dual<double, N, NS> DI;
    typedef Dual<DI, 1, NS> DO;
    const double* x; int64_t ostart; int ocs; int64_t istart; int ics; int k;
    FDB_DEV DO operator()(int64_t pos) const {
        DO d; d.v = seeded<N, NS>(x[pos], pos, ostart, ocs);
        const bool hot = (pos - istart == k) && (k < ics);
        d.p[0].v = hot ? 1.0 : 0.0;
#pragma unroll
        for (int j = 0; j < N; j++) d.p[0].p[j] = 0.0;
        return d;
    }
};

template <class F, int N, bool NS, int THREADS>
__global__ void __launch_bounds__(THREADS) hessian_sweep_kernel(const double* __restrict__ x, int64_t n, int64_t outer_lo,
                                                                int64_t nchunks_total, int lastchunksize, double* __restrict__ H,
                                                                int64_t ldH, double* __restrict__ grad, double* __restrict__ value_out,
                                                                const __grid_constant__ fdb_mirrors M) {
    typedef Dual<double, N, NS> DI;
    typedef Dual<DI, 1, NS> DG;
    typedef Dual<Dual<double, 1, NS>, 1, NS> DF;
    constexpr int MAXNEAR = 2 * (N + F::HALO);
    __shared__ double s_p[MAXNEAR][N][N + 1];   // per near term, per inner lane k: partials(y,k) as Dual{N}
    __shared__ double s_v[MAXNEAR][N + 1];      // per near term: value(y) as Dual{N}
    __shared__ double s_ff[4];
    __shared__ double s_red[THREADS / 32][4];

    const int64_t ci = blockIdx.x, co = outer_lo + blockIdx.y;
    const bool ilast = (ci == nchunks_total - 1), olast = (co == nchunks_total - 1);
    const int ics = ilast ? lastchunksize : N, ocs = olast ? lastchunksize : N;
    const int64_t istart = ilast ? n - lastchunksize : ci * N;
    const int64_t ostart = olast ? n - lastchunksize : co * N;
    const int64_t nt = F::nterms(n);

    // near-term ranges (terms reading a seeded position): A for the inner window, B for the outer
    int64_t a0 = istart - F::HALO, a1 = istart + ics, b0 = ostart - F::HALO, b1 = ostart + ocs;
    if (a0 < 0) a0 = 0; if (b0 < 0) b0 = 0; if (a1 > nt) a1 = nt; if (b1 > nt) b1 = nt;
    if (a1 < a0) a1 = a0; if (b1 < b0) b1 = b0;
    if (b0 <= a1 && a0 <= b1) { a0 = a0 < b0 ? a0 : b0; a1 = a1 > b1 ? a1 : b1; b0 = b1 = a1; }   // merge overlapping ranges
    const int na = (int)(a1 - a0), nb = (int)(b1 - b0), nnear = na + nb;

    // ---- phase 1: terms far from both windows, four shared lanes --------------------------------
    DF accff[1]; accff[0].v.v = 0.0; accff[0].v.p[0] = 0.0; accff[0].p[0].v = 0.0; accff[0].p[0].p[0] = 0.0;
    FFLoader<NS> xf{x};
    for (int64_t i = threadIdx.x; i < nt; i += THREADS) {
        const bool near = (i >= a0 && i < a1) || (i >= b0 && i < b1);
        if (!near) F::term(xf, i, n, accff);
    }
    {
        double r[4] = {accff[0].v.v, accff[0].v.p[0], accff[0].p[0].v, accff[0].p[0].p[0]};
#pragma unroll
        for (int q = 0; q < 4; q++) {
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) r[q] = r[q] + shfl_down_d(r[q], off);
            if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5][q] = r[q];
        }
    }
    // ---- phase 2: near terms, one work item per (term, inner lane k) ------------------------------
    for (int w = threadIdx.x; w < nnear * N; w += THREADS) {
        const int t = w / N, k = w - t * N;
        const int64_t i = t < na ? a0 + t : b0 + (t - na);
        NearLoader<N, NS> xn{x, ostart, ocs, istart, ics, k};
        DG acc[1];
        acc[0].v.v = 0.0; acc[0].p[0].v = 0.0;
#pragma unroll
        for (int j = 0; j < N; j++) { acc[0].v.p[j] = 0.0; acc[0].p[0].p[j] = 0.0; }
        F::term(xn, i, n, acc);
        s_p[t][k][0] = acc[0].p[0].v;
#pragma unroll
        for (int j = 0; j < N; j++) s_p[t][k][j + 1] = acc[0].p[0].p[j];
        if (k == 0) {
            s_v[t][0] = acc[0].v.v;
#pragma unroll
            for (int j = 0; j < N; j++) s_v[t][j + 1] = acc[0].v.p[j];
        }
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        double s = 0.0;
        for (int wq = 0; wq < THREADS / 32; wq++) s = s + s_red[wq][threadIdx.x];
        s_ff[threadIdx.x] = s;
    }
    __syncthreads();
    // ---- phase 3: fold and store (fused extract_gradient_chunk! + extract_jacobian_chunk!) --------
    // lanes: s_ff[0] = value.value, [1] = value.outer-partial, [2] = inner-partial.value, [3] = inner-partial.outer-partial
    for (int e = threadIdx.x; e < N * (N + 1); e += THREADS) {
        const int k = e / (N + 1), j = e - k * (N + 1);     // j = 0: value of partials(y,k); j >= 1: outer partial j
        if (k >= ics) continue;
        double s = 0.0;
        for (int t = 0; t < nnear; t++) s = s + s_p[t][k][j];
        s = s + (j == 0 ? s_ff[2] : s_ff[3]);
        if (j == 0) { if (olast && grad) mstore(M, &grad[istart + k], s); }
        else if (j - 1 < ocs && H) mstore(M, &H[(istart + k) + (ostart + j - 1) * ldH], s);
    }
    if (threadIdx.x == 0 && olast && ilast && value_out) {
        double s = 0.0;
        for (int t = 0; t < nnear; t++) s = s + s_v[t][0];
        mstore(M, value_out, s + s_ff[0]);
    }
}

template <class F>
static int hessian_launch(fdb_ctx* ctx, const fdb_array* x, int N, int64_t lo, int64_t hi, int64_t nchunks, int lcs, fdb_array* H,
                          int64_t ldH, fdb_array* grad, fdb_array* value_dev, const char* what) {
    constexpr int T = 128;
    dim3 grid((unsigned)nchunks, (unsigned)(hi - lo));
    double* Hp = H ? H->data : nullptr; double* gp = grad ? grad->data : nullptr; double* vp = value_dev ? value_dev->data : nullptr;
    fdb_mirrors M;
    const fdb_array* outs[3] = {H, grad, value_dev};
    { int mrc = mirrors_for(ctx, &M, outs, 3, what); if (mrc) return mrc; }
    FDB_DISPATCH_N(N, FDB_DISPATCH_NS(ctx->nansafe,
        (hessian_sweep_kernel<F, NN, NS, T><<<grid, T, 0, ctx->stream>>>(x->data, x->n, lo, nchunks, lcs, Hp, ldH, gp, vp, M))));
    return finish(ctx, 1, what);
}

}  // namespace fdb

using namespace fdb;

extern "C" int fdb_hessian_chunked(fdb_ctx* ctx, int fn, const fdb_array* x, int N, int64_t chunk_lo, int64_t chunk_hi,
                                   fdb_array* H, int64_t ldH, fdb_array* grad, fdb_array* value_dev) {
    if (!ctx || !x) return fail(ctx, FDB_ERR_BADARG, "fdb_hessian_chunked: null argument");
    if (x->level != 0 || (H && H->level != 0) || (grad && grad->level != 0) || (value_dev && value_dev->level != 0))
        return fail(ctx, FDB_ERR_BADARG, "fdb_hessian_chunked: x, H, grad, value must be plain arrays");
    const int64_t n = x->n;
    if (N < 1 || N > FDB_MAX_CHUNK) return fail(ctx, FDB_ERR_BADARG, "chunk size N=%d outside 1..12", N);
    if (n < N) return fail(ctx, FDB_ERR_CHUNK, "chunk size cannot be greater than ForwardDiff.structural_length(x) (%d > %lld)", N, (long long)n);
    if (H && (ldH < n || H->n < ldH * (n - 1) + n)) return fail(ctx, FDB_ERR_SHAPE, "DimensionMismatch: Hessian result must hold %lld x %lld", (long long)n, (long long)n);
    if (grad && grad->n != n) return fail(ctx, FDB_ERR_SHAPE, "DimensionMismatch: gradient result has %lld entries, x has %lld", (long long)grad->n, (long long)n);
    Plan p = make_plan(n, N);
    const int64_t nchunks = (n == N) ? 1 : p.nchunks;
    if (chunk_hi < 0) chunk_hi = nchunks;
    if (chunk_lo < 0 || chunk_lo > chunk_hi || chunk_hi > nchunks) return fail(ctx, FDB_ERR_BADARG, "bad chunk range");
    if (chunk_lo == chunk_hi) return FDB_OK;
    if (nchunks > 2147483647LL || chunk_hi - chunk_lo > 65535) return fail(ctx, FDB_ERR_UNSUPPORTED, "fdb_hessian_chunked: more than 65535 outer chunks per call");
    switch (fn) {
        case FDB_FN_ROSENBROCK: return hessian_launch<Rosenbrock>(ctx, x, N, chunk_lo, chunk_hi, nchunks, (int)p.lastchunksize, H, ldH, grad, value_dev, "fdb_hessian_chunked(rosenbrock)");
        case FDB_FN_SUMSQ: return hessian_launch<SumSq>(ctx, x, N, chunk_lo, chunk_hi, nchunks, (int)p.lastchunksize, H, ldH, grad, value_dev, "fdb_hessian_chunked(sumsq)");
    }
    return fail(ctx, FDB_ERR_UNSUPPORTED, "fdb_hessian_chunked: target %d has no nested-dual kernel (additive targets only)", fn);
}

extern "C" int fdb_hessian_host(fdb_ctx* ctx, int fn, const double* x_host, int64_t xlen, int N, int64_t chunk_lo, int64_t chunk_hi,
                                double* H_host, int64_t ldH, double* grad_host, double* value_host) {
    if (!ctx || !x_host || !H_host) return fail(ctx, FDB_ERR_BADARG, "fdb_hessian_host: null argument");
    if (xlen == 0) { if (value_host) *value_host = 0.0; return FDB_OK; }
    if (N < 1 || N > FDB_MAX_CHUNK) return fail(ctx, FDB_ERR_BADARG, "chunk size N=%d outside 1..12", N);
    if (xlen < N) return fail(ctx, FDB_ERR_CHUNK, "chunk size cannot be greater than ForwardDiff.structural_length(x) (%d > %lld)", N, (long long)xlen);
    if (ldH < xlen) return fail(ctx, FDB_ERR_SHAPE, "DimensionMismatch: ldH < n");
    Plan p = make_plan(xlen, N);
    int64_t nchunks = (xlen == N) ? 1 : p.nchunks;
    if (chunk_hi < 0) chunk_hi = nchunks;
    if (chunk_lo < 0 || chunk_lo > chunk_hi || chunk_hi > nchunks) return fail(ctx, FDB_ERR_BADARG, "bad chunk range");
    int64_t ld = round_up(xlen, 32);
    size_t need = (size_t)(2 * ld + 32 + ld * xlen) * sizeof(double);
    int rc = ensure_scratch(ctx, need); if (rc) return rc;
    fdb_array xa, ga, va, ha;
    xa.ctx = ga.ctx = va.ctx = ha.ctx = ctx; xa.owned = ga.owned = va.owned = ha.owned = false;
    xa.data = ctx->scratch; xa.n = xlen; xa.ld = ld;
    ga.data = ctx->scratch + ld; ga.n = xlen; ga.ld = ld;
    va.data = ctx->scratch + 2 * ld; va.n = 1; va.ld = 32;
    ha.data = ctx->scratch + 2 * ld + 32; ha.n = ld * xlen; ha.ld = ld * xlen;
    FDB_CUDA(ctx, cudaMemcpyAsync(xa.data, x_host, (size_t)xlen * 8, cudaMemcpyHostToDevice, ctx->stream));
    bool was_async = ctx->async; ctx->async = true;
    rc = fdb_hessian_chunked(ctx, fn, &xa, N, chunk_lo, chunk_hi, &ha, ld, &ga, &va);
    ctx->async = was_async;
    if (rc) return rc;
    int64_t c0 = chunk_lo * N, c1 = (chunk_hi == nchunks) ? xlen : chunk_hi * N;   // columns owned by the requested outer chunks
    if (c1 > c0)
        FDB_CUDA(ctx, cudaMemcpy2DAsync(H_host + c0 * ldH, (size_t)ldH * 8, ha.data + c0 * ld, (size_t)ld * 8, (size_t)xlen * 8, (size_t)(c1 - c0),
                                        cudaMemcpyDeviceToHost, ctx->stream));
    if (chunk_hi == nchunks) {
        if (grad_host) FDB_CUDA(ctx, cudaMemcpyAsync(grad_host, ga.data, (size_t)xlen * 8, cudaMemcpyDeviceToHost, ctx->stream));
        if (value_host) FDB_CUDA(ctx, cudaMemcpyAsync(value_host, va.data, 8, cudaMemcpyDeviceToHost, ctx->stream));
    }
    FDB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FDB_OK;
}
