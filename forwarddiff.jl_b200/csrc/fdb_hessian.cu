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
#include "fdb_hessian_kernel.cuh"

namespace fdb {

template <class F>
static int hessian_launch(fdb_ctx* ctx, const fdb_array* x, int N, int64_t lo, int64_t hi, int64_t nchunks, int lcs, fdb_array* H,
                          int64_t ldH, fdb_array* grad, fdb_array* value_dev, const char* what) {
    constexpr int T = 128;      // 160 threads (one near-phase pass at N = 8 instead of two) measured slower: 0.36 vs 0.31 ms at C5
    // plain sums: a block takes HESS_GROUP inner chunks (the far terms are shared between them); general targets: one pair per block
    dim3 grid((unsigned)(F::PLAIN_SUM ? cdiv(nchunks, (int64_t)HESS_GROUP) : nchunks), (unsigned)(hi - lo));
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
    FDB_ENTER(ctx);
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
        case FDB_FN_ACKLEY: return hessian_launch<Ackley>(ctx, x, N, chunk_lo, chunk_hi, nchunks, (int)p.lastchunksize, H, ldH, grad, value_dev, "fdb_hessian_chunked(ackley)");
        case FDB_FN_SUMSQ: return hessian_launch<SumSq>(ctx, x, N, chunk_lo, chunk_hi, nchunks, (int)p.lastchunksize, H, ldH, grad, value_dev, "fdb_hessian_chunked(sumsq)");
    }
    return fail(ctx, FDB_ERR_UNSUPPORTED, "fdb_hessian_chunked: target %d has no nested-dual kernel (additive targets only)", fn);
}

extern "C" int fdb_hessian_host(fdb_ctx* ctx, int fn, const double* x_host, int64_t xlen, int N, int64_t chunk_lo, int64_t chunk_hi,
                                double* H_host, int64_t ldH, double* grad_host, double* value_host) {
    FDB_ENTER(ctx);
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
