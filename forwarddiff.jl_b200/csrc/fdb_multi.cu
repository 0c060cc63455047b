// fdb_multi.cu -- ONE host process driving several GPUs, and the host-buffer Jacobian entry point.
//
// A Julia session that calls `ForwardDiff.gradient!(out, f, x, cfg)` (src/gradient.jl:35-44),
// `jacobian!` (src/jacobian.jl:57-87) or `hessian!` (src/hessian.jl:31-37) is a single process.  The chunk
// sweeps are independent, so a group of contexts -- one per visible GPU, all owned by this
// process -- takes contiguous chunk ranges (fdb_plan_chunks), and
//   * the `fdb_multi_*_host` drivers give every device its copy of the inputs, launch every device's sweeps
//     before waiting for any of them, and copy each device's result slice (gradient entries / Jacobian or Hessian
//     column block) straight into the caller's host result: no inter-GPU traffic at all;
//   * the `fdb_multi_*_resident` drivers leave the COMPLETE result in every device's exchange arena: the sweep
//     kernels store each extracted entry into all peers' arenas (fdb_mirror.cuh), the arenas being linked with
//     cudaDeviceEnablePeerAccess (fdb_exchange_connect_local) instead of CUDA IPC, followed by the flag barrier.
// Every launch is asynchronous until all devices have been issued; only then does the host wait.
#include <string.h>

#include "fdb_internal.h"

struct fdb_multi {
    int n = 0;
    fdb_ctx* ctx[FDB_MAX_RANKS] = {nullptr};
    std::string err;
};

namespace fdb {

static int mfail(fdb_multi* g, int code, const char* fmt, ...) __attribute__((format(printf, 3, 4)));
static int mfail(fdb_multi* g, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (g) g->err = buf; else g_init_error = buf;
    return code;
}
// a per-context failure surfaces on the group with the device it happened on
static int mpropagate(fdb_multi* g, int i, int rc) {
    if (rc) g->err = "device " + std::to_string(g->ctx[i]->device) + ": " + g->ctx[i]->err;
    return rc;
}

static int ensure_hostws(fdb_ctx* ctx, size_t bytes) {
    if (ctx->hostws_bytes >= bytes) return FDB_OK;
    if (ctx->hostws) { cudaStreamSynchronize(ctx->stream); cudaFree(ctx->hostws); ctx->hostws = nullptr; ctx->hostws_bytes = 0; }
    FDB_CUDA(ctx, cudaMalloc(&ctx->hostws, bytes + 4096));
    ctx->hostws_bytes = bytes + 4096;
    return FDB_OK;
}

static fdb_array plain(fdb_ctx* ctx, double* p, int64_t n, int64_t ld) {
    fdb_array a;
    a.ctx = ctx; a.data = p; a.n = n; a.ld = ld; a.N = 0; a.level = 0; a.owned = false;
    return a;
}

static int64_t total_chunks(int64_t n, int N) { return n == N ? 1 : make_plan(n, N).nchunks; }   // N == n: vector mode, one sweep

// rank r's contiguous chunk range and the structural positions (result entries / columns) it owns: fdb_plan_chunks
static void rank_range(int64_t n, int N, int r, int w, int64_t* lo, int64_t* hi, int64_t* c0, int64_t* c1) {
    const int64_t nchunks = total_chunks(n, N), per = cdiv(nchunks, w);
    int64_t a = (int64_t)r * per, b = a + per;
    if (a > nchunks) a = nchunks;
    if (b > nchunks) b = nchunks;
    *lo = a; *hi = b;
    *c0 = a * N < n ? a * N : n;
    *c1 = (b == nchunks) ? n : b * N;
    if (a == b) *c1 = *c0;
}

struct AsyncScope {      // drivers called inside never wait for their own stream
    fdb_ctx* ctx; bool saved;
    explicit AsyncScope(fdb_ctx* c) : ctx(c), saved(c->async) { c->async = true; }
    ~AsyncScope() { ctx->async = saved; }
};

// One device's share of a host-buffer call, staged in ctx->hostws.
struct HostJob {
    fdb_ctx* ctx = nullptr;
    int64_t lo = 0, hi = 0, nchunks = 0, c0 = 0, c1 = 0, rows = 0, ld = 0;
    fdb_array x, A, b, out, y, g, v;
    double* block = nullptr;     // device copy of result columns [c0, c1)
};

// ---- gradient -------------------------------------------------------------------------------------------
static int grad_begin(fdb_ctx* ctx, int fn, const double* x_host, int64_t n, int N, int r, int w, HostJob& j) {
    DeviceGuard guard(ctx->device);
    j.ctx = ctx; j.nchunks = total_chunks(n, N);
    rank_range(n, N, r, w, &j.lo, &j.hi, &j.c0, &j.c1);
    if (j.lo == j.hi) return FDB_OK;
    const int64_t ld = round_up(n, 32);
    int rc = ensure_hostws(ctx, (size_t)(2 * ld + 32) * sizeof(double)); if (rc) return rc;
    j.x = plain(ctx, ctx->hostws, n, ld); j.g = plain(ctx, ctx->hostws + ld, n, ld); j.v = plain(ctx, ctx->hostws + 2 * ld, 1, 32);
    FDB_CUDA(ctx, cudaMemcpyAsync(j.x.data, x_host, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    AsyncScope as(ctx);
    return fdb_gradient_chunked(ctx, fn, &j.x, N, j.lo, j.hi, &j.g, &j.v);
}
static int grad_end(HostJob& j, double* grad_host, double* value_host) {
    fdb_ctx* ctx = j.ctx;
    if (j.lo == j.hi) return FDB_OK;
    DeviceGuard guard(ctx->device);
    if (j.c1 > j.c0) FDB_CUDA(ctx, cudaMemcpyAsync(grad_host + j.c0, j.g.data + j.c0, (size_t)(j.c1 - j.c0) * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (value_host && j.hi == j.nchunks) FDB_CUDA(ctx, cudaMemcpyAsync(value_host, j.v.data, 8, cudaMemcpyDeviceToHost, ctx->stream));   // gradient.jl:155
    FDB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return check_poison(ctx, "fdb_multi_gradient_host");
}

// ---- Jacobian -------------------------------------------------------------------------------------------
// Device staging: A (lda_dev = ld) | x | y | b | result column block [c0, c1) (m x ncols, leading dimension ld).
// The drivers address columns absolutely, so the block's base pointer is shifted back by c0 columns.
static int jac_begin(fdb_ctx* ctx, int fn, const double* x_host, int64_t n, int64_t m, int N, const double* A_host, int64_t lda,
                     const double* b_host, double alpha, int reuse_params, int64_t lo, int64_t hi, HostJob& j) {
    DeviceGuard guard(ctx->device);
    j.ctx = ctx; j.nchunks = total_chunks(n, N); j.rows = m;
    j.lo = lo; j.hi = hi; j.c0 = lo * N < n ? lo * N : n; j.c1 = (hi == j.nchunks) ? n : hi * N;
    if (j.lo == j.hi) return FDB_OK;
    const int64_t ldn = round_up(n, 32), ld = round_up(m, 32), ncols = j.c1 - j.c0;
    j.ld = ld;
    const bool hasA = A_host != nullptr;
    // with parameters the staging is sized for all n result columns, so that a later call with another chunk range (and
    // reuse_params = 1) finds A and b where they are
    const size_t need = (size_t)(ldn + 2 * ld + ld * (hasA ? n : ncols) + (hasA ? ld * n : 0)) * sizeof(double);
    if (reuse_params && ctx->hostws_bytes < need)
        return fail(ctx, FDB_ERR_BADARG, "fdb_jacobian_host: reuse_params = 1 but no previous call of this shape staged A and b on device %d", ctx->device);
    int rc = ensure_hostws(ctx, need); if (rc) return rc;
    double* p = ctx->hostws;
    if (hasA) {                          // A first: its place depends on (m, n) only, so reuse_params survives a change of chunk range
        j.A = plain(ctx, p, ld * (n - 1) + m, ld * (n - 1) + m); p += ld * n;
    }
    j.x = plain(ctx, p, n, ldn); p += ldn;
    j.y = plain(ctx, p, m, ld); p += ld;
    j.b = plain(ctx, p, m, ld); p += ld;
    j.block = p;
    j.out = plain(ctx, j.block - j.c0 * ld, ld * (n - 1) + m, ld * (n - 1) + m);
    FDB_CUDA(ctx, cudaMemcpyAsync(j.x.data, x_host, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    if (hasA && !reuse_params) {
        FDB_CUDA(ctx, cudaMemcpy2DAsync(j.A.data, (size_t)ld * 8, A_host, (size_t)lda * 8, (size_t)m * 8, (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
        if (b_host) FDB_CUDA(ctx, cudaMemcpyAsync(j.b.data, b_host, (size_t)m * 8, cudaMemcpyHostToDevice, ctx->stream));
    }
    AsyncScope as(ctx);
    return fdb_jacobian_chunked(ctx, fn, &j.x, m, N, hasA ? &j.A : nullptr, ld, (hasA && b_host) ? &j.b : nullptr, alpha, j.lo, j.hi, &j.out, ld, &j.y);
}
static int jac_end(HostJob& j, double* J_host, int64_t ldJ, double* y_host, const char* what) {
    fdb_ctx* ctx = j.ctx;
    if (j.lo == j.hi) return FDB_OK;
    DeviceGuard guard(ctx->device);
    if (j.c1 > j.c0)
        FDB_CUDA(ctx, cudaMemcpy2DAsync(J_host + j.c0 * ldJ, (size_t)ldJ * 8, j.block, (size_t)j.ld * 8, (size_t)j.rows * 8, (size_t)(j.c1 - j.c0),
                                        cudaMemcpyDeviceToHost, ctx->stream));
    if (y_host && j.hi == j.nchunks)    // map!(d -> value(T,d), y, ydual) from the last sweep, jacobian.jl:229
        FDB_CUDA(ctx, cudaMemcpyAsync(y_host, j.y.data, (size_t)j.rows * 8, cudaMemcpyDeviceToHost, ctx->stream));
    FDB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return check_poison(ctx, what);
}

// ---- Hessian --------------------------------------------------------------------------------------------
static int hess_begin(fdb_ctx* ctx, int fn, const double* x_host, int64_t n, int N, int r, int w, HostJob& j) {
    DeviceGuard guard(ctx->device);
    j.ctx = ctx; j.nchunks = total_chunks(n, N); j.rows = n;
    rank_range(n, N, r, w, &j.lo, &j.hi, &j.c0, &j.c1);
    if (j.lo == j.hi) return FDB_OK;
    const int64_t ld = round_up(n, 32), ncols = j.c1 - j.c0;
    j.ld = ld;
    int rc = ensure_hostws(ctx, (size_t)(2 * ld + 32 + ld * ncols) * sizeof(double)); if (rc) return rc;
    double* p = ctx->hostws;
    j.x = plain(ctx, p, n, ld); p += ld;
    j.g = plain(ctx, p, n, ld); p += ld;
    j.v = plain(ctx, p, 1, 32); p += 32;
    j.block = p;
    j.out = plain(ctx, j.block - j.c0 * ld, ld * (n - 1) + n, ld * (n - 1) + n);
    FDB_CUDA(ctx, cudaMemcpyAsync(j.x.data, x_host, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    AsyncScope as(ctx);
    return fdb_hessian_chunked(ctx, fn, &j.x, N, j.lo, j.hi, &j.out, ld, &j.g, &j.v);
}
static int hess_end(HostJob& j, double* H_host, int64_t ldH, double* grad_host, double* value_host) {
    fdb_ctx* ctx = j.ctx;
    if (j.lo == j.hi) return FDB_OK;
    DeviceGuard guard(ctx->device);
    if (j.c1 > j.c0)
        FDB_CUDA(ctx, cudaMemcpy2DAsync(H_host + j.c0 * ldH, (size_t)ldH * 8, j.block, (size_t)j.ld * 8, (size_t)j.rows * 8, (size_t)(j.c1 - j.c0),
                                        cudaMemcpyDeviceToHost, ctx->stream));
    if (j.hi == j.nchunks) {            // DiffResult value and gradient fall out of the last outer chunk, hessian.jl:50-55
        if (grad_host) FDB_CUDA(ctx, cudaMemcpyAsync(grad_host, j.g.data, (size_t)j.rows * 8, cudaMemcpyDeviceToHost, ctx->stream));
        if (value_host) FDB_CUDA(ctx, cudaMemcpyAsync(value_host, j.v.data, 8, cudaMemcpyDeviceToHost, ctx->stream));
    }
    FDB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return check_poison(ctx, "fdb_multi_hessian_host");
}

static int check_sizes(fdb_ctx* ctx, int64_t n, int N) {
    if (N < 1 || N > FDB_MAX_CHUNK) return fail(ctx, FDB_ERR_BADARG, "chunk size N=%d outside 1..12", N);
    if (n < N) return fail(ctx, FDB_ERR_CHUNK, "chunk size cannot be greater than ForwardDiff.structural_length(x) (%d > %lld)", N, (long long)n);
    return FDB_OK;
}

// after a failed begin some devices may already be running: drain them so the staging buffers can be reused
static void drain(fdb_multi* g) {
    for (int i = 0; i < g->n; i++) { DeviceGuard guard(g->ctx[i]->device); cudaStreamSynchronize(g->ctx[i]->stream); }
}

}  // namespace fdb

using namespace fdb;

// ===========================================================================================================
// single context: jacobian!(result, f, x, cfg) / jacobian!(result, f!, y, x, cfg) on host arrays
// ===========================================================================================================
extern "C" int fdb_jacobian_host(fdb_ctx* ctx, int fn, const double* x_host, int64_t xlen, int64_t m, int N, const double* A_host, int64_t lda,
                                 const double* b_host, double alpha, int reuse_params, int64_t chunk_lo, int64_t chunk_hi, double* J_host,
                                 int64_t ldJ, double* y_host) {
    FDB_ENTER(ctx);
    if (!ctx || !x_host || !J_host) return fail(ctx, FDB_ERR_BADARG, "fdb_jacobian_host: null argument");
    if (m <= 0 || ldJ < m || (A_host && lda < m)) return fail(ctx, FDB_ERR_SHAPE, "DimensionMismatch: result / A leading dimension smaller than m");
    int rc = check_sizes(ctx, xlen, N); if (rc) return rc;
    const int64_t nchunks = total_chunks(xlen, N);
    if (chunk_hi < 0) chunk_hi = nchunks;
    if (chunk_lo < 0 || chunk_lo > chunk_hi || chunk_hi > nchunks) return fail(ctx, FDB_ERR_BADARG, "bad chunk range");
    if (chunk_lo == chunk_hi) return FDB_OK;
    HostJob j;
    rc = jac_begin(ctx, fn, x_host, xlen, m, N, A_host, lda, b_host, alpha, reuse_params, chunk_lo, chunk_hi, j); if (rc) return rc;
    return jac_end(j, J_host, ldJ, y_host, "fdb_jacobian_host");
}

// pin / unpin a caller-owned host buffer (x, result): cudaMemcpyAsync on pinned memory is truly asynchronous, so the
// per-device copies of the fdb_multi_*_host drivers overlap instead of running one after the other
extern "C" int fdb_host_register(void* ptr, int64_t bytes) {
    if (!ptr || bytes <= 0) return FDB_ERR_BADARG;
    cudaError_t e = cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterPortable);
    if (e != cudaSuccess) return fail(nullptr, FDB_ERR_CUDA, "cudaHostRegister: %s", cudaGetErrorString(e));
    return FDB_OK;
}
extern "C" int fdb_host_unregister(void* ptr) {
    if (!ptr) return FDB_ERR_BADARG;
    cudaError_t e = cudaHostUnregister(ptr);
    if (e != cudaSuccess) return fail(nullptr, FDB_ERR_CUDA, "cudaHostUnregister: %s", cudaGetErrorString(e));
    return FDB_OK;
}

// ===========================================================================================================
// the group
// ===========================================================================================================
extern "C" int fdb_multi_init(int ndev, const int* devices, fdb_multi** out) {
    if (!out) return FDB_ERR_BADARG;
    *out = nullptr;
    int have = 0;
    cudaError_t e = cudaGetDeviceCount(&have);
    if (e != cudaSuccess || have == 0)
        return mfail(nullptr, FDB_ERR_CUDA, "no CUDA device available (%s); fdb200 has no CPU fallback", e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    if (ndev <= 0) { ndev = have < FDB_MAX_RANKS ? have : FDB_MAX_RANKS; devices = nullptr; }
    if (ndev > FDB_MAX_RANKS) return mfail(nullptr, FDB_ERR_BADARG, "fdb_multi_init: at most %d devices (one NVSwitch domain)", FDB_MAX_RANKS);
    fdb_multi* g = new fdb_multi();
    for (int i = 0; i < ndev; i++) {
        const int dev = devices ? devices[i] : i;
        int rc = fdb_init(dev, &g->ctx[i]);
        if (rc) {
            for (int q = 0; q < i; q++) fdb_shutdown(g->ctx[q]);
            delete g;
            return rc;                                           // message in fdb_last_error(NULL)
        }
        g->ctx[i]->multi = g;
        g->n = i + 1;
    }
    *out = g;
    return FDB_OK;
}
extern "C" int fdb_multi_shutdown(fdb_multi* g) {
    if (!g) return FDB_ERR_BADARG;
    for (int i = 0; i < g->n; i++) fdb_shutdown(g->ctx[i]);
    delete g;
    return FDB_OK;
}
extern "C" int fdb_multi_size(const fdb_multi* g) { return g ? g->n : 0; }
extern "C" fdb_ctx* fdb_multi_ctx(fdb_multi* g, int i) { return (g && i >= 0 && i < g->n) ? g->ctx[i] : nullptr; }
extern "C" const char* fdb_multi_last_error(fdb_multi* g) { return g ? g->err.c_str() : g_init_error.c_str(); }
extern "C" int fdb_multi_sync(fdb_multi* g) {
    if (!g) return FDB_ERR_BADARG;
    int first = FDB_OK;
    for (int i = 0; i < g->n; i++) { int rc = fdb_sync(g->ctx[i]); if (rc && !first) first = mpropagate(g, i, rc); }
    return first;
}

// ---- host results: every device sweeps its chunk range and returns its slice ---------------------------
extern "C" int fdb_multi_gradient_host(fdb_multi* g, int fn, const double* x_host, int64_t xlen, int N, double* grad_host, double* value_host) {
    if (!g || !x_host || !grad_host) return mfail(g, FDB_ERR_BADARG, "fdb_multi_gradient_host: null argument");
    if (xlen == 0) { if (value_host) *value_host = 0.0; return FDB_OK; }
    int rc = check_sizes(g->ctx[0], xlen, N); if (rc) return mpropagate(g, 0, rc);
    HostJob jobs[FDB_MAX_RANKS];
    for (int i = 0; i < g->n; i++) { rc = grad_begin(g->ctx[i], fn, x_host, xlen, N, i, g->n, jobs[i]); if (rc) { drain(g); return mpropagate(g, i, rc); } }
    int first = FDB_OK;
    for (int i = 0; i < g->n; i++) { rc = grad_end(jobs[i], grad_host, value_host); if (rc && !first) first = mpropagate(g, i, rc); }
    return first;
}

extern "C" int fdb_multi_jacobian_host(fdb_multi* g, int fn, const double* x_host, int64_t xlen, int64_t m, int N, const double* A_host, int64_t lda,
                                       const double* b_host, double alpha, int reuse_params, double* J_host, int64_t ldJ, double* y_host) {
    if (!g || !x_host || !J_host) return mfail(g, FDB_ERR_BADARG, "fdb_multi_jacobian_host: null argument");
    if (m <= 0 || ldJ < m || (A_host && lda < m)) return mfail(g, FDB_ERR_SHAPE, "DimensionMismatch: result / A leading dimension smaller than m");
    int rc = check_sizes(g->ctx[0], xlen, N); if (rc) return mpropagate(g, 0, rc);
    HostJob jobs[FDB_MAX_RANKS];
    for (int i = 0; i < g->n; i++) {
        int64_t lo, hi, c0, c1;
        rank_range(xlen, N, i, g->n, &lo, &hi, &c0, &c1);
        rc = jac_begin(g->ctx[i], fn, x_host, xlen, m, N, A_host, lda, b_host, alpha, reuse_params, lo, hi, jobs[i]);
        if (rc) { drain(g); return mpropagate(g, i, rc); }
    }
    int first = FDB_OK;
    for (int i = 0; i < g->n; i++) { rc = jac_end(jobs[i], J_host, ldJ, y_host, "fdb_multi_jacobian_host"); if (rc && !first) first = mpropagate(g, i, rc); }
    return first;
}

extern "C" int fdb_multi_hessian_host(fdb_multi* g, int fn, const double* x_host, int64_t xlen, int N, double* H_host, int64_t ldH,
                                      double* grad_host, double* value_host) {
    if (!g || !x_host || !H_host) return mfail(g, FDB_ERR_BADARG, "fdb_multi_hessian_host: null argument");
    if (xlen == 0) { if (value_host) *value_host = 0.0; return FDB_OK; }
    if (ldH < xlen) return mfail(g, FDB_ERR_SHAPE, "DimensionMismatch: ldH < n");
    int rc = check_sizes(g->ctx[0], xlen, N); if (rc) return mpropagate(g, 0, rc);
    HostJob jobs[FDB_MAX_RANKS];
    for (int i = 0; i < g->n; i++) { rc = hess_begin(g->ctx[i], fn, x_host, xlen, N, i, g->n, jobs[i]); if (rc) { drain(g); return mpropagate(g, i, rc); } }
    int first = FDB_OK;
    for (int i = 0; i < g->n; i++) { rc = hess_end(jobs[i], H_host, ldH, grad_host, value_host); if (rc && !first) first = mpropagate(g, i, rc); }
    return first;
}

// ---- device-resident results: complete on every device through the in-process peer exchange -----------------
extern "C" int fdb_multi_exchange_create(fdb_multi* g, int64_t n_doubles, fdb_exchange** xs) {
    if (!g || !xs) return mfail(g, FDB_ERR_BADARG, "fdb_multi_exchange_create: null argument");
    for (int i = 0; i < g->n; i++) xs[i] = nullptr;
    for (int i = 0; i < g->n; i++) {
        int rc = fdb_exchange_create(g->ctx[i], n_doubles, i, g->n, &xs[i]);
        if (rc) { for (int q = 0; q < i; q++) { fdb_exchange_destroy(xs[q]); xs[q] = nullptr; } return mpropagate(g, i, rc); }
    }
    int rc = fdb_exchange_connect_local(xs, g->n);
    if (rc) { for (int q = 0; q < g->n; q++) { fdb_exchange_destroy(xs[q]); xs[q] = nullptr; } return mpropagate(g, 0, rc); }
    return FDB_OK;
}
extern "C" int fdb_multi_exchange_destroy(fdb_multi* g, fdb_exchange** xs) {
    if (!g || !xs) return FDB_ERR_BADARG;
    drain(g);                                     // nobody may still be storing into an arena that is about to be freed
    for (int i = 0; i < g->n; i++) if (xs[i]) { fdb_exchange_destroy(xs[i]); xs[i] = nullptr; }
    return FDB_OK;
}
// barrier kernels on every device first, then the host waits for all of them (a synchronous barrier on one device would
// wait for peers whose barrier has not been launched yet)
extern "C" int fdb_multi_barrier(fdb_multi* g, fdb_exchange* const* xs) {
    if (!g || !xs) return mfail(g, FDB_ERR_BADARG, "fdb_multi_barrier: null argument");
    for (int i = 0; i < g->n; i++) { AsyncScope as(g->ctx[i]); int rc = fdb_exchange_barrier(xs[i]); if (rc) { drain(g); return mpropagate(g, i, rc); } }
    return fdb_multi_sync(g);
}
// copy `count` doubles from the host to offset `offset` of EVERY device's arena (replicated inputs: x, parameters)
extern "C" int fdb_multi_upload(fdb_multi* g, fdb_exchange* const* xs, int64_t offset, const double* host, int64_t count) {
    if (!g || !xs || !host || offset < 0 || count < 0) return mfail(g, FDB_ERR_BADARG, "fdb_multi_upload: bad argument");
    for (int i = 0; i < g->n; i++) {
        fdb_ctx* ctx = g->ctx[i];
        DeviceGuard guard(ctx->device);
        if ((size_t)(offset + count) * sizeof(double) > xs[i]->data_bytes) return mfail(g, FDB_ERR_BADARG, "fdb_multi_upload: outside the arena");
        cudaError_t e = cudaMemcpyAsync(reinterpret_cast<double*>(xs[i]->base) + offset, host, (size_t)count * 8, cudaMemcpyHostToDevice, ctx->stream);
        if (e != cudaSuccess) return mfail(g, FDB_ERR_CUDA, "device %d: cudaMemcpyAsync: %s", ctx->device, cudaGetErrorString(e));
    }
    return FDB_OK;
}
extern "C" int fdb_multi_download(fdb_multi* g, fdb_exchange* const* xs, int device_index, int64_t offset, double* host, int64_t count) {
    if (!g || !xs || !host || device_index < 0 || device_index >= g->n || offset < 0 || count < 0) return mfail(g, FDB_ERR_BADARG, "fdb_multi_download: bad argument");
    fdb_ctx* ctx = g->ctx[device_index];
    DeviceGuard guard(ctx->device);
    if ((size_t)(offset + count) * sizeof(double) > xs[device_index]->data_bytes) return mfail(g, FDB_ERR_BADARG, "fdb_multi_download: outside the arena");
    cudaError_t e = cudaMemcpyAsync(host, reinterpret_cast<double*>(xs[device_index]->base) + offset, (size_t)count * 8, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) return mfail(g, FDB_ERR_CUDA, "device %d: %s", ctx->device, cudaGetErrorString(e));
    return mpropagate(g, device_index, check_poison(ctx, "fdb_multi_download"));
}

namespace fdb {
static fdb_array arena_view(fdb_exchange* x, int64_t off, int64_t n, int64_t ld) {
    return plain(x->ctx, reinterpret_cast<double*>(x->base) + off, n, ld);
}
static bool in_arena(const fdb_exchange* x, int64_t off, int64_t count) {
    return off >= 0 && count >= 0 && (size_t)(off + count) * sizeof(double) <= x->data_bytes;
}
// run `launch(i, lo, hi)` for every device's chunk range with its exchange enabled, then barrier + wait
template <class Launch>
static int resident_run(fdb_multi* g, fdb_exchange* const* xs, int64_t n, int N, Launch launch) {
    int rc = FDB_OK;
    for (int i = 0; i < g->n && !rc; i++) {
        int64_t lo, hi, c0, c1;
        rank_range(n, N, i, g->n, &lo, &hi, &c0, &c1);
        fdb_ctx* ctx = g->ctx[i];
        AsyncScope as(ctx);
        rc = fdb_exchange_enable(ctx, xs[i]);
        if (!rc && lo < hi) rc = launch(i, lo, hi);
        fdb_exchange_enable(ctx, nullptr);
        if (rc) { drain(g); return mpropagate(g, i, rc); }
    }
    return fdb_multi_barrier(g, xs);
}
}  // namespace fdb

extern "C" int fdb_multi_gradient_resident(fdb_multi* g, fdb_exchange* const* xs, int fn, int64_t x_off, int64_t xlen, int N, int64_t grad_off,
                                           int64_t value_off) {
    if (!g || !xs) return mfail(g, FDB_ERR_BADARG, "fdb_multi_gradient_resident: null argument");
    int rc = check_sizes(g->ctx[0], xlen, N); if (rc) return mpropagate(g, 0, rc);
    for (int i = 0; i < g->n; i++)
        if (!xs[i] || !in_arena(xs[i], x_off, xlen) || !in_arena(xs[i], grad_off, xlen) || (value_off >= 0 && !in_arena(xs[i], value_off, 1)))
            return mfail(g, FDB_ERR_BADARG, "fdb_multi_gradient_resident: x / gradient / value outside the arena");
    return resident_run(g, xs, xlen, N, [&](int i, int64_t lo, int64_t hi) {
        fdb_array x = arena_view(xs[i], x_off, xlen, xlen), gr = arena_view(xs[i], grad_off, xlen, xlen), v = arena_view(xs[i], value_off < 0 ? 0 : value_off, 1, 1);
        return fdb_gradient_chunked(g->ctx[i], fn, &x, N, lo, hi, &gr, value_off >= 0 ? &v : nullptr);
    });
}

extern "C" int fdb_multi_jacobian_resident(fdb_multi* g, fdb_exchange* const* xs, int fn, int64_t x_off, int64_t xlen, int64_t m, int N,
                                           const fdb_array* const* A, int64_t lda, const fdb_array* const* b, double alpha, int64_t J_off, int64_t ldJ,
                                           int64_t y_off) {
    if (!g || !xs) return mfail(g, FDB_ERR_BADARG, "fdb_multi_jacobian_resident: null argument");
    int rc = check_sizes(g->ctx[0], xlen, N); if (rc) return mpropagate(g, 0, rc);
    if (m <= 0 || ldJ < m) return mfail(g, FDB_ERR_SHAPE, "DimensionMismatch: ldJ < m");
    const int64_t Jn = ldJ * (xlen - 1) + m;
    for (int i = 0; i < g->n; i++)
        if (!xs[i] || !in_arena(xs[i], x_off, xlen) || !in_arena(xs[i], J_off, Jn) || (y_off >= 0 && !in_arena(xs[i], y_off, m)))
            return mfail(g, FDB_ERR_BADARG, "fdb_multi_jacobian_resident: x / J / y outside the arena");
    return resident_run(g, xs, xlen, N, [&](int i, int64_t lo, int64_t hi) {
        fdb_array x = arena_view(xs[i], x_off, xlen, xlen), J = arena_view(xs[i], J_off, Jn, Jn), y = arena_view(xs[i], y_off < 0 ? 0 : y_off, m, m);
        return fdb_jacobian_chunked(g->ctx[i], fn, &x, m, N, A ? A[i] : nullptr, lda, b ? b[i] : nullptr, alpha, lo, hi, &J, ldJ, y_off >= 0 ? &y : nullptr);
    });
}

extern "C" int fdb_multi_hessian_resident(fdb_multi* g, fdb_exchange* const* xs, int fn, int64_t x_off, int64_t xlen, int N, int64_t H_off, int64_t ldH,
                                          int64_t grad_off, int64_t value_off) {
    if (!g || !xs) return mfail(g, FDB_ERR_BADARG, "fdb_multi_hessian_resident: null argument");
    int rc = check_sizes(g->ctx[0], xlen, N); if (rc) return mpropagate(g, 0, rc);
    if (ldH < xlen) return mfail(g, FDB_ERR_SHAPE, "DimensionMismatch: ldH < n");
    const int64_t Hn = ldH * (xlen - 1) + xlen;
    for (int i = 0; i < g->n; i++)
        if (!xs[i] || !in_arena(xs[i], x_off, xlen) || !in_arena(xs[i], H_off, Hn) || (grad_off >= 0 && !in_arena(xs[i], grad_off, xlen)) ||
            (value_off >= 0 && !in_arena(xs[i], value_off, 1)))
            return mfail(g, FDB_ERR_BADARG, "fdb_multi_hessian_resident: x / H / gradient / value outside the arena");
    return resident_run(g, xs, xlen, N, [&](int i, int64_t lo, int64_t hi) {
        fdb_array x = arena_view(xs[i], x_off, xlen, xlen), H = arena_view(xs[i], H_off, Hn, Hn), gr = arena_view(xs[i], grad_off < 0 ? 0 : grad_off, xlen, xlen),
                  v = arena_view(xs[i], value_off < 0 ? 0 : value_off, 1, 1);
        return fdb_hessian_chunked(g->ctx[i], fn, &x, N, lo, hi, &H, ldH, grad_off >= 0 ? &gr : nullptr, value_off >= 0 ? &v : nullptr);
    });
}
