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
#include "fdb_plane_io.cuh"

namespace fdb {

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
        case FDB_NANMAX: return nanmax_v(x, y);
        case FDB_NANMIN: return nanmin_v(x, y);
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
    FDB_ENTER(ctx);
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
    FDB_ENTER(ctx);
    if (!ctx || !out || !a || !b) return fail(ctx, FDB_ERR_BADARG, "fdb_map2: null argument");
    if (op < FDB_ADD || op > FDB_NANMIN) return fail(ctx, FDB_ERR_UNSUPPORTED, "fdb_map2: unknown op %d", op);
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
    FDB_ENTER(ctx);
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
    FDB_ENTER(ctx);
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
