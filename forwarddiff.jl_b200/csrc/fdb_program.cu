// fdb_program.cu -- fused broadcast trees without a JIT: a per-thread interpreter for small
// "op-programs", so that an arbitrary user `f` made of elementwise Dual arithmetic on shifted views of
// x plus sum reductions (the shape Julia's broadcast + sum gives: SURVEY.md section 8f-1) runs through the
// SAME batched chunk sweep as the catalogue targets -- one launch for all chunks, implicit seeds, partials
// in registers/local memory, fused extraction -- instead of one kernel per array operation per chunk.
//
// A program has two parts.  The term program runs for every term i in [0, n - halo): it loads seeded
// Duals x[i + off], applies src/dual.jl ops (same device functions as every other kernel, so the
// arithmetic per op is bit-identical to fdb_map1/fdb_map2), and adds results into K sum accumulators
// (`sum(...)`).  The epilogue program runs once per chunk on the K reduced Duals (scalar Dual math after the
// reductions, e.g. Ackley's exp/sqrt) and names the result register; the chunk's partials go straight into
// the gradient slice.  Programs travel as kernel parameters (constant bank: every thread reads the same
// instruction, a broadcast).  Lane sharing as in fdb_sweep.cu: terms outside the seeded window are
// interpreted on Dual{1}.
#include <string.h>

#include "fdb_internal.h"
#include "fdb_targets.cuh"

namespace fdb {

constexpr int PG_MAXI = 48;      // instructions per program part
constexpr int PG_MAXC = 24;      // Float64 constants
constexpr int PG_MAXR = 12;      // virtual Dual registers
constexpr int PG_MAXK = 4;       // sum accumulators

struct Program {
    int n_term, n_epi, n_acc, halo, result_reg, pad;
    short term[PG_MAXI][4];      // {op, dst, a, b}
    short epi[PG_MAXI][4];
    double consts[PG_MAXC];
};

enum { PK_LOADX = 0, PK_UNARY = 1, PK_DD = 2, PK_DC = 3, PK_CD = 4, PK_ACC = 5, PK_MOV = 6 };

template <int N, bool NS> FDB_DEV Dual<double, N, NS> pg_unary(int code, const Dual<double, N, NS>& x) {
    switch (code) {
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
template <int N, bool NS> FDB_DEV Dual<double, N, NS> pg_dd(int code, const Dual<double, N, NS>& x, const Dual<double, N, NS>& y) {
    switch (code) {
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
template <int N, bool NS> FDB_DEV Dual<double, N, NS> pg_dc(int code, const Dual<double, N, NS>& x, double s) {
    switch (code) {
        case FDB_ADD: return x + s;
        case FDB_SUB: return x - s;
        case FDB_MUL: return x * s;
        case FDB_DIV: return x / s;
        default: return pow_v(x, s);
    }
}
template <int N, bool NS> FDB_DEV Dual<double, N, NS> pg_cd(int code, double s, const Dual<double, N, NS>& y) {
    switch (code) {
        case FDB_ADD: return s + y;
        case FDB_SUB: return s - y;
        case FDB_MUL: return s * y;
        case FDB_DIV: return s / y;
        default: return pow_v(s, y);
    }
}

// one pass over a program part; L yields the Dual at a position (term part only)
template <class D, int NL, bool NS, class L>
FDB_DEV void pg_run(const short (*code)[4], int ninstr, const double* consts, const L& x, int64_t i, D* R, D* acc) {
    for (int pc = 0; pc < ninstr; pc++) {
        const int op = code[pc][0], dst = code[pc][1], a = code[pc][2], b = code[pc][3];
        const int kind = op >> 6, c = op & 63;
        switch (kind) {
            case PK_LOADX: R[dst] = x(i + b); break;
            case PK_UNARY: R[dst] = pg_unary<NL, NS>(c, R[a]); break;
            case PK_DD: R[dst] = pg_dd<NL, NS>(c, R[a], R[b]); break;
            case PK_DC: R[dst] = pg_dc<NL, NS>(c, R[a], consts[b]); break;
            case PK_CD: R[dst] = pg_cd<NL, NS>(c, consts[a], R[b]); break;
            case PK_ACC: acc[dst] = acc[dst] + R[a]; break;
            default: R[dst] = R[a]; break;
        }
    }
}
struct NoLoader {
    template <class T> FDB_DEV T operator()(int64_t) const { return T(); }
};

struct SumK {   // block_combine policy: K sum accumulators
    static constexpr int K = PG_MAXK;
    template <class D> static FDB_DEV D init(int, const D& like) { return zero_of(like); }
    template <class D> static FDB_DEV D combine(int, const D& a, const D& b) { return a + b; }
};

template <int N, bool NS, int THREADS, bool SHARE>
__global__ void __launch_bounds__(THREADS) program_gradient_kernel(const __grid_constant__ Program P, const double* __restrict__ x, int64_t n,
                                                                   int64_t chunk_lo, int64_t nchunks_total, int lastchunksize,
                                                                   double* __restrict__ grad, double* __restrict__ value_out) {
    typedef Dual<double, N, NS> D;
    typedef Dual<double, 1, NS> D1;
    const int64_t c = chunk_lo + blockIdx.x;
    const bool last = (c == nchunks_total - 1);
    const int chunksize = last ? lastchunksize : N;
    const int64_t start = last ? (n - lastchunksize) : c * N;
    SeedLoader<N, NS> xs{x, start, chunksize};
    ZeroLoader<1, NS> xz{x};
    D acc[PG_MAXK]; D1 acc1[PG_MAXK];
#pragma unroll
    for (int k = 0; k < PG_MAXK; k++) {
        acc[k].v = 0.0; acc1[k].v = 0.0; acc1[k].p[0] = 0.0;
#pragma unroll
        for (int j = 0; j < N; j++) acc[k].p[j] = 0.0;
    }
    D R[PG_MAXR]; D1 R1[PG_MAXR];
    const int64_t nt = n - P.halo;
    for (int64_t i = threadIdx.x; i < nt; i += THREADS) {
        const int64_t w0 = i - (threadIdx.x & 31);
        const bool far = SHARE && ((w0 + 31 + P.halo < start) || (w0 >= start + chunksize));   // warp-uniform
        if (far) pg_run<D1, 1, NS>(P.term, P.n_term, P.consts, xz, i, R1, acc1);
        else pg_run<D, N, NS>(P.term, P.n_term, P.consts, xs, i, R, acc);
    }
    if (SHARE) {
#pragma unroll
        for (int k = 0; k < PG_MAXK; k++) {
            D e; e.v = acc1[k].v;
#pragma unroll
            for (int j = 0; j < N; j++) e.p[j] = acc1[k].p[0];
            acc[k] = acc[k] + e;
        }
    }
    block_combine<SumK, D, THREADS>(acc);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < PG_MAXK; k++) R[k] = acc[k];
        pg_run<D, N, NS>(P.epi, P.n_epi, P.consts, xs, 0, R, acc);
        const D y = R[P.result_reg];
#pragma unroll
        for (int k = 0; k < N; k++)
            if (k < chunksize) grad[start + k] = y.p[k];
        if (last && value_out) value_out[0] = y.v;
    }
}

// elementwise / stencil array targets: y[i] = program(x[i .. i+halo]) for i in [0, n - halo);
// J[i, start+k] = partials(y[i], k) -- the f(x)-form Jacobian chunk sweep (src/jacobian.jl:169-216)
template <int N, bool NS, int THREADS, bool SHARE>
__global__ void __launch_bounds__(THREADS) program_jacobian_kernel(const __grid_constant__ Program P, const double* __restrict__ x, int64_t n,
                                                                   int64_t chunk_lo, int64_t nchunks_total, int lastchunksize,
                                                                   double* __restrict__ J, int64_t ldJ, double* __restrict__ yout) {
    typedef Dual<double, N, NS> D;
    typedef Dual<double, 1, NS> D1;
    const int64_t c = chunk_lo + blockIdx.x;
    const bool last = (c == nchunks_total - 1);
    const int chunksize = last ? lastchunksize : N;
    const int64_t start = last ? (n - lastchunksize) : c * N;
    const int64_t i = (int64_t)blockIdx.y * THREADS + threadIdx.x;
    const int64_t m = n - P.halo;
    const int64_t w0 = i - (threadIdx.x & 31);
    const bool far = SHARE && ((w0 + 31 + P.halo < start) || (w0 >= start + chunksize));
    if (i >= m) return;
    if (far) {
        D1 R1[PG_MAXR]; D1 acc1[1];
        pg_run<D1, 1, NS>(P.term, P.n_term, P.consts, ZeroLoader<1, NS>{x}, i, R1, acc1);
        const D1 y = R1[P.result_reg];
        for (int k = 0; k < chunksize; k++) __stcs(J + i + (start + k) * ldJ, y.p[0]);
        if (yout && last) yout[i] = y.v;
    } else {
        D R[PG_MAXR]; D acc[1];
        pg_run<D, N, NS>(P.term, P.n_term, P.consts, SeedLoader<N, NS>{x, start, chunksize}, i, R, acc);
        const D y = R[P.result_reg];
#pragma unroll
        for (int k = 0; k < N; k++)
            if (k < chunksize) __stcs(J + i + (start + k) * ldJ, y.p[k]);
        if (yout && last) yout[i] = y.v;
    }
}

static int build_program(fdb_ctx* ctx, Program& P, const int32_t* term, int n_term, const int32_t* epi, int n_epi, const double* consts,
                         int n_consts, int n_acc, int halo, int result_reg, bool array_form) {
    if (n_term < 0 || n_term > PG_MAXI || n_epi < 0 || n_epi > PG_MAXI || n_consts < 0 || n_consts > PG_MAXC || n_acc < 0 || n_acc > PG_MAXK ||
        halo < 0 || result_reg < 0 || result_reg >= PG_MAXR || (n_term && !term) || (n_epi && !epi) || (n_consts && !consts))
        return fail(ctx, FDB_ERR_UNSUPPORTED, "op-program exceeds interpreter limits (%d instr, %d consts, %d registers, %d accumulators)",
                    PG_MAXI, PG_MAXC, PG_MAXR, PG_MAXK);
    memset(&P, 0, sizeof P);
    P.n_term = n_term; P.n_epi = n_epi; P.n_acc = n_acc; P.halo = halo; P.result_reg = result_reg;
    for (int part = 0; part < 2; part++) {
        const int32_t* src = part ? epi : term; const int cnt = part ? n_epi : n_term;
        for (int i = 0; i < cnt; i++) {
            const int op = src[4 * i], dst = src[4 * i + 1], a = src[4 * i + 2], b = src[4 * i + 3];
            const int kind = op >> 6, code = op & 63;
            bool ok = dst >= 0 && dst < PG_MAXR && kind >= PK_LOADX && kind <= PK_MOV;
            if (kind == PK_LOADX) ok = ok && !part && b >= 0 && b <= halo;
            if (kind == PK_UNARY) ok = ok && code >= FDB_NEG && code <= FDB_CBRT && a >= 0 && a < PG_MAXR;
            if (kind == PK_DD) ok = ok && code >= FDB_ADD && code <= FDB_HYPOT && a >= 0 && a < PG_MAXR && b >= 0 && b < PG_MAXR;
            if (kind == PK_DC) ok = ok && code >= FDB_ADD && code <= FDB_POW && a >= 0 && a < PG_MAXR && b >= 0 && b < n_consts;
            if (kind == PK_CD) ok = ok && code >= FDB_ADD && code <= FDB_POW && a >= 0 && a < n_consts && b >= 0 && b < PG_MAXR;
            if (kind == PK_ACC) ok = ok && !part && !array_form && dst < n_acc && a >= 0 && a < PG_MAXR;
            if (kind == PK_MOV) ok = ok && a >= 0 && a < PG_MAXR;
            if (!ok) return fail(ctx, FDB_ERR_BADARG, "op-program: malformed instruction %d of the %s part", i, part ? "epilogue" : "term");
            short* d = part ? P.epi[i] : P.term[i];
            d[0] = (short)op; d[1] = (short)dst; d[2] = (short)a; d[3] = (short)b;
        }
    }
    for (int i = 0; i < n_consts; i++) P.consts[i] = consts[i];
    return FDB_OK;
}

}  // namespace fdb

using namespace fdb;

extern "C" int fdb_program_gradient_chunked(fdb_ctx* ctx, const int32_t* term_prog, int n_term, const int32_t* epi_prog, int n_epi,
                                            const double* consts, int n_consts, int n_acc, int halo, int result_reg, const fdb_array* x, int N,
                                            int64_t chunk_lo, int64_t chunk_hi, fdb_array* grad, fdb_array* value_dev) {
    if (!ctx || !x || !grad) return fail(ctx, FDB_ERR_BADARG, "fdb_program_gradient_chunked: null argument");
    if (x->level != 0 || grad->level != 0 || (value_dev && value_dev->level != 0)) return fail(ctx, FDB_ERR_BADARG, "x, grad, value must be plain arrays");
    if (grad->n != x->n) return fail(ctx, FDB_ERR_SHAPE, "DimensionMismatch: gradient result has %lld entries, x has %lld", (long long)grad->n, (long long)x->n);
    const int64_t n = x->n;
    if (N < 1 || N > FDB_MAX_CHUNK) return fail(ctx, FDB_ERR_BADARG, "chunk size N=%d outside 1..12", N);
    if (n < N) return fail(ctx, FDB_ERR_CHUNK, "chunk size cannot be greater than ForwardDiff.structural_length(x) (%d > %lld)", N, (long long)n);
    Program P;
    int rc = build_program(ctx, P, term_prog, n_term, epi_prog, n_epi, consts, n_consts, n_acc, halo, result_reg, false); if (rc) return rc;
    Plan p = make_plan(n, N);
    const int64_t nchunks = (n == N) ? 1 : p.nchunks;
    if (chunk_hi < 0) chunk_hi = nchunks;
    if (chunk_lo < 0 || chunk_lo > chunk_hi || chunk_hi > nchunks) return fail(ctx, FDB_ERR_BADARG, "bad chunk range");
    if (chunk_lo == chunk_hi) return FDB_OK;
    constexpr int T = 128;
    const unsigned grid = (unsigned)(chunk_hi - chunk_lo);
    double* vout = value_dev ? value_dev->data : nullptr;
    if (ctx->lane_sharing) {
        FDB_DISPATCH_N(N, FDB_DISPATCH_NS(ctx->nansafe,
            (program_gradient_kernel<NN, NS, T, true><<<grid, T, 0, ctx->stream>>>(P, x->data, n, chunk_lo, nchunks, (int)p.lastchunksize, grad->data, vout))));
    } else {
        FDB_DISPATCH_N(N, FDB_DISPATCH_NS(ctx->nansafe,
            (program_gradient_kernel<NN, NS, T, false><<<grid, T, 0, ctx->stream>>>(P, x->data, n, chunk_lo, nchunks, (int)p.lastchunksize, grad->data, vout))));
    }
    return finish(ctx, 1, "fdb_program_gradient_chunked");
}

extern "C" int fdb_program_jacobian_chunked(fdb_ctx* ctx, const int32_t* prog, int n_instr, const double* consts, int n_consts, int halo,
                                            int result_reg, const fdb_array* x, int N, int64_t chunk_lo, int64_t chunk_hi, fdb_array* J,
                                            int64_t ldJ, fdb_array* y) {
    if (!ctx || !x || !J) return fail(ctx, FDB_ERR_BADARG, "fdb_program_jacobian_chunked: null argument");
    if (x->level != 0 || J->level != 0 || (y && y->level != 0)) return fail(ctx, FDB_ERR_BADARG, "x, J, y must be plain arrays");
    const int64_t n = x->n, m = n - halo;
    if (N < 1 || N > FDB_MAX_CHUNK) return fail(ctx, FDB_ERR_BADARG, "chunk size N=%d outside 1..12", N);
    if (n < N) return fail(ctx, FDB_ERR_CHUNK, "chunk size cannot be greater than ForwardDiff.structural_length(x) (%d > %lld)", N, (long long)n);
    if (m <= 0) return fail(ctx, FDB_ERR_SHAPE, "op-program Jacobian: empty output");
    if (ldJ < m || J->n < ldJ * (n - 1) + m) return fail(ctx, FDB_ERR_SHAPE, "DimensionMismatch: Jacobian result must hold %lld x %lld", (long long)m, (long long)n);
    if (y && y->n != m) return fail(ctx, FDB_ERR_SHAPE, "DimensionMismatch: y has %lld entries, f(x) has %lld", (long long)y->n, (long long)m);
    Program P;
    int rc = build_program(ctx, P, prog, n_instr, nullptr, 0, consts, n_consts, 0, halo, result_reg, true); if (rc) return rc;
    Plan p = make_plan(n, N);
    const int64_t nchunks = (n == N) ? 1 : p.nchunks;
    if (chunk_hi < 0) chunk_hi = nchunks;
    if (chunk_lo < 0 || chunk_lo > chunk_hi || chunk_hi > nchunks) return fail(ctx, FDB_ERR_BADARG, "bad chunk range");
    if (chunk_lo == chunk_hi) return FDB_OK;
    constexpr int T = 128;
    dim3 grid((unsigned)(chunk_hi - chunk_lo), (unsigned)cdiv(m, T));
    double* yp = y ? y->data : nullptr;
    if (ctx->lane_sharing) {
        FDB_DISPATCH_N(N, FDB_DISPATCH_NS(ctx->nansafe,
            (program_jacobian_kernel<NN, NS, T, true><<<grid, T, 0, ctx->stream>>>(P, x->data, n, chunk_lo, nchunks, (int)p.lastchunksize, J->data, ldJ, yp))));
    } else {
        FDB_DISPATCH_N(N, FDB_DISPATCH_NS(ctx->nansafe,
            (program_jacobian_kernel<NN, NS, T, false><<<grid, T, 0, ctx->stream>>>(P, x->data, n, chunk_lo, nchunks, (int)p.lastchunksize, J->data, ldJ, yp))));
    }
    return finish(ctx, 1, "fdb_program_jacobian_chunked");
}
