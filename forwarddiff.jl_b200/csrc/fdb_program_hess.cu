// fdb_program_hess.cu -- Hessians of flattened user functions: hessian(f, x, HessianConfig(f, x, Chunk{N}()))
// = jacobian(y -> gradient(f, y), x) (src/hessian.jl:14-19) for any f of the form sum_i term(x[i .. i+halo])
// expressed as an op-program (fdb_program.cuh).  Same structure as fdb_hessian.cu -- block (ci, co) evaluates f
// once on the nested duals Dual{T,Dual{T,Float64,N},N} of inner chunk ci and outer chunk co; terms far from both
// windows are interpreted on the four shared lanes Dual{Dual{1},1}, the few near terms per inner lane on
// Dual{Dual{N},1} -- with the interpreter in place of the compiled term.
// Restrictions (checked on the host): one sum accumulator, no epilogue, and the op subset whose rules exist one
// dual level up (+ - * / neg, literal powers, exp log sin cos tanh sqrt abs abs2 inv).
#include "fdb_program.cuh"
#include "fdb_hessian_kernel.cuh"   // HESS_GROUP: grid of the specialised plain-sum kernel

namespace fdb {

template <class D> FDB_DEV D pgn_unary(int code, const D& x) {
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
        default: return pow3(x);
    }
}
template <class D> FDB_DEV D pgn_dd(int code, const D& x, const D& y) {
    switch (code) {
        case FDB_ADD: return x + y;
        case FDB_SUB: return x - y;
        case FDB_MUL: return x * y;
        default: return x / y;
    }
}
template <class D> FDB_DEV D pgn_dc(int code, const D& x, double s) {
    switch (code) {
        case FDB_ADD: return x + s;
        case FDB_SUB: return x - s;
        case FDB_MUL: return x * s;
        default: return x / s;
    }
}
template <class D> FDB_DEV D pgn_cd(int code, double s, const D& y) {
    switch (code) {
        case FDB_ADD: return s + y;
        case FDB_SUB: return s - y;
        case FDB_MUL: return s * y;
        default: return s / y;
    }
}
template <class D, class L>
FDB_DEV void pgn_run(const short (*code)[4], int ninstr, const double* consts, const L& x, int64_t i, D* R, D& acc) {
    for (int pc = 0; pc < ninstr; pc++) {
        const int op = code[pc][0], dst = code[pc][1], a = code[pc][2], b = code[pc][3];
        const int kind = op >> 6, c = op & 63;
        switch (kind) {
            case PK_LOADX: R[dst] = x(i + b); break;
            case PK_UNARY: R[dst] = pgn_unary<D>(c, R[a]); break;
            case PK_DD: R[dst] = pgn_dd<D>(c, R[a], R[b]); break;
            case PK_DC: R[dst] = pgn_dc<D>(c, R[a], consts[b]); break;
            case PK_CD: R[dst] = pgn_cd<D>(c, consts[a], R[b]); break;
            case PK_ACC: acc = acc + R[a]; break;
            default: R[dst] = R[a]; break;
        }
    }
}

// loaders of the nested duals (same as fdb_hessian.cu)
template <bool NS> struct PFFLoader {
    typedef Dual<Dual<double, 1, NS>, 1, NS> DO;
    const double* x;
    FDB_DEV DO operator()(int64_t i) const { DO d; d.v.v = x[i]; d.v.p[0] = 0.0; d.p[0].v = 0.0; d.p[0].p[0] = 0.0; return d; }
};
template <int N, bool NS> struct PNearLoader {
    typedef Dual<double, N, NS> DI;
    typedef Dual<DI, 1, NS> DO;
    const double* x; int64_t ostart; int ocs; int64_t istart; int ics; int k;
    FDB_DEV DO operator()(int64_t pos) const {
        DO d; d.v = seeded<N, NS>(x[pos], pos, ostart, ocs);
        d.p[0].v = ((pos - istart == k) && (k < ics)) ? 1.0 : 0.0;
#pragma unroll
        for (int j = 0; j < N; j++) d.p[0].p[j] = 0.0;
        return d;
    }
};

constexpr int PH_MAXHALO = 4;

template <int N, bool NS, int THREADS>
__global__ void __launch_bounds__(THREADS) program_hessian_kernel(const __grid_constant__ Program P, const double* __restrict__ x, int64_t n,
                                                                  int64_t outer_lo, int64_t nchunks_total, int lastchunksize,
                                                                  double* __restrict__ H, int64_t ldH, double* __restrict__ grad,
                                                                  double* __restrict__ value_out, const __grid_constant__ fdb_mirrors M) {
    typedef Dual<double, N, NS> DI;
    typedef Dual<DI, 1, NS> DG;
    typedef Dual<Dual<double, 1, NS>, 1, NS> DF;
    constexpr int MAXNEAR = 2 * (N + PH_MAXHALO);
    __shared__ double s_p[MAXNEAR][N][N + 1];
    __shared__ double s_v[MAXNEAR][N + 1];
    __shared__ double s_ff[4];
    __shared__ double s_red[THREADS / 32][4];
    const int halo = P.halo;
    const int64_t ci = blockIdx.x, co = outer_lo + blockIdx.y;
    const bool ilast = (ci == nchunks_total - 1), olast = (co == nchunks_total - 1);
    const int ics = ilast ? lastchunksize : N, ocs = olast ? lastchunksize : N;
    const int64_t istart = ilast ? n - lastchunksize : ci * N;
    const int64_t ostart = olast ? n - lastchunksize : co * N;
    const int64_t nt = n - halo;
    int64_t a0 = istart - halo, a1 = istart + ics, b0 = ostart - halo, b1 = ostart + ocs;
    if (a0 < 0) a0 = 0; if (b0 < 0) b0 = 0; if (a1 > nt) a1 = nt; if (b1 > nt) b1 = nt;
    if (a1 < a0) a1 = a0; if (b1 < b0) b1 = b0;
    if (b0 <= a1 && a0 <= b1) { a0 = a0 < b0 ? a0 : b0; a1 = a1 > b1 ? a1 : b1; b0 = b1 = a1; }
    const int na = (int)(a1 - a0), nb = (int)(b1 - b0), nnear = na + nb;

    // phase 1: far terms on the four shared lanes
    DF accff; accff.v.v = 0.0; accff.v.p[0] = 0.0; accff.p[0].v = 0.0; accff.p[0].p[0] = 0.0;
    {
        DF R[PG_MAXR];
        PFFLoader<NS> xf{x};
        for (int64_t i = threadIdx.x; i < nt; i += THREADS) {
            const bool near = (i >= a0 && i < a1) || (i >= b0 && i < b1);
            if (!near) pgn_run<DF>(P.term, P.n_term, P.consts, xf, i, R, accff);
        }
    }
    {
        double r[4] = {accff.v.v, accff.v.p[0], accff.p[0].v, accff.p[0].p[0]};
#pragma unroll
        for (int q = 0; q < 4; q++) {
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) r[q] = r[q] + shfl_down_d(r[q], off);
            if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5][q] = r[q];
        }
    }
    // phase 2: near terms, one work item per (term, inner lane k)
    for (int w = threadIdx.x; w < nnear * N; w += THREADS) {
        const int t = w / N, k = w - t * N;
        const int64_t i = t < na ? a0 + t : b0 + (t - na);
        PNearLoader<N, NS> xn{x, ostart, ocs, istart, ics, k};
        DG R[PG_MAXR];
        DG acc; acc.v.v = 0.0; acc.p[0].v = 0.0;
#pragma unroll
        for (int j = 0; j < N; j++) { acc.v.p[j] = 0.0; acc.p[0].p[j] = 0.0; }
        pgn_run<DG>(P.term, P.n_term, P.consts, xn, i, R, acc);
        s_p[t][k][0] = acc.p[0].v;
#pragma unroll
        for (int j = 0; j < N; j++) s_p[t][k][j + 1] = acc.p[0].p[j];
        if (k == 0) {
            s_v[t][0] = acc.v.v;
#pragma unroll
            for (int j = 0; j < N; j++) s_v[t][j + 1] = acc.v.p[j];
        }
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        double s = 0.0;
        for (int wq = 0; wq < THREADS / 32; wq++) s = s + s_red[wq][threadIdx.x];
        s_ff[threadIdx.x] = s;
    }
    __syncthreads();
    // phase 3: fold and store
    for (int e = threadIdx.x; e < N * (N + 1); e += THREADS) {
        const int k = e / (N + 1), j = e - k * (N + 1);
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

}  // namespace fdb

using namespace fdb;

extern "C" int fdb_program_hessian_chunked(fdb_ctx* ctx, const int32_t* term_prog, int n_term, const int32_t* epi_prog, int n_epi,
                                           const double* consts, int n_consts, int n_acc, int halo, int result_reg, const fdb_array* x, int N,
                                           int64_t chunk_lo, int64_t chunk_hi, fdb_array* H, int64_t ldH, fdb_array* grad,
                                           fdb_array* value_dev) {
    FDB_ENTER(ctx);
    if (!ctx || !x) return fail(ctx, FDB_ERR_BADARG, "fdb_program_hessian_chunked: null argument");
    if (x->level != 0 || (H && H->level != 0) || (grad && grad->level != 0) || (value_dev && value_dev->level != 0))
        return fail(ctx, FDB_ERR_BADARG, "x, H, grad, value must be plain arrays");
    const int64_t n = x->n;
    if (N < 1 || N > FDB_MAX_CHUNK) return fail(ctx, FDB_ERR_BADARG, "chunk size N=%d outside 1..12", N);
    if (n < N) return fail(ctx, FDB_ERR_CHUNK, "chunk size cannot be greater than ForwardDiff.structural_length(x) (%d > %lld)", N, (long long)n);
    if (H && (ldH < n || H->n < ldH * (n - 1) + n)) return fail(ctx, FDB_ERR_SHAPE, "DimensionMismatch: Hessian result must hold %lld x %lld", (long long)n, (long long)n);
    if (grad && grad->n != n) return fail(ctx, FDB_ERR_SHAPE, "DimensionMismatch: gradient result has %lld entries, x has %lld", (long long)grad->n, (long long)n);
    if (halo > PH_MAXHALO) return fail(ctx, FDB_ERR_UNSUPPORTED, "op-program Hessian: stencil reach %d > %d", halo, PH_MAXHALO);
    Program P;
    int rc = build_program(ctx, P, term_prog, n_term, epi_prog, n_epi, consts, n_consts, n_acc, halo, result_reg, false, ctx->n_prm); if (rc) return rc;
    rc = check_params(ctx, P, n, "fdb_program_hessian_chunked"); if (rc) return rc;
    for (int part = 0; part < 2; part++)
        for (int i = 0; i < (part ? n_epi : n_term); i++) {                 // rules that exist one dual level up
            const short* ins = part ? P.epi[i] : P.term[i];
            const int kind = ins[0] >> 6, code = ins[0] & 63;
            const bool ok = (kind == PK_LOADX) || (kind == PK_ACC) || (kind == PK_MOV) || (kind == PK_GUARD) || (kind == PK_UNARY && code <= FDB_POW3) ||
                            ((kind == PK_DD || kind == PK_DC || kind == PK_CD || kind == PK_DP || kind == PK_PD) && code <= FDB_DIV);
            if (!ok) return fail(ctx, FDB_ERR_UNSUPPORTED, "op-program Hessian: instruction %d uses an op without a nested-dual rule", i);
        }
    // f = one plain sum of terms: both executors.  K sums + arithmetic after the reductions: the specialised kernel only
    if (P.two_pass) return fail(ctx, FDB_ERR_UNSUPPORTED, "op-program Hessian: two-pass programs are not supported");
    if (P.prod_mask) return fail(ctx, FDB_ERR_UNSUPPORTED, "op-program Hessian: product accumulators have no nested-dual batched kernel");
    const bool plain = (n_epi == 0 && n_acc == 1 && result_reg == 0 && P.n_params == 0 && !P.has_guard);
    Plan p = make_plan(n, N);
    const int64_t nchunks = (n == N) ? 1 : p.nchunks;
    if (chunk_hi < 0) chunk_hi = nchunks;
    if (chunk_lo < 0 || chunk_lo > chunk_hi || chunk_hi > nchunks) return fail(ctx, FDB_ERR_BADARG, "bad chunk range");
    if (chunk_lo == chunk_hi) return FDB_OK;
    if (chunk_hi - chunk_lo > 65535) return fail(ctx, FDB_ERR_UNSUPPORTED, "more than 65535 outer chunks per call");
    constexpr int T = 128;
    dim3 grid((unsigned)nchunks, (unsigned)(chunk_hi - chunk_lo));
    double* Hp = H ? H->data : nullptr; double* gp = grad ? grad->data : nullptr; double* vp = value_dev ? value_dev->data : nullptr;
    fdb_mirrors M;
    const fdb_array* outs[3] = {H, grad, value_dev};
    { int mrc = mirrors_for(ctx, &M, outs, 3, "fdb_program_hessian_chunked"); if (mrc) return mrc; }
    {   // specialised kernel for this program (fdb_jit.cu): hessian_sweep_body instantiated for the user's functor
        void* fn = nullptr;
        if (jit_get(ctx, 2, P, N, ctx->nansafe, false, false, &fn) == FDB_OK) {
            const double* xp = x->data; int64_t nn = n, lo = chunk_lo, nc = nchunks, ld = ldH; int lcs = (int)p.lastchunksize;
            void* args[] = {&xp, &nn, &lo, &nc, &lcs, &Hp, &ld, &gp, &vp, &M};
            // UserFn::PLAIN_SUM (fdb_jit.cu) selects hessian_sweep_body, whose blocks take HESS_GROUP inner chunks each
            const bool jit_plain = (n_epi == 0 && n_acc == 1 && result_reg == 0 && !P.prod_mask);
            const unsigned gx = jit_plain ? (unsigned)cdiv(nchunks, (int64_t)HESS_GROUP) : grid.x;
            rc = jit_launch(ctx, fn, gx, grid.y, T, args, "fdb_program_hessian_chunked(specialised)"); if (rc) return rc;
            return finish(ctx, 1, "fdb_program_hessian_chunked(specialised)");
        }
    }
    if (!plain || !P.fits_interpreter)
        return fail(ctx, FDB_ERR_UNSUPPORTED, "op-program Hessian: arithmetic after the reduction, array parameters or programs beyond the interpreter's limits need the run-time specialised kernel (%s)",
                    ctx->use_jit ? fdb_jit_last_log(ctx) : "disabled by fdb_set_jit");
    FDB_DISPATCH_N(N, FDB_DISPATCH_NS(ctx->nansafe,
        (program_hessian_kernel<NN, NS, T><<<grid, T, 0, ctx->stream>>>(P, x->data, n, chunk_lo, nchunks, (int)p.lastchunksize, Hp, ldH, gp, vp, M))));
    return finish(ctx, 1, "fdb_program_hessian_chunked");
}
