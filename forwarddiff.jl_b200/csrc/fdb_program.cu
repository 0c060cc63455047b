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
#include "fdb_program.cuh"

namespace fdb {

// scalar Dual{1} pass for far terms in partial blocks; not inlined for the same reason as pg_run_shared_vec
template <bool NS>
__device__ __noinline__ void pg_run_far_scalar(const short (*code)[4], int ninstr, const double* consts, const ZeroLoader<1, NS>& x, int64_t i,
                                               Dual<double, 1, NS>* R, Dual<double, 1, NS>* acc) {
    pg_run<Dual<double, 1, NS>, 1, NS>(code, ninstr, consts, x, i, R, acc);
}

// the N-lane pass is used twice per kernel (near terms, epilogue): one out-of-line copy per (N, NS)
template <int N, bool NS>
__device__ __noinline__ void pg_run_near(const short (*code)[4], int ninstr, const double* consts, const SeedLoader<N, NS>& x, int64_t i,
                                         Dual<double, N, NS>* R, Dual<double, N, NS>* acc) {
    pg_run<Dual<double, N, NS>, N, NS>(code, ninstr, consts, x, i, R, acc);
}

template <int N, bool NS, int THREADS>
__global__ void __launch_bounds__(THREADS) program_gradient_kernel(const __grid_constant__ Program P, const bool SHARE, const double* __restrict__ x, int64_t n,
                                                                   int64_t chunk_lo, int64_t nchunks_total, int lastchunksize,
                                                                   double* __restrict__ grad, double* __restrict__ value_out,
                                                                   const __grid_constant__ fdb_mirrors M) {
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
    extern __shared__ double rfv[];                      // Dual{1} register file of the vectorised lane-shared path
    constexpr int V = PG_VEC;
    const int64_t nt = n - P.halo;
    const int64_t blk = (int64_t)V * THREADS;
    for (int64_t i0 = 0; i0 < nt; i0 += blk) {
        // block-uniform: a full block of V*THREADS terms none of which reads a seeded position
        const bool far_blk = SHARE && (i0 + blk <= nt) && ((i0 + blk - 1 + P.halo < start) || (i0 >= start + chunksize));
        if (far_blk) {
            pg_run_shared_vec<NS, THREADS, V>(P.term, P.n_term, P.consts, xz, i0 + threadIdx.x, rfv, acc1);
        } else {
            for (int v = 0; v < V; v++) {
                const int64_t i = i0 + (int64_t)v * THREADS + threadIdx.x;
                if (i >= nt) break;
                const int64_t w0 = i - (threadIdx.x & 31);
                const bool far = SHARE && ((w0 + 31 + P.halo < start) || (w0 >= start + chunksize));   // warp-uniform
                if (far) pg_run_far_scalar<NS>(P.term, P.n_term, P.consts, xz, i, R1, acc1);
                else pg_run_near<N, NS>(P.term, P.n_term, P.consts, xs, i, R, acc);
            }
        }
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
        pg_run_near<N, NS>(P.epi, P.n_epi, P.consts, xs, 0, R, acc);
        const D y = R[P.result_reg];
#pragma unroll
        for (int k = 0; k < N; k++)
            if (k < chunksize) mstore(M, &grad[start + k], y.p[k]);
        if (last && value_out) mstore(M, value_out, y.v);
    }
}

template <int N, bool NS, int T>
static int launch_program_gradient(fdb_ctx* ctx, unsigned grid, size_t smem, const Program& P, const double* x, int64_t n, int64_t chunk_lo,
                                   int64_t nchunks, int lcs, double* grad, double* vout, const fdb_mirrors& M) {
    static unsigned long long configured = 0;      // per device: several contexts (GPUs) may live in one process
    if (first_use_on_device(configured, ctx->device)) {
        cudaError_t e = cudaFuncSetAttribute(program_gradient_kernel<N, NS, T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)(PG_MAXR * 2 * PG_VEC * T * sizeof(double)));
        if (e != cudaSuccess) return fail(ctx, FDB_ERR_CUDA, "cudaFuncSetAttribute(program_gradient_kernel): %s", cudaGetErrorString(e));
    }
    program_gradient_kernel<N, NS, T><<<grid, T, smem, ctx->stream>>>(P, ctx->lane_sharing, x, n, chunk_lo, nchunks, lcs, grad, vout, M);
    return FDB_OK;
}

int build_program(fdb_ctx* ctx, Program& P, const int32_t* term, int n_term, const int32_t* epi, int n_epi, const double* consts,
                         int n_consts, int n_acc, int halo, int result_reg, bool array_form, int n_params) {
    if (n_term < 0 || n_term > PJ_MAXI || n_epi < 0 || n_epi > PJ_MAXI || n_consts < 0 || n_consts > PJ_MAXC || n_acc < 0 || n_acc > PJ_MAXK ||
        halo < 0 || result_reg < 0 || result_reg >= PJ_MAXR || (n_term && !term) || (n_epi && !epi) || (n_consts && !consts))
        return fail(ctx, FDB_ERR_UNSUPPORTED, "op-program exceeds the program limits (%d instr, %d consts, %d registers, %d accumulators)",
                    PJ_MAXI, PJ_MAXC, PJ_MAXR, PJ_MAXK);
    memset(&P, 0, sizeof P);
    P.n_term = n_term; P.n_epi = n_epi; P.n_acc = n_acc; P.halo = halo; P.result_reg = result_reg;
    int used_params = 0;
    P.term_halo = halo;
    for (int k = 0; k < FDB_MAX_PARAMS; k++) P.prm_halo[k] = (short)halo;
    int guard_left = 0, guard_h = halo, min_guard = halo, stage = 0, max_acc = -1, acc_seen = 0, max_reg = result_reg;
    bool unguarded_reads = false;
    for (int part = 0; part < 2; part++) {
        const int32_t* src = part ? epi : term; const int cnt = part ? n_epi : n_term;
        for (int i = 0; i < cnt; i++) {
            const int op = src[4 * i], dst = src[4 * i + 1], a = src[4 * i + 2], b = src[4 * i + 3];
            const int kind = op >> 6, code = op & 63;
            bool ok = dst >= 0 && dst < PJ_MAXR && kind >= PK_LOADX && kind <= PK_LOADS;
            if (kind != PK_ACC && kind != PK_GUARD && kind != PK_STAGE && dst > max_reg) max_reg = dst;
            if (kind == PK_STAGE) {
                ok = !part && guard_left == 0 && a == stage + 1 && a <= 2;
                if (ok) { stage = a; P.two_pass = 1; if (a == 1) P.k1 = max_acc + 1; }
            }
            if (kind == PK_LOADS) ok = ok && a >= 0 && a < PJ_MAXR && P.two_pass && (part || stage == 2);
            if (!part && stage == 1 && (kind == PK_LOADX || kind == PK_ACC || kind == PK_DP || kind == PK_PD || kind == PK_GUARD)) ok = false;   // scalar code only
            if (kind == PK_ACC && dst > max_acc) max_acc = dst;
            if (kind == PK_GUARD) {
                ok = !part && !array_form && guard_left == 0 && a >= 0 && a <= halo && b >= 1 && i + b <= n_term - 1;
                if (ok) { guard_left = b + 1; guard_h = a; P.has_guard = 1; if (a < min_guard) min_guard = a; }
            }
            const int h_here = (!part && guard_left > 0 && kind != PK_GUARD) ? guard_h : halo;
            if (!part && kind != PK_GUARD && guard_left == 0 && (kind == PK_LOADX || kind == PK_ACC || kind == PK_DP || kind == PK_PD)) unguarded_reads = true;
            if (kind == PK_LOADX && !part) ok = ok && b <= h_here;
            if (kind == PK_DP && b >= 0 && b < FDB_MAX_PARAMS && h_here < P.prm_halo[b]) P.prm_halo[b] = (short)h_here;
            if (kind == PK_PD && a >= 0 && a < FDB_MAX_PARAMS && h_here < P.prm_halo[a]) P.prm_halo[a] = (short)h_here;
            if (guard_left > 0) guard_left--;
            if (kind == PK_DP) { ok = ok && !part && code >= FDB_ADD && code <= FDB_POW && a >= 0 && a < PJ_MAXR && b >= 0 && b < n_params; if (b + 1 > used_params) used_params = b + 1; }
            if (kind == PK_PD) { ok = ok && !part && code >= FDB_ADD && code <= FDB_POW && a >= 0 && a < n_params && b >= 0 && b < PJ_MAXR; if (a + 1 > used_params) used_params = a + 1; }
            if (kind == PK_LOADX) ok = ok && !part && b >= 0 && b <= halo;
            if (kind == PK_UNARY) ok = ok && code >= FDB_NEG && code <= FDB_CBRT && a >= 0 && a < PJ_MAXR;
            if (kind == PK_DD) ok = ok && code >= FDB_ADD && code <= FDB_NANMIN && a >= 0 && a < PJ_MAXR && b >= 0 && b < PJ_MAXR;
            if (kind == PK_DC) ok = ok && code >= FDB_ADD && code <= FDB_POW && a >= 0 && a < PJ_MAXR && b >= 0 && b < n_consts;
            if (kind == PK_CD) ok = ok && code >= FDB_ADD && code <= FDB_POW && a >= 0 && a < n_consts && b >= 0 && b < PJ_MAXR;
            if (kind == PK_ACC) {          // b = 1: product accumulator (prod); an accumulator is either a sum or a product throughout
                // array-valued form: sums exist only in pass 1 of a two-pass program (softmax: exp.(x) ./ sum(exp.(x))), pass 2 yields y[i]
                ok = ok && !part && !(array_form && stage != 0) && dst < n_acc && a >= 0 && a < PJ_MAXR && (b == 0 || (b == 1 && stage == 0));
                if (ok) {
                    const int bit = 1 << dst;
                    if ((acc_seen & bit) && (((P.prod_mask & bit) != 0) != (b == 1))) ok = false;
                    acc_seen |= bit; if (b == 1) P.prod_mask |= bit;
                }
            }
            if (kind == PK_MOV) ok = ok && a >= 0 && a < PJ_MAXR;
            if (!ok) return fail(ctx, FDB_ERR_BADARG, "op-program: malformed instruction %d of the %s part", i, part ? "epilogue" : "term");
            short* d = part ? P.epi[i] : P.term[i];
            d[0] = (short)op; d[1] = (short)dst; d[2] = (short)a; d[3] = (short)b;
        }
    }
    for (int i = 0; i < n_consts; i++) P.consts[i] = consts[i];
    P.max_reg = max_reg;
    P.fits_interpreter = n_term <= PG_MAXI && n_epi <= PG_MAXI && n_consts <= PG_MAXC && n_acc <= PG_MAXK && max_reg < PG_MAXR && result_reg < PG_MAXR;
    P.n_params = used_params;
    if (array_form && max_acc >= 0 && !P.two_pass) return fail(ctx, FDB_ERR_BADARG, "op-program: an array-valued program with reductions needs the two-pass form");
    if (P.two_pass && (stage != 2 || P.has_guard))
        return fail(ctx, FDB_ERR_BADARG, "op-program: a two-pass program needs both stage markers and no guarded groups");
    if (P.has_guard) {
        if (unguarded_reads) return fail(ctx, FDB_ERR_BADARG, "op-program: with guards every term instruction must belong to a guarded group");
        P.term_halo = min_guard;
    }
    return FDB_OK;
}

int check_params(fdb_ctx* ctx, const Program& P, int64_t xlen, const char* what) {
    if (P.n_params > ctx->n_prm)
        return fail(ctx, FDB_ERR_BADARG, "%s: the program reads %d parameter arrays, %d are bound (fdb_program_bind_params)", what, P.n_params, ctx->n_prm);
    for (int k = 0; k < P.n_params; k++)
        if (ctx->prm_len[k] < xlen - P.prm_halo[k])
            return fail(ctx, FDB_ERR_SHAPE, "DimensionMismatch: parameter array %d has %lld entries, the broadcast has %lld", k,
                        (long long)ctx->prm_len[k], (long long)(xlen - P.prm_halo[k]));
    return FDB_OK;
}

}  // namespace fdb

using namespace fdb;

extern "C" int fdb_program_gradient_chunked(fdb_ctx* ctx, const int32_t* term_prog, int n_term, const int32_t* epi_prog, int n_epi,
                                            const double* consts, int n_consts, int n_acc, int halo, int result_reg, const fdb_array* x, int N,
                                            int64_t chunk_lo, int64_t chunk_hi, fdb_array* grad, fdb_array* value_dev) {
    FDB_ENTER(ctx);
    if (!ctx || !x || !grad) return fail(ctx, FDB_ERR_BADARG, "fdb_program_gradient_chunked: null argument");
    if (x->level != 0 || grad->level != 0 || (value_dev && value_dev->level != 0)) return fail(ctx, FDB_ERR_BADARG, "x, grad, value must be plain arrays");
    if (grad->n != x->n) return fail(ctx, FDB_ERR_SHAPE, "DimensionMismatch: gradient result has %lld entries, x has %lld", (long long)grad->n, (long long)x->n);
    const int64_t n = x->n;
    if (N < 1 || N > FDB_MAX_CHUNK) return fail(ctx, FDB_ERR_BADARG, "chunk size N=%d outside 1..12", N);
    if (n < N) return fail(ctx, FDB_ERR_CHUNK, "chunk size cannot be greater than ForwardDiff.structural_length(x) (%d > %lld)", N, (long long)n);
    Program P;
    int rc = build_program(ctx, P, term_prog, n_term, epi_prog, n_epi, consts, n_consts, n_acc, halo, result_reg, false, ctx->n_prm); if (rc) return rc;
    rc = check_params(ctx, P, x->n, "fdb_program_gradient_chunked"); if (rc) return rc;
    Plan p = make_plan(n, N);
    const int64_t nchunks = (n == N) ? 1 : p.nchunks;
    if (chunk_hi < 0) chunk_hi = nchunks;
    if (chunk_lo < 0 || chunk_lo > chunk_hi || chunk_hi > nchunks) return fail(ctx, FDB_ERR_BADARG, "bad chunk range");
    if (chunk_lo == chunk_hi) return FDB_OK;
    constexpr int T = 128;
    const unsigned grid = (unsigned)(chunk_hi - chunk_lo);
    double* vout = value_dev ? value_dev->data : nullptr;
    fdb_mirrors M;
    const fdb_array* outs[2] = {grad, value_dev};
    rc = mirrors_for(ctx, &M, outs, 2, "fdb_program_gradient_chunked"); if (rc) return rc;
    {   // specialised kernel for this program (fdb_jit.cu): the catalogue sweep kernel instantiated for the user's functor
        void* fn = nullptr;
        if (jit_get(ctx, 0, P, N, ctx->nansafe, ctx->lane_sharing, false, &fn) == FDB_OK) {
            const double* xp = x->data; int64_t nn = n, lo = chunk_lo, nc = nchunks; int lcs = (int)p.lastchunksize;
            double* gp = grad->data;
            void* args[] = {&xp, &nn, &lo, &nc, &lcs, &gp, &vout, &M};
            rc = jit_launch(ctx, fn, grid, 1, 256, args, "fdb_program_gradient_chunked(specialised)"); if (rc) return rc;
            return finish(ctx, 1, "fdb_program_gradient_chunked(specialised)");
        }
    }
    if (!P.fits_interpreter) return fail(ctx, FDB_ERR_UNSUPPORTED, "op-program beyond the interpreter's limits (%d instr, %d consts, %d registers, %d accumulators) needs the run-time specialised kernel (%s)",
                                         PG_MAXI, PG_MAXC, PG_MAXR, PG_MAXK, ctx->use_jit ? fdb_jit_last_log(ctx) : "disabled by fdb_set_jit");
    if (P.n_params || P.has_guard || P.two_pass || P.prod_mask) return fail(ctx, FDB_ERR_UNSUPPORTED, "op-program with array parameters, reductions of different lengths, products or two passes needs the run-time specialised kernel (%s)",
                                ctx->use_jit ? fdb_jit_last_log(ctx) : "disabled by fdb_set_jit");
    int nregs = 1;
    for (int i = 0; i < n_term; i++) { const int d = term_prog[4 * i + 1]; if ((term_prog[4 * i] >> 6) != PK_ACC && d + 1 > nregs) nregs = d + 1; }
    const size_t smem = (size_t)nregs * 2 * PG_VEC * T * sizeof(double);
    FDB_DISPATCH_N(N, FDB_DISPATCH_NS(ctx->nansafe,
        (rc = launch_program_gradient<NN, NS, T>(ctx, grid, smem, P, x->data, n, chunk_lo, nchunks, (int)p.lastchunksize, grad->data, vout, M))));
    if (rc) return rc;
    return finish(ctx, 1, "fdb_program_gradient_chunked");
}

// gradient(f, x::UpperTriangular / LowerTriangular / Diagonal, cfg) of a flattened f (src/apiutils.jl:44-71): x and grad hold the
// rows x rows column-major storage, chunks run over the structural positions, only those entries of grad are written.
// Single-pass programs, specialised kernels only.
extern "C" int fdb_program_gradient_chunked_structured(fdb_ctx* ctx, const int32_t* term_prog, int n_term, const int32_t* epi_prog, int n_epi,
                                                       const double* consts, int n_consts, int n_acc, int halo, int result_reg, int structure,
                                                       int64_t rows, const fdb_array* x, int N, int64_t chunk_lo, int64_t chunk_hi, fdb_array* grad,
                                                       fdb_array* value_dev) {
    FDB_ENTER(ctx);
    if (!ctx || !x || !grad) return fail(ctx, FDB_ERR_BADARG, "fdb_program_gradient_chunked_structured: null argument");
    if (x->level != 0 || grad->level != 0 || (value_dev && value_dev->level != 0)) return fail(ctx, FDB_ERR_BADARG, "x, grad, value must be plain arrays");
    if (structure < FDB_UPPER || structure > FDB_DIAGONAL) return fail(ctx, FDB_ERR_BADARG, "unknown structure %d (dense inputs: fdb_program_gradient_chunked)", structure);
    if (rows < 1 || x->n != rows * rows || grad->n != rows * rows) return fail(ctx, FDB_ERR_SHAPE, "DimensionMismatch: structured arrays need rows*rows storage");
    const int64_t ns = fdb_structural_length(structure, rows);
    if (N < 1 || N > FDB_MAX_CHUNK) return fail(ctx, FDB_ERR_BADARG, "chunk size N=%d outside 1..12", N);
    if (ns < N) return fail(ctx, FDB_ERR_CHUNK, "chunk size cannot be greater than ForwardDiff.structural_length(x) (%d > %lld)", N, (long long)ns);
    Program P;
    int rc = build_program(ctx, P, term_prog, n_term, epi_prog, n_epi, consts, n_consts, n_acc, halo, result_reg, false, ctx->n_prm); if (rc) return rc;
    rc = check_params(ctx, P, x->n, "fdb_program_gradient_chunked_structured"); if (rc) return rc;
    if (P.two_pass) return fail(ctx, FDB_ERR_UNSUPPORTED, "structured inputs: two-pass programs are not supported");
    Plan p = make_plan(ns, N);
    const int64_t nchunks = (ns == N) ? 1 : p.nchunks;
    if (chunk_hi < 0) chunk_hi = nchunks;
    if (chunk_lo < 0 || chunk_lo > chunk_hi || chunk_hi > nchunks) return fail(ctx, FDB_ERR_BADARG, "bad chunk range");
    if (chunk_lo == chunk_hi) return FDB_OK;
    fdb_mirrors M;
    const fdb_array* outs[2] = {grad, value_dev};
    rc = mirrors_for(ctx, &M, outs, 2, "fdb_program_gradient_chunked_structured"); if (rc) return rc;
    void* fn = nullptr;
    if (jit_get(ctx, 5, P, N, ctx->nansafe, ctx->lane_sharing, false, &fn) != FDB_OK)
        return fail(ctx, FDB_ERR_UNSUPPORTED, "structured inputs on the batched sweep need the run-time specialised kernel (%s)",
                    ctx->use_jit ? fdb_jit_last_log(ctx) : "disabled by fdb_set_jit");
    const double* xp = x->data; int64_t nx = x->n, rw = rows, nst = ns, lo = chunk_lo, nc = nchunks; int kind = structure, lcs = (int)p.lastchunksize;
    double* gp = grad->data; double* vout = value_dev ? value_dev->data : nullptr;
    void* args[] = {&xp, &nx, &kind, &rw, &nst, &lo, &nc, &lcs, &gp, &vout, &M};
    rc = jit_launch(ctx, fn, (unsigned)(chunk_hi - chunk_lo), 1, 256, args, "fdb_program_gradient_chunked_structured"); if (rc) return rc;
    return finish(ctx, 1, "fdb_program_gradient_chunked_structured");
}

extern "C" int fdb_program_bind_params(fdb_ctx* ctx, int count, const fdb_array* const* arrays) {
    FDB_ENTER(ctx);
    if (!ctx || count < 0 || count > FDB_MAX_PARAMS || (count && !arrays))
        return fail(ctx, FDB_ERR_BADARG, "fdb_program_bind_params: 0..%d plain arrays", FDB_MAX_PARAMS);
    for (int k = 0; k < count; k++)
        if (!arrays[k] || arrays[k]->level != 0) return fail(ctx, FDB_ERR_BADARG, "fdb_program_bind_params: parameter %d must be a plain (level-0) array", k);
    ctx->n_prm = count;
    for (int k = 0; k < FDB_MAX_PARAMS; k++) {
        ctx->prm_ptr[k] = k < count ? arrays[k]->data : nullptr;
        ctx->prm_len[k] = k < count ? arrays[k]->n : 0;
    }
    return FDB_OK;
}
