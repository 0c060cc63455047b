// fdb_program.cuh -- op-program representation and the per-thread interpreter shared by fdb_program.cu
// (gradient sweeps) and fdb_program_jac.cu (Jacobian sweeps).  See fdb_program.cu for the description.
#pragma once
#include <string.h>

#include "fdb_internal.h"
#include "fdb_mirror.cuh"
#include "fdb_targets.cuh"

namespace fdb {

// limits of the per-thread INTERPRETER (register file in local / shared memory, instruction words in the constant bank)
constexpr int PG_MAXI = 96;      // instructions per program part
constexpr int PG_MAXC = 24;      // Float64 constants
constexpr int PG_MAXR = 12;      // virtual Dual registers
constexpr int PG_MAXK = 4;       // sum accumulators
// limits of the program representation itself = what the run-time specialised kernels accept (there a "register" is a C++
// variable and spilling is the compiler's business)
constexpr int PJ_MAXI = 256, PJ_MAXC = 64, PJ_MAXR = 32, PJ_MAXK = 8;
constexpr int PG_VEC = 4;        // terms per thread per decoded instruction on the lane-shared path

struct Program {
    int n_term, n_epi, n_acc, halo, result_reg, n_params;
    int term_halo, has_guard;    // terms run for i in [0, n - term_halo); guards restrict groups whose views are shorter
    int prod_mask;               // bit k: accumulator k is a product (`prod`), PK_ACC with b = 1; specialised kernels only
    int two_pass, k1;            // PK_STAGE markers split the term part into term1 | epilogue1 | term2; accumulators [0, k1) belong to pass 1
    short prm_halo[FDB_MAX_PARAMS];   // parameter k is read for i < n - prm_halo[k]
    int fits_interpreter, max_reg;   // within the PG_* limits; highest register number used
    short term[PJ_MAXI][4];      // {op, dst, a, b}
    short epi[PJ_MAXI][4];
    double consts[PJ_MAXC];
};

enum { PK_LOADX = 0, PK_UNARY = 1, PK_DD = 2, PK_DC = 3, PK_CD = 4, PK_ACC = 5, PK_MOV = 6,
       PK_DP = 7, PK_PD = 8,     // Dual (x) param[b][i], param[a][i] (x) Dual: specialised kernels only (fdb_jit.cu)
       PK_GUARD = 9,             // the next b instructions run only while i + a < n (a reduction over a shorter view); ditto
       PK_STAGE = 10,            // a = 1: the first-stage epilogue follows (scalar Dual code on the pass-1 sums); a = 2: the
                                 // second-stage terms follow (they may read the epilogue's registers with PK_LOADS); ditto
       PK_LOADS = 11 };          // R[dst] = S[a], S = the register file as the first-stage epilogue left it

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
        case FDB_NANMAX: return nanmax_v(x, y);
        case FDB_NANMIN: return nanmin_v(x, y);
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
// The same pass for the lane-shared path (Dual{1}: value + one representative partial), with the virtual
// register file in SHARED memory -- rf[reg][component][thread], conflict free -- instead of a dynamically
// indexed local array (whose traffic spilled to L2/DRAM: ncu showed 1.4 GB of DRAM traffic per launch for
// the local-memory version), and the K accumulators in real registers (predicated update).
template <bool NS, int THREADS, class L>
FDB_DEV void pg_run_shared1(const short (*code)[4], int ninstr, const double* consts, const L& x, int64_t i,
                            double (*rf)[2][THREADS], Dual<double, 1, NS> (&acc)[PG_MAXK]) {
    typedef Dual<double, 1, NS> D1;
    const int t = threadIdx.x;
    auto get = [&](int r) { D1 d; d.v = rf[r][0][t]; d.p[0] = rf[r][1][t]; return d; };
    auto put = [&](int r, const D1& d) { rf[r][0][t] = d.v; rf[r][1][t] = d.p[0]; };
    for (int pc = 0; pc < ninstr; pc++) {
        const int op = code[pc][0], dst = code[pc][1], a = code[pc][2], b = code[pc][3];
        const int kind = op >> 6, c = op & 63;
        switch (kind) {
            case PK_LOADX: put(dst, x(i + b)); break;
            case PK_UNARY: put(dst, pg_unary<1, NS>(c, get(a))); break;
            case PK_DD: put(dst, pg_dd<1, NS>(c, get(a), get(b))); break;
            case PK_DC: put(dst, pg_dc<1, NS>(c, get(a), consts[b])); break;
            case PK_CD: put(dst, pg_cd<1, NS>(c, consts[a], get(b))); break;
            case PK_ACC: {
                const D1 v = get(a);
#pragma unroll
                for (int k = 0; k < PG_MAXK; k++) if (dst == k) acc[k] = acc[k] + v;
            } break;
            default: put(dst, get(a)); break;
        }
    }
}

// (not inlined: the Dual{1} paths do not depend on the chunk size N, one copy per NS serves all 12 kernels)
// Vector form of the lane-shared pass: each thread carries V terms (i, i+T, ..., i+(V-1)T) through one decode
// of every instruction, so the interpreter's per-instruction overhead (fetch, two-level dispatch, loop) is
// paid once per V terms.  rf[reg][component][v][thread] in dynamic shared memory (nregs registers only).
template <bool NS, int THREADS, int V, class L>
__device__ __noinline__ void pg_run_shared_vec(const short (*code)[4], int ninstr, const double* consts, const L& x, int64_t i0, double* rf,
                               Dual<double, 1, NS> (&acc)[PG_MAXK]) {
    typedef Dual<double, 1, NS> D1;
    const int t = threadIdx.x;
    auto at = [&](int r, int comp, int v) -> double& { return rf[((r * 2 + comp) * V + v) * THREADS + t]; };
    for (int pc = 0; pc < ninstr; pc++) {
        const int op = code[pc][0], dst = code[pc][1], a = code[pc][2], b = code[pc][3];
        const int kind = op >> 6, c = op & 63;
        D1 xa[V], xb[V], r[V];
        if (kind != PK_LOADX) {
#pragma unroll
            for (int v = 0; v < V; v++) { const int ra = (kind == PK_CD) ? b : a; xa[v].v = at(ra, 0, v); xa[v].p[0] = at(ra, 1, v); }
        }
        if (kind == PK_DD) {
#pragma unroll
            for (int v = 0; v < V; v++) { xb[v].v = at(b, 0, v); xb[v].p[0] = at(b, 1, v); }
        }
        // one flat dispatch per instruction; every case runs its V terms back to back
#define PGV(EXPR) { _Pragma("unroll") for (int v = 0; v < V; v++) r[v] = (EXPR); } break;
        const double s = (kind == PK_DC) ? consts[b] : ((kind == PK_CD) ? consts[a] : 0.0);
        switch (op) {
            case (PK_LOADX << 6): PGV(x(i0 + (int64_t)v * THREADS + b))
            case (PK_UNARY << 6) | FDB_NEG: PGV(-xa[v])
            case (PK_UNARY << 6) | FDB_EXP: PGV(exp_v(xa[v]))
            case (PK_UNARY << 6) | FDB_LOG: PGV(log_v(xa[v]))
            case (PK_UNARY << 6) | FDB_SIN: PGV(sin_v(xa[v]))
            case (PK_UNARY << 6) | FDB_COS: PGV(cos_v(xa[v]))
            case (PK_UNARY << 6) | FDB_TANH: PGV(tanh_v(xa[v]))
            case (PK_UNARY << 6) | FDB_SQRT: PGV(sqrt_v(xa[v]))
            case (PK_UNARY << 6) | FDB_ABS: PGV(abs_v(xa[v]))
            case (PK_UNARY << 6) | FDB_ABS2: PGV(abs2_v(xa[v]))
            case (PK_UNARY << 6) | FDB_INV: PGV(inv_v(xa[v]))
            case (PK_UNARY << 6) | FDB_POW0: PGV(pow0(xa[v]))
            case (PK_UNARY << 6) | FDB_POW1: PGV(pow1(xa[v]))
            case (PK_UNARY << 6) | FDB_POW2: PGV(pow2(xa[v]))
            case (PK_UNARY << 6) | FDB_POW3: PGV(pow3(xa[v]))
            case (PK_UNARY << 6) | FDB_TAN: PGV(tan_v(xa[v]))
            case (PK_UNARY << 6) | FDB_EXP2: PGV(exp2_v(xa[v]))
            case (PK_UNARY << 6) | FDB_LOG2: PGV(log2_v(xa[v]))
            case (PK_UNARY << 6) | FDB_LOG10: PGV(log10_v(xa[v]))
            case (PK_UNARY << 6) | FDB_LOG1P: PGV(log1p_v(xa[v]))
            case (PK_UNARY << 6) | FDB_EXPM1: PGV(expm1_v(xa[v]))
            case (PK_UNARY << 6) | FDB_SINH: PGV(sinh_v(xa[v]))
            case (PK_UNARY << 6) | FDB_COSH: PGV(cosh_v(xa[v]))
            case (PK_UNARY << 6) | FDB_ASIN: PGV(asin_v(xa[v]))
            case (PK_UNARY << 6) | FDB_ACOS: PGV(acos_v(xa[v]))
            case (PK_UNARY << 6) | FDB_ATAN: PGV(atan_v(xa[v]))
            case (PK_UNARY << 6) | FDB_CBRT: PGV(cbrt_v(xa[v]))
            case (PK_DD << 6) | FDB_ADD: PGV(xa[v] + xb[v])
            case (PK_DD << 6) | FDB_SUB: PGV(xa[v] - xb[v])
            case (PK_DD << 6) | FDB_MUL: PGV(xa[v] * xb[v])
            case (PK_DD << 6) | FDB_DIV: PGV(xa[v] / xb[v])
            case (PK_DD << 6) | FDB_POW: PGV(pow_v(xa[v], xb[v]))
            case (PK_DD << 6) | FDB_MAX: PGV(max_v(xa[v], xb[v]))
            case (PK_DD << 6) | FDB_MIN: PGV(min_v(xa[v], xb[v]))
            case (PK_DD << 6) | FDB_ATAN2: PGV(atan2_v(xa[v], xb[v]))
            case (PK_DD << 6) | FDB_HYPOT: PGV(hypot_v(xa[v], xb[v]))
            case (PK_DD << 6) | FDB_NANMAX: PGV(nanmax_v(xa[v], xb[v]))
            case (PK_DD << 6) | FDB_NANMIN: PGV(nanmin_v(xa[v], xb[v]))
            case (PK_DC << 6) | FDB_ADD: PGV(xa[v] + s)
            case (PK_DC << 6) | FDB_SUB: PGV(xa[v] - s)
            case (PK_DC << 6) | FDB_MUL: PGV(xa[v] * s)
            case (PK_DC << 6) | FDB_DIV: PGV(xa[v] / s)
            case (PK_DC << 6) | FDB_POW: PGV(pow_v(xa[v], s))
            case (PK_CD << 6) | FDB_ADD: PGV(s + xa[v])
            case (PK_CD << 6) | FDB_SUB: PGV(s - xa[v])
            case (PK_CD << 6) | FDB_MUL: PGV(s * xa[v])
            case (PK_CD << 6) | FDB_DIV: PGV(s / xa[v])
            case (PK_CD << 6) | FDB_POW: PGV(pow_v(s, xa[v]))
            case (PK_ACC << 6):
#pragma unroll
                for (int k = 0; k < PG_MAXK; k++)
                    if (dst == k) {
#pragma unroll
                        for (int v = 0; v < V; v++) acc[k] = acc[k] + xa[v];
                    }
                continue;
            default: PGV(xa[v])
        }
#undef PGV
#pragma unroll
        for (int v = 0; v < V; v++) { at(dst, 0, v) = r[v].v; at(dst, 1, v) = r[v].p[0]; }
    }
}

struct SumK {   // block_combine policy: K sum accumulators
    static constexpr int K = PG_MAXK;
    static constexpr bool PURE_X = false;
    template <class D> static FDB_DEV D init(int, const D& like) { return zero_of(like); }
    template <class D> static FDB_DEV D combine(int, const D& a, const D& b) { return a + b; }
};


// fdb_jit.cu: the kernel specialised for this program (kind 0 gradient, 1 Jacobian, 2 Hessian), or non-zero
int jit_get(fdb_ctx* ctx, int kind, const Program& P, int N, bool ns, bool share, bool mir, void** fn);
// the bound parameter arrays cover `nterms` entries each (and there are as many as the program reads)
int check_params(fdb_ctx* ctx, const Program& P, int64_t xlen, const char* what);

int build_program(fdb_ctx* ctx, Program& P, const int32_t* term, int n_term, const int32_t* epi, int n_epi, const double* consts,
                  int n_consts, int n_acc, int halo, int result_reg, bool array_form, int n_params);

}  // namespace fdb
