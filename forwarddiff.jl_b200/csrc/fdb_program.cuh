// fdb_program.cuh -- op-program representation and the per-thread interpreter shared by fdb_program.cu
// (gradient sweeps) and fdb_program_jac.cu (Jacobian sweeps).  See fdb_program.cu for the description.
#pragma once
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

struct SumK {   // block_combine policy: K sum accumulators
    static constexpr int K = PG_MAXK;
    template <class D> static FDB_DEV D init(int, const D& like) { return zero_of(like); }
    template <class D> static FDB_DEV D combine(int, const D& a, const D& b) { return a + b; }
};


int build_program(fdb_ctx* ctx, Program& P, const int32_t* term, int n_term, const int32_t* epi, int n_epi, const double* consts,
                  int n_consts, int n_acc, int halo, int result_reg, bool array_form);

}  // namespace fdb
