// fdb_jit.cu -- run-time specialisation of the batched sweep kernels for a user function.
//
// The reference is a JIT-specialised system: Julia compiles `f` for Vector{Dual{T,Float64,N}} the first time
// gradient(f, x, cfg) runs (src/gradient.jl:135,144,151 call a type-specialised f).  The op-program drivers
// (fdb_program_*_chunked) receive the same information -- f flattened to a short list of src/dual.jl
// operations -- and by default interpret it per thread (fdb_program.cu).  Here the program is instead turned
// into a functor with the interface of the catalogue targets in fdb_targets.cuh (K, HALO, term, finalize,
// apply): one C++ statement per instruction, calling the same Dual device functions as every other kernel, so
// the arithmetic per operation is identical.  The hand-written kernel templates -- gradient_sweep_body
// (fdb_sweep_kernel.cuh), hessian_sweep_body (fdb_hessian_kernel.cuh), elementwise_jacobian_body
// (fdb_jac_kernel.cuh) -- are then instantiated for that functor by NVRTC for sm_100a (-fmad=false, like the
// library) and the cubin is loaded through the driver API.  Specialised kernels are cached per context by
// (kind, program, constants, N, nansafe, lane sharing, mirroring).  No kernel is generated: the generated text
// is only the functor body.  If libnvrtc is not installed or a compile fails, the drivers keep interpreting.
#include <cuda.h>
#include <dlfcn.h>
#include <nvrtc.h>
#include <string.h>

#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "fdb_program.cuh"

#include "../build/fdb_rtc_headers.inc"

struct fdb_jit_cache {
    struct Entry { CUmodule mod; CUfunction fn; CUdeviceptr prm; };   // prm: the module's fdb_prm[] (parameter array pointers)
    std::unordered_map<std::string, Entry> map;
    int64_t compiled = 0, hits = 0, failures = 0;
    std::string last_log;
};

namespace fdb {

// ---- libnvrtc through dlopen (no link-time dependency) and the driver's module API through the runtime ----------
struct Nvrtc {
    void* so = nullptr;
    nvrtcResult (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
    nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
    nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t*) = nullptr;
    nvrtcResult (*GetCUBIN)(nvrtcProgram, char*) = nullptr;
    nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
    nvrtcResult (*GetProgramLog)(nvrtcProgram, char*) = nullptr;
    nvrtcResult (*DestroyProgram)(nvrtcProgram*) = nullptr;
    const char* (*GetErrorString)(nvrtcResult) = nullptr;
    bool ok = false;
};
static Nvrtc& nvrtc() {
    static Nvrtc n;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so"};
        for (const char* nm : names) { n.so = dlopen(nm, RTLD_NOW | RTLD_LOCAL); if (n.so) break; }
        if (!n.so) return;
#define FDB_SYM(field, sym) *(void**)(&n.field) = dlsym(n.so, sym); if (!n.field) return;
        FDB_SYM(CreateProgram, "nvrtcCreateProgram") FDB_SYM(CompileProgram, "nvrtcCompileProgram")
        FDB_SYM(GetCUBINSize, "nvrtcGetCUBINSize") FDB_SYM(GetCUBIN, "nvrtcGetCUBIN")
        FDB_SYM(GetProgramLogSize, "nvrtcGetProgramLogSize") FDB_SYM(GetProgramLog, "nvrtcGetProgramLog")
        FDB_SYM(DestroyProgram, "nvrtcDestroyProgram") FDB_SYM(GetErrorString, "nvrtcGetErrorString")
#undef FDB_SYM
        n.ok = true;
    });
    return n;
}

struct Driver {
    CUresult (*ModuleLoadData)(CUmodule*, const void*) = nullptr;
    CUresult (*ModuleGetFunction)(CUfunction*, CUmodule, const char*) = nullptr;
    CUresult (*ModuleUnload)(CUmodule) = nullptr;
    CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void**, void**) = nullptr;
    CUresult (*GetErrorString)(CUresult, const char**) = nullptr;
    CUresult (*ModuleGetGlobal)(CUdeviceptr*, size_t*, CUmodule, const char*) = nullptr;
    CUresult (*MemcpyHtoDAsync)(CUdeviceptr, const void*, size_t, CUstream) = nullptr;
    bool ok = false;
};
static Driver& driver() {
    static Driver d;
    static std::once_flag once;
    std::call_once(once, [] {
        cudaDriverEntryPointQueryResult q;
#define FDB_DRV(field, sym) if (cudaGetDriverEntryPoint(sym, (void**)&d.field, cudaEnableDefault, &q) != cudaSuccess || !d.field) return;
        FDB_DRV(ModuleLoadData, "cuModuleLoadData") FDB_DRV(ModuleGetFunction, "cuModuleGetFunction")
        FDB_DRV(ModuleUnload, "cuModuleUnload") FDB_DRV(LaunchKernel, "cuLaunchKernel") FDB_DRV(GetErrorString, "cuGetErrorString")
        FDB_DRV(ModuleGetGlobal, "cuModuleGetGlobal") FDB_DRV(MemcpyHtoDAsync, "cuMemcpyHtoDAsync")
#undef FDB_DRV
        d.ok = true;
    });
    return d;
}

// ---- program -> functor text ------------------------------------------------------------------------------------
static const char* unary_fn(int c) {
    switch (c) {
        case FDB_EXP: return "exp_v"; case FDB_LOG: return "log_v"; case FDB_SIN: return "sin_v"; case FDB_COS: return "cos_v";
        case FDB_TANH: return "tanh_v"; case FDB_SQRT: return "sqrt_v"; case FDB_ABS: return "abs_v"; case FDB_ABS2: return "abs2_v";
        case FDB_INV: return "inv_v"; case FDB_POW0: return "pow0"; case FDB_POW1: return "pow1"; case FDB_POW2: return "pow2";
        case FDB_POW3: return "pow3"; case FDB_TAN: return "tan_v"; case FDB_EXP2: return "exp2_v"; case FDB_LOG2: return "log2_v";
        case FDB_LOG10: return "log10_v"; case FDB_LOG1P: return "log1p_v"; case FDB_EXPM1: return "expm1_v"; case FDB_SINH: return "sinh_v";
        case FDB_COSH: return "cosh_v"; case FDB_ASIN: return "asin_v"; case FDB_ACOS: return "acos_v"; case FDB_ATAN: return "atan_v";
        default: return "cbrt_v";
    }
}
static std::string binary_expr(int c, const std::string& a, const std::string& b) {
    switch (c) {
        case FDB_ADD: return a + " + " + b;
        case FDB_SUB: return a + " - " + b;
        case FDB_MUL: return a + " * " + b;
        case FDB_DIV: return a + " / " + b;
        case FDB_POW: return "pow_v(" + a + ", " + b + ")";
        case FDB_MAX: return "max_v(" + a + ", " + b + ")";
        case FDB_MIN: return "min_v(" + a + ", " + b + ")";
        case FDB_ATAN2: return "atan2_v(" + a + ", " + b + ")";
        case FDB_NANMAX: return "nanmax_v(" + a + ", " + b + ")";
        case FDB_NANMIN: return "nanmin_v(" + a + ", " + b + ")";
        default: return "hypot_v(" + a + ", " + b + ")";
    }
}
static std::string R(int r) { return "r" + std::to_string(r); }
static std::string C(int c) { return "c" + std::to_string(c); }

// one statement per instruction, the same evaluation order as pg_run (fdb_program.cuh)
static std::string emit_part(const short (*code)[4], int n, int from = 0) {
    std::string s;
    int close_at = -1;
    for (int pc = from; pc < n; pc++) {
        const int op = code[pc][0], dst = code[pc][1], a = code[pc][2], b = code[pc][3];
        const int kind = op >> 6, c = op & 63;
        s += "        ";
        switch (kind) {
            case PK_GUARD: s += "if (i + " + std::to_string(a) + " < n) {"; close_at = pc + b; break;
            case PK_LOADS: s += R(dst) + " = S[" + std::to_string(a) + "];"; break;
            case PK_LOADX: s += R(dst) + " = x(i + " + std::to_string(b) + ");"; break;
            case PK_UNARY: s += R(dst) + " = " + (c == FDB_NEG ? "-" + R(a) : std::string(unary_fn(c)) + "(" + R(a) + ")") + ";"; break;
            case PK_DD: s += R(dst) + " = " + binary_expr(c, R(a), R(b)) + ";"; break;
            case PK_DC: s += R(dst) + " = " + binary_expr(c, R(a), C(b)) + ";"; break;
            case PK_CD: s += R(dst) + " = " + binary_expr(c, C(a), R(b)) + ";"; break;
            case PK_DP: s += R(dst) + " = " + binary_expr(c, R(a), "fdb_prm[" + std::to_string(b) + "][i]") + ";"; break;
            case PK_PD: s += R(dst) + " = " + binary_expr(c, "fdb_prm[" + std::to_string(a) + "][i]", R(b)) + ";"; break;
            case PK_ACC: s += "acc[" + std::to_string(dst) + "] = acc[" + std::to_string(dst) + (b == 1 ? "] * " : "] + ") + R(a) + ";"; break;
            default: s += R(dst) + " = " + R(a) + ";"; break;
        }
        s += "\n";
        if (pc == close_at) { s += "        }\n"; close_at = -1; }
    }
    return s;
}
static std::string decl_regs() {
    std::string s = "        D";
    for (int r = 0; r < PJ_MAXR; r++) s += std::string(r ? ", " : " ") + R(r) + "{}";
    return s + ";\n";
}

enum { JIT_GRADIENT = 0, JIT_JACOBIAN = 1, JIT_HESSIAN = 2, JIT_MATVEC_JACOBIAN = 3, JIT_JACOBIAN2 = 4, JIT_GRADIENT_STRUCT = 5 };

static std::string make_source(int kind, const Program& P, int N, bool ns, bool share, bool mir) {
    std::string s = "#include \"fdb_sweep_kernel.cuh\"\n#include \"fdb_hessian_kernel.cuh\"\n#include \"fdb_jac_kernel.cuh\"\n";
    // closed-over data arrays of the user's f: pointers uploaded into the module before each launch (jit_launch)
    s += "extern \"C\" { __constant__ const double* fdb_prm[" + std::to_string(FDB_MAX_PARAMS) + "]; }\n";
    s += "namespace fdb {\n";
    char buf[96];
    for (int i = 0; i < PJ_MAXC; i++) {            // constants by bit pattern: exact for every double including Inf/NaN
        long long bits; memcpy(&bits, &P.consts[i], 8);
        snprintf(buf, sizeof buf, "#define c%d __longlong_as_double(0x%016llxLL)\n", i, (unsigned long long)bits);
        s += buf;
    }
    const int K = P.n_acc > 0 ? P.n_acc : 1;
    s += "struct UserFn {\n";
    s += std::string("    static constexpr bool PURE_X = ") + ((P.n_params == 0 && !P.has_guard) ? "true" : "false") + ";\n";
    s += std::string("    static constexpr bool PLAIN_SUM = ") + ((P.n_epi == 0 && K == 1 && P.result_reg == 0 && !P.prod_mask) ? "true" : "false") + ";\n";
    s += "    static constexpr int K = " + std::to_string(K) + ";\n    static constexpr int HALO = " + std::to_string(P.halo) + ";\n";
    s += "    static __device__ int64_t nterms(int64_t n) { return n > " + std::to_string(P.term_halo) + " ? n - " + std::to_string(P.term_halo) + " : 0; }\n";
    if (P.prod_mask) {          // accumulators of `prod` start at one(Dual) and combine by Dual*Dual (src/dual.jl:260-300)
        const std::string pm = std::to_string(P.prod_mask);
        s += "    template <class D> static FDB_DEV D init(int k, const D& like) { return ((" + pm + " >> k) & 1) ? one_of(like) : zero_of(like); }\n";
        s += "    template <class D> static FDB_DEV D combine(int k, const D& a, const D& b) { return ((" + pm + " >> k) & 1) ? a * b : a + b; }\n";
    } else {
        s += "    template <class D> static FDB_DEV D init(int, const D& like) { return zero_of(like); }\n";
        s += "    template <class D> static FDB_DEV D combine(int, const D& a, const D& b) { return a + b; }\n";
    }
    int st1 = -1, st2 = -1;                         // positions of the stage markers of a two-pass program
    for (int pc = 0; pc < P.n_term; pc++)
        if ((P.term[pc][0] >> 6) == PK_STAGE) { if (P.term[pc][2] == 1) st1 = pc; else st2 = pc; }
    s += std::string("    static constexpr bool TWO_PASS = ") + (P.two_pass ? "true" : "false") + ";\n";
    s += "    static constexpr int NSCAL = " + std::to_string(PJ_MAXR) + ";\n";
    // term: the user's broadcast body for one i, adding into the sum accumulators (pass 1 of a two-pass program)
    s += "    template <class D, class L> static FDB_DEV void term(const L& x, int64_t i, int64_t n, D (&acc)[K]) {\n" + decl_regs() +
         emit_part(P.term, P.two_pass ? st1 : P.n_term) + "    }\n";
    if (P.two_pass) {
        // finalize1: scalar Dual code on the pass-1 sums; the whole register file is published as S
        s += "    template <class D> static FDB_DEV void finalize1(const D (&acc)[K], D* S) {\n" + decl_regs();
        for (int k = 0; k < P.k1 && k < PJ_MAXK; k++) s += "        " + R(k) + " = acc[" + std::to_string(k) + "];\n";
        s += emit_part(P.term, st2, st1 + 1);
        for (int r = 0; r < PJ_MAXR; r++) s += "        S[" + std::to_string(r) + "] = " + R(r) + ";\n";
        s += "    }\n";
        // term2: the second broadcast, reading x and the scalars S
        s += "    template <class D, class L> static FDB_DEV void term2(const L& x, int64_t i, int64_t n, const D* S, D (&acc)[K]) {\n" + decl_regs() +
             emit_part(P.term, P.n_term, st2 + 1) + "    }\n";
        // apply2: the array-valued second broadcast, y[i] = result register (jacobian_sweep2_body)
        s += "    template <class D, class L> static FDB_DEV D apply2(const L& x, int64_t i, int64_t n, const D* S) {\n        D acc[K]; (void)acc;\n" + decl_regs() +
             emit_part(P.term, P.n_term, st2 + 1) + "        return " + R(P.result_reg) + ";\n    }\n";
        // finalize2: registers k1..K-1 start as the pass-2 sums
        s += "    template <class D> static FDB_DEV D finalize2(const D (&acc)[K], const D* S, int64_t) {\n" + decl_regs();
        for (int k = P.k1; k < K && k < PJ_MAXK; k++) s += "        " + R(k) + " = acc[" + std::to_string(k) + "];\n";
        s += emit_part(P.epi, P.n_epi) + "        return " + R(P.result_reg) + ";\n    }\n";
    }
    // finalize: scalar Dual arithmetic after the reductions; registers 0..K-1 start as the reduced accumulators
    s += "    template <class D> static FDB_DEV D finalize(const D (&acc)[K], int64_t) {\n" + decl_regs();
    for (int k = 0; k < K && k < PJ_MAXK; k++) s += "        " + R(k) + " = acc[" + std::to_string(k) + "];\n";
    s += (P.two_pass ? std::string() : emit_part(P.epi, P.n_epi)) + "        return " + R(P.result_reg) + ";\n    }\n";
    // apply: array-valued form, y[i] = result register after the term part
    if (!P.two_pass)
        s += "    template <class D, class L> static FDB_DEV D apply(const L& x, int64_t i) {\n        D acc[K];\n        const int64_t n = 0; (void)n;\n" + decl_regs() +
             emit_part(P.term, P.n_term) + "        (void)acc;\n        return " + R(P.result_reg) + ";\n    }\n";
    s += "};\n}  // namespace fdb\n";
    const std::string targs = "fdb::UserFn, " + std::to_string(N) + ", " + (ns ? "true" : "false");
    if (kind == JIT_GRADIENT) {
        s += "extern \"C\" __global__ void __launch_bounds__(256) fdb_user_kernel(const double* __restrict__ x, int64_t n, int64_t chunk_lo,\n"
             "        int64_t nchunks_total, int lastchunksize, double* __restrict__ grad, double* __restrict__ value_out,\n"
             "        const __grid_constant__ fdb_mirrors M) {\n"
             "    fdb::" + std::string(P.two_pass ? "gradient_sweep2_body<" : "gradient_sweep_body<") + targs + ", 256, " + (share ? "true" : "false") +
             ">(x, n, chunk_lo, nchunks_total, lastchunksize, grad, value_out, M);\n}\n";
    } else if (kind == JIT_JACOBIAN) {
        s += "extern \"C\" __global__ void __launch_bounds__(128) fdb_user_kernel(const bool share, const double* __restrict__ x, int64_t n,\n"
             "        int64_t chunk_lo, int64_t nchunks_total, int lastchunksize, double* __restrict__ J, int64_t ldJ, double* __restrict__ yout,\n"
             "        const unsigned nrb, const unsigned nch, const bool rowfast, const __grid_constant__ fdb_mirrors M) {\n"
             "    fdb::elementwise_jacobian_body<" + targs + ", 128, " + (mir ? "true" : "false") +
             ">(share, x, n, chunk_lo, nchunks_total, lastchunksize, J, ldJ, yout, nrb, nch, rowfast, M);\n}\n";
    } else if (kind == JIT_GRADIENT_STRUCT) {
        s += "extern \"C\" __global__ void __launch_bounds__(256) fdb_user_kernel(const double* __restrict__ x, int64_t nx, int kind, int64_t rows,\n"
             "        int64_t nstruct, int64_t chunk_lo, int64_t nchunks_total, int lastchunksize, double* __restrict__ grad,\n"
             "        double* __restrict__ value_out, const __grid_constant__ fdb_mirrors M) {\n"
             "    fdb::gradient_sweep_struct_body<" + targs + ", 256, " + (share ? "true" : "false") +
             ">(x, nx, kind, rows, nstruct, chunk_lo, nchunks_total, lastchunksize, grad, value_out, M);\n}\n";
    } else if (kind == JIT_JACOBIAN2) {
        s += "extern \"C\" __global__ void __launch_bounds__(256) fdb_user_kernel(const double* __restrict__ x, int64_t n, int64_t chunk_lo,\n"
             "        int64_t nchunks_total, int lastchunksize, double* __restrict__ J, int64_t ldJ, double* __restrict__ yout,\n"
             "        const __grid_constant__ fdb_mirrors M) {\n"
             "    fdb::jacobian_sweep2_body<" + targs + ", 256, " + (share ? "true" : "false") + ", " + (mir ? "true" : "false") +
             ">(x, n, chunk_lo, nchunks_total, lastchunksize, J, ldJ, yout, M);\n}\n";
    } else if (kind == JIT_MATVEC_JACOBIAN) {
        s += "extern \"C\" __global__ void __launch_bounds__(128) fdb_user_kernel(const bool share, const double* __restrict__ Y, int64_t ldy,\n"
             "        const double* __restrict__ A, int64_t lda, int64_t m, int64_t n, int64_t chunk_lo, int64_t nchunks_total, int lastchunksize,\n"
             "        double* __restrict__ J, int64_t ldJ, double* __restrict__ yout, const unsigned nrb, const unsigned nch, const bool rowfast,\n"
             "        const __grid_constant__ fdb_mirrors M) {\n"
             "    fdb::matvec_jacobian_body<" + targs + ", 128, " + (mir ? "true" : "false") +
             ">(share, Y, ldy, A, lda, m, n, chunk_lo, nchunks_total, lastchunksize, J, ldJ, yout, nrb, nch, rowfast, M);\n}\n";
    } else {
        s += "extern \"C\" __global__ void __launch_bounds__(128) fdb_user_kernel(const double* __restrict__ x, int64_t n, int64_t outer_lo,\n"
             "        int64_t nchunks_total, int lastchunksize, double* __restrict__ H, int64_t ldH, double* __restrict__ grad,\n"
             "        double* __restrict__ value_out, const __grid_constant__ fdb_mirrors M) {\n"
             "    if constexpr (fdb::UserFn::PLAIN_SUM) fdb::hessian_sweep_body<" + targs + ", 128>(x, n, outer_lo, nchunks_total, lastchunksize, H, ldH, grad, value_out, M);\n"
             "    else fdb::hessian_general_body<" + targs + ", 128>(x, n, outer_lo, nchunks_total, lastchunksize, H, ldH, grad, value_out, M);\n}\n";
    }
    return s;
}

// NVRTC: source -> sm_100a cubin.  Pure host code (no GPU needed), so the CPU test suite exercises it too.
static int compile_source(const std::string& src, std::vector<char>& cubin, std::string& log) {
    Nvrtc& n = nvrtc();
    if (!n.ok) { log = "libnvrtc not available"; return FDB_ERR_UNSUPPORTED; }
    const char* hdr_src[] = {rtc_fdb_dual_cuh, rtc_fdb_targets_cuh, rtc_fdb_mirror_cuh, rtc_fdb_sweep_kernel_cuh, rtc_fdb_hessian_kernel_cuh,
                             rtc_fdb_jac_kernel_cuh};
    const char* hdr_name[] = {"fdb_dual.cuh", "fdb_targets.cuh", "fdb_mirror.cuh", "fdb_sweep_kernel.cuh", "fdb_hessian_kernel.cuh", "fdb_jac_kernel.cuh"};
    nvrtcProgram prog;
    nvrtcResult r = n.CreateProgram(&prog, src.c_str(), "fdb_user.cu", 6, hdr_src, hdr_name);
    if (r != NVRTC_SUCCESS) { log = std::string("nvrtcCreateProgram: ") + n.GetErrorString(r); return FDB_ERR_CUDA; }
    // -fmad=false: the reference never contracts a*b+c (src/partials.jl:222-224), same flag as the library build
    const char* opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-fmad=false", "-lineinfo"};
    r = n.CompileProgram(prog, 4, opts);
    size_t ls = 0;
    if (n.GetProgramLogSize(prog, &ls) == NVRTC_SUCCESS && ls > 1) { log.resize(ls); n.GetProgramLog(prog, &log[0]); }
    if (r != NVRTC_SUCCESS) {
        log = std::string("nvrtcCompileProgram: ") + n.GetErrorString(r) + "\n" + log;
        n.DestroyProgram(&prog);
        return FDB_ERR_UNSUPPORTED;
    }
    size_t cs = 0;
    r = n.GetCUBINSize(prog, &cs);
    if (r == NVRTC_SUCCESS && cs > 0) { cubin.resize(cs); r = n.GetCUBIN(prog, cubin.data()); }
    n.DestroyProgram(&prog);
    if (r != NVRTC_SUCCESS || cs == 0) { log = "nvrtcGetCUBIN failed"; return FDB_ERR_CUDA; }
    return FDB_OK;
}

static std::string cache_key(int kind, const Program& P, int N, bool ns, bool share, bool mir) {
    std::string k(reinterpret_cast<const char*>(&P), sizeof P);
    const char tail[6] = {(char)kind, (char)N, (char)ns, (char)share, (char)mir, 0};
    return k.append(tail, 6);
}

// the specialised kernel for (kind, P, N, ...): from the context's cache, or compiled now.  A non-zero return means
// "not available" (the caller interprets instead); the reason is kept in the cache's log.
int jit_get(fdb_ctx* ctx, int kind, const Program& P, int N, bool ns, bool share, bool mir, void** fn) {
    *fn = nullptr;
    if (!ctx->use_jit) return FDB_ERR_UNSUPPORTED;
    if (!ctx->jit) ctx->jit = new fdb_jit_cache();
    fdb_jit_cache& J = *ctx->jit;
    if (kind != JIT_GRADIENT && kind != JIT_JACOBIAN2 && kind != JIT_GRADIENT_STRUCT) share = false;          // run-time flag (Jacobian) or not applicable (Hessian)
    if (kind != JIT_JACOBIAN && kind != JIT_MATVEC_JACOBIAN && kind != JIT_JACOBIAN2) mir = false;
    const std::string key = cache_key(kind, P, N, ns, share, mir);
    auto it = J.map.find(key);
    if (it != J.map.end()) { J.hits++; *fn = it->second.fn ? &it->second : nullptr; return it->second.fn ? FDB_OK : FDB_ERR_UNSUPPORTED; }
    fdb_jit_cache::Entry e{nullptr, nullptr, 0};
    Driver& d = driver();
    if (J.map.size() >= 512) {                         // e.g. a caller that changes a literal constant on every call
        cudaStreamSynchronize(ctx->stream);            // nothing queued may still run a module that is about to go
        for (auto& kv : J.map) if (kv.second.mod && d.ok) d.ModuleUnload(kv.second.mod);
        J.map.clear();
    }
    std::vector<char> cubin;
    std::string log;
    int rc = d.ok ? compile_source(make_source(kind, P, N, ns, share, mir), cubin, log) : FDB_ERR_UNSUPPORTED;
    if (!d.ok) log = "driver module API not available";
    if (rc == FDB_OK) {
        cudaSetDevice(ctx->device);
        cudaFree(0);                                   // the runtime's primary context is current for the driver calls
        CUresult cr = d.ModuleLoadData(&e.mod, cubin.data());
        if (cr == CUDA_SUCCESS) cr = d.ModuleGetFunction(&e.fn, e.mod, "fdb_user_kernel");
        size_t pb = 0;
        if (cr == CUDA_SUCCESS && P.n_params > 0) cr = d.ModuleGetGlobal(&e.prm, &pb, e.mod, "fdb_prm");
        if (cr != CUDA_SUCCESS) {
            const char* es = nullptr; d.GetErrorString(cr, &es);
            log = std::string("cuModuleLoadData/GetFunction: ") + (es ? es : "?");
            if (e.mod) d.ModuleUnload(e.mod);
            e.mod = nullptr; e.fn = nullptr; rc = FDB_ERR_CUDA;
        }
    }
    if (rc == FDB_OK) J.compiled++; else { J.failures++; J.last_log = log; }
    auto ins = J.map.emplace(key, e);                  // failures are cached too: no recompilation per call
    *fn = e.fn ? &ins.first->second : nullptr;         // (node addresses of an unordered_map are stable)
    return rc;
}

int jit_launch(fdb_ctx* ctx, void* entry, unsigned gridx, unsigned gridy, unsigned threads, void** args, const char* what) {
    Driver& d = driver();
    fdb_jit_cache::Entry* e = static_cast<fdb_jit_cache::Entry*>(entry);
    CUresult cr = CUDA_SUCCESS;
    if (ctx->n_prm > 0 && e->prm)      // stream-ordered, so a launch already queued keeps the pointers it was given
        cr = d.MemcpyHtoDAsync(e->prm, ctx->prm_ptr, sizeof(const double*) * FDB_MAX_PARAMS, (CUstream)ctx->stream);
    if (cr == CUDA_SUCCESS) cr = d.LaunchKernel(e->fn, gridx, gridy, 1, threads, 1, 1, 0, (CUstream)ctx->stream, args, nullptr);
    if (cr != CUDA_SUCCESS) {
        const char* es = nullptr; d.GetErrorString(cr, &es);
        return fail(ctx, FDB_ERR_CUDA, "%s: cuLaunchKernel: %s", what, es ? es : "?");
    }
    return FDB_OK;
}

void jit_release(fdb_ctx* ctx) {
    if (!ctx->jit) return;
    Driver& d = driver();
    for (auto& kv : ctx->jit->map) if (kv.second.mod && d.ok) d.ModuleUnload(kv.second.mod);
    delete ctx->jit;
    ctx->jit = nullptr;
}

}  // namespace fdb

using namespace fdb;

extern "C" int fdb_set_jit(fdb_ctx* ctx, int enabled) {
    if (!ctx) return FDB_ERR_BADARG;
    ctx->use_jit = enabled != 0;
    return FDB_OK;
}

extern "C" int fdb_jit_stats(fdb_ctx* ctx, int64_t* compiled, int64_t* cache_hits, int64_t* failures) {
    if (!ctx) return FDB_ERR_BADARG;
    if (compiled) *compiled = ctx->jit ? ctx->jit->compiled : 0;
    if (cache_hits) *cache_hits = ctx->jit ? ctx->jit->hits : 0;
    if (failures) *failures = ctx->jit ? ctx->jit->failures : 0;
    return FDB_OK;
}

extern "C" const char* fdb_jit_last_log(fdb_ctx* ctx) { return (ctx && ctx->jit) ? ctx->jit->last_log.c_str() : ""; }

extern "C" int fdb_jit_compile_check(int kind, const int32_t* term_prog, int n_term, const int32_t* epi_prog, int n_epi, const double* consts,
                                     int n_consts, int n_acc, int halo, int result_reg, int n_params, int N, int nansafe, int lane_sharing,
                                     char* log, int64_t log_capacity, int64_t* cubin_bytes) {
    if (kind < JIT_GRADIENT || kind > JIT_GRADIENT_STRUCT || N < 1 || N > FDB_MAX_CHUNK) return FDB_ERR_BADARG;
    Program P;
    int rc = build_program(nullptr, P, term_prog, n_term, epi_prog, n_epi, consts, n_consts, n_acc, halo, result_reg, kind == JIT_JACOBIAN || kind == JIT_MATVEC_JACOBIAN || kind == JIT_JACOBIAN2, n_params);
    std::vector<char> cubin;
    std::string lg;
    if (rc == FDB_OK) rc = compile_source(make_source(kind, P, N, nansafe != 0, lane_sharing != 0, false), cubin, lg);
    else lg = "malformed program";
    if (log && log_capacity > 0) { strncpy(log, lg.c_str(), (size_t)log_capacity - 1); log[log_capacity - 1] = 0; }
    if (cubin_bytes) *cubin_bytes = (int64_t)cubin.size();
    return rc;
}
