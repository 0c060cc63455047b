// oracle.cpp -- CPU restatement of ForwardDiff.jl's chunked dual-number sweep.
//
// TEST INFRASTRUCTURE ONLY.  This file is the parity checker and the timed CPU
// baseline.  Nothing under forwarddiff.jl_b200/ (the product) may include, link
// or call it; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs do.
//
// Pinning status: the chunk loops, seeding, extraction and the + - * / ^ sin cos
// fma arithmetic are pinned against the reference's own golden vectors
// (tests/test_oracle_golden.py: test/GradientTest.jl:25-54, test/JacobianTest.jl:23-37,
// test/HessianTest.jl:18-54, test/MiscTest.jl:39-49, test/SeedTest.jl:22-91,
// benchmarks/cpp/benchmarks.cpp:35-96) and against the reference's benchmarks/cpp
// program compiled from /root/reference into oracle/_ref (oracle/build_ref.sh).
// The DiffRules-derived formula choices (tanh, sqrt, exp, log, inv, abs ...) and
// the Brusselator / tanh.(A*x .+ b) targets do not appear in the reference tree
// with golden values: for those the oracle is "parity unpinned" beyond the
// finite-difference checks in tests/test_oracle_golden.py.
//
// The Julia reference cannot run in this image (no julia binary), so every
// function below cites the reference file:line it restates.  All paths are
// relative to /root/reference.
//
// Layout: AoS, exactly the reference's `struct Dual{T,V,N}; value::V;
// partials::Partials{N,V}` (src/dual.jl:14-16, src/partials.jl:1-3; wire order
// src/dual.jl:357-360): element e of an array of Dual{N} occupies doubles
// [e*(N+1) .. e*(N+1)+N] = value, p1..pN.  A nested Dual{T,Dual{T,V,N},N}
// occupies (N+1)^2 doubles: [v.v, v.p1..pN, P1.v, P1.p1..pN, ...].
//
// Build: g++ -O3 -march=native -ffp-contract=off (the reference never contracts
// a*b+c: src/partials.jl:222-224 are separate mul and add); -DORC_NANSAFE=1
// builds the `nansafe_mode` preference variant (src/prelude.jl:1,
// src/partials.jl:106-123).

#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include <algorithm>

#ifndef ORC_NANSAFE
#define ORC_NANSAFE 0
#endif
#define ORC_MAXN 16

namespace orc {

// ---------------------------------------------------------------------------
// Partials / Dual   (src/partials.jl, src/dual.jl)
// ---------------------------------------------------------------------------
template <class V, int N> struct Dual {
    V v;
    V p[N];
};

// value-type traits
template <class V> struct is_dual { static const bool value = false; };
template <class V, int N> struct is_dual<Dual<V, N>> { static const bool value = true; };

// zero / one of a value type (src/dual.jl:362-366)
inline double zero_of(const double&) { return 0.0; }
inline double one_of(const double&) { return 1.0; }
template <class V, int N> inline Dual<V, N> zero_of(const Dual<V, N>& d) {
    Dual<V, N> r; r.v = zero_of(d.v); for (int i = 0; i < N; i++) r.p[i] = zero_of(d.v); return r;
}
template <class V, int N> inline Dual<V, N> one_of(const Dual<V, N>& d) {
    Dual<V, N> r; r.v = one_of(d.v); for (int i = 0; i < N; i++) r.p[i] = zero_of(d.v); return r;
}
// iszero on a value type: Base.iszero(x) = x == zero(x); for Dual `==` compares value
// AND partials (src/dual.jl:397-405 defines == on value then partials)
inline bool iszero_v(const double& x) { return x == 0.0; }
template <class V, int N> inline bool iszero_v(const Dual<V, N>& d) {
    if (!iszero_v(d.v)) return false;
    for (int i = 0; i < N; i++) if (!iszero_v(d.p[i])) return false;
    return true;
}
inline double primal(const double& x) { return x; }
template <class V, int N> inline double primal(const Dual<V, N>& d) { return primal(d.v); }

// _mul_partial / _div_partial (src/partials.jl:106-123).  NaN-safe mode returns
// zero(y) whenever the partial is zero.
template <class P, class X> inline auto mul_partial(const P& partial, const X& x) -> decltype(partial * x) {
    auto y = partial * x;
#if ORC_NANSAFE
    if (iszero_v(partial)) return zero_of(y);
#endif
    return y;
}
template <class P, class X> inline auto div_partial(const P& partial, const X& x) -> decltype(partial / x) {
    auto y = partial / x;
#if ORC_NANSAFE
    if (iszero_v(partial)) return zero_of(y);
#endif
    return y;
}

// ---- + and - (src/dual.jl:499-519, src/partials.jl:82-84,210-220) ----------
template <class V, int N> inline Dual<V, N> operator+(const Dual<V, N>& x, const Dual<V, N>& y) {
    Dual<V, N> r; r.v = x.v + y.v; for (int i = 0; i < N; i++) r.p[i] = x.p[i] + y.p[i]; return r;
}
template <class V, int N> inline Dual<V, N> operator+(const Dual<V, N>& x, double s) {
    Dual<V, N> r = x; r.v = x.v + s; return r;
}
template <class V, int N> inline Dual<V, N> operator+(double s, const Dual<V, N>& y) {
    Dual<V, N> r = y; r.v = s + y.v; return r;
}
template <class V, int N> inline Dual<V, N> operator-(const Dual<V, N>& x, const Dual<V, N>& y) {
    Dual<V, N> r; r.v = x.v - y.v; for (int i = 0; i < N; i++) r.p[i] = x.p[i] - y.p[i]; return r;
}
template <class V, int N> inline Dual<V, N> operator-(const Dual<V, N>& x, double s) {
    Dual<V, N> r = x; r.v = x.v - s; return r;
}
template <class V, int N> inline Dual<V, N> operator-(const Dual<V, N>& x) {
    Dual<V, N> r; r.v = -x.v; for (int i = 0; i < N; i++) r.p[i] = -x.p[i]; return r;
}
template <class V, int N> inline Dual<V, N> operator-(double s, const Dual<V, N>& y) {
    Dual<V, N> r; r.v = s - y.v; for (int i = 0; i < N; i++) r.p[i] = -y.p[i]; return r;  // src/dual.jl:516
}

// ---- * (DiffRules rule, pinned by test/DualTest.jl:470-472) ---------------
// Dual*Dual: value vx*vy, partials _mul_partials(px, py, vy, vx) = px*vy + py*vx
// (src/dual.jl:216-218, src/partials.jl:95-97,222-224)
template <class V, int N> inline Dual<V, N> operator*(const Dual<V, N>& x, const Dual<V, N>& y) {
    Dual<V, N> r; r.v = x.v * y.v;
    for (int i = 0; i < N; i++) r.p[i] = mul_partial(x.p[i], y.v) + mul_partial(y.p[i], x.v);
    return r;
}
// Dual*Real and Real*Dual: partials * s (src/partials.jl:85-89,202-204)
template <class V, int N> inline Dual<V, N> operator*(const Dual<V, N>& x, double s) {
    Dual<V, N> r; r.v = x.v * s; for (int i = 0; i < N; i++) r.p[i] = mul_partial(x.p[i], s); return r;
}
template <class V, int N> inline Dual<V, N> operator*(double s, const Dual<V, N>& y) {
    // DiffRules y-body: val = s*vy, dvy = s; partials(y)*s.  (value s*vy == vy*s in IEEE)
    Dual<V, N> r; r.v = s * y.v; for (int i = 0; i < N; i++) r.p[i] = mul_partial(y.p[i], s); return r;
}

// ---- / (src/dual.jl:532-544, src/partials.jl:99-101) ----------------------
inline double inv_v(double x) { return 1.0 / x; }
template <class V, int N> inline Dual<V, N> operator/(const Dual<V, N>& x, const Dual<V, N>& y);
template <class V, int N> inline Dual<V, N> operator/(double s, const Dual<V, N>& y);
template <class V, int N> inline Dual<V, N> inv_v(const Dual<V, N>& x);  // DiffRules inv, below

template <class V, int N> inline Dual<V, N> operator/(const Dual<V, N>& x, const Dual<V, N>& y) {
    Dual<V, N> r; r.v = x.v / y.v;
    V fa = inv_v(y.v);                    // inv(bval)
    V fb = -(x.v / (y.v * y.v));          // -(aval / (bval*bval))
    for (int i = 0; i < N; i++) r.p[i] = mul_partial(x.p[i], fa) + mul_partial(y.p[i], fb);
    return r;
}
template <class V, int N> inline Dual<V, N> operator/(const Dual<V, N>& x, double s) {
    Dual<V, N> r; r.v = x.v / s; for (int i = 0; i < N; i++) r.p[i] = div_partial(x.p[i], s); return r;
}
template <class V, int N> inline Dual<V, N> operator/(double s, const Dual<V, N>& y) {
    Dual<V, N> r; V divv = s / y.v; r.v = divv;
    V c = -(divv / y.v);                  // -(divv / v) * partials(y)
    for (int i = 0; i < N; i++) r.p[i] = mul_partial(y.p[i], c);
    return r;
}

// ---- unary DiffRules ops: val = f(v); deriv = rule(v) after CSE; partials*deriv
// (src/dual.jl:244-258, 213-215).  Formulas restated from DiffRules 1.x, see
// SURVEY.md App. A -- formula choice is parity-unpinned in-tree.
template <class V, int N, class DV> inline Dual<V, N> chain(const V& val, const DV& deriv, const Dual<V, N>& d) {
    Dual<V, N> r; r.v = val; for (int i = 0; i < N; i++) r.p[i] = mul_partial(d.p[i], deriv); return r;
}
// literal_pow (src/dual.jl:587-597); Float64 v^2 = v*v, v^3 = v*v*v (Base.literal_pow)
inline double pow2(double v) { return v * v; }
inline double pow3(double v) { return v * v * v; }
inline double pow1(double v) { return v; }
template <class V, int N> inline Dual<V, N> pow1(const Dual<V, N>& x) {
    // expv = v^1; deriv = 1 * v^0 = 1*one(v)
    return chain(pow1(x.v), 1.0 * one_of(x.v), x);
}
template <class V, int N> inline Dual<V, N> pow2(const Dual<V, N>& x) {
    return chain(pow2(x.v), 2.0 * pow1(x.v), x);   // deriv = 2 * v^1
}
template <class V, int N> inline Dual<V, N> pow3(const Dual<V, N>& x) {
    return chain(pow3(x.v), 3.0 * pow2(x.v), x);   // deriv = 3 * v^2
}
template <class V, int N> inline Dual<V, N> pow0(const Dual<V, N>& x) { return one_of(x); }

inline double exp_v(double x) { return std::exp(x); }
inline double log_v(double x) { return std::log(x); }
inline double sqrt_v(double x) { return std::sqrt(x); }
inline double tanh_v(double x) { return std::tanh(x); }
inline double sin_v(double x) { return std::sin(x); }
inline double cos_v(double x) { return std::cos(x); }
inline double abs_v(double x) { return std::fabs(x); }
inline double abs2_v(double x) { return x * x; }
inline bool signbit_v(double x) { return std::signbit(x); }
template <class V, int N> inline bool signbit_v(const Dual<V, N>& x) { return signbit_v(x.v); }

template <class V, int N> inline Dual<V, N> exp_v(const Dual<V, N>& x) { V val = exp_v(x.v); return chain(val, val, x); }
template <class V, int N> inline Dual<V, N> log_v(const Dual<V, N>& x) { return chain(log_v(x.v), inv_v(x.v), x); }
template <class V, int N> inline Dual<V, N> sqrt_v(const Dual<V, N>& x) { V val = sqrt_v(x.v); return chain(val, inv_v(2.0 * val), x); }
template <class V, int N> inline Dual<V, N> tanh_v(const Dual<V, N>& x) { V val = tanh_v(x.v); return chain(val, 1.0 - pow2(val), x); }
template <class V, int N> inline Dual<V, N> inv_v(const Dual<V, N>& x) { V val = inv_v(x.v); return chain(val, -abs2_v(val), x); }
template <class V, int N> inline Dual<V, N> abs2_v(const Dual<V, N>& x) { return chain(abs2_v(x.v), x.v + x.v, x); }
template <class V, int N> inline Dual<V, N> abs_v(const Dual<V, N>& x) {
    // DiffRules: ifelse(signbit(x), -one(x), one(x)); pinned by test/MiscTest.jl:183 (abs'(0)=+1)
    V o = one_of(x.v);
    return signbit_v(x.v) ? chain(abs_v(x.v), -o, x) : chain(abs_v(x.v), o, x);
}
// further DiffRules unary rules (V = double only)
template <int N> inline Dual<double, N> tan_v(const Dual<double, N>& x) { double val = std::tan(x.v); return chain(val, 1.0 + val * val, x); }
template <int N> inline Dual<double, N> exp2_v(const Dual<double, N>& x) { double val = std::exp2(x.v); return chain(val, val * 0.6931471805599453, x); }
template <int N> inline Dual<double, N> log2_v(const Dual<double, N>& x) { return chain(std::log2(x.v), (1.0 / x.v) / 0.6931471805599453, x); }
template <int N> inline Dual<double, N> log10_v(const Dual<double, N>& x) { return chain(std::log10(x.v), (1.0 / x.v) / 2.302585092994046, x); }
template <int N> inline Dual<double, N> log1p_v(const Dual<double, N>& x) { return chain(std::log1p(x.v), 1.0 / (x.v + 1.0), x); }
template <int N> inline Dual<double, N> expm1_v(const Dual<double, N>& x) { return chain(std::expm1(x.v), std::exp(x.v), x); }
template <int N> inline Dual<double, N> sinh_v(const Dual<double, N>& x) { return chain(std::sinh(x.v), std::cosh(x.v), x); }
template <int N> inline Dual<double, N> cosh_v(const Dual<double, N>& x) { return chain(std::cosh(x.v), std::sinh(x.v), x); }
template <int N> inline Dual<double, N> asin_v(const Dual<double, N>& x) { return chain(std::asin(x.v), 1.0 / std::sqrt(1.0 - x.v * x.v), x); }
template <int N> inline Dual<double, N> acos_v(const Dual<double, N>& x) { return chain(std::acos(x.v), -(1.0 / std::sqrt(1.0 - x.v * x.v)), x); }
template <int N> inline Dual<double, N> atan_v(const Dual<double, N>& x) { return chain(std::atan(x.v), 1.0 / (1.0 + x.v * x.v), x); }
template <int N> inline Dual<double, N> cbrt_v(const Dual<double, N>& x) { double val = std::cbrt(x.v); return chain(val, 1.0 / (3.0 * (val * val)), x); }
// binary atan(y, x) and hypot(x, y) (DiffRules)
template <int N> inline Dual<double, N> atan2_v(const Dual<double, N>& y, const Dual<double, N>& x) {
    Dual<double, N> r; r.v = std::atan2(y.v, x.v);
    double den = y.v * y.v + x.v * x.v; double dy = x.v / den, dx = -y.v / den;
    for (int i = 0; i < N; i++) r.p[i] = mul_partial(y.p[i], dy) + mul_partial(x.p[i], dx);
    return r;
}
template <int N> inline Dual<double, N> hypot_v(const Dual<double, N>& x, const Dual<double, N>& y) {
    Dual<double, N> r; double h = std::hypot(x.v, y.v); r.v = h;
    double dx = x.v / h, dy = y.v / h;
    for (int i = 0; i < N; i++) r.p[i] = mul_partial(x.p[i], dx) + mul_partial(y.p[i], dy);
    return r;
}
// sin / cos share one sincos (src/dual.jl:709-717): cos partial is (-s)*p
template <class V, int N> inline Dual<V, N> sin_v(const Dual<V, N>& x) { V s = sin_v(x.v), c = cos_v(x.v); return chain(s, c, x); }
template <class V, int N> inline Dual<V, N> cos_v(const Dual<V, N>& x) { V s = sin_v(x.v), c = cos_v(x.v); return chain(c, -s, x); }

// ---- ^ with non-literal exponent (src/dual.jl:549-585) ---------------------
inline double pow_v(double x, double y) { return std::pow(x, y); }
template <class V, int N> inline bool partials_iszero(const Dual<V, N>& x) {
    for (int i = 0; i < N; i++) if (!iszero_v(x.p[i])) return false;
    return true;
}
// Dual^Real (x-body, src/dual.jl:567-576); only V=double needed by the op set
template <int N> inline Dual<double, N> pow_v(const Dual<double, N>& x, double y) {
    Dual<double, N> r; r.v = std::pow(x.v, y);
    if (y == 0.0 || partials_iszero(x)) { for (int i = 0; i < N; i++) r.p[i] = 0.0; }
    else { double w = std::pow(x.v, y - 1.0); for (int i = 0; i < N; i++) r.p[i] = mul_partial(mul_partial(x.p[i], y), w); }
    return r;
}
// Dual^Dual (xy-body, src/dual.jl:553-566)
template <int N> inline Dual<double, N> pow_v(const Dual<double, N>& x, const Dual<double, N>& y) {
    Dual<double, N> r; double expv = std::pow(x.v, y.v); r.v = expv;
    double powval = y.v * std::pow(x.v, y.v - 1.0);
    double logval;
    if (partials_iszero(y)) logval = 1.0;                // isconstant(y)
    else if (x.v == 0.0 && y.v > 0) logval = 0.0;
    else logval = expv * std::log(x.v);
    for (int i = 0; i < N; i++) r.p[i] = mul_partial(x.p[i], powval) + mul_partial(y.p[i], logval);
    return r;
}
// Real^Dual (y-body, src/dual.jl:577-582)
template <int N> inline Dual<double, N> pow_v(double x, const Dual<double, N>& y) {
    double expv = std::pow(x, y.v);
    double deriv = (x == 0.0 && y.v > 0) ? 0.0 : expv * std::log(x);
    return chain(expv, deriv, y);
}

// ---- fma (src/dual.jl:625-665) ---------------------------------------------
template <int N> inline Dual<double, N> fma_xyz(const Dual<double, N>& x, const Dual<double, N>& y, const Dual<double, N>& z) {
    Dual<double, N> r; r.v = std::fma(x.v, y.v, z.v);
    for (int i = 0; i < N; i++) r.p[i] = std::fma(x.v, y.p[i], std::fma(y.v, x.p[i], z.p[i]));
    return r;
}
template <int N> inline Dual<double, N> fma_xy(const Dual<double, N>& x, const Dual<double, N>& y, double z) {
    Dual<double, N> r; r.v = std::fma(x.v, y.v, z);
    for (int i = 0; i < N; i++) r.p[i] = mul_partial(x.p[i], y.v) + mul_partial(y.p[i], x.v);
    return r;
}
template <int N> inline Dual<double, N> fma_xz(const Dual<double, N>& x, double y, const Dual<double, N>& z) {
    Dual<double, N> r; r.v = std::fma(x.v, y, z.v);
    for (int i = 0; i < N; i++) r.p[i] = std::fma(x.p[i], y, z.p[i]);
    return r;
}

// binary max/min (DiffRules: derivative 1 for the selected argument; x > y picks x)
template <int N> inline Dual<double, N> max_v(const Dual<double, N>& x, const Dual<double, N>& y) {
    Dual<double, N> r; r.v = std::fmax(x.v, y.v);   // parity-unpinned for NaN
    double dx = x.v > y.v ? 1.0 : 0.0, dy = x.v > y.v ? 0.0 : 1.0;
    for (int i = 0; i < N; i++) r.p[i] = mul_partial(x.p[i], dx) + mul_partial(y.p[i], dy);
    return r;
}
template <int N> inline Dual<double, N> min_v(const Dual<double, N>& x, const Dual<double, N>& y) {
    Dual<double, N> r; r.v = std::fmin(x.v, y.v);
    double dx = x.v > y.v ? 0.0 : 1.0, dy = x.v > y.v ? 1.0 : 0.0;
    for (int i = 0; i < N; i++) r.p[i] = mul_partial(x.p[i], dx) + mul_partial(y.p[i], dy);
    return r;
}

// NaNMath.max / NaNMath.min (NaNMath.jl: a NaN operand is ignored); derivative 1 for the argument returned.  The DiffRules rule for
// these (src/dual.jl:476-490 generates the Dual methods from it) is not pinned by any test in the reference tree: parity unpinned.
template <int N> inline Dual<double, N> nanmax_v(const Dual<double, N>& x, const Dual<double, N>& y) {
    const bool takex = std::isnan(y.v) || (!std::isnan(x.v) && x.v > y.v);
    Dual<double, N> r; r.v = takex ? x.v : y.v;
    const double dx = takex ? 1.0 : 0.0, dy = takex ? 0.0 : 1.0;
    for (int i = 0; i < N; i++) r.p[i] = mul_partial(x.p[i], dx) + mul_partial(y.p[i], dy);
    return r;
}
template <int N> inline Dual<double, N> nanmin_v(const Dual<double, N>& x, const Dual<double, N>& y) {
    const bool takex = std::isnan(y.v) || (!std::isnan(x.v) && !(x.v > y.v));
    Dual<double, N> r; r.v = takex ? x.v : y.v;
    const double dx = takex ? 1.0 : 0.0, dy = takex ? 0.0 : 1.0;
    for (int i = 0; i < N; i++) r.p[i] = mul_partial(x.p[i], dx) + mul_partial(y.p[i], dy);
    return r;
}

// ---------------------------------------------------------------------------
// seeds and windows  (src/apiutils.jl:39-41, 76-143; src/partials.jl:9-12)
// ---------------------------------------------------------------------------
// `sidx` maps structural position (0-based) -> storage index (0-based);
// nullptr means the identity (dense arrays, apiutils.jl:45-48).
inline long sidx_at(const long* sidx, long k) { return sidx ? sidx[k] : k; }

// seed!(duals, x, index, seeds, chunksize)  apiutils.jl:125-143 ; index is 1-based.
// zip(1:chunksize, drop(idxs, index-1)) stops at whichever ends first.
template <class V, int N>
void seed(Dual<V, N>* duals, const V* x, long nstruct, const long* sidx, long index, long chunksize) {
    for (long i = 1; i <= chunksize; i++) {
        long pos = index - 1 + (i - 1);
        if (pos >= nstruct) break;
        long idx = sidx_at(sidx, pos);
        Dual<V, N> d; d.v = x[idx];
        for (int j = 1; j <= N; j++) d.p[j - 1] = (j == i) ? one_of(x[idx]) : zero_of(x[idx]);  // single_seed, partials.jl:9-12
        duals[idx] = d;
    }
}
// seed_zero_partials!(duals, x, index, count)  apiutils.jl:83-87 -> :89-105
template <class V, int N>
void seed_zero(Dual<V, N>* duals, const V* x, long nstruct, const long* sidx, long index, long count) {
    for (long i = 0; i < count; i++) {                 // Iterators.take(drop(idxs, index-1), count)
        long pos = index - 1 + i;
        if (pos >= nstruct) break;                     // clamped, test/SeedTest.jl:66-72
        long idx = sidx_at(sidx, pos);
        Dual<V, N> d; d.v = x[idx];
        for (int j = 0; j < N; j++) d.p[j] = zero_of(x[idx]);
        duals[idx] = d;
    }
}

// chunk bookkeeping (src/gradient.jl:121-125 == src/jacobian.jl:177-181)
struct Plan { long xlen, remainder, lastchunksize, lastchunkindex, middle_hi, nchunks; };
inline Plan make_plan(long xlen, long N) {
    Plan p; p.xlen = xlen;
    p.remainder = xlen % N;
    p.lastchunksize = p.remainder == 0 ? N : p.remainder;
    p.lastchunkindex = xlen - p.lastchunksize + 1;
    p.middle_hi = (xlen - p.lastchunksize) / N;        // middlechunks = 2:middle_hi
    p.nchunks = p.middle_hi + 1;                       // first + middle + last (first==last never: N<xlen here)
    return p;
}

// ---------------------------------------------------------------------------
// target functions on dual arrays
// ---------------------------------------------------------------------------
// Rosenbrock exactly as docs/src/user/advanced.md:36-44 (= DiffTests.rosenbrock_1):
// a = one(eltype(x)); b = 100*a are Duals with zero partials.
template <class D> D rosenbrock(const D* x, long n) {
    D a = one_of(x[0]);
    D b = 100.0 * a;
    D result = zero_of(x[0]);
    for (long i = 0; i + 1 < n; i++) {
        D t = pow2(a - x[i]) + b * pow2(x[i + 1] - pow2(x[i]));
        result = result + t;
    }
    return result;
}
// Ackley: DiffTests.ackley; C++ twin at benchmarks/cpp/benchmarks.cpp:10-21.
template <class D> D ackley(const D* x, long n) {
    const double a = 20.0, b = -0.2, c = 2.0 * M_PI;
    double len_recip = 1.0 / (double)n;
    D sum_sqrs = zero_of(x[0]);
    D sum_cos = sum_sqrs;
    for (long i = 0; i < n; i++) {
        sum_cos = sum_cos + cos_v(c * x[i]);
        sum_sqrs = sum_sqrs + pow2(x[i]);
    }
    return (-a) * exp_v(b * sqrt_v(len_recip * sum_sqrs)) - exp_v(len_recip * sum_cos) + a + M_E;
}
// sum(abs2, x)-style quadratic: sequential sum of x[i]^2 (for generic-kernel parity)
template <class D> D sumsq(const D* x, long n) {
    D r = zero_of(x[0]);
    for (long i = 0; i < n; i++) r = r + pow2(x[i]);
    return r;
}
// prod(x): sequential product
template <class D> D prodall(const D* x, long n) {
    D r = one_of(x[0]);
    for (long i = 0; i < n; i++) r = r * x[i];
    return r;
}
template <class D> D eval_scalar_fn(int fn, const D* x, long n) {
    switch (fn) {
        case 1: return rosenbrock(x, n);
        case 2: return ackley(x, n);
        case 3: return sumsq(x, n);
        case 4: return prodall(x, n);
    }
    return zero_of(x[0]);
}

// Array-valued targets.  `par` carries the plain-Float64 parameters.
struct JacParams { const double* A; long lda; const double* b; double alpha;
                   const double* A2 = nullptr; long lda2 = 0; const double* b2 = nullptr; long h = 0; };

// f(x) = tanh.(A*x .+ b);  A*x is Julia's generic column sweep
// (LinearAlgebra.generic_matvecmul!: C[i] = zero; for k: C[i] += A[i,k]*x[k]).
template <class D> void tanh_affine(D* y, long m, const D* x, long n, const JacParams& par) {
    for (long i = 0; i < m; i++) y[i] = zero_of(x[0]);
    for (long k = 0; k < n; k++) {
        const D& b = x[k];
        for (long i = 0; i < m; i++) y[i] = y[i] + par.A[i + k * par.lda] * b;
    }
    for (long i = 0; i < m; i++) y[i] = tanh_v(y[i] + par.b[i]);
}
// 1-D Brusselator RHS, in-place form f!(dx, x); defined in SURVEY.md section 8(d)
// (not in the reference).  x = [u; v], M = n/2 grid points, Dirichlet u=1, v=3.
template <class D> void brusselator(D* dx, const D* x, long n, const JacParams& par) {
    long M = n / 2; const double al = par.alpha;
    const D* u = x; const D* v = x + M;
    for (long i = 0; i < M; i++) {
        D lapu, lapv;
        if (M == 1) { lapu = (1.0 - 2.0 * u[i]) + 1.0; lapv = (3.0 - 2.0 * v[i]) + 3.0; }
        else if (i == 0) { lapu = (1.0 - 2.0 * u[i]) + u[i + 1]; lapv = (3.0 - 2.0 * v[i]) + v[i + 1]; }
        else if (i == M - 1) { lapu = (u[i - 1] - 2.0 * u[i]) + 1.0; lapv = (v[i - 1] - 2.0 * v[i]) + 3.0; }
        else { lapu = (u[i - 1] - 2.0 * u[i]) + u[i + 1]; lapv = (v[i - 1] - 2.0 * v[i]) + v[i + 1]; }
        dx[i] = ((1.0 + pow2(u[i]) * v[i]) - 4.0 * u[i]) + al * lapu;
        dx[M + i] = (3.0 * u[i] - pow2(u[i]) * v[i]) + al * lapv;
    }
}
// two-layer target f(x) = tanh.(A2 * tanh.(A1*x .+ b1) .+ b2): the second A*xdual sees dense partials
// (SURVEY.md section 8f-4).  A1 is h x n (par.A, par.b), A2 is m x h (par.A2, par.b2).
template <class D> void mlp2(D* y, long m, const D* x, long n, const JacParams& par) {
    std::vector<D> hid(par.h);
    JacParams p1 = {par.A, par.lda, par.b, 0.0};
    tanh_affine(hid.data(), par.h, x, n, p1);
    JacParams p2 = {par.A2, par.lda2, par.b2, 0.0};
    tanh_affine(y, m, hid.data(), par.h, p2);
}
// test/JacobianTest.jl:23-31 hard-coded target (n=3, m=4)
template <class D> void jactest(D* y, const D* x) {
    y[0] = x[0] * x[1];
    y[0] = y[0] * sin_v(pow2(x[2]));
    y[1] = y[0] + x[2];
    y[2] = y[0] / y[1];
    y[3] = x[2];
}
template <class D> void eval_array_fn(int fn, D* y, long m, const D* x, long n, const JacParams& par) {
    switch (fn) {
        case 1: tanh_affine(y, m, x, n, par); break;
        case 2: brusselator(y, x, n, par); break;
        case 3: jactest(y, x); break;
        case 4: mlp2(y, m, x, n, par); break;
    }
}

// ---------------------------------------------------------------------------
// chunk-mode drivers
// ---------------------------------------------------------------------------
// chunk_mode_gradient! (src/gradient.jl:114-159).  f is any callable D(const D*, long).
// Chunks [clo, chi) (0-based) are evaluated; the full call is clo=0, chi=nchunks,
// which follows the reference statement by statement.  A partial range (used only
// to time a bounded sample) starts from a fully zero-seeded buffer, which is the
// state the reference is in at the top of every middle-chunk iteration.
template <class V, int N, class F>
void chunk_mode_gradient(F&& f, const V* x, long n, V* result, V* value_out, long clo, long chi,
                         Dual<V, N>* xdual) {
    typedef Dual<V, N> D;
    Plan pl = make_plan(n, N);
    D ydual;
    bool full = (clo == 0 && chi == pl.nchunks);
    if (!full) seed_zero(xdual, x, n, (const long*)nullptr, 1, n);
    for (long c = clo; c < chi; c++) {
        long c1 = c + 1;                                   // 1-based chunk number
        if (c1 == pl.nchunks) {                            // final chunk, gradient.jl:150-152
            seed(xdual, x, n, (const long*)nullptr, pl.lastchunkindex, (long)pl.lastchunksize);
            ydual = f(xdual, n);
            for (long k = 1; k <= pl.lastchunksize; k++) result[pl.lastchunkindex - 1 + k - 1] = ydual.p[k - 1];
        } else if (c1 == 1 && full) {                      // first chunk, gradient.jl:133-138
            seed(xdual, x, n, (const long*)nullptr, 1, (long)N);
            seed_zero(xdual, x, n, (const long*)nullptr, N + 1, n - N);
            ydual = f(xdual, n);
            for (long k = 1; k <= N; k++) result[k - 1] = ydual.p[k - 1];     // extract_gradient_chunk!, gradient.jl:74-81
            seed_zero(xdual, x, n, (const long*)nullptr, 1, (long)N);
        } else {                                           // middle chunks, gradient.jl:141-147
            long i = (c1 - 1) * N + 1;
            seed(xdual, x, n, (const long*)nullptr, i, (long)N);
            ydual = f(xdual, n);
            for (long k = 1; k <= N; k++) result[i - 1 + k - 1] = ydual.p[k - 1];
            seed_zero(xdual, x, n, (const long*)nullptr, i, (long)N);
        }
    }
    if (value_out) *value_out = ydual.v;                   // extract_value!, gradient.jl:155 (last sweep)
}

// vector_mode_gradient (src/gradient.jl:97-108, apiutils.jl:21-25)
template <class V, int N, class F>
void vector_mode_gradient(F&& f, const V* x, long n, V* result, V* value_out, Dual<V, N>* xdual) {
    seed(xdual, x, n, (const long*)nullptr, 1, (long)N);
    Dual<V, N> ydual = f(xdual, n);
    for (int k = 0; k < N; k++) result[k] = ydual.p[k];    // extract_gradient!, gradient.jl:66-72
    if (value_out) *value_out = ydual.v;
}

template <class V, int N, class F>
void gradient_any(F&& f, const V* x, long n, V* result, V* value_out, long clo, long chi, Dual<V, N>* xdual) {
    if (n == N) vector_mode_gradient<V, N>(f, x, n, result, value_out, xdual);   // gradient.jl:19-23
    else chunk_mode_gradient<V, N>(f, x, n, result, value_out, clo, chi, xdual);
}

// chunk_mode_jacobian (src/jacobian.jl:169-216), f-form and f!-form.  J is
// column-major m x n with leading dimension ldJ (jacobian.jl:221); if J is null
// only the sweeps run (bounded CPU timing of configs whose J would not fit).
template <int N>
void chunk_mode_jacobian(int fn, bool inplace, const double* x, long n, long m, const JacParams& par,
                         double* J, long ldJ, double* y, long clo, long chi) {
    typedef Dual<double, N> D;
    Plan pl = make_plan(n, N);
    std::vector<D> xdual(n), ydual(m);
    std::vector<double> y0(m, 0.0);
    if (y && inplace) for (long i = 0; i < m; i++) y0[i] = y[i];
    bool full = (clo == 0 && chi == pl.nchunks);
    if (!full) seed_zero(xdual.data(), x, n, (const long*)nullptr, 1, n);
    auto compute = [&]() {
        if (inplace) seed_zero(ydual.data(), y0.data(), m, (const long*)nullptr, 1, m);   // jacobian.jl:227
        eval_array_fn(fn, ydual.data(), m, xdual.data(), n, par);
    };
    auto extract = [&](long index, long chunksize) {       // extract_jacobian_chunk!, jacobian.jl:109-118
        if (!J) return;
        for (long k = 1; k <= chunksize; k++)
            for (long r = 0; r < m; r++) J[r + (index - 1 + k - 1) * ldJ] = ydual[r].p[k - 1];
    };
    for (long c = clo; c < chi; c++) {
        long c1 = c + 1;
        if (c1 == pl.nchunks) {
            seed(xdual.data(), x, n, (const long*)nullptr, pl.lastchunkindex, pl.lastchunksize);
            compute(); extract(pl.lastchunkindex, pl.lastchunksize);
        } else if (c1 == 1 && full) {
            seed(xdual.data(), x, n, (const long*)nullptr, 1, (long)N);
            seed_zero(xdual.data(), x, n, (const long*)nullptr, N + 1, n - N);
            compute(); extract(1, N);
            seed_zero(xdual.data(), x, n, (const long*)nullptr, 1, (long)N);
        } else {
            long i = (c1 - 1) * N + 1;
            seed(xdual.data(), x, n, (const long*)nullptr, i, (long)N);
            compute(); extract(i, N);
            seed_zero(xdual.data(), x, n, (const long*)nullptr, i, (long)N);
        }
    }
    if (y) for (long i = 0; i < m; i++) y[i] = ydual[i].v;   // map!(d->value(T,d), y, ydual), jacobian.jl:229
}
template <int N>
void vector_mode_jacobian(int fn, bool inplace, const double* x, long n, long m, const JacParams& par,
                          double* J, long ldJ, double* y) {
    typedef Dual<double, N> D;
    std::vector<D> xdual(n), ydual(m);
    seed(xdual.data(), x, n, (const long*)nullptr, 1, (long)N);
    if (inplace) { std::vector<double> y0(m, 0.0); if (y) for (long i = 0; i < m; i++) y0[i] = y[i];
                   seed_zero(ydual.data(), y0.data(), m, (const long*)nullptr, 1, m); }
    eval_array_fn(fn, ydual.data(), m, xdual.data(), n, par);
    if (J) for (long k = 0; k < N; k++) for (long r = 0; r < m; r++) J[r + k * ldJ] = ydual[r].p[k];   // jacobian.jl:95-102
    if (y) for (long i = 0; i < m; i++) y[i] = ydual[i].v;
}

// hessian = jacobian(y -> gradient(f, y, gcfg), x, jcfg)  (src/hessian.jl:14-19);
// DiffResult form also yields value and gradient from the last inner sweep
// (hessian.jl:50-55,66-71).  Outer chunks [olo, ohi).
template <int N>
void chunk_mode_hessian(int fn, const double* x, long n, double* H, long ldH, double* grad, double* value,
                        long olo, long ohi) {
    typedef Dual<double, N> DI;          // element type of the outer xdual
    typedef Dual<DI, N> DO;              // element type of the inner (gradient) xdual
    std::vector<DI> xdual(n), gdual(n);
    std::vector<DO> inner(n);
    auto f = [&](const DO* z, long nn) { return eval_scalar_fn<DO>(fn, z, nn); };
    DI inner_value;
    auto grad_of = [&]() {               // ydual = gradient(f, xdual, gcfg): a full inner chunk loop
        Plan ip = make_plan(n, N);
        gradient_any<DI, N>(f, xdual.data(), n, gdual.data(), &inner_value, 0, ip.nchunks, inner.data());
    };
    auto extract = [&](long index, long chunksize) {
        if (!H) return;
        for (long k = 1; k <= chunksize; k++)
            for (long r = 0; r < n; r++) H[r + (index - 1 + k - 1) * ldH] = gdual[r].p[k - 1];
    };
    if (n == N) {                        // vector mode at both levels
        seed(xdual.data(), x, n, (const long*)nullptr, 1, (long)N);
        grad_of(); extract(1, N);
    } else {
        Plan pl = make_plan(n, N);
        bool full = (olo == 0 && ohi == pl.nchunks);
        if (!full) seed_zero(xdual.data(), x, n, (const long*)nullptr, 1, n);
        for (long c = olo; c < ohi; c++) {
            long c1 = c + 1;
            if (c1 == pl.nchunks) {
                seed(xdual.data(), x, n, (const long*)nullptr, pl.lastchunkindex, pl.lastchunksize);
                grad_of(); extract(pl.lastchunkindex, pl.lastchunksize);
            } else if (c1 == 1 && full) {
                seed(xdual.data(), x, n, (const long*)nullptr, 1, (long)N);
                seed_zero(xdual.data(), x, n, (const long*)nullptr, N + 1, n - N);
                grad_of(); extract(1, N);
                seed_zero(xdual.data(), x, n, (const long*)nullptr, 1, (long)N);
            } else {
                long i = (c1 - 1) * N + 1;
                seed(xdual.data(), x, n, (const long*)nullptr, i, (long)N);
                grad_of(); extract(i, N);
                seed_zero(xdual.data(), x, n, (const long*)nullptr, i, (long)N);
            }
        }
    }
    if (grad) for (long i = 0; i < n; i++) grad[i] = gdual[i].v;      // y = value.(ydual)
    if (value) *value = inner_value.v;                                 // hessian.jl:53
}

// third-order tensor test (test/MiscTest.jl:39-49): jacobian(y -> hessian(f, y), x)
// = three nested levels, vector mode (n == N) at every level.
template <int N>
void third_order_vector_mode(int fn, const double* x, double* T3) {
    typedef Dual<double, N> D1; typedef Dual<D1, N> D2; typedef Dual<D2, N> D3;
    D1 x1[N]; seed(x1, x, (long)N, (const long*)nullptr, 1, (long)N);
    D2 x2[N]; seed(x2, x1, (long)N, (const long*)nullptr, 1, (long)N);
    D3 x3[N]; seed(x3, x2, (long)N, (const long*)nullptr, 1, (long)N);
    D3 y = eval_scalar_fn<D3>(fn, x3, N);
    // y.p[i] : D2 = d/dx_i at innermost level; .p[j] : D1 ; .p[k] : double
    // hessian(f,y)[i,j] -> jacobian row (i + j*N), column k
    for (int k = 0; k < N; k++) for (int j = 0; j < N; j++) for (int i = 0; i < N; i++)
        T3[(i + j * N) + (long)k * N * N] = y.p[i].p[j].p[k];
}

// ---------------------------------------------------------------------------
// runtime-N dispatch
// ---------------------------------------------------------------------------
#define ORC_DISPATCH_N(N_, CALL)                                                           \
    switch (N_) {                                                                          \
        case 1: { const int NN = 1; CALL; } break;  case 2: { const int NN = 2; CALL; } break;   \
        case 3: { const int NN = 3; CALL; } break;  case 4: { const int NN = 4; CALL; } break;   \
        case 5: { const int NN = 5; CALL; } break;  case 6: { const int NN = 6; CALL; } break;   \
        case 7: { const int NN = 7; CALL; } break;  case 8: { const int NN = 8; CALL; } break;   \
        case 9: { const int NN = 9; CALL; } break;  case 10: { const int NN = 10; CALL; } break; \
        case 11: { const int NN = 11; CALL; } break; case 12: { const int NN = 12; CALL; } break;\
        default: return -1;                                                                \
    }
#define ORC_DISPATCH_N8(N_, CALL)                                                          \
    switch (N_) {                                                                          \
        case 1: { const int NN = 1; CALL; } break;  case 2: { const int NN = 2; CALL; } break;   \
        case 3: { const int NN = 3; CALL; } break;  case 4: { const int NN = 4; CALL; } break;   \
        case 5: { const int NN = 5; CALL; } break;  case 6: { const int NN = 6; CALL; } break;   \
        case 7: { const int NN = 7; CALL; } break;  case 8: { const int NN = 8; CALL; } break;   \
        case 9: { const int NN = 9; CALL; } break;  case 10: { const int NN = 10; CALL; } break; \
        case 11: { const int NN = 11; CALL; } break; case 12: { const int NN = 12; CALL; } break;\
        default: return -1;                                                                \
    }

// elementwise maps over AoS arrays (the array-level ops of a user `f`)
enum Op1 { OP_NEG = 1, OP_EXP, OP_LOG, OP_SIN, OP_COS, OP_TANH, OP_SQRT, OP_ABS, OP_ABS2, OP_INV,
           OP_POW0, OP_POW1, OP_POW2, OP_POW3, OP_TAN, OP_EXP2, OP_LOG2, OP_LOG10, OP_LOG1P, OP_EXPM1, OP_SINH, OP_COSH,
           OP_ASIN, OP_ACOS, OP_ATAN, OP_CBRT };
enum Op2 { OP_ADD = 1, OP_SUB, OP_MUL, OP_DIV, OP_POW, OP_MAX, OP_MIN, OP_ATAN2, OP_HYPOT, OP_NANMAX, OP_NANMIN };

template <int N> int map1(int op, Dual<double, N>* out, const Dual<double, N>* a, long n) {
    for (long i = 0; i < n; i++) {
        const Dual<double, N>& x = a[i];
        switch (op) {
            case OP_NEG: out[i] = -x; break;          case OP_EXP: out[i] = exp_v(x); break;
            case OP_LOG: out[i] = log_v(x); break;    case OP_SIN: out[i] = sin_v(x); break;
            case OP_COS: out[i] = cos_v(x); break;    case OP_TANH: out[i] = tanh_v(x); break;
            case OP_SQRT: out[i] = sqrt_v(x); break;  case OP_ABS: out[i] = abs_v(x); break;
            case OP_ABS2: out[i] = abs2_v(x); break;  case OP_INV: out[i] = inv_v(x); break;
            case OP_POW0: out[i] = pow0(x); break;    case OP_POW1: out[i] = pow1(x); break;
            case OP_POW2: out[i] = pow2(x); break;    case OP_POW3: out[i] = pow3(x); break;
            case OP_TAN: out[i] = tan_v(x); break;    case OP_EXP2: out[i] = exp2_v(x); break;
            case OP_LOG2: out[i] = log2_v(x); break;  case OP_LOG10: out[i] = log10_v(x); break;
            case OP_LOG1P: out[i] = log1p_v(x); break; case OP_EXPM1: out[i] = expm1_v(x); break;
            case OP_SINH: out[i] = sinh_v(x); break;  case OP_COSH: out[i] = cosh_v(x); break;
            case OP_ASIN: out[i] = asin_v(x); break;  case OP_ACOS: out[i] = acos_v(x); break;
            case OP_ATAN: out[i] = atan_v(x); break;  case OP_CBRT: out[i] = cbrt_v(x); break;
            default: return -2;
        }
    }
    return 0;
}
// kind: 0 = Dual op Dual, 1 = Dual op Real(s), 2 = Real(s) op Dual
template <int N> int map2(int op, int kind, Dual<double, N>* out, const Dual<double, N>* a,
                          const Dual<double, N>* b, double s, long n) {
    typedef Dual<double, N> D;
    for (long i = 0; i < n; i++) {
        D r;
        if (kind == 0) {
            const D &x = a[i], &y = b[i];
            switch (op) { case OP_ADD: r = x + y; break; case OP_SUB: r = x - y; break; case OP_MUL: r = x * y; break;
                          case OP_DIV: r = x / y; break; case OP_POW: r = pow_v(x, y); break;
                          case OP_MAX: r = max_v(x, y); break; case OP_MIN: r = min_v(x, y); break;
                          case OP_ATAN2: r = atan2_v(x, y); break; case OP_HYPOT: r = hypot_v(x, y); break;
                          case OP_NANMAX: r = nanmax_v(x, y); break; case OP_NANMIN: r = nanmin_v(x, y); break; default: return -2; }
        } else if (kind == 1) {
            const D& x = a[i];
            switch (op) { case OP_ADD: r = x + s; break; case OP_SUB: r = x - s; break; case OP_MUL: r = x * s; break;
                          case OP_DIV: r = x / s; break; case OP_POW: r = pow_v(x, s); break; default: return -2; }
        } else {
            const D& y = a[i];
            switch (op) { case OP_ADD: r = s + y; break; case OP_SUB: r = s - y; break; case OP_MUL: r = s * y; break;
                          case OP_DIV: r = s / y; break; case OP_POW: r = pow_v(s, y); break; default: return -2; }
        }
        out[i] = r;
    }
    return 0;
}

}  // namespace orc

using namespace orc;

// ===========================================================================
// C entry points (ctypes)
// ===========================================================================
extern "C" {

int orc_nansafe(void) { return ORC_NANSAFE; }

// pickchunksize (src/prelude.jl:29-36): round(Int, x/threshold, RoundUp) on Float64 division
long orc_pickchunksize(long input_length, long threshold) {
    if (input_length <= threshold) return input_length;
    long nchunks = (long)std::ceil((double)input_length / (double)threshold);
    return (long)std::ceil((double)input_length / (double)nchunks);
}
// out = {remainder, lastchunksize, lastchunkindex(1-based), middle_hi, nchunks}
int orc_chunk_plan(long xlen, long N, long* out) {
    if (N <= 0 || xlen < N) return -1;      // ArgumentError, gradient.jl:116-118
    Plan p = make_plan(xlen, N);
    out[0] = p.remainder; out[1] = p.lastchunksize; out[2] = p.lastchunkindex; out[3] = p.middle_hi; out[4] = p.nchunks;
    return 0;
}
// structural_eachindex (src/apiutils.jl:44-71) as 0-based column-major storage
// indices; kind 0 dense (rows = length), 1 UpperTriangular, 2 LowerTriangular, 3 Diagonal.
long orc_structural_indices(int kind, long rows, long* out) {
    long k = 0;
    if (kind == 0) { for (long i = 0; i < rows; i++) out[k++] = i; }
    else if (kind == 1) { for (long j = 0; j < rows; j++) for (long i = 0; i <= j; i++) out[k++] = i + j * rows; }
    else if (kind == 2) { for (long j = 0; j < rows; j++) for (long i = j; i < rows; i++) out[k++] = i + j * rows; }
    else if (kind == 3) { for (long i = 0; i < rows; i++) out[k++] = i + i * rows; }
    else return -1;
    return k;
}
long orc_structural_length(int kind, long rows) {   // src/prelude.jl:15-20
    if (kind == 0 || kind == 3) return rows;
    return (rows * (rows + 1)) >> 1;
}

// index <= 0 selects the un-windowed forms: seed!(duals,x,seeds) (apiutils.jl:107-123,
// first N structural positions) and seed_zero_partials!(duals,x) (apiutils.jl:76-77, all).
int orc_seed(double* duals, const double* x, long nstruct, const long* sidx, int N, long index, long chunksize) {
    if (index <= 0) { index = 1; chunksize = N; }
    ORC_DISPATCH_N(N, (seed<double, NN>((Dual<double, NN>*)duals, x, nstruct, sidx, index, chunksize)));
    return 0;
}
int orc_seed_zero(double* duals, const double* x, long nstruct, const long* sidx, int N, long index, long count) {
    if (index <= 0) { index = 1; count = nstruct; }
    ORC_DISPATCH_N(N, (seed_zero<double, NN>((Dual<double, NN>*)duals, x, nstruct, sidx, index, count)));
    return 0;
}
// extract_gradient_chunk!(T, result, dual, index, chunksize)  src/gradient.jl:74-81
int orc_extract_gradient_chunk(double* result, long nstruct, const long* sidx, const double* ydual, int N,
                               long index, long chunksize) {
    for (long i = 1; i <= chunksize; i++) {
        long pos = index - 1 + (i - 1);
        if (pos >= nstruct) break;
        result[sidx_at(sidx, pos)] = ydual[i];           // partials(T, dual, i); ydual[0] is the value
    }
    (void)N; return 0;
}
// extract_jacobian_chunk!(T, result, ydual, index, chunksize)  src/jacobian.jl:109-118
int orc_extract_jacobian_chunk(double* result, long ld, const double* ydual, long m, int N, long index, long chunksize) {
    for (long k = 1; k <= chunksize; k++)
        for (long r = 0; r < m; r++) result[r + (index - 1 + k - 1) * ld] = ydual[r * (N + 1) + k];
    return 0;
}
// extract_jacobian!(T, result, ydual, n)  src/jacobian.jl:95-102
int orc_extract_jacobian(double* result, long ld, const double* ydual, long m, int N, long n) {
    for (long k = 1; k <= n; k++)
        for (long r = 0; r < m; r++) result[r + (k - 1) * ld] = ydual[r * (N + 1) + k];
    return 0;
}
// extract_value!: map!(d -> value(T,d), y, ydual)  src/apiutils.jl:9-12
int orc_extract_value(double* y, const double* ydual, long m, int N) {
    for (long r = 0; r < m; r++) y[r] = ydual[r * (N + 1)];
    return 0;
}

int orc_map1(int op, double* out, const double* a, long n, int N) {
    int rc = 0;
    ORC_DISPATCH_N(N, (rc = map1<NN>(op, (Dual<double, NN>*)out, (const Dual<double, NN>*)a, n)));
    return rc;
}
int orc_map2(int op, int kind, double* out, const double* a, const double* b, double s, long n, int N) {
    int rc = 0;
    ORC_DISPATCH_N(N, (rc = map2<NN>(op, kind, (Dual<double, NN>*)out, (const Dual<double, NN>*)a,
                                     (const Dual<double, NN>*)b, s, n)));
    return rc;
}
int orc_fma3(int kind, double* out, const double* a, const double* b, const double* c, double s, long n, int N) {
    // kind 0: fma(x,y,z) all dual; 1: fma(x,y,s); 2: fma(x,s,z)
    ORC_DISPATCH_N(N, ({
        typedef Dual<double, NN> D; D* o = (D*)out; const D* x = (const D*)a; const D* y = (const D*)b; const D* z = (const D*)c;
        for (long i = 0; i < n; i++) {
            if (kind == 0) o[i] = fma_xyz(x[i], y[i], z[i]);
            else if (kind == 1) o[i] = fma_xy(x[i], y[i], s);
            else o[i] = fma_xz(x[i], s, z[i]);
        }
    }));
    return 0;
}
// sequential reductions of an AoS dual array: op 1 = sum (zero + a1 + a2 ...), 2 = prod
int orc_reduce(int op, double* out, const double* a, long n, int N) {
    ORC_DISPATCH_N(N, ({
        typedef Dual<double, NN> D; const D* x = (const D*)a; D r;
        if (n == 0) return -1;
        r = x[0];
        for (long i = 1; i < n; i++) r = (op == 1) ? (r + x[i]) : (r * x[i]);
        std::memcpy(out, &r, sizeof(D));
    }));
    return 0;
}
// C = A * B on column-major matrices of Duals (AoS): Julia's generic matrix product, C[i,j] = sum over l of A[i,l]*B[l,j]
// with the Dual*Dual body of src/dual.jl:260-300 (val = va*vb; p = pa*vb + pb*va) and Dual+Dual, accumulated in ascending l.
int orc_dual_matmul(double* c, const double* a, const double* b, long m, long k, long n, int N) {
    ORC_DISPATCH_N(N, ({
        typedef Dual<double, NN> D; const D* A = (const D*)a; const D* B = (const D*)b; D* C = (D*)c;
        for (long j = 0; j < n; j++)
            for (long i = 0; i < m; i++) {
                D r = A[i] * B[j * k];
                for (long l = 1; l < k; l++) r = r + A[i + l * m] * B[l + j * k];
                C[i + j * m] = r;
            }
    }));
    return 0;
}
// evaluate a scalar target on an AoS dual array: out = f(xdual)
int orc_eval_scalar_fn(int fn, double* out, const double* xdual, long n, int N) {
    ORC_DISPATCH_N(N, ({
        typedef Dual<double, NN> D; D r = eval_scalar_fn<D>(fn, (const D*)xdual, n); std::memcpy(out, &r, sizeof(D));
    }));
    return 0;
}
double orc_eval_scalar_fn_value(int fn, const double* x, long n) {
    // plain Float64 evaluation = Dual<double,1> with zero partials' value lane
    std::vector<Dual<double, 1>> xd(n);
    for (long i = 0; i < n; i++) { xd[i].v = x[i]; xd[i].p[0] = 0.0; }
    return eval_scalar_fn<Dual<double, 1>>(fn, xd.data(), n).v;
}
// evaluate an array target on an AoS dual array
int orc_eval_array_fn(int fn, double* ydual, long m, const double* xdual, long n, int N,
                      const double* A, long lda, const double* b, double alpha) {
    JacParams par = {A, lda, b, alpha};
    ORC_DISPATCH_N(N, ({ typedef Dual<double, NN> D; eval_array_fn<D>(fn, (D*)ydual, m, (const D*)xdual, n, par); }));
    return 0;
}

// ForwardDiff.gradient!(result, f, x, GradientConfig(f, x, Chunk{N}()))  -- chunks [clo,chi)
// (pass clo=0, chi=-1 for all).  Returns -3 for N > n (ArgumentError, gradient.jl:116-118).
int orc_gradient(int fn, const double* x, long n, int N, double* grad, double* value, long clo, long chi) {
    if (n == 0) { if (value) *value = 0.0; return 0; }         // Chunk{0}: fill!(result, 0) on an empty result
    if (N > n) return -3;
    ORC_DISPATCH_N(N, ({
        typedef Dual<double, NN> D;
        std::vector<D> xdual(n);
        auto f = [&](const D* z, long nn) { return eval_scalar_fn<D>(fn, z, nn); };
        if (n != NN && chi < 0) chi = make_plan(n, NN).nchunks;
        gradient_any<double, NN>(f, x, n, grad, value, clo, chi, xdual.data());
    }));
    return 0;
}
// jacobian of the two-layer target (fn 4): A1 h x n, A2 m x h
int orc_jacobian_mlp2(const double* x, long n, long h, long m, int N, const double* A1, long lda1, const double* b1, const double* A2,
                      long lda2, const double* b2, double* J, long ldJ, double* y, long clo, long chi) {
    if (N > n) return -3;
    JacParams par = {A1, lda1, b1, 0.0, A2, lda2, b2, h};
    ORC_DISPATCH_N(N, ({
        if (n == NN) vector_mode_jacobian<NN>(4, false, x, n, m, par, J, ldJ, y);
        else { if (chi < 0) chi = make_plan(n, NN).nchunks;
               chunk_mode_jacobian<NN>(4, false, x, n, m, par, J, ldJ, y, clo, chi); }
    }));
    return 0;
}
// ForwardDiff.jacobian / jacobian!  f-form (inplace=0) or f!-form (inplace=1)
int orc_jacobian(int fn, int inplace, const double* x, long n, long m, int N, const double* A, long lda,
                 const double* b, double alpha, double* J, long ldJ, double* y, long clo, long chi) {
    if (N > n) return -3;
    JacParams par = {A, lda, b, alpha};
    ORC_DISPATCH_N(N, ({
        if (n == NN) vector_mode_jacobian<NN>(fn, inplace != 0, x, n, m, par, J, ldJ, y);
        else { if (chi < 0) chi = make_plan(n, NN).nchunks;
               chunk_mode_jacobian<NN>(fn, inplace != 0, x, n, m, par, J, ldJ, y, clo, chi); }
    }));
    return 0;
}
// ForwardDiff.hessian! with a HessianResult: H (column-major n x n), gradient, value
int orc_hessian(int fn, const double* x, long n, int N, double* H, long ldH, double* grad, double* value,
                long olo, long ohi) {
    if (N > n) return -3;
    ORC_DISPATCH_N8(N, ({
        if (n != NN && ohi < 0) ohi = make_plan(n, NN).nchunks;
        chunk_mode_hessian<NN>(fn, x, n, H, ldH, grad, value, olo, ohi);
    }));
    return 0;
}
// jacobian(y -> hessian(f, y), x) in vector mode, n == N <= 4 (test/MiscTest.jl:39-49)
int orc_third_order(int fn, const double* x, int n, double* T3) {
    switch (n) {
        case 1: third_order_vector_mode<1>(fn, x, T3); break;
        case 2: third_order_vector_mode<2>(fn, x, T3); break;
        case 3: third_order_vector_mode<3>(fn, x, T3); break;
        case 4: third_order_vector_mode<4>(fn, x, T3); break;
        default: return -1;
    }
    return 0;
}

}  // extern "C"
