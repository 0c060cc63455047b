// fdb_dual.cuh -- the register-resident dual number of the B200 engine.
//
// One thread owns one element: its value and its N partials live in registers
// (fully unrolled arrays), and the element's planes are read from / written to
// SoA device arrays so that a warp touches 32 consecutive doubles of one plane.
//
// Arithmetic contract (what must match the reference bit for bit, given the same
// primal values): ForwardDiff.jl src/partials.jl:82-123,202-224 and
// src/dual.jl:213-218,499-597,625-722; the DiffRules-derived formulas are listed
// in SURVEY.md Appendix A.  The reference never contracts a*b+c, so this
// translation unit MUST be compiled with -fmad=false; the only fused operations
// are the explicit fma() calls that mirror Base.fma in src/dual.jl:625-665.
//
// The value type V is `double` or, for Hessians, another Dual<double,N,NS>
// (Dual{T,Dual{T,V,N},N} with one tag on both levels, src/config.jl:198-201).
#pragma once
#ifndef __CUDACC_RTC__     // NVRTC (fdb_jit.cu) has the device math built in and no host headers
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#else
typedef long long int64_t;
#endif

namespace fdb {

#define FDB_DEV __device__ __forceinline__

// NS = nansafe_mode preference (src/prelude.jl:1; src/partials.jl:106-123)
template <class V, int N, bool NS = false> struct Dual {
    V v;
    V p[N];
};

// ----- value-type helpers ---------------------------------------------------
FDB_DEV double zero_of(const double&) { return 0.0; }
FDB_DEV double one_of(const double&) { return 1.0; }
template <class V, int N, bool NS> FDB_DEV Dual<V, N, NS> zero_of(const Dual<V, N, NS>& d) {
    Dual<V, N, NS> r; r.v = zero_of(d.v);
#pragma unroll
    for (int i = 0; i < N; i++) r.p[i] = zero_of(d.v);
    return r;
}
template <class V, int N, bool NS> FDB_DEV Dual<V, N, NS> one_of(const Dual<V, N, NS>& d) {
    Dual<V, N, NS> r; r.v = one_of(d.v);
#pragma unroll
    for (int i = 0; i < N; i++) r.p[i] = zero_of(d.v);
    return r;
}
FDB_DEV bool iszero_v(const double& x) { return x == 0.0; }
template <class V, int N, bool NS> FDB_DEV bool iszero_v(const Dual<V, N, NS>& d) {
    bool z = iszero_v(d.v);
#pragma unroll
    for (int i = 0; i < N; i++) z = z && iszero_v(d.p[i]);
    return z;
}

// _mul_partial / _div_partial, src/partials.jl:106-123
template <bool NS> FDB_DEV double mulp(double partial, double x) {
    double y = partial * x;
    if (NS) return partial == 0.0 ? 0.0 : y;
    return y;
}
template <bool NS> FDB_DEV double divp(double partial, double x) {
    double y = partial / x;
    if (NS) return partial == 0.0 ? 0.0 : y;
    return y;
}
template <bool NS, class V, int N, class X> FDB_DEV Dual<V, N, NS> mulp(const Dual<V, N, NS>& partial, const X& x) {
    Dual<V, N, NS> y = partial * x;
    if (NS) { if (iszero_v(partial)) return zero_of(y); }
    return y;
}

// ----- + and -  (src/dual.jl:499-519) ---------------------------------------
template <class V, int N, bool NS> FDB_DEV Dual<V, N, NS> operator+(const Dual<V, N, NS>& x, const Dual<V, N, NS>& y) {
    Dual<V, N, NS> r; r.v = x.v + y.v;
#pragma unroll
    for (int i = 0; i < N; i++) r.p[i] = x.p[i] + y.p[i];
    return r;
}
template <class V, int N, bool NS> FDB_DEV Dual<V, N, NS> operator+(const Dual<V, N, NS>& x, double s) {
    Dual<V, N, NS> r = x; r.v = x.v + s; return r;
}
template <class V, int N, bool NS> FDB_DEV Dual<V, N, NS> operator+(double s, const Dual<V, N, NS>& y) {
    Dual<V, N, NS> r = y; r.v = s + y.v; return r;
}
template <class V, int N, bool NS> FDB_DEV Dual<V, N, NS> operator-(const Dual<V, N, NS>& x, const Dual<V, N, NS>& y) {
    Dual<V, N, NS> r; r.v = x.v - y.v;
#pragma unroll
    for (int i = 0; i < N; i++) r.p[i] = x.p[i] - y.p[i];
    return r;
}
template <class V, int N, bool NS> FDB_DEV Dual<V, N, NS> operator-(const Dual<V, N, NS>& x, double s) {
    Dual<V, N, NS> r = x; r.v = x.v - s; return r;
}
template <class V, int N, bool NS> FDB_DEV Dual<V, N, NS> operator-(const Dual<V, N, NS>& x) {
    Dual<V, N, NS> r; r.v = -x.v;
#pragma unroll
    for (int i = 0; i < N; i++) r.p[i] = -x.p[i];
    return r;
}
template <class V, int N, bool NS> FDB_DEV Dual<V, N, NS> operator-(double s, const Dual<V, N, NS>& y) {
    Dual<V, N, NS> r; r.v = s - y.v;
#pragma unroll
    for (int i = 0; i < N; i++) r.p[i] = -y.p[i];
    return r;
}

// ----- *  (DiffRules; test/DualTest.jl:470-472) -----------------------------
template <class V, int N, bool NS> FDB_DEV Dual<V, N, NS> operator*(const Dual<V, N, NS>& x, const Dual<V, N, NS>& y) {
    Dual<V, N, NS> r; r.v = x.v * y.v;
#pragma unroll
    for (int i = 0; i < N; i++) r.p[i] = mulp<NS>(x.p[i], y.v) + mulp<NS>(y.p[i], x.v);
    return r;
}
template <class V, int N, bool NS> FDB_DEV Dual<V, N, NS> operator*(const Dual<V, N, NS>& x, double s) {
    Dual<V, N, NS> r; r.v = x.v * s;
#pragma unroll
    for (int i = 0; i < N; i++) r.p[i] = mulp<NS>(x.p[i], s);
    return r;
}
template <class V, int N, bool NS> FDB_DEV Dual<V, N, NS> operator*(double s, const Dual<V, N, NS>& y) {
    Dual<V, N, NS> r; r.v = s * y.v;
#pragma unroll
    for (int i = 0; i < N; i++) r.p[i] = mulp<NS>(y.p[i], s);
    return r;
}

// ----- /  (src/dual.jl:532-544; src/partials.jl:99-101) ---------------------
FDB_DEV double inv_v(double x) { return 1.0 / x; }
template <int N, bool NS> FDB_DEV Dual<double, N, NS> operator/(const Dual<double, N, NS>& x, const Dual<double, N, NS>& y) {
    Dual<double, N, NS> r; r.v = x.v / y.v;
    double fa = 1.0 / y.v;
    double fb = -(x.v / (y.v * y.v));
#pragma unroll
    for (int i = 0; i < N; i++) r.p[i] = mulp<NS>(x.p[i], fa) + mulp<NS>(y.p[i], fb);
    return r;
}
template <int N, bool NS> FDB_DEV Dual<double, N, NS> operator/(const Dual<double, N, NS>& x, double s) {
    Dual<double, N, NS> r; r.v = x.v / s;
#pragma unroll
    for (int i = 0; i < N; i++) r.p[i] = divp<NS>(x.p[i], s);
    return r;
}
template <int N, bool NS> FDB_DEV Dual<double, N, NS> operator/(double s, const Dual<double, N, NS>& y) {
    Dual<double, N, NS> r; double divv = s / y.v; r.v = divv;
    double c = -(divv / y.v);
#pragma unroll
    for (int i = 0; i < N; i++) r.p[i] = mulp<NS>(y.p[i], c);
    return r;
}

// ----- chain rule: Dual(val, partials * deriv)  (src/dual.jl:213-215) -------
template <class V, int N, bool NS, class DV>
FDB_DEV Dual<V, N, NS> chain(const V& val, const DV& deriv, const Dual<V, N, NS>& d) {
    Dual<V, N, NS> r; r.v = val;
#pragma unroll
    for (int i = 0; i < N; i++) r.p[i] = mulp<NS>(d.p[i], deriv);
    return r;
}

// ----- literal_pow 0..3  (src/dual.jl:587-597) ------------------------------
FDB_DEV double pow1(double v) { return v; }
FDB_DEV double pow2(double v) { return v * v; }
FDB_DEV double pow3(double v) { return v * v * v; }
template <class V, int N, bool NS> FDB_DEV Dual<V, N, NS> pow0(const Dual<V, N, NS>& x) { return one_of(x); }
template <class V, int N, bool NS> FDB_DEV Dual<V, N, NS> pow1(const Dual<V, N, NS>& x) {
    return chain(pow1(x.v), 1.0 * one_of(x.v), x);
}
template <class V, int N, bool NS> FDB_DEV Dual<V, N, NS> pow2(const Dual<V, N, NS>& x) {
    return chain(pow2(x.v), 2.0 * pow1(x.v), x);
}
template <class V, int N, bool NS> FDB_DEV Dual<V, N, NS> pow3(const Dual<V, N, NS>& x) {
    return chain(pow3(x.v), 3.0 * pow2(x.v), x);
}

// ----- DiffRules unary set -------------------------------
template <int N, bool NS> FDB_DEV Dual<double, N, NS> exp_v(const Dual<double, N, NS>& x) {
    double val = exp(x.v); return chain(val, val, x);
}
template <int N, bool NS> FDB_DEV Dual<double, N, NS> log_v(const Dual<double, N, NS>& x) {
    return chain(log(x.v), 1.0 / x.v, x);
}
template <int N, bool NS> FDB_DEV Dual<double, N, NS> sqrt_v(const Dual<double, N, NS>& x) {
    double val = sqrt(x.v); return chain(val, 1.0 / (2.0 * val), x);
}
template <int N, bool NS> FDB_DEV Dual<double, N, NS> tanh_v(const Dual<double, N, NS>& x) {
    double val = tanh(x.v); return chain(val, 1.0 - val * val, x);
}
template <int N, bool NS> FDB_DEV Dual<double, N, NS> inv_v(const Dual<double, N, NS>& x) {
    double val = 1.0 / x.v; return chain(val, -(val * val), x);
}
template <int N, bool NS> FDB_DEV Dual<double, N, NS> abs2_v(const Dual<double, N, NS>& x) {
    return chain(x.v * x.v, x.v + x.v, x);
}
template <int N, bool NS> FDB_DEV Dual<double, N, NS> abs_v(const Dual<double, N, NS>& x) {
    return chain(fabs(x.v), signbit(x.v) ? -1.0 : 1.0, x);     // abs'(0) = +1, test/MiscTest.jl:183
}
// further DiffRules unary rules: val = f(v), deriv shares sub-expressions with val
template <int N, bool NS> FDB_DEV Dual<double, N, NS> tan_v(const Dual<double, N, NS>& x) { double val = tan(x.v); return chain(val, 1.0 + val * val, x); }
template <int N, bool NS> FDB_DEV Dual<double, N, NS> exp2_v(const Dual<double, N, NS>& x) { double val = exp2(x.v); return chain(val, val * 0.6931471805599453, x); }
template <int N, bool NS> FDB_DEV Dual<double, N, NS> log2_v(const Dual<double, N, NS>& x) { return chain(log2(x.v), (1.0 / x.v) / 0.6931471805599453, x); }
template <int N, bool NS> FDB_DEV Dual<double, N, NS> log10_v(const Dual<double, N, NS>& x) { return chain(log10(x.v), (1.0 / x.v) / 2.302585092994046, x); }
template <int N, bool NS> FDB_DEV Dual<double, N, NS> log1p_v(const Dual<double, N, NS>& x) { return chain(log1p(x.v), 1.0 / (x.v + 1.0), x); }
template <int N, bool NS> FDB_DEV Dual<double, N, NS> expm1_v(const Dual<double, N, NS>& x) { return chain(expm1(x.v), exp(x.v), x); }
template <int N, bool NS> FDB_DEV Dual<double, N, NS> sinh_v(const Dual<double, N, NS>& x) { return chain(sinh(x.v), cosh(x.v), x); }
template <int N, bool NS> FDB_DEV Dual<double, N, NS> cosh_v(const Dual<double, N, NS>& x) { return chain(cosh(x.v), sinh(x.v), x); }
template <int N, bool NS> FDB_DEV Dual<double, N, NS> asin_v(const Dual<double, N, NS>& x) { return chain(asin(x.v), 1.0 / sqrt(1.0 - x.v * x.v), x); }
template <int N, bool NS> FDB_DEV Dual<double, N, NS> acos_v(const Dual<double, N, NS>& x) { return chain(acos(x.v), -(1.0 / sqrt(1.0 - x.v * x.v)), x); }
template <int N, bool NS> FDB_DEV Dual<double, N, NS> atan_v(const Dual<double, N, NS>& x) { return chain(atan(x.v), 1.0 / (1.0 + x.v * x.v), x); }
template <int N, bool NS> FDB_DEV Dual<double, N, NS> cbrt_v(const Dual<double, N, NS>& x) { double val = cbrt(x.v); return chain(val, 1.0 / (3.0 * (val * val)), x); }
// binary atan(y, x) and hypot(x, y)
template <int N, bool NS> FDB_DEV Dual<double, N, NS> atan2_v(const Dual<double, N, NS>& y, const Dual<double, N, NS>& x) {
    Dual<double, N, NS> r; r.v = atan2(y.v, x.v);
    double den = y.v * y.v + x.v * x.v; double dy = x.v / den, dx = -y.v / den;
#pragma unroll
    for (int i = 0; i < N; i++) r.p[i] = mulp<NS>(y.p[i], dy) + mulp<NS>(x.p[i], dx);
    return r;
}
template <int N, bool NS> FDB_DEV Dual<double, N, NS> hypot_v(const Dual<double, N, NS>& x, const Dual<double, N, NS>& y) {
    Dual<double, N, NS> r; double h = hypot(x.v, y.v); r.v = h;
    double dx = x.v / h, dy = y.v / h;
#pragma unroll
    for (int i = 0; i < N; i++) r.p[i] = mulp<NS>(x.p[i], dx) + mulp<NS>(y.p[i], dy);
    return r;
}
// sin / cos: one sincos, cos partial is (-s)*p  (src/dual.jl:709-717)
template <int N, bool NS> FDB_DEV Dual<double, N, NS> sin_v(const Dual<double, N, NS>& x) {
    double s, c; sincos(x.v, &s, &c); return chain(s, c, x);
}
template <int N, bool NS> FDB_DEV Dual<double, N, NS> cos_v(const Dual<double, N, NS>& x) {
    double s, c; sincos(x.v, &s, &c); return chain(c, -s, x);
}

// ----- the same rules one level up: value type V = Dual{T,Float64,M} (Hessians) ------------------------------
// Same definitions as above with every scalar operation replaced by the inner Dual operation, exactly what
// Julia's dispatch does for Dual{T,Dual{T,V,M},N} (same tag on both levels, src/config.jl:198-201).
#define FDB_NESTED template <int M, int N, bool NS> FDB_DEV Dual<Dual<double, M, NS>, N, NS>
#define FDB_NARG const Dual<Dual<double, M, NS>, N, NS>&
FDB_NESTED exp_v(FDB_NARG x) { Dual<double, M, NS> val = exp_v(x.v); return chain(val, val, x); }
FDB_NESTED log_v(FDB_NARG x) { return chain(log_v(x.v), inv_v(x.v), x); }
FDB_NESTED sqrt_v(FDB_NARG x) { Dual<double, M, NS> val = sqrt_v(x.v); return chain(val, inv_v(2.0 * val), x); }
FDB_NESTED tanh_v(FDB_NARG x) { Dual<double, M, NS> val = tanh_v(x.v); return chain(val, 1.0 - pow2(val), x); }
FDB_NESTED inv_v(FDB_NARG x) { Dual<double, M, NS> val = inv_v(x.v); return chain(val, -abs2_v(val), x); }
FDB_NESTED abs2_v(FDB_NARG x) { return chain(abs2_v(x.v), x.v + x.v, x); }
FDB_NESTED abs_v(FDB_NARG x) {
    Dual<double, M, NS> o = one_of(x.v);
    return signbit(x.v.v) ? chain(abs_v(x.v), -o, x) : chain(abs_v(x.v), o, x);
}
FDB_NESTED sin_v(FDB_NARG x) { return chain(sin_v(x.v), cos_v(x.v), x); }         // Base.sincos(::Dual), src/dual.jl:719-722
FDB_NESTED cos_v(FDB_NARG x) { return chain(cos_v(x.v), -sin_v(x.v), x); }
FDB_NESTED operator/(FDB_NARG x, FDB_NARG y) {
    Dual<Dual<double, M, NS>, N, NS> r; r.v = x.v / y.v;
    Dual<double, M, NS> fa = inv_v(y.v);
    Dual<double, M, NS> fb = -(x.v / (y.v * y.v));
#pragma unroll
    for (int i = 0; i < N; i++) r.p[i] = mulp<NS>(x.p[i], fa) + mulp<NS>(y.p[i], fb);
    return r;
}
FDB_NESTED operator/(FDB_NARG x, double s) {
    Dual<Dual<double, M, NS>, N, NS> r; r.v = x.v / s;
#pragma unroll
    for (int i = 0; i < N; i++) {
        r.p[i] = x.p[i] / s;
        if (NS) { if (iszero_v(x.p[i])) r.p[i] = zero_of(x.p[i]); }
    }
    return r;
}
FDB_NESTED operator/(double s, FDB_NARG y) {
    Dual<Dual<double, M, NS>, N, NS> r; Dual<double, M, NS> divv = s / y.v; r.v = divv;
    Dual<double, M, NS> c = -(divv / y.v);
#pragma unroll
    for (int i = 0; i < N; i++) r.p[i] = mulp<NS>(y.p[i], c);
    return r;
}
#undef FDB_NESTED
#undef FDB_NARG

// ----- ^ with a run-time exponent  (src/dual.jl:549-585) --------------------
template <int N, bool NS> FDB_DEV bool partials_iszero(const Dual<double, N, NS>& x) {
    bool z = true;
#pragma unroll
    for (int i = 0; i < N; i++) z = z && (x.p[i] == 0.0);
    return z;
}
template <int N, bool NS> FDB_DEV Dual<double, N, NS> pow_v(const Dual<double, N, NS>& x, double y) {
    Dual<double, N, NS> r; r.v = pow(x.v, y);
    if (y == 0.0 || partials_iszero(x)) {
#pragma unroll
        for (int i = 0; i < N; i++) r.p[i] = 0.0;
    } else {
        double w = pow(x.v, y - 1.0);
#pragma unroll
        for (int i = 0; i < N; i++) r.p[i] = mulp<NS>(mulp<NS>(x.p[i], y), w);
    }
    return r;
}
template <int N, bool NS> FDB_DEV Dual<double, N, NS> pow_v(const Dual<double, N, NS>& x, const Dual<double, N, NS>& y) {
    Dual<double, N, NS> r; double expv = pow(x.v, y.v); r.v = expv;
    double powval = y.v * pow(x.v, y.v - 1.0);
    double logval;
    if (partials_iszero(y)) logval = 1.0;
    else if (x.v == 0.0 && y.v > 0) logval = 0.0;
    else logval = expv * log(x.v);
#pragma unroll
    for (int i = 0; i < N; i++) r.p[i] = mulp<NS>(x.p[i], powval) + mulp<NS>(y.p[i], logval);
    return r;
}
template <int N, bool NS> FDB_DEV Dual<double, N, NS> pow_v(double x, const Dual<double, N, NS>& y) {
    double expv = pow(x, y.v);
    double deriv = (x == 0.0 && y.v > 0) ? 0.0 : expv * log(x);
    return chain(expv, deriv, y);
}

// ----- fma  (src/dual.jl:625-665) -------------------------------------------
template <int N, bool NS> FDB_DEV Dual<double, N, NS> fma_xyz(const Dual<double, N, NS>& x, const Dual<double, N, NS>& y, const Dual<double, N, NS>& z) {
    Dual<double, N, NS> r; r.v = fma(x.v, y.v, z.v);
#pragma unroll
    for (int i = 0; i < N; i++) r.p[i] = fma(x.v, y.p[i], fma(y.v, x.p[i], z.p[i]));
    return r;
}
template <int N, bool NS> FDB_DEV Dual<double, N, NS> fma_xy(const Dual<double, N, NS>& x, const Dual<double, N, NS>& y, double z) {
    Dual<double, N, NS> r; r.v = fma(x.v, y.v, z);
#pragma unroll
    for (int i = 0; i < N; i++) r.p[i] = mulp<NS>(x.p[i], y.v) + mulp<NS>(y.p[i], x.v);
    return r;
}
template <int N, bool NS> FDB_DEV Dual<double, N, NS> fma_xz(const Dual<double, N, NS>& x, double y, const Dual<double, N, NS>& z) {
    Dual<double, N, NS> r; r.v = fma(x.v, y, z.v);
#pragma unroll
    for (int i = 0; i < N; i++) r.p[i] = fma(x.p[i], y, z.p[i]);
    return r;
}

// ----- binary max / min (DiffRules: 1 for the selected argument) ------------
template <int N, bool NS> FDB_DEV Dual<double, N, NS> max_v(const Dual<double, N, NS>& x, const Dual<double, N, NS>& y) {
    Dual<double, N, NS> r; r.v = fmax(x.v, y.v);
    double dx = x.v > y.v ? 1.0 : 0.0, dy = x.v > y.v ? 0.0 : 1.0;
#pragma unroll
    for (int i = 0; i < N; i++) r.p[i] = mulp<NS>(x.p[i], dx) + mulp<NS>(y.p[i], dy);
    return r;
}
// NaNMath.max / NaNMath.min: a NaN operand is ignored (the other one is returned); derivative 1 for the argument returned
template <int N, bool NS> FDB_DEV Dual<double, N, NS> nanmax_v(const Dual<double, N, NS>& x, const Dual<double, N, NS>& y) {
    const bool takex = isnan(y.v) || (!isnan(x.v) && x.v > y.v);
    Dual<double, N, NS> r; r.v = takex ? x.v : y.v;
    const double dx = takex ? 1.0 : 0.0, dy = takex ? 0.0 : 1.0;
#pragma unroll
    for (int i = 0; i < N; i++) r.p[i] = mulp<NS>(x.p[i], dx) + mulp<NS>(y.p[i], dy);
    return r;
}
template <int N, bool NS> FDB_DEV Dual<double, N, NS> nanmin_v(const Dual<double, N, NS>& x, const Dual<double, N, NS>& y) {
    const bool takex = isnan(y.v) || (!isnan(x.v) && !(x.v > y.v));
    Dual<double, N, NS> r; r.v = takex ? x.v : y.v;
    const double dx = takex ? 1.0 : 0.0, dy = takex ? 0.0 : 1.0;
#pragma unroll
    for (int i = 0; i < N; i++) r.p[i] = mulp<NS>(x.p[i], dx) + mulp<NS>(y.p[i], dy);
    return r;
}
template <int N, bool NS> FDB_DEV Dual<double, N, NS> min_v(const Dual<double, N, NS>& x, const Dual<double, N, NS>& y) {
    Dual<double, N, NS> r; r.v = fmin(x.v, y.v);
    double dx = x.v > y.v ? 0.0 : 1.0, dy = x.v > y.v ? 1.0 : 0.0;
#pragma unroll
    for (int i = 0; i < N; i++) r.p[i] = mulp<NS>(x.p[i], dx) + mulp<NS>(y.p[i], dy);
    return r;
}

// ----- structural index sets (src/apiutils.jl:44-71): UpperTriangular, LowerTriangular, Diagonal over rows x rows column-major storage
// kind 0 dense, 1 upper ((i,j) for j in 1:n for i in 1:j), 2 lower ((i,j) for j in 1:n for i in j:n), 3 diagonal
// 0-based storage index of structural position pos
__host__ __device__ inline long struct_index_of(int kind, long rows, long pos) {
    switch (kind) {
        case 3: return pos * (rows + 1);
        case 1: {
            long j = (long)((sqrt(8.0 * (double)pos + 1.0) - 1.0) * 0.5);
            while (j * (j + 1) / 2 > pos) j--;
            while ((j + 1) * (j + 2) / 2 <= pos) j++;
            return (pos - j * (j + 1) / 2) + j * rows;
        }
        case 2: {
            const double r = (double)rows + 0.5;
            long j = (long)(r - sqrt(r * r - 2.0 * (double)pos));
            if (j < 0) j = 0;
            if (j > rows - 1) j = rows - 1;
            while (j > 0 && j * rows - j * (j - 1) / 2 > pos) j--;
            while (j + 1 < rows && (j + 1) * rows - (j + 1) * j / 2 <= pos) j++;
            return (j + (pos - (j * rows - j * (j - 1) / 2))) + j * rows;
        }
    }
    return pos;
}
// structural position of storage index idx, or -1 when the entry is structurally zero
__host__ __device__ inline long struct_pos_of(int kind, long rows, long idx) {
    if (kind == 0) return idx;
    const long c = idx / rows, r = idx - c * rows;
    if (kind == 3) return r == c ? r : -1;
    if (kind == 1) return r <= c ? c * (c + 1) / 2 + r : -1;
    return r >= c ? c * rows - c * (c - 1) / 2 + (r - c) : -1;
}

// ----- implicit seeds --------------------------------------------------------
// The Dual that seed!/seed_zero_partials! would have stored at structural
// position `pos` (0-based) while the chunk [start, start+chunksize) is seeded:
// partial k is one(V) iff pos == start + k  (src/apiutils.jl:125-143,
// src/partials.jl:9-12), zero otherwise (apiutils.jl:89-105).
template <int N, bool NS> FDB_DEV Dual<double, N, NS> seeded(double xv, long pos, long start, int chunksize) {
    Dual<double, N, NS> d; d.v = xv;
    long off = pos - start;
#pragma unroll
    for (int k = 0; k < N; k++) d.p[k] = (off == k && k < chunksize) ? 1.0 : 0.0;
    return d;
}

// ----- warp / block reductions of a Dual -------------------------------------
FDB_DEV double shfl_down_d(double x, int delta) { return __shfl_down_sync(0xffffffffu, x, delta); }
FDB_DEV double shfl_xor_d(double x, int m) { return __shfl_xor_sync(0xffffffffu, x, m); }
template <int N, bool NS> FDB_DEV Dual<double, N, NS> shfl_down(const Dual<double, N, NS>& a, int delta) {
    Dual<double, N, NS> r; r.v = shfl_down_d(a.v, delta);
#pragma unroll
    for (int i = 0; i < N; i++) r.p[i] = shfl_down_d(a.p[i], delta);
    return r;
}

}  // namespace fdb
