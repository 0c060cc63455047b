// ref_driver.cpp -- thin C entry points around the reference's OWN native
// benchmark program (benchmarks/cpp of ForwardDiff.jl), compiled from the
// sources where they lie under /root/reference (see Makefile target `ref`).
// Test infrastructure only: used to cross-check oracle.cpp for chunk sizes 1..5.
//
// The reference's translation unit benchmarks.cpp holds the `rosenbrock` and
// `ackley` templates, `testgrad()` and a `main`; it is included verbatim at
// compile time (not copied) with its `main` renamed so it can live in a library.
#define main fdref_benchmark_main
#include "benchmarks.cpp"   // resolved through -I/root/reference/benchmarks/cpp
#undef main

extern "C" {

// runs the reference's own known-answer asserts (benchmarks.cpp:35-96)
int fdref_testgrad(void) { testgrad(); return 0; }

// gradient<F<DualK>>(result, dualvec, input) for fn 1 = rosenbrock, 2 = ackley.
// The reference requires len % K == 0 (benchmarks.h:120-227).
int fdref_gradient(int fn, int K, const double* x, long n, double* out) {
    if (K < 1 || K > 5 || n % K != 0) return -1;
    std::vector<double> input(x, x + n), result(n);
    switch (K) {
        case 1: { std::vector<Dual1> dv(n);
                  if (fn == 1) gradient<rosenbrock<Dual1>>(result, dv, input); else gradient<ackley<Dual1>>(result, dv, input); } break;
        case 2: { std::vector<Dual2> dv(n);
                  if (fn == 1) gradient<rosenbrock<Dual2>>(result, dv, input); else gradient<ackley<Dual2>>(result, dv, input); } break;
        case 3: { std::vector<Dual3> dv(n);
                  if (fn == 1) gradient<rosenbrock<Dual3>>(result, dv, input); else gradient<ackley<Dual3>>(result, dv, input); } break;
        case 4: { std::vector<Dual4> dv(n);
                  if (fn == 1) gradient<rosenbrock<Dual4>>(result, dv, input); else gradient<ackley<Dual4>>(result, dv, input); } break;
        case 5: { std::vector<Dual5> dv(n);
                  if (fn == 1) gradient<rosenbrock<Dual5>>(result, dv, input); else gradient<ackley<Dual5>>(result, dv, input); } break;
    }
    for (long i = 0; i < n; i++) out[i] = result[i];
    return 0;
}

double fdref_value(int fn, const double* x, long n) {
    std::vector<double> input(x, x + n);
    return fn == 1 ? rosenbrock(input) : ackley(input);
}

}  // extern "C"
