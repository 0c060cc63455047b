/* c_abi_demo.c -- the drop-in boundary exercised from plain C (no Python, no torch): what a cgo / JNI / ccall
 * binding does.  Computes ForwardDiff.gradient!(out, rosenbrock, x, GradientConfig(rosenbrock, x, Chunk{N}())) and
 * ForwardDiff.hessian! through include/fdb200.h and checks the reference's golden vectors
 * (test/GradientTest.jl:25-54, test/HessianTest.jl:18-54, benchmarks/cpp/benchmarks.cpp:35-96).
 *
 *   gcc -O2 -Iinclude examples/c_abi_demo.c -Lforwarddiff.jl_b200/lib -lfdb200 -Wl,-rpath,$PWD/forwarddiff.jl_b200/lib -lm -o /tmp/c_abi_demo
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "fdb200.h"

#define CHECK(call)                                                                      \
    do {                                                                                 \
        int rc__ = (call);                                                               \
        if (rc__ != FDB_OK) {                                                            \
            fprintf(stderr, "%s failed (%d): %s\n", #call, rc__, fdb_last_error(ctx));  \
            return 1;                                                                    \
        }                                                                                \
    } while (0)

int main(void) {
    fdb_ctx* ctx = NULL;
    if (fdb_init(0, &ctx) != FDB_OK) { fprintf(stderr, "fdb_init: %s\n", fdb_last_error(NULL)); return 2; }
    printf("%s\n", fdb_version());

    /* golden gradient: x = [0.1, 0.2, 0.3] -> [-9.4, 15.6, 52.0], chunk sizes 1, 2, 3 */
    const double x3[3] = {0.1, 0.2, 0.3}, g3[3] = {-9.4, 15.6, 52.0};
    for (int N = 1; N <= 3; N++) {
        double g[3], v = 0.0;
        CHECK(fdb_gradient_host(ctx, FDB_FN_ROSENBROCK, x3, 3, N, 0, -1, g, &v));
        for (int i = 0; i < 3; i++)
            if (fabs(g[i] - g3[i]) > 1e-12) { fprintf(stderr, "gradient mismatch N=%d i=%d: %.17g\n", N, i, g[i]); return 1; }
        if (fabs(v - 11.82) > 1e-12) { fprintf(stderr, "value mismatch %.17g\n", v); return 1; }
    }
    /* exact integer gradient of benchmarks.cpp testgrad(): x = 0..9 */
    const double r10[10] = {-2, -200, 1002, 5804, 16606, 35808, 65810, 109012, 167814, -11000};
    double x10[10], g10[10];
    for (int i = 0; i < 10; i++) x10[i] = i;
    CHECK(fdb_gradient_host(ctx, FDB_FN_ROSENBROCK, x10, 10, 5, 0, -1, g10, NULL));
    for (int i = 0; i < 10; i++)
        if (g10[i] != r10[i]) { fprintf(stderr, "integer gradient mismatch at %d: %.17g\n", i, g10[i]); return 1; }
    /* golden Hessian with value and gradient (HessianResult) */
    const double h3[9] = {-66, -40, 0, -40, 130, -80, 0, -80, 200};
    double H[9], gh[3], vh = 0.0;
    CHECK(fdb_hessian_host(ctx, FDB_FN_ROSENBROCK, x3, 3, 2, 0, -1, H, 3, gh, &vh));
    for (int i = 0; i < 9; i++)
        if (fabs(H[i] - h3[i]) > 1e-11) { fprintf(stderr, "hessian mismatch at %d: %.17g\n", i, H[i]); return 1; }
    /* error behaviour: chunk size larger than the input is the reference's ArgumentError */
    if (fdb_gradient_host(ctx, FDB_FN_ROSENBROCK, x3, 3, 4, 0, -1, H, NULL) != FDB_ERR_CHUNK) { fprintf(stderr, "expected FDB_ERR_CHUNK\n"); return 1; }
    printf("error path ok: %s\n", fdb_last_error(ctx));
    /* a larger run: n = 100 000, Chunk{12}, compared with the closed-form Rosenbrock gradient */
    const long n = 100000;
    double* x = malloc(n * sizeof *x); double* g = malloc(n * sizeof *g);
    for (long i = 0; i < n; i++) x[i] = 0.5 + 0.5 * sin(0.001 * (double)i);
    CHECK(fdb_gradient_host(ctx, FDB_FN_ROSENBROCK, x, n, 12, 0, -1, g, NULL));
    double worst = 0.0;
    for (long i = 0; i < n; i++) {
        double a = 0.0;
        if (i < n - 1) a += -2.0 * (1.0 - x[i]) - 400.0 * x[i] * (x[i + 1] - x[i] * x[i]);
        if (i > 0) a += 200.0 * (x[i] - x[i - 1] * x[i - 1]);
        double d = fabs(g[i] - a) / fmax(1.0, fabs(a));
        if (d > worst) worst = d;
    }
    int64_t launches = 0;
    CHECK(fdb_kernel_launches(ctx, &launches));
    printf("n=%ld Chunk{12}: max rel deviation from the closed form %.3g, %lld kernel launches\n", n, worst, (long long)launches);
    free(x); free(g);
    fdb_shutdown(ctx);
    if (worst > 1e-12) return 1;
    puts("c_abi_demo ok");
    return 0;
}
