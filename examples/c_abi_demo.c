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
    if (worst > 1e-12) return 1;

    /* ---- one process, all GPUs: what a single Julia session binds to (fdb_multi_*).  With one visible GPU the group is
     * built from two contexts on it, which runs the same code except cudaDeviceEnablePeerAccess. ---- */
    int ndev = 0;
    CHECK(fdb_device_count(&ndev));
    fdb_multi* grp = NULL;
    const int two_on_one[2] = {0, 0};
    if ((ndev >= 2 ? fdb_multi_init(0, NULL, &grp) : fdb_multi_init(2, two_on_one, &grp)) != FDB_OK) {
        fprintf(stderr, "fdb_multi_init: %s\n", fdb_multi_last_error(NULL)); return 2;
    }
    const int gsz = fdb_multi_size(grp);
#define MCHECK(call)                                                                          \
    do {                                                                                      \
        int rc__ = (call);                                                                    \
        if (rc__ != FDB_OK) {                                                                 \
            fprintf(stderr, "%s failed (%d): %s\n", #call, rc__, fdb_multi_last_error(grp)); \
            return 1;                                                                         \
        }                                                                                     \
    } while (0)
    /* gradient!(out, rosenbrock, x, cfg): every device sweeps its chunk range, the host result is complete and
     * bit-identical to the single-context call */
    double* gm = malloc(n * sizeof *gm); double vm = 0.0, v1 = 0.0;
    MCHECK(fdb_multi_gradient_host(grp, FDB_FN_ROSENBROCK, x, n, 12, gm, &vm));
    CHECK(fdb_gradient_host(ctx, FDB_FN_ROSENBROCK, x, n, 12, 0, -1, g, &v1));
    for (long i = 0; i < n; i++)
        if (gm[i] != g[i]) { fprintf(stderr, "multi-device gradient differs at %ld: %.17g vs %.17g\n", i, gm[i], g[i]); return 1; }
    if (vm != v1) { fprintf(stderr, "multi-device value differs\n"); return 1; }
    /* hessian! with column blocks sharded over the devices */
    double Hm[9], ghm[3], vhm = 0.0;
    MCHECK(fdb_multi_hessian_host(grp, FDB_FN_ROSENBROCK, x3, 3, 1, Hm, 3, ghm, &vhm));
    for (int i = 0; i < 9; i++)
        if (fabs(Hm[i] - h3[i]) > 1e-11) { fprintf(stderr, "multi-device hessian mismatch at %d: %.17g\n", i, Hm[i]); return 1; }
    /* jacobian!(J, f!, y, x, cfg) of the Brusselator right-hand side on host arrays: tridiagonal blocks */
    enum { MB = 40, NB = 2 * MB };
    double xb[NB], Jb[NB * NB], yb[NB], Jb1[NB * NB], yb1[NB];
    for (int i = 0; i < MB; i++) { xb[i] = 1.0 + sin(2.0 * 3.141592653589793 * (i + 1) / (MB + 1)); xb[MB + i] = 3.0; }
    const double alpha = (MB + 1.0) * (MB + 1.0) / 50.0;
    MCHECK(fdb_multi_jacobian_host(grp, FDB_JFN_BRUSSELATOR, xb, NB, NB, 8, NULL, NB, NULL, alpha, 0, Jb, NB, yb));
    CHECK(fdb_jacobian_host(ctx, FDB_JFN_BRUSSELATOR, xb, NB, NB, 8, NULL, NB, NULL, alpha, 0, 0, -1, Jb1, NB, yb1));
    for (int i = 0; i < NB * NB; i++) if (Jb[i] != Jb1[i]) { fprintf(stderr, "multi-device jacobian differs at %d\n", i); return 1; }
    for (int i = 0; i < NB; i++) if (yb[i] != yb1[i]) { fprintf(stderr, "multi-device y differs at %d\n", i); return 1; }
    if (fabs(Jb[0 + 0 * NB] - (2.0 * xb[0] * xb[MB] - 4.0 - 2.0 * alpha)) > 1e-12 * fabs(Jb[0])) { fprintf(stderr, "d(du_1)/d(u_1) wrong: %.17g\n", Jb[0]); return 1; }
    /* device-resident: the complete gradient ends up in EVERY device's arena (peer stores fused into the sweep kernel,
     * arenas linked in-process, device-side flag barrier) */
    const int64_t npad = (n + 31) / 32 * 32;
    fdb_exchange* xs[FDB_MAX_RANKS] = {0};
    MCHECK(fdb_multi_exchange_create(grp, 2 * npad + 32, xs));
    MCHECK(fdb_multi_upload(grp, xs, 0, x, n));
    MCHECK(fdb_multi_gradient_resident(grp, xs, FDB_FN_ROSENBROCK, 0, n, 12, npad, 2 * npad));
    for (int d = 0; d < gsz; d++) {
        MCHECK(fdb_multi_download(grp, xs, d, npad, gm, n));
        for (long i = 0; i < n; i++)
            if (gm[i] != g[i]) { fprintf(stderr, "resident gradient on device %d differs at %ld\n", d, i); return 1; }
    }
    MCHECK(fdb_multi_exchange_destroy(grp, xs));
    printf("one process, %d context(s) on %d device(s): host and resident results match the single-context call bit for bit\n", gsz, ndev);
    fdb_multi_shutdown(grp);
    free(gm);
    free(x); free(g);
    fdb_shutdown(ctx);
    puts("c_abi_demo ok");
    return 0;
}
