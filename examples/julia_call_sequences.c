/* julia_call_sequences.c -- every C-ABI call SEQUENCE that forwarddiff.jl_b200/julia/ForwardDiffB200.jl makes, executed from
 * plain C with the argument order and C types of the `ccall`s (Julia is not installed in this image, so the shim itself cannot
 * run; what it would ask of the library is run here instead).  Sections carry the shim's headings.  Results are checked against
 * closed forms.
 *
 *   gcc -O2 -Iinclude examples/julia_call_sequences.c -Lforwarddiff.jl_b200/lib -lfdb200 -Wl,-rpath,$PWD/forwarddiff.jl_b200/lib -lm -o /tmp/jcs
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "fdb200.h"

static fdb_ctx* ctx;
#define CHECK(call)                                                                                   \
    do {                                                                                              \
        int rc__ = (call);                                                                            \
        if (rc__ != FDB_OK) { fprintf(stderr, "%s:%d %s failed (%d): %s\n", __FILE__, __LINE__, #call, rc__, fdb_last_error(ctx)); exit(1); } \
    } while (0)
#define REQUIRE(cond) do { if (!(cond)) { fprintf(stderr, "%s:%d requirement failed: %s\n", __FILE__, __LINE__, #cond); exit(1); } } while (0)

static double close_to(double a, double b) { return fabs(a - b) <= 1e-12 * fmax(1.0, fabs(b)); }

/* B200Array(x::Vector{Float64}) */
static fdb_array* upload(const double* x, int64_t n) {
    fdb_array* a = NULL;
    CHECK(fdb_array_alloc(ctx, n, 0, 0, &a));            /* B200Array{Float64}(undef, n): chunk_of = 0, dual_level = 0 */
    CHECK(fdb_upload_values(a, x));
    return a;
}
static fdb_array* plain(int64_t n) { fdb_array* a = NULL; CHECK(fdb_array_alloc(ctx, n, 0, 0, &a)); return a; }

int main(void) {
    /* ---- context: one per process and GPU ---- */
    REQUIRE(fdb_init(0, &ctx) == FDB_OK);

    enum { n = 50, N = 4 };
    double xh[n], g[n], ref[n];
    for (int i = 0; i < n; i++) xh[i] = 0.5 + 0.01 * i;

    /* ---- device arrays: B200Array(x), similar(x, Dual{T,Float64,N}) (the seam), Array(a) ---- */
    fdb_array* x = upload(xh, n);
    fdb_array* xdual = NULL;
    CHECK(fdb_array_alloc(ctx, n, N, 1, &xdual));        /* similar(x, Dual{T,Float64,4}): chunk_of = 4, dual_level = 1 */
    double back[n];
    CHECK(fdb_download_values(x, back));
    REQUIRE(memcmp(back, xh, sizeof xh) == 0);

    /* ---- sweep utilities + array arithmetic: the reference chunk loop (src/gradient.jl:114-159) that
     *      `invoke(chunk_mode_gradient!, ...)` runs on device arrays for f(x) = sum(3 .* x.^2) ---- */
    fdb_array* result = plain(n);
    fdb_array *t1 = NULL, *t2 = NULL, *ydual = NULL;
    CHECK(fdb_array_alloc(ctx, n, N, 1, &t1)); CHECK(fdb_array_alloc(ctx, n, N, 1, &t2)); CHECK(fdb_array_alloc(ctx, 1, N, 1, &ydual));
    const int64_t nchunks = (n + N - 1) / N, lastchunksize = n % N == 0 ? N : n % N, lastchunkindex = n - lastchunksize + 1;
    CHECK(fdb_seed_zero(ctx, xdual, x, 0, 0));                                   /* seed_zero_partials!(xdual, x) */
    for (int64_t c = 1; c <= nchunks; c++) {
        const int64_t index = c == nchunks ? lastchunkindex : (c - 1) * N + 1, cs = c == nchunks ? lastchunksize : N;
        CHECK(fdb_seed(ctx, xdual, x, index, cs));                               /* seed!(xdual, x, index, seeds, chunksize) */
        CHECK(fdb_map1(ctx, 13, t1, xdual));                                     /* map1(:pow2, xdual)  (x .^ 2 -> literal_pow) */
        CHECK(fdb_map2_scalar(ctx, 3, t2, t1, 3.0, 1));                          /* map2(:*, t1, 3.0, scalar_first = true) */
        CHECK(fdb_reduce(ctx, 1, ydual, t2));                                    /* sum(t2) */
        CHECK(fdb_extract_gradient_chunk(ctx, result, ydual, index, cs));        /* extract_gradient_chunk!(T, result, ydual, index, cs) */
        CHECK(fdb_seed_zero(ctx, xdual, x, index, cs));                          /* seed_zero_partials!(xdual, x, index) */
    }
    CHECK(fdb_download_values(result, g));
    for (int i = 0; i < n; i++) REQUIRE(g[i] == 6.0 * xh[i]);
    fdb_array* yval = plain(1);
    CHECK(fdb_extract_value(ctx, yval, ydual));                                  /* extract_value!(T, out, ydual) */
    double v = 0, vs = 0;
    CHECK(fdb_download_values(yval, &v));
    for (int i = 0; i < n; i++) vs += 3.0 * xh[i] * xh[i];
    REQUIRE(close_to(v, vs));
    { fdb_array* p = NULL; CHECK(fdb_array_alloc(ctx, 1, N, 1, &p)); CHECK(fdb_reduce(ctx, 2, p, xdual)); CHECK(fdb_array_free(p)); }   /* prod(xdual) */
    CHECK(fdb_map2(ctx, 1, t2, t1, xdual));                                      /* map2(:+, a, b) */

    /* ---- whole-path drivers: chunk_mode_gradient!(result, rosenbrock, x, cfg) with catalogue_id = 1 ---- */
    fdb_chunk_plan plan;                                                          /* Ref{NTuple{11,Int64}}: fields 8, 9 (1-based) = chunk_lo, chunk_hi */
    REQUIRE(sizeof plan == 11 * sizeof(int64_t));
    REQUIRE(fdb_plan_chunks(n, N, 0, 1, &plan) == FDB_OK);
    REQUIRE(((int64_t*)&plan)[7] == plan.chunk_lo && ((int64_t*)&plan)[8] == plan.chunk_hi && plan.chunk_hi == nchunks);
    fdb_array* value = plain(1);
    CHECK(fdb_gradient_chunked(ctx, FDB_FN_ROSENBROCK, x, N, plan.chunk_lo, plan.chunk_hi, result, value));
    CHECK(fdb_download_values(result, g));
    for (int i = 0; i < n; i++) {
        ref[i] = 0.0;
        if (i < n - 1) ref[i] += -2.0 * (1.0 - xh[i]) - 400.0 * xh[i] * (xh[i + 1] - xh[i] * xh[i]);
        if (i > 0) ref[i] += 200.0 * (xh[i] - xh[i - 1] * xh[i - 1]);
        REQUIRE(close_to(g[i], ref[i]));
    }
    /* gradient_host!(out, f, x, Chunk{N}()) */
    double vh = NAN;
    memset(g, 0, sizeof g);
    CHECK(fdb_gradient_host(ctx, 1, xh, n, N, 0, -1, g, &vh));
    for (int i = 0; i < n; i++) REQUIRE(close_to(g[i], ref[i]));

    /* ---- chunk_mode_jacobian!(result, f, x, cfg) with a JacobianTarget: tanh.(A*x .+ b), then the f!(y, x) form (Brusselator) ---- */
    enum { m = 6 };
    double Ah[m * n], bh[m], Jh[m * n], yh[m];
    for (int i = 0; i < m * n; i++) Ah[i] = sin(0.37 * i) / 7.0;
    for (int i = 0; i < m; i++) bh[i] = 0.1 * i;
    fdb_array *A = upload(Ah, m * n), *b = upload(bh, m), *J = plain(m * n), *y = plain(m);
    CHECK(fdb_jacobian_chunked(ctx, FDB_JFN_TANH_AFFINE, x, m, N, A, m, b, 0.0, 0, -1, J, m, y));
    CHECK(fdb_download_values(J, Jh)); CHECK(fdb_download_values(y, yh));
    for (int r = 0; r < m; r++) {
        double z = bh[r];
        for (int k = 0; k < n; k++) z += Ah[r + k * m] * xh[k];
        const double t = tanh(z);
        REQUIRE(close_to(yh[r], t));
        for (int k = 0; k < n; k++) REQUIRE(close_to(Jh[r + k * m], (1.0 - t * t) * Ah[r + k * m]));
    }
    fdb_array *Jb = plain(n * n), *yb = plain(n);
    CHECK(fdb_jacobian_chunked(ctx, FDB_JFN_BRUSSELATOR, x, n, N, NULL, n, NULL, 2.5, 0, -1, Jb, n, yb));        /* handle(nothing) = C_NULL */
    /* jacobian_host!(J, y, id, x, Chunk{N}(); A, b) and its reuse_params form */
    double Jh2[m * n], yh2[m];
    CHECK(fdb_jacobian_host(ctx, FDB_JFN_TANH_AFFINE, xh, n, m, N, Ah, m, bh, 0.0, 0, 0, -1, Jh2, m, yh2));
    for (int i = 0; i < m * n; i++) REQUIRE(close_to(Jh2[i], Jh[i]));
    CHECK(fdb_jacobian_host(ctx, FDB_JFN_TANH_AFFINE, xh, n, m, N, Ah, m, bh, 0.0, 1, 0, -1, Jh2, m, yh2));
    for (int i = 0; i < m * n; i++) REQUIRE(close_to(Jh2[i], Jh[i]));

    /* ---- hessian!(result, f, x, cfg; grad, value) and hessian_host! ---- */
    fdb_array *H = plain(n * n), *hg = plain(n);
    CHECK(fdb_hessian_chunked(ctx, FDB_FN_ROSENBROCK, x, N, 0, -1, H, n, hg, value));
    static double Hh[n * n], Hh2[n * n];
    double hgh[n], hv = 0;
    CHECK(fdb_download_values(H, Hh));
    CHECK(fdb_hessian_host(ctx, 1, xh, n, N, 0, -1, Hh2, n, hgh, &hv));
    for (int i = 0; i < n * n; i++) REQUIRE(close_to(Hh[i], Hh2[i]));
    REQUIRE(close_to(Hh[1 + 0 * n], -400.0 * xh[0]) && close_to(Hh[(n - 1) + (n - 1) * n], 200.0));
    for (int i = 0; i < n; i++) REQUIRE(close_to(hgh[i], ref[i]));

    /* ---- op-programs: program_gradient! / program_hessian! for f(x) = sum(x.^2) as trace_scalar would emit it ---- */
    const int32_t term[3][4] = {{0 << 6, 0, 0, 0}, {(1 << 6) | 13, 0, 0, 0}, {5 << 6, 0, 0, 0}};      /* LOADX r0 ; r0 = pow2(r0) ; acc0 += r0 */
    CHECK(fdb_program_gradient_chunked(ctx, &term[0][0], 3, NULL, 0, NULL, 0, 1, 0, 0, x, N, 0, -1, result, value));
    CHECK(fdb_download_values(result, g));
    for (int i = 0; i < n; i++) REQUIRE(g[i] == 2.0 * xh[i]);
    CHECK(fdb_program_hessian_chunked(ctx, &term[0][0], 3, NULL, 0, NULL, 0, 1, 0, 0, x, N, 0, -1, H, n, hg, value));
    CHECK(fdb_download_values(H, Hh));
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) REQUIRE(Hh[i + j * n] == (i == j ? 2.0 : 0.0));
    /* bind_params!(ctx, arrays): f(x) = sum(w .* x) with w a data array -> kind 7 (Dual (x) param) */
    double wh[n];
    for (int i = 0; i < n; i++) wh[i] = 1.0 + i;
    fdb_array* w = upload(wh, n);
    const fdb_array* params[1] = {w};
    CHECK(fdb_program_bind_params(ctx, 1, params));
    const int32_t termw[3][4] = {{0 << 6, 0, 0, 0}, {(7 << 6) | 3, 0, 0, 0}, {5 << 6, 0, 0, 0}};      /* r0 = x(i) ; r0 = r0 * param[0][i] ; acc0 += r0 */
    int rc = fdb_program_gradient_chunked(ctx, &termw[0][0], 3, NULL, 0, NULL, 0, 1, 0, 0, x, N, 0, -1, result, value);
    if (rc == FDB_OK) {                                   /* array parameters need the run-time specialised kernel (libnvrtc) */
        CHECK(fdb_download_values(result, g));
        for (int i = 0; i < n; i++) REQUIRE(g[i] == wh[i]);
    } else REQUIRE(rc == FDB_ERR_UNSUPPORTED);
    CHECK(fdb_program_bind_params(ctx, 0, NULL));

    /* ---- the generic chunk loop as a replayed CUDA graph: chunk_mode_gradient_graph! for f(x) = sum(3 .* x.^2) ---- */
    CHECK(fdb_seed(ctx, xdual, x, 1, N)); CHECK(fdb_seed_zero(ctx, xdual, x, N + 1, n - N));
    CHECK(fdb_map1(ctx, 13, t1, xdual)); CHECK(fdb_map2_scalar(ctx, 3, t2, t1, 3.0, 1)); CHECK(fdb_reduce(ctx, 1, ydual, t2));
    CHECK(fdb_extract_gradient_chunk(ctx, result, ydual, 1, N)); CHECK(fdb_seed_zero(ctx, xdual, x, 1, N));
    CHECK(fdb_cursor_set(ctx, N + 1, N < n - N ? N : n - N));
    CHECK(fdb_graph_begin(ctx));
    CHECK(fdb_seed_at_cursor(ctx, xdual, x));
    CHECK(fdb_map1(ctx, 13, t1, xdual)); CHECK(fdb_map2_scalar(ctx, 3, t2, t1, 3.0, 1)); CHECK(fdb_reduce(ctx, 1, ydual, t2));
    CHECK(fdb_extract_gradient_chunk_at_cursor(ctx, result, ydual));
    CHECK(fdb_seed_zero_at_cursor(ctx, xdual, x));
    CHECK(fdb_cursor_advance(ctx, N, n));
    void* exec = NULL;
    CHECK(fdb_graph_end(ctx, &exec));
    CHECK(fdb_graph_launch(ctx, exec, nchunks - 1));
    CHECK(fdb_graph_destroy(exec));
    CHECK(fdb_download_values(result, g));
    for (int i = 0; i < n; i++) REQUIRE(g[i] == 6.0 * xh[i]);

    /* ---- structured inputs: seed_structured! / extract_gradient_chunk_structured! on an UpperTriangular 3 x 3 ---- */
    enum { rows = 3 };
    double uh[rows * rows] = {1, 0, 0, 2, 3, 0, 4, 5, 6};
    fdb_array *u = upload(uh, rows * rows), *ud = NULL, *ures = plain(rows * rows), *uy = NULL;
    CHECK(fdb_array_alloc(ctx, rows * rows, 3, 1, &ud)); CHECK(fdb_array_alloc(ctx, 1, 3, 1, &uy));
    CHECK(fdb_fill(ctx, ud, 0.0));
    for (int64_t index = 1; index <= 6; index += 3) {                             /* structural_length = 6: two chunks of 3 */
        CHECK(fdb_seed_structured(ctx, ud, u, 1, rows, index, 3));
        CHECK(fdb_reduce(ctx, 1, uy, ud));                                        /* f = sum */
        CHECK(fdb_extract_gradient_chunk_structured(ctx, ures, uy, 1, rows, index, 3));
        CHECK(fdb_seed_zero_structured(ctx, ud, u, 1, rows, index, 3));
    }
    double ug[rows * rows];
    CHECK(fdb_download_values(ures, ug));
    for (int j = 0; j < rows; j++) for (int i = 0; i < rows; i++) REQUIRE(ug[i + j * rows] == (i <= j ? 1.0 : 0.0));   /* gradient(sum, UpperTriangular) == UpperTriangular(ones) */

    /* ---- multi-GPU, one process per GPU: PeerExchange + sharded_gradient! (a world of one here; tools/check_sharded.py runs the
     *      same sequence under torchrun on 2/4/8 GPUs) ---- */
    fdb_exchange* xch = NULL;
    CHECK(fdb_exchange_create(ctx, 3 * 64, 0, 1, &xch));
    unsigned char mine[FDB_IPC_HANDLE_BYTES];
    CHECK(fdb_exchange_handle(xch, mine));
    CHECK(fdb_exchange_connect(xch, mine));                                       /* allgather64(mine) of a single rank */
    fdb_array *ar = NULL, *av = NULL;
    CHECK(fdb_exchange_view(xch, 0, n, 64, &ar)); CHECK(fdb_exchange_view(xch, 64, 1, 32, &av));   /* arena_view(x, offset, n, ld) */
    CHECK(fdb_exchange_barrier(xch));
    CHECK(fdb_exchange_enable(ctx, xch));
    CHECK(fdb_gradient_chunked(ctx, 1, x, N, plan.chunk_lo, plan.chunk_hi, ar, av));
    CHECK(fdb_exchange_enable(ctx, NULL));
    CHECK(fdb_exchange_barrier(xch));
    CHECK(fdb_download_values(ar, g));
    for (int i = 0; i < n; i++) REQUIRE(close_to(g[i], ref[i]));
    CHECK(fdb_array_free(ar)); CHECK(fdb_array_free(av));
    CHECK(fdb_exchange_destroy(xch));

    /* ---- ONE Julia session, all GPUs: DeviceGroup, gradient_host! / jacobian_host! / hessian_host!(g, ...) ---- */
    fdb_multi* grp = NULL;
    int ndev = 0;
    CHECK(fdb_device_count(&ndev));
    const int two_on_one[2] = {0, 0};
    REQUIRE((ndev >= 2 ? fdb_multi_init(0, NULL, &grp) : fdb_multi_init(2, two_on_one, &grp)) == FDB_OK);
#define MCHECK(call) do { int rc__ = (call); if (rc__ != FDB_OK) { fprintf(stderr, "%s:%d %s failed (%d): %s\n", __FILE__, __LINE__, #call, rc__, fdb_multi_last_error(grp)); exit(1); } } while (0)
    double vm = 0;
    memset(g, 0, sizeof g);
    MCHECK(fdb_multi_gradient_host(grp, 1, xh, n, N, g, &vm));
    for (int i = 0; i < n; i++) REQUIRE(close_to(g[i], ref[i]));
    REQUIRE(close_to(vm, vh));
    MCHECK(fdb_multi_jacobian_host(grp, FDB_JFN_TANH_AFFINE, xh, n, m, N, Ah, m, bh, 0.0, 0, Jh2, m, yh2));
    for (int i = 0; i < m * n; i++) REQUIRE(close_to(Jh2[i], Jh[i]));
    MCHECK(fdb_multi_hessian_host(grp, 1, xh, n, N, Hh2, n, hgh, &hv));
    CHECK(fdb_hessian_host(ctx, 1, xh, n, N, 0, -1, Hh, n, hgh, &hv));
    for (int i = 0; i < n * n; i++) REQUIRE(Hh[i] == Hh2[i]);
    REQUIRE(fdb_multi_shutdown(grp) == FDB_OK);

    /* finalizers: fdb_array_free for every array, fdb_shutdown for the context */
    fdb_array* all[] = {x, xdual, result, t1, t2, ydual, yval, value, A, b, J, y, Jb, yb, H, hg, w, u, ud, ures, uy};
    for (size_t i = 0; i < sizeof all / sizeof all[0]; i++) CHECK(fdb_array_free(all[i]));
    REQUIRE(fdb_shutdown(ctx) == FDB_OK);
    puts("julia_call_sequences ok");
    return 0;
}
