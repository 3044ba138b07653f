/* Plain-C consumer of include/reach.h: the ABI exercised without ctypes or any Python in between.
 *
 *   consumer            load-and-probe: version, device count; without a GPU reach_create must fail with a message
 *   consumer --run      needs a B200: reach_lgg09 (inhomogeneous, one-shot, host buffers), reach_glgm06 (homogeneous)
 *                       and reach_multi_lgg09 on a 3-dimensional diagonal system whose support values / generators
 *                       have closed forms, checked here in C.
 * Built by tests/test_c_consumer.py with `gcc -std=c99 -Wall -Werror -I include consumer.c -L<pkg> -lreach_cuda`.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "reach.h"

#define N 3
#define NSTEPS 50
#define NDIRS 6

static int fail(const char* what, const reach_ctx* ctx) {
    fprintf(stderr, "FAIL %s: %s\n", what, reach_last_error(ctx));
    return 1;
}

int main(int argc, char** argv) {
    printf("reach_version %d\n", reach_version());
    if (reach_version() < 100) return 1;
    const int ndev = reach_device_count();
    printf("devices %d\n", ndev);
    reach_ctx* ctx = NULL;
    int rc = reach_create(&ctx, 0);
    if (argc < 2 || strcmp(argv[1], "--run") != 0) {
        if (ndev <= 0) {                       /* no GPU: loud failure, never a fallback */
            if (rc == REACH_OK || ctx != NULL) { fprintf(stderr, "reach_create succeeded without a device\n"); return 1; }
            printf("no device: reach_create -> %d (%s)\n", rc, reach_last_error(NULL));
            return strlen(reach_last_error(NULL)) > 0 ? 0 : 1;
        }
        if (rc != REACH_OK) return fail("reach_create", NULL);
        reach_destroy(ctx);
        return 0;
    }
    if (rc != REACH_OK) return fail("reach_create", NULL);

    /* x' = diag(-1, -0.5, 0.25) x + u;  Phi = diag(a_i), column-major */
    const double a[N] = {0.9, 0.97, 1.01};
    double Phi[N * N] = {0};
    for (int i = 0; i < N; ++i) Phi[i + N * i] = a[i];
    double dirs[N * NDIRS] = {0};              /* [+e_1..+e_3, -e_1..-e_3] */
    for (int i = 0; i < N; ++i) { dirs[i + N * i] = 1.0; dirs[i + N * (N + i)] = -1.0; }
    const double c0[N] = {1.0, -2.0, 0.5}, r0[N] = {0.1, 0.2, 0.0};
    const double cu[N] = {0.01, 0.0, -0.02}, ru[N] = {0.001, 0.002, 0.003};
    reach_term tx = {REACH_TERM_BOX, 0, 0, 0, NULL, c0, r0, NULL};
    reach_term tu = {REACH_TERM_BOX, 0, 0, 0, NULL, cu, ru, NULL};
    const int32_t one = 1;
    reach_setexpr Omega0 = {1, &one, &tx}, V = {1, &one, &tu};
    double* rho = (double*)malloc(sizeof(double) * NDIRS * NSTEPS);
    int64_t done = -1;
    rc = reach_lgg09(ctx, N, NDIRS, NSTEPS, Phi, dirs, &Omega0, &V, NULL, NULL, rho, &done);
    if (rc != REACH_OK) return fail("reach_lgg09", ctx);
    if (done != NSTEPS) { fprintf(stderr, "nsteps_done %lld\n", (long long)done); return 1; }
    /* reach_inhomog.jl:49-64 for a diagonal Phi: r_k = a^k e_i, rho = rho(r_k, X0) + s_k, s_{k+1} = s_k + rho(r_k, U) */
    double worst = 0.0;
    for (int i = 0; i < N; ++i)
        for (int sgn = 0; sgn < 2; ++sgn) {
            double r = sgn ? -1.0 : 1.0, s = 0.0;
            for (int k = 0; k < NSTEPS; ++k) {
                const double want = r * c0[i] + fabs(r) * r0[i] + s;
                const double got = rho[(sgn * N + i) + NDIRS * k];
                const double err = fabs(got - want) / fmax(1.0, fabs(want));
                if (err > worst) worst = err;
                s += r * cu[i] + fabs(r) * ru[i];
                r *= a[i];
            }
        }
    printf("reach_lgg09: max scaled error %.3e\n", worst);
    if (!(worst <= 1e-12)) return 1;

    /* the same run through reach_multi_lgg09 on (0, 0) or (0, 1): bit-identical */
    {
        reach_multi* m = NULL;
        const int ids[2] = {0, ndev > 1 ? 1 : 0};
        if (reach_multi_create(&m, ids, 2) != REACH_OK) { fprintf(stderr, "FAIL reach_multi_create: %s\n", reach_multi_last_error(NULL)); return 1; }
        double* rho2 = (double*)malloc(sizeof(double) * NDIRS * NSTEPS);
        rc = reach_multi_lgg09(m, N, NDIRS, NSTEPS, Phi, dirs, &Omega0, &V, NULL, rho2, NULL, NULL);
        if (rc != REACH_OK) { fprintf(stderr, "FAIL reach_multi_lgg09: %s\n", reach_multi_last_error(m)); return 1; }
        if (memcmp(rho, rho2, sizeof(double) * NDIRS * NSTEPS) != 0) { fprintf(stderr, "multi-device result differs\n"); return 1; }
        printf("reach_multi_lgg09: identical on %d devices\n", reach_multi_ndev(m));
        free(rho2);
        reach_multi_destroy(m);
    }

    /* GLGM06 homogeneous (reach_homog.jl:33-73): Z_k = Phi^k Z_0; step 0 is Z_0 itself */
    {
        enum { P0 = 4, GCAP = 15 };
        const double G0[N * P0] = {0.1, 0, 0, 0, 0.2, 0, 0, 0, 0.3, 0.05, -0.05, 0.02};
        double* C = (double*)malloc(sizeof(double) * N * NSTEPS);
        double* G = (double*)malloc(sizeof(double) * N * GCAP * NSTEPS);
        int32_t* ng = (int32_t*)malloc(sizeof(int32_t) * NSTEPS);
        rc = reach_glgm06(ctx, N, NSTEPS, 1, 5, Phi, c0, G0, P0, NULL, NULL, -1, C, G, GCAP, ng);
        if (rc != REACH_OK) return fail("reach_glgm06", ctx);
        worst = 0.0;
        double pw[N] = {1.0, 1.0, 1.0};
        for (int k = 0; k < NSTEPS; ++k) {
            if (ng[k] != P0) { fprintf(stderr, "step %d has %d generators\n", k, ng[k]); return 1; }
            for (int i = 0; i < N; ++i) {
                double e = fabs(C[i + N * k] - pw[i] * c0[i]);
                if (e > worst) worst = e;
                for (int j = 0; j < P0; ++j) {
                    e = fabs(G[i + N * (j + (size_t)GCAP * k)] - pw[i] * G0[i + N * j]);
                    if (e > worst) worst = e;
                }
            }
            for (int i = 0; i < N; ++i) pw[i] *= a[i];
        }
        printf("reach_glgm06: max abs error %.3e\n", worst);
        if (!(worst <= 1e-12)) return 1;
        free(C); free(G); free(ng);
    }
    /* error path: a bad argument comes back as REACH_EINVAL with a message, nothing aborts */
    rc = reach_lgg09(ctx, N, 0, NSTEPS, Phi, dirs, &Omega0, &V, NULL, NULL, rho, NULL);
    if (rc != REACH_EINVAL || strlen(reach_last_error(ctx)) == 0) { fprintf(stderr, "expected REACH_EINVAL, got %d\n", rc); return 1; }
    free(rho);
    reach_destroy(ctx);
    printf("OK\n");
    return 0;
}
