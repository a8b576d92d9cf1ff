/* c_abi_demo.c -- the C ABI from plain C (include/b200adj.h compiled as C99, no CUDA or C++ headers):
 * one small Lorenz ensemble gradient (GaussAdjoint, fixed-step Tsit5, in-kernel cost dgdu = u - 2) with HOST buffers.
 *
 *   gcc -std=c99 -Wall -Wextra -I include examples/c_abi_demo.c -o examples/c_abi_demo -L scimlsensitivity.jl_b200 -lb200adj \
 *       -Wl,-rpath,'$ORIGIN/../scimlsensitivity.jl_b200' -lm
 *
 * Without a usable CUDA device the library has no fallback: b200adj_create returns B200ADJ_ERR_NO_DEVICE and the program
 * says so (exit code 3) -- that contract is what tests/test_abi_and_host.py checks on the CPU-only build machine.  This is the
 * call sequence the Julia glue (julia/B200AdjointExt.jl) issues through ccall. */
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include "b200adj.h"

int main(void) {
    enum { N = 256, D = 3, P = 3, K = 11 };
    static double u0[D][N], du0[D][N], saved[K][D][N];
    double p[P] = {10.0, 28.0, 8.0 / 3.0}, dp[P], saveat[K];
    int32_t status[N];
    for (int k = 0; k < K; k++) saveat[k] = 0.1 * k;
    for (int i = 0; i < N; i++) { u0[0][i] = 1.0 + 0.001 * i; u0[1][i] = 0.0; u0[2][i] = 0.0; }

    if (b200adj_sizeof_cfg() != sizeof(b200adj_cfg)) { fprintf(stderr, "header / library mismatch\n"); return 2; }
    b200adj_cfg cfg = {0};
    cfg.rhs_family = B200ADJ_FAM_LORENZ; cfg.sensealg = B200ADJ_SA_GAUSS; cfg.stepper = B200ADJ_ST_TSIT5_FIXED; cfg.dtype = B200ADJ_F64;
    cfg.d = D; cfg.P = P; cfg.m = 0; cfg.K = K; cfg.N = N;
    cfg.t0 = 0.0; cfg.t1 = 1.0; cfg.dt = 0.01;
    cfg.saveat = saveat; cfg.shared_p = 1; cfg.buffers_on_device = 0; cfg.device = 0;
    cfg.cost_kind = B200ADJ_COST_AFFINE; cfg.cost_a = 1.0; cfg.cost_b = -2.0;      /* dgdu(t_k) = u - 2 */

    void* h = NULL;
    int32_t rc = b200adj_create(&cfg, &h);
    if (rc == B200ADJ_ERR_NO_DEVICE) { printf("no CUDA device: %s\n", b200adj_last_error(NULL)); return 3; }
    if (rc != B200ADJ_OK) { fprintf(stderr, "create failed (%d): %s\n", (int)rc, b200adj_last_error(NULL)); return 1; }
    rc = b200adj_forward(h, u0, p, NULL, saved, status);
    if (rc == B200ADJ_OK) rc = b200adj_reverse(h, NULL, du0, dp);
    if (rc != B200ADJ_OK) { fprintf(stderr, "solve failed (%d): %s\n", (int)rc, b200adj_last_error(h)); b200adj_destroy(h); return 1; }
    int bad = 0;
    for (int i = 0; i < N; i++) bad += status[i] != 0;
    printf("version %#x, launches %lld, members with a non-zero status %d\n", (unsigned)b200adj_version(), (long long)b200adj_launch_count(h), bad);
    printf("u(1)[member 0] = (%.12g, %.12g, %.12g)\n", saved[K - 1][0][0], saved[K - 1][1][0], saved[K - 1][2][0]);
    printf("dG/dp = (%.12g, %.12g, %.12g)   dG/du0[member 0] = (%.12g, %.12g, %.12g)\n", dp[0], dp[1], dp[2], du0[0][0], du0[1][0], du0[2][0]);
    b200adj_destroy(h);
    return (bad == 0 && isfinite(dp[0])) ? 0 : 1;
}
