// ode_tsit5.cuh -- fixed-step Tsit5 ensemble kernels: forward solve with per-step checkpoints and the fused
// reverse (adjoint) pass.  One ensemble member per thread, SoA [step][dim][member] so every global access is a
// fully coalesced 8 B x 32 lanes = 256 B row; the whole time loop runs inside the kernel so lambda, the dG/dp
// accumulators and the FSAL stages never leave registers (a per-step launch would be
// launch-latency bound at N = 65536).  The body of one loop iteration is exactly "one fused reverse time step":
//   checkpoint load -> forward-stage recompute (dense output data) -> interpolate y at the 6 adjoint stage times
//   -> batched RHS + VJPs -> Tsit5 stage update -> 3-pt Gauss dG/dp accumulate -> jump at save times.
//
// Reference functions replaced (per stage, per member):
//   sense functors   src/interpolating_adjoint.jl:150-174, src/gauss_adjoint.jl:118-128, src/backsolve_adjoint.jl:32-61
//   split_states     sol(y,t,continuity=:right)  src/interpolating_adjoint.jl:190-205, src/gauss_adjoint.jl:158-166
//   vecjacobian!     src/derivative_wrappers.jl:256-267         vec_pjac!  src/gauss_adjoint.jl:629-743
//   GaussIntegrand   src/gauss_adjoint.jl:745-759 (+ upstream IntegratingSumCallback, 3-pt Gauss-Legendre per step)
//   ReverseLossCallback  src/adjoint_common.jl:754-821 (lambda += dgdu at t_k, FSAL k1 recomputed)
//   backsolve_checkpoint_callbacks  src/backsolve_adjoint.jl:523-546
// and the upstream Tsit5 perform_step! / dense interpolant.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "families.cuh"

namespace b200adj {

struct Tsit5Consts {
    double A[7][6];     // A[s][j] (row 6 = b weights)
    double C[7];
    double Bst[4][7];   // dense-output weights b_j(theta) at theta = 1 - c_s for adjoint stages s = 2..5
    double Bq[3][7];    // dense-output weights at theta = (1 -/+ sqrt(.6))/2, 1/2  (3-pt Gauss-Legendre nodes)
    double GW[3];       // Gauss-Legendre weights 5/9, 8/9, 5/9
};
__constant__ Tsit5Consts c_ts;

enum { SA_INTERP = 0, SA_GAUSS = 1, SA_QUAD = 2, SA_BACKSOLVE = 3 };
enum { COST_EXPLICIT = 0, COST_AFFINE = 1 };

struct OdeFwdArgs {
    const double* u0;        // [D][N]
    const double* p;         // [P] or [P][N]
    double* ckpt;            // [S+1][D][N]
    double* saved;           // [K][D][N] or null
    const int32_t* save_of_step;  // [S+1]: save index k at grid point n, or -1
    int32_t* status;         // [N] or null
    int64_t N;
    int32_t S;
    double h;
};

struct OdeRevArgs {
    const double* ckpt;      // [S+1][D][N]
    const double* p;         // [P] or [P][N]
    const double* dLdu;      // [K][D][N] (COST_EXPLICIT)
    const int32_t* save_of_step;
    double* du0;             // [D][N]
    double* dp_members;      // [P][N] when !shared_p
    double* partials;        // [gridDim][P] block partial sums (shared_p)
    double* dp;              // [P] final (shared_p)
    unsigned int* ticket;    // last-block-done counter
    int64_t N;
    int32_t S;
    double h;
    double cost_a, cost_b;
    uint32_t flags;          // bit0 no_start, bit1 no checkpointing (backsolve), bit2 ckpt every step
};

template <int D> __device__ __forceinline__ void load_state(const double* base, int64_t N, int64_t i, double* u) {
#pragma unroll
    for (int j = 0; j < D; j++) u[j] = __ldg(base + (int64_t)j * N + i);
}
template <int D> __device__ __forceinline__ void store_state(double* base, int64_t N, int64_t i, const double* u) {
#pragma unroll
    for (int j = 0; j < D; j++) base[(int64_t)j * N + i] = u[j];
}

// stage value  u + h * sum_{j<s} A[s][j] k_j   (same association order as the oracle)
template <int D, int S_> __device__ __forceinline__ void tsit5_stage(const double* u, const double (*k)[D], double h, double* out) {
#pragma unroll
    for (int i = 0; i < D; i++) {
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < S_; j++) acc = fma(c_ts.A[S_][j], k[j][i], acc);
        out[i] = fma(h, acc, u[i]);
    }
}
// dense output  u + h * sum_j w[j] k_j
template <int D> __device__ __forceinline__ void tsit5_dense(const double* u, const double (*k)[D], double h, const double* w, double* out) {
#pragma unroll
    for (int i = 0; i < D; i++) {
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < 7; j++) acc = fma(w[j], k[j][i], acc);
        out[i] = fma(h, acc, u[i]);
    }
}

// ------------------------------------------------------------------------------------------------------------
// Forward ensemble solve, fixed-step Tsit5, writes every step's state (the dense solution is NOT stored: the
// reverse pass recomputes the 6 stages from u_n, 24 B/step instead of 192 B/step of HBM traffic).
// ------------------------------------------------------------------------------------------------------------
template <class Fam, bool SHARED_P, int BLOCK>
__global__ void __launch_bounds__(BLOCK) tsit5_forward_kernel(OdeFwdArgs a) {
    constexpr int D = Fam::D, P = Fam::P;
    const int64_t gi = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    const bool active = gi < a.N;
    const int64_t i = active ? gi : a.N - 1;
    double p[P];
#pragma unroll
    for (int q = 0; q < P; q++) p[q] = SHARED_P ? __ldg(a.p + q) : __ldg(a.p + (int64_t)q * a.N + i);
    double u[D], k[7][D], tmp[D];
    load_state<D>(a.u0, a.N, i, u);
    const int64_t stride = (int64_t)D * a.N;
    if (active) {
        store_state<D>(a.ckpt, a.N, i, u);
        if (a.saved) { int ks = a.save_of_step[0]; if (ks >= 0) store_state<D>(a.saved + (int64_t)ks * stride, a.N, i, u); }
    }
    Fam::f(u, p, k[0]);
    const double h = a.h;
    for (int n = 0; n < a.S; n++) {
        tsit5_stage<D, 1>(u, k, h, tmp); Fam::f(tmp, p, k[1]);
        tsit5_stage<D, 2>(u, k, h, tmp); Fam::f(tmp, p, k[2]);
        tsit5_stage<D, 3>(u, k, h, tmp); Fam::f(tmp, p, k[3]);
        tsit5_stage<D, 4>(u, k, h, tmp); Fam::f(tmp, p, k[4]);
        tsit5_stage<D, 5>(u, k, h, tmp); Fam::f(tmp, p, k[5]);
        tsit5_stage<D, 6>(u, k, h, tmp);
#pragma unroll
        for (int j = 0; j < D; j++) u[j] = tmp[j];
        Fam::f(u, p, k[0]);                      // FSAL: k7 of this step = k1 of the next
        if (active) {
            store_state<D>(a.ckpt + (int64_t)(n + 1) * stride, a.N, i, u);
            if (a.saved) { int ks = a.save_of_step[n + 1]; if (ks >= 0) store_state<D>(a.saved + (int64_t)ks * stride, a.N, i, u); }
        }
    }
    if (active && a.status) {
        bool ok = true;
#pragma unroll
        for (int j = 0; j < D; j++) ok = ok && isfinite(u[j]);
        a.status[i] = ok ? 0 : 1;
    }
}

// deterministic block reduction of P per-thread values -> partials[block][P]; the last block to finish sums the
// partials in index order (fixed order => bitwise reproducible for a given grid), no floating-point atomics.
template <int P, int BLOCK>
__device__ __forceinline__ void reduce_dp(const double* acc, double* partials, double* dp, unsigned int* ticket) {
    __shared__ double s_red[(BLOCK / 32) * P];
    __shared__ bool s_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < P; q++) {
        double v = acc[q];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if (lane == 0) s_red[warp * P + q] = v;
    }
    __syncthreads();
    if (threadIdx.x < P) {
        double v = 0.0;
        for (int w = 0; w < BLOCK / 32; w++) v += s_red[w * P + threadIdx.x];
        partials[(int64_t)blockIdx.x * P + threadIdx.x] = v;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int t = atomicAdd(ticket, 1u);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        // P x gridDim sums; each warp-lane strides over blocks in a fixed pattern, then a fixed shuffle tree
        for (int q = warp; q < P; q += BLOCK / 32) {
            double v = 0.0;
            for (unsigned int b = lane; b < gridDim.x; b += 32) v += __ldcg(partials + (int64_t)b * P + q);
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
            if (lane == 0) dp[q] = v;
        }
        if (threadIdx.x == 0) *ticket = 0;   // re-arm for the next launch
    }
}

// ------------------------------------------------------------------------------------------------------------
// Fused reverse pass.  SA in {SA_INTERP, SA_GAUSS, SA_BACKSOLVE}.
// ------------------------------------------------------------------------------------------------------------
template <class Fam, int SA, bool SHARED_P, int COST, int BLOCK>
__global__ void __launch_bounds__(BLOCK) tsit5_reverse_kernel(OdeRevArgs a) {
    constexpr int D = Fam::D, P = Fam::P;
    const int64_t gi = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    const bool active = gi < a.N;
    const int64_t i = active ? gi : a.N - 1;
    const int64_t N = a.N, stride = (int64_t)D * N;
    const double h = a.h, hr = -a.h;      // reverse step
    double p[P];
#pragma unroll
    for (int q = 0; q < P; q++) p[q] = SHARED_P ? __ldg(a.p + q) : __ldg(a.p + (int64_t)q * N + i);

    double lam[D], mu[P];                 // mu: dG/dp accumulator (Gauss quadrature sum, or the augmented state)
#pragma unroll
    for (int j = 0; j < D; j++) lam[j] = 0.0;
#pragma unroll
    for (int q = 0; q < P; q++) mu[q] = 0.0;

    double uhi[D];                        // u_{n+1} (checkpoint) -- for Backsolve: the backward-integrated y
    load_state<D>(a.ckpt + (int64_t)a.S * stride, N, i, uhi);

    // jump at t = T (PresetTimeCallback fires at initialisation when T is a save time)
    {
        int ks = a.save_of_step[a.S];
        if (ks >= 0) {
            if (COST == COST_EXPLICIT) {
#pragma unroll
                for (int j = 0; j < D; j++) lam[j] += __ldg(a.dLdu + (int64_t)ks * stride + (int64_t)j * N + i);
            } else {
#pragma unroll
                for (int j = 0; j < D; j++) lam[j] += fma(a.cost_a, uhi[j], a.cost_b);
            }
        }
    }

    if (SA == SA_BACKSOLVE) {
        // z = [lam; mu; y]; dy/dt = f(y) integrated backwards (src/backsolve_adjoint.jl:32-61)
        double ky[7][D], kl[7][D], ys[D], ls[D], dg[P];
        const bool ckpt_on = !(a.flags & 2u), every = (a.flags & 4u);
        bool fsal = false;
        for (int n = a.S - 1; n >= 0; n--) {
            if (!fsal) {
                Fam::f(uhi, p, ky[0]);
                Fam::vjp_u(uhi, p, lam, kl[0]);
#pragma unroll
                for (int j = 0; j < D; j++) kl[0][j] = -kl[0][j];
            }
            // stage 1 contribution to mu: -(df/dp)' lam
            double mus[P];
            Fam::vjp_p(uhi, p, lam, dg);
#pragma unroll
            for (int q = 0; q < P; q++) mus[q] = -c_ts.A[6][0] * dg[q];
#define B200_BS_STAGE(S_)                                                         \
            tsit5_stage<D, S_>(uhi, ky, hr, ys); tsit5_stage<D, S_>(lam, kl, hr, ls);   \
            Fam::f(ys, p, ky[S_]); Fam::vjp_u(ys, p, ls, kl[S_]);                  \
            _Pragma("unroll") for (int j = 0; j < D; j++) kl[S_][j] = -kl[S_][j];  \
            if (S_ < 6) { Fam::vjp_p(ys, p, ls, dg);                               \
                _Pragma("unroll") for (int q = 0; q < P; q++) mus[q] = fma(-c_ts.A[6][S_ < 6 ? S_ : 0], dg[q], mus[q]); }
            B200_BS_STAGE(1) B200_BS_STAGE(2) B200_BS_STAGE(3) B200_BS_STAGE(4) B200_BS_STAGE(5) B200_BS_STAGE(6)
#undef B200_BS_STAGE
            // after stage 6: ys, ls hold the new state (c7 = 1, row 6 = b), ky[6], kl[6] are the FSAL derivatives
#pragma unroll
            for (int j = 0; j < D; j++) { uhi[j] = ys[j]; lam[j] = ls[j]; ky[0][j] = ky[6][j]; kl[0][j] = kl[6][j]; }
#pragma unroll
            for (int q = 0; q < P; q++) mu[q] = fma(hr, mus[q], mu[q]);
            fsal = true;
            // callbacks at t_n: checkpoint reset first, then the loss jump (CallbackSet order, backsolve_adjoint.jl:545)
            const int ks = a.save_of_step[n];
            if (ckpt_on && (every || ks >= 0)) { load_state<D>(a.ckpt + (int64_t)n * stride, N, i, uhi); fsal = false; }
            if (ks >= 0 && !((a.flags & 1u) && n == 0 && false)) {   // no_start never skips for Backsolve (adjoint_common.jl:761)
                if (COST == COST_EXPLICIT) {
#pragma unroll
                    for (int j = 0; j < D; j++) lam[j] += __ldg(a.dLdu + (int64_t)ks * stride + (int64_t)j * N + i);
                } else {
#pragma unroll
                    for (int j = 0; j < D; j++) lam[j] += fma(a.cost_a, uhi[j], a.cost_b);
                }
                fsal = false;
            }
        }
    } else {
        double kf7[D];                         // f(u_{n+1}) = forward k7 of step n (= forward k1 of step n+1)
        Fam::f(uhi, p, kf7);
        double ka1[D];                         // adjoint FSAL stage
        bool fsal = false;
        double ulo[D];
        load_state<D>(a.ckpt + (int64_t)(a.S - 1) * stride, N, i, ulo);
        for (int n = a.S - 1; n >= 0; n--) {
            // prefetch the next checkpoint one full step ahead (hides HBM latency behind ~450 DFMAs)
            double unext[D];
            {
                const int nn = n > 0 ? n - 1 : 0;
                load_state<D>(a.ckpt + (int64_t)nn * stride, N, i, unext);
            }
            // ---- forward stage recompute on [t_n, t_{n+1}]: the dense-output data of this step ----
            double kf[7][D], tmp[D];
            Fam::f(ulo, p, kf[0]);
            tsit5_stage<D, 1>(ulo, kf, h, tmp); Fam::f(tmp, p, kf[1]);
            tsit5_stage<D, 2>(ulo, kf, h, tmp); Fam::f(tmp, p, kf[2]);
            tsit5_stage<D, 3>(ulo, kf, h, tmp); Fam::f(tmp, p, kf[3]);
            tsit5_stage<D, 4>(ulo, kf, h, tmp); Fam::f(tmp, p, kf[4]);
            tsit5_stage<D, 5>(ulo, kf, h, tmp); Fam::f(tmp, p, kf[5]);
#pragma unroll
            for (int j = 0; j < D; j++) kf[6][j] = kf7[j];

            // ---- adjoint Tsit5 step t_{n+1} -> t_n; stage s evaluated at y(t_{n+1} - c_s h) ----
            double ka[7][D], ls[D], y[D], dg[P];
            if (!fsal) {
                Fam::vjp_u(uhi, p, lam, ka[0]);           // y(t_{n+1}) = u_{n+1} (right-continuous lookup at a knot)
#pragma unroll
                for (int j = 0; j < D; j++) ka[0][j] = -ka[0][j];
            } else {
#pragma unroll
                for (int j = 0; j < D; j++) ka[0][j] = ka1[j];
            }
            double mus[P];
            if (SA == SA_INTERP) {
                Fam::vjp_p(uhi, p, lam, dg);
#pragma unroll
                for (int q = 0; q < P; q++) mus[q] = -c_ts.A[6][0] * dg[q];
            }
#define B200_ADJ_STAGE(S_, YEXPR)                                                      \
            tsit5_stage<D, S_>(lam, ka, hr, ls);                                       \
            YEXPR;                                                                     \
            Fam::vjp_u(y, p, ls, ka[S_]);                                              \
            _Pragma("unroll") for (int j = 0; j < D; j++) ka[S_][j] = -ka[S_][j];      \
            if (SA == SA_INTERP && S_ < 6) { Fam::vjp_p(y, p, ls, dg);                 \
                _Pragma("unroll") for (int q = 0; q < P; q++) mus[q] = fma(-c_ts.A[6][S_ < 6 ? S_ : 0], dg[q], mus[q]); }
            B200_ADJ_STAGE(1, tsit5_dense<D>(ulo, kf, h, c_ts.Bst[0], y))
            B200_ADJ_STAGE(2, tsit5_dense<D>(ulo, kf, h, c_ts.Bst[1], y))
            B200_ADJ_STAGE(3, tsit5_dense<D>(ulo, kf, h, c_ts.Bst[2], y))
            B200_ADJ_STAGE(4, tsit5_dense<D>(ulo, kf, h, c_ts.Bst[3], y))
            B200_ADJ_STAGE(5, _Pragma("unroll") for (int j = 0; j < D; j++) y[j] = ulo[j])
            B200_ADJ_STAGE(6, _Pragma("unroll") for (int j = 0; j < D; j++) y[j] = ulo[j])
#undef B200_ADJ_STAGE
            // ls = lambda(t_n) (row 6 of A is b), ka[6] = FSAL derivative at (t_n, ls, u_n)

            if (SA == SA_GAUSS) {
                // 3-point Gauss-Legendre over this step, pre-jump lambda from the adjoint step's own dense output,
                // y from the forward dense output: dp += (h/2) w_q (df/dp)'(y_q) lam_q  (gauss_adjoint.jl:745-759)
                double lq[D], acc[P];
#pragma unroll
                for (int q = 0; q < P; q++) acc[q] = 0.0;
#pragma unroll
                for (int g = 0; g < 3; g++) {
                    tsit5_dense<D>(lam, ka, hr, c_ts.Bq[g], lq);
                    tsit5_dense<D>(ulo, kf, h, c_ts.Bq[2 - g], y);
                    Fam::vjp_p(y, p, lq, dg);
#pragma unroll
                    for (int q = 0; q < P; q++) acc[q] = fma(c_ts.GW[g], dg[q], acc[q]);
                }
#pragma unroll
                for (int q = 0; q < P; q++) mu[q] = fma(0.5 * h, acc[q], mu[q]);
            } else {
#pragma unroll
                for (int q = 0; q < P; q++) mu[q] = fma(hr, mus[q], mu[q]);
            }
#pragma unroll
            for (int j = 0; j < D; j++) { lam[j] = ls[j]; ka1[j] = ka[6][j]; }
            fsal = true;

            // ---- jump at t_n (ReverseLossCallback): lam += dgdu(t_k), FSAL invalidated ----
            const int ks = a.save_of_step[n];
            if (ks >= 0 && !((a.flags & 1u) && n == 0)) {
                if (COST == COST_EXPLICIT) {
#pragma unroll
                    for (int j = 0; j < D; j++) lam[j] += __ldg(a.dLdu + (int64_t)ks * stride + (int64_t)j * N + i);
                } else {
#pragma unroll
                    for (int j = 0; j < D; j++) lam[j] += fma(a.cost_a, ulo[j], a.cost_b);
                }
                fsal = false;
            }
#pragma unroll
            for (int j = 0; j < D; j++) { uhi[j] = ulo[j]; kf7[j] = kf[0][j]; ulo[j] = unext[j]; }
        }
    }

    if (active) store_state<D>(a.du0, N, i, lam);
    if (SHARED_P) {
        if (!active) {
#pragma unroll
            for (int q = 0; q < P; q++) mu[q] = 0.0;
        }
        reduce_dp<P, BLOCK>(mu, a.partials, a.dp, a.ticket);
    } else if (active) {
#pragma unroll
        for (int q = 0; q < P; q++) a.dp_members[(int64_t)q * N + i] = mu[q];
    }
}

}  // namespace b200adj
