// families.cuh -- named RHS families: f(u,p,t) and the hand-differentiated VJPs lam'(df/du), lam'(df/dp).
// This is the device replacement of the AD back-ends behind vecjacobian! / vec_pjac!
// (/root/reference/src/derivative_wrappers.jl:256-267, src/gauss_adjoint.jl:629-743); formulas: SURVEY.md App. C.
// All functions are templated on the real type and fully unrolled (D, P compile-time) so state lives in registers.
#pragma once
#include <cuda_runtime.h>

namespace b200adj {

struct LotkaVolterra {
    static constexpr int D = 2, P = 4, M = 0;
    template <class T> __device__ __forceinline__ static void f(const T* u, const T* p, T* du) {
        du[0] = p[0] * u[0] - p[1] * u[0] * u[1];
        du[1] = -p[2] * u[1] + p[3] * u[0] * u[1];
    }
    // dl = (df/du)' l
    template <class T> __device__ __forceinline__ static void vjp_u(const T* u, const T* p, const T* l, T* dl) {
        dl[0] = l[0] * (p[0] - p[1] * u[1]) + l[1] * p[3] * u[1];
        dl[1] = -l[0] * p[1] * u[0] + l[1] * (-p[2] + p[3] * u[0]);
    }
    // dg = (df/dp)' l
    template <class T> __device__ __forceinline__ static void vjp_p(const T* u, const T* p, const T* l, T* dg) {
        T xy = u[0] * u[1];
        dg[0] = u[0] * l[0]; dg[1] = -xy * l[0]; dg[2] = -u[1] * l[1]; dg[3] = xy * l[1];
    }
    template <class T> __device__ __forceinline__ static void jac(const T* u, const T* p, T (*J)[2]) {
        J[0][0] = p[0] - p[1] * u[1]; J[0][1] = -p[1] * u[0];
        J[1][0] = p[3] * u[1];        J[1][1] = -p[2] + p[3] * u[0];
    }
    template <class T> __device__ __forceinline__ static void djac(const T* p, const T* yd, T (*J)[2]) {
        J[0][0] = -p[1] * yd[1]; J[0][1] = -p[1] * yd[0];
        J[1][0] = p[3] * yd[1];  J[1][1] = p[3] * yd[0];
    }
    // directional derivative of vjp_p with respect to u along ud (Rosenbrock23 on the augmented adjoint states: the time
    // derivative -(dF/dt)'lam with ud = ydot, and the Jacobian block d(F'lam)/dy of BacksolveAdjoint)
    template <class T> __device__ __forceinline__ static void dvjp_p(const T* u, const T* p, const T* ud, const T* l, T* dg) {
        const T dxy = ud[0] * u[1] + u[0] * ud[1];
        dg[0] = ud[0] * l[0]; dg[1] = -dxy * l[0]; dg[2] = -ud[1] * l[1]; dg[3] = dxy * l[1];
    }
};

struct Lorenz {
    static constexpr int D = 3, P = 3, M = 0;
    template <class T> __device__ __forceinline__ static void f(const T* u, const T* p, T* du) {
        du[0] = p[0] * (u[1] - u[0]);
        du[1] = u[0] * (p[1] - u[2]) - u[1];
        du[2] = u[0] * u[1] - p[2] * u[2];
    }
    template <class T> __device__ __forceinline__ static void vjp_u(const T* u, const T* p, const T* l, T* dl) {
        dl[0] = -p[0] * l[0] + (p[1] - u[2]) * l[1] + u[1] * l[2];
        dl[1] = p[0] * l[0] - l[1] + u[0] * l[2];
        dl[2] = -u[0] * l[1] - p[2] * l[2];
    }
    template <class T> __device__ __forceinline__ static void vjp_p(const T* u, const T* p, const T* l, T* dg) {
        dg[0] = (u[1] - u[0]) * l[0]; dg[1] = u[0] * l[1]; dg[2] = -u[2] * l[2];
    }
    template <class T> __device__ __forceinline__ static void jac(const T* u, const T* p, T (*J)[3]) {
        J[0][0] = -p[0];       J[0][1] = p[0]; J[0][2] = 0;
        J[1][0] = p[1] - u[2]; J[1][1] = -1;   J[1][2] = -u[0];
        J[2][0] = u[1];        J[2][1] = u[0]; J[2][2] = -p[2];
    }
    template <class T> __device__ __forceinline__ static void djac(const T* p, const T* yd, T (*J)[3]) {
        J[0][0] = 0;      J[0][1] = 0;     J[0][2] = 0;
        J[1][0] = -yd[2]; J[1][1] = 0;     J[1][2] = -yd[0];
        J[2][0] = yd[1];  J[2][1] = yd[0]; J[2][2] = 0;
    }
    template <class T> __device__ __forceinline__ static void dvjp_p(const T* u, const T* p, const T* ud, const T* l, T* dg) {
        dg[0] = (ud[1] - ud[0]) * l[0]; dg[1] = ud[0] * l[1]; dg[2] = -ud[2] * l[2];
    }
};

struct Robertson {
    static constexpr int D = 3, P = 3, M = 0;
    template <class T> __device__ __forceinline__ static void f(const T* y, const T* k, T* dy) {
        dy[0] = -k[0] * y[0] + k[2] * y[1] * y[2];
        dy[1] = k[0] * y[0] - k[1] * y[1] * y[1] - k[2] * y[1] * y[2];
        dy[2] = k[1] * y[1] * y[1];
    }
    template <class T> __device__ __forceinline__ static void vjp_u(const T* y, const T* k, const T* l, T* dl) {
        dl[0] = -k[0] * l[0] + k[0] * l[1];
        dl[1] = k[2] * y[2] * l[0] - (2 * k[1] * y[1] + k[2] * y[2]) * l[1] + 2 * k[1] * y[1] * l[2];
        dl[2] = k[2] * y[1] * l[0] - k[2] * y[1] * l[1];
    }
    template <class T> __device__ __forceinline__ static void vjp_p(const T* y, const T* k, const T* l, T* dg) {
        dg[0] = -y[0] * l[0] + y[0] * l[1];
        dg[1] = -y[1] * y[1] * l[1] + y[1] * y[1] * l[2];
        dg[2] = y[1] * y[2] * l[0] - y[1] * y[2] * l[1];
    }
    // Jacobian J[i][j] = df_i/dy_j and its directional derivative along yd (constant Hessian)
    template <class T> __device__ __forceinline__ static void jac(const T* y, const T* k, T (*J)[3]) {
        J[0][0] = -k[0]; J[0][1] = k[2] * y[2];                    J[0][2] = k[2] * y[1];
        J[1][0] = k[0];  J[1][1] = -2 * k[1] * y[1] - k[2] * y[2]; J[1][2] = -k[2] * y[1];
        J[2][0] = 0;     J[2][1] = 2 * k[1] * y[1];                J[2][2] = 0;
    }
    template <class T> __device__ __forceinline__ static void djac(const T* k, const T* yd, T (*J)[3]) {
        J[0][0] = 0; J[0][1] = k[2] * yd[2];                     J[0][2] = k[2] * yd[1];
        J[1][0] = 0; J[1][1] = -2 * k[1] * yd[1] - k[2] * yd[2]; J[1][2] = -k[2] * yd[1];
        J[2][0] = 0; J[2][1] = 2 * k[1] * yd[1];                 J[2][2] = 0;
    }
    template <class T> __device__ __forceinline__ static void dvjp_p(const T* y, const T* k, const T* yd, const T* l, T* dg) {
        dg[0] = -yd[0] * l[0] + yd[0] * l[1];
        dg[1] = 2 * y[1] * yd[1] * (l[2] - l[1]);
        dg[2] = (yd[1] * y[2] + y[1] * yd[2]) * (l[0] - l[1]);
    }
};

// bouncing ball (docs/src/examples/hybrid_jump/bouncing_ball.md; test/callbacks/continuous_callbacks.jl): x' = v, v' = -p0;
// p = [gravity, restitution] -- the restitution coefficient enters through the event affect v <- -p1 v only
struct BouncingBall {
    static constexpr int D = 2, P = 2, M = 0;
    template <class T> __device__ __forceinline__ static void f(const T* u, const T* p, T* du) { du[0] = u[1]; du[1] = -p[0]; }
    template <class T> __device__ __forceinline__ static void vjp_u(const T* u, const T* p, const T* l, T* dl) { dl[0] = T(0); dl[1] = l[0]; }
    template <class T> __device__ __forceinline__ static void vjp_p(const T* u, const T* p, const T* l, T* dg) { dg[0] = -l[1]; dg[1] = T(0); }
    template <class T> __device__ __forceinline__ static void jac(const T* u, const T* p, T (*J)[2]) { J[0][0] = 0; J[0][1] = 1; J[1][0] = 0; J[1][1] = 0; }
    template <class T> __device__ __forceinline__ static void djac(const T* p, const T* yd, T (*J)[2]) { J[0][0] = 0; J[0][1] = 0; J[1][0] = 0; J[1][1] = 0; }
    template <class T> __device__ __forceinline__ static void dvjp_p(const T* u, const T* p, const T* ud, const T* l, T* dg) { dg[0] = T(0); dg[1] = T(0); }
};

// relaxation towards p0 (test/Callbacks2/continuous_callbacks.jl:317-324: f(D,u,p,t) = (D[1] = p[1] - u[1])); p1 enters through
// the event only (condition u - 3/4 p[1], affect u += p[2]: b200adj_set_continuous_callback_params)
struct Relax {
    static constexpr int D = 1, P = 2, M = 0;
    template <class T> __device__ __forceinline__ static void f(const T* u, const T* p, T* du) { du[0] = p[0] - u[0]; }
    template <class T> __device__ __forceinline__ static void vjp_u(const T* u, const T* p, const T* l, T* dl) { dl[0] = -l[0]; }
    template <class T> __device__ __forceinline__ static void vjp_p(const T* u, const T* p, const T* l, T* dg) { dg[0] = l[0]; dg[1] = T(0); }
    template <class T> __device__ __forceinline__ static void jac(const T* u, const T* p, T (*J)[1]) { J[0][0] = -1; }
    template <class T> __device__ __forceinline__ static void djac(const T* p, const T* yd, T (*J)[1]) { J[0][0] = 0; }
    template <class T> __device__ __forceinline__ static void dvjp_p(const T* u, const T* p, const T* ud, const T* l, T* dg) { dg[0] = T(0); dg[1] = T(0); }
};

// SDE Lotka-Volterra with diagonal noise g_i = p[4+i] u_i.  ITO selects the reference's transformed drift
// f - (dg/du)' g (src/sde_tools.jl:29-66, chosen at src/backsolve_adjoint.jl:327-345 for Ito solvers like EM).
template <bool ITO>
struct SdeLotkaVolterra {
    static constexpr int D = 2, P = 6, M = 2;
    template <class T> __device__ __forceinline__ static void f_plain(const T* u, const T* p, T* du) {
        du[0] = p[0] * u[0] - p[1] * u[0] * u[1];
        du[1] = -p[2] * u[1] + p[3] * u[0] * u[1];
    }
    template <class T> __device__ __forceinline__ static void f(const T* u, const T* p, T* du) {
        f_plain(u, p, du);
        if (ITO) { du[0] -= p[4] * p[4] * u[0]; du[1] -= p[5] * p[5] * u[1]; }
    }
    template <class T> __device__ __forceinline__ static void g(const T* u, const T* p, T* g) {
        g[0] = p[4] * u[0]; g[1] = p[5] * u[1];
    }
    template <class T> __device__ __forceinline__ static void vjp_u(const T* u, const T* p, const T* l, T* dl) {
        dl[0] = l[0] * (p[0] - p[1] * u[1]) + l[1] * p[3] * u[1];
        dl[1] = -l[0] * p[1] * u[0] + l[1] * (-p[2] + p[3] * u[0]);
        if (ITO) { dl[0] -= p[4] * p[4] * l[0]; dl[1] -= p[5] * p[5] * l[1]; }
    }
    template <class T> __device__ __forceinline__ static void vjp_p(const T* u, const T* p, const T* l, T* dg) {
        T xy = u[0] * u[1];
        dg[0] = u[0] * l[0]; dg[1] = -xy * l[0]; dg[2] = -u[1] * l[1]; dg[3] = xy * l[1];
        dg[4] = ITO ? -2 * p[4] * u[0] * l[0] : T(0);
        dg[5] = ITO ? -2 * p[5] * u[1] * l[1] : T(0);
    }
    // diagonal-noise VJPs (src/derivative_wrappers.jl:1197-1199): dl_i = l_i dg_i/du_i; the P x m block has
    // a single non-zero per column i: row 4+i, value l_i u_i.
    template <class T> __device__ __forceinline__ static void gvjp_u(const T* u, const T* p, const T* l, T* dl) {
        dl[0] = l[0] * p[4]; dl[1] = l[1] * p[5];
    }
    // dgw = sum_i (l_i dg_i/dp) w_i  for increments w[m]
    template <class T> __device__ __forceinline__ static void gvjp_p_apply(const T* u, const T* p, const T* l, const T* w, T* dgw) {
        dgw[0] = 0; dgw[1] = 0; dgw[2] = 0; dgw[3] = 0;
        dgw[4] = l[0] * u[0] * w[0]; dgw[5] = l[1] * u[1] * w[1];
    }
};

// linear SDE du_i = p0 u_i dt + p1 u_i dW_i (d = 2 instantiation used by the closed-form parity tests)
template <bool ITO>
struct SdeLinear2 {
    static constexpr int D = 2, P = 2, M = 2;
    template <class T> __device__ __forceinline__ static void f_plain(const T* u, const T* p, T* du) { du[0] = p[0] * u[0]; du[1] = p[0] * u[1]; }
    template <class T> __device__ __forceinline__ static void f(const T* u, const T* p, T* du) {
        T a = ITO ? p[0] - p[1] * p[1] : p[0];
        du[0] = a * u[0]; du[1] = a * u[1];
    }
    template <class T> __device__ __forceinline__ static void g(const T* u, const T* p, T* g) { g[0] = p[1] * u[0]; g[1] = p[1] * u[1]; }
    template <class T> __device__ __forceinline__ static void vjp_u(const T* u, const T* p, const T* l, T* dl) {
        T a = ITO ? p[0] - p[1] * p[1] : p[0];
        dl[0] = a * l[0]; dl[1] = a * l[1];
    }
    template <class T> __device__ __forceinline__ static void vjp_p(const T* u, const T* p, const T* l, T* dg) {
        T s = u[0] * l[0] + u[1] * l[1];
        dg[0] = s; dg[1] = ITO ? -2 * p[1] * s : T(0);
    }
    template <class T> __device__ __forceinline__ static void gvjp_u(const T* u, const T* p, const T* l, T* dl) { dl[0] = l[0] * p[1]; dl[1] = l[1] * p[1]; }
    template <class T> __device__ __forceinline__ static void gvjp_p_apply(const T* u, const T* p, const T* l, const T* w, T* dgw) {
        dgw[0] = 0; dgw[1] = l[0] * u[0] * w[0] + l[1] * u[1] * w[1];
    }
};

}  // namespace b200adj
