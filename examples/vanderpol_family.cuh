// vanderpol_family.cuh -- a USER RHS family for libb200adj.so (not part of the library): the van der Pol oscillator
//   u1' = u2,  u2' = p1 (1 - u1^2) u2 - p2 u1        d = 2, P = 2
// with its hand-written VJPs, in the shape of csrc/families.cuh.  The same role as the user-supplied
// ODEFunction(f; vjp, vjp_p, jac, paramjac) of the reference (test/Core3/user_vjp.jl:14-38).
// Build + register:  python -m scimlsensitivity_jl_b200.family_plugin examples/vanderpol_family.cuh VanDerPol vanderpol --jac
#pragma once

struct VanDerPol {
    static constexpr int D = 2, P = 2, M = 0;
    template <class T> __device__ __forceinline__ static void f(const T* u, const T* p, T* du) {
        du[0] = u[1];
        du[1] = p[0] * (1 - u[0] * u[0]) * u[1] - p[1] * u[0];
    }
    // dl = (df/du)' l
    template <class T> __device__ __forceinline__ static void vjp_u(const T* u, const T* p, const T* l, T* dl) {
        dl[0] = (-2 * p[0] * u[0] * u[1] - p[1]) * l[1];
        dl[1] = l[0] + p[0] * (1 - u[0] * u[0]) * l[1];
    }
    // dg = (df/dp)' l
    template <class T> __device__ __forceinline__ static void vjp_p(const T* u, const T* p, const T* l, T* dg) {
        dg[0] = (1 - u[0] * u[0]) * u[1] * l[1];
        dg[1] = -u[0] * l[1];
    }
    // Rosenbrock23 support: J[i][j] = df_i/du_j, its directional derivative along ud, and that of vjp_p
    template <class T> __device__ __forceinline__ static void jac(const T* u, const T* p, T (*J)[2]) {
        J[0][0] = 0;                                J[0][1] = 1;
        J[1][0] = -2 * p[0] * u[0] * u[1] - p[1];   J[1][1] = p[0] * (1 - u[0] * u[0]);
    }
    // NOTE: the Hessian of this family is not constant -- djac needs the point u, which the library passes only for
    // quadratic families; van der Pol therefore builds WITHOUT --jac (Tsit5 steppers).  Kept here as documentation of the
    // interface for quadratic families (see csrc/families.cuh::LotkaVolterra::djac / dvjp_p).
};
