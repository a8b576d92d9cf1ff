// lv_clone_family.cuh -- the Lotka-Volterra family written as a USER plug-in (tests: a plug-in family must reproduce the built-in
// one bit for bit, and carries jac / djac / dvjp_p for the Rosenbrock23 kernels).
#pragma once

struct LvClone {
    static constexpr int D = 2, P = 4, M = 0;
    template <class T> __device__ __forceinline__ static void f(const T* u, const T* p, T* du) {
        du[0] = p[0] * u[0] - p[1] * u[0] * u[1];
        du[1] = -p[2] * u[1] + p[3] * u[0] * u[1];
    }
    template <class T> __device__ __forceinline__ static void vjp_u(const T* u, const T* p, const T* l, T* dl) {
        dl[0] = l[0] * (p[0] - p[1] * u[1]) + l[1] * p[3] * u[1];
        dl[1] = -l[0] * p[1] * u[0] + l[1] * (-p[2] + p[3] * u[0]);
    }
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
    template <class T> __device__ __forceinline__ static void dvjp_p(const T* u, const T* p, const T* ud, const T* l, T* dg) {
        const T dxy = ud[0] * u[1] + u[0] * ud[1];
        dg[0] = ud[0] * l[0]; dg[1] = -dxy * l[0]; dg[2] = -ud[1] * l[1]; dg[3] = dxy * l[1];
    }
};
