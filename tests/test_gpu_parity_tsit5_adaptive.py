"""Error-controlled Tsit5 on the device (BASELINE config C1: Lotka-Volterra d=2, single trajectory, InterpolatingAdjoint,
Tsit5, u0=[1,1], p=[1.5,1,3,1], T=10, saveat=0.1, loss=sum(sol); test/Core1/concrete_solve_derivatives.jl:106-157) and
ensembles of it, against the oracle's adaptive Tsit5 (PI controller) and against differentiation through the solver."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import scimlsensitivity_jl_b200 as b
from oracle import oracle as O


def _rel(a, ref):
    return np.abs(np.asarray(a) - ref).max() / max(np.abs(ref).max(), 1e-300)


@pytest.mark.parametrize("sensealg", ["interpolating", "gauss", "quadrature"])
@pytest.mark.parametrize("N", [1, 100])
def test_c1_lotka_volterra_adaptive(sensealg, N):
    T = 10.0
    saveat = np.linspace(0.0, T, 101)
    rng = np.random.default_rng(0)
    u0 = np.ones((2, N)) * (np.exp(0.1 * rng.standard_normal((2, N))) if N > 1 else 1.0)
    p = np.array([1.5, 1.0, 3.0, 1.0])
    tol = dict(abstol=1e-10, reltol=1e-10)
    cfg = O.make_cfg("lv", sensealg, "tsit5_adaptive", N, saveat, 0.0, T, cost=("affine", 0.0, 1.0), quad_abstol=1e-10, quad_reltol=1e-10, **tol)
    ref = O.gradient(cfg, saveat, u0, p)
    eng = b.DeviceEnsemble("lv", sensealg, "tsit5_adaptive", N, saveat, (0.0, T), 0.0, cost=b.AffineCost(0.0, 1.0),
                           quad_abstol=1e-10, quad_reltol=1e-10, max_steps=8192, **tol)
    saved, status = eng.forward(u0, p)
    assert (status == 0).all()
    fsteps, _ = eng.step_counts()
    assert np.array_equal(fsteps, ref["steps"])                       # same accept/reject sequence as the oracle
    assert np.abs(saved - ref["saved"]).max() < 1e-11
    du0, dp = eng.reverse()
    assert _rel(du0, ref["du0"]) < 1e-8 and _rel(dp, ref["dp"]) < (1e-7 if sensealg == "quadrature" else 1e-8)
    if N == 1:
        # C1's own bar: the gradient through the solver (finite differences of the oracle's forward solve), <= 1e-8 rel
        cfgl = O.make_cfg("lv", "interpolating", "tsit5_adaptive", 1, saveat, 0.0, T, abstol=1e-13, reltol=1e-13, cost=("affine", 0.0, 1.0))
        g = np.zeros(4)
        for i in range(4):
            e = np.zeros(4); e[i] = 1e-5
            L = lambda q: O.loss(cfgl, saveat, u0, q)[0]
            g[i] = (-L(p + 2 * e) + 8 * L(p + e) - 8 * L(p - e) + L(p - 2 * e)) / (12e-5)
        assert _rel(dp, g) < 1e-8
        assert abs(dp[0] - 8.3053) < 5e-4                             # test/Core6/forward_prob_kwargs.jl:28-30
    eng.close()


def test_lorenz_checkpointed_backsolve_equals_interpolating():
    """The reference's own Lorenz check (test/Core3/adjoint.jl:1157-1241): adaptive Tsit5 forward solve, dg = u - 2 at
    0:0.1:10, BacksolveAdjoint (checkpoints = sol.t, and = every 10th ... here: the save times) == InterpolatingAdjoint
    at rtol 1e-5 / 1e-4, on the device and against the oracle."""
    N, T = 16, 10.0
    rng = np.random.default_rng(2)
    u0 = np.array([1.0, 0.0, 0.0])[:, None] + 0.001 * rng.standard_normal((3, N))
    p = np.array([10.0, 28.0, 8.0 / 3.0])
    t = np.linspace(0.0, T, 101)
    tol = dict(abstol=1e-9, reltol=1e-9)
    res = {}
    for sa, every in (("interpolating", False), ("backsolve", True), ("backsolve", False)):
        eng = b.DeviceEnsemble("lorenz", sa, "tsit5_adaptive", N, t, (0.0, T), 0.0, cost=b.AffineCost(1.0, -2.0), ckpt_every_step=every,
                               max_steps=16384, **tol)
        eng.forward(u0, p)
        du0, dp = eng.reverse()
        cfg = O.make_cfg("lorenz", sa, "tsit5_adaptive", N, t, 0.0, T, cost=("affine", 1.0, -2.0), ckpt_every_step=every, **tol)
        ref = O.gradient(cfg, t, u0, p)
        assert _rel(du0, ref["du0"]) < 1e-6 and _rel(dp, ref["dp"]) < 1e-6, (sa, every)
        res[(sa, every)] = (du0, dp)
        eng.close()
    assert _rel(res[("backsolve", True)][1], res[("interpolating", False)][1]) < 1e-5
    assert _rel(res[("backsolve", True)][0], res[("interpolating", False)][0]) < 1e-5
    assert _rel(res[("backsolve", False)][1], res[("interpolating", False)][1]) < 1e-4


def test_adaptive_tsit5_public_api_lorenz():
    """test/Core3/adjoint.jl:1157-1241 (Lorenz, adaptive Tsit5, dg = u - 2 at 0:0.1:10): Interpolating == Gauss == Quadrature."""
    N, T = 32, 10.0
    rng = np.random.default_rng(1)
    u0 = np.array([1.0, 0.0, 0.0])[:, None] + 0.01 * rng.standard_normal((3, N))
    p = np.array([10.0, 28.0, 8.0 / 3.0])
    t = np.linspace(0.0, T, 101)
    prob = b.EnsembleProblem(b.ODEProblem("lorenz", u0[:, 0], (0.0, T), p), u0s=u0)
    alg = b.Tsit5(adaptive=True)
    sol = b.solve(prob, alg, saveat=t, abstol=1e-12, reltol=1e-12, maxiters=16384)
    res = {}
    for inner, name in ((b.InterpolatingAdjoint(), "interpolating"), (b.GaussAdjoint(), "gauss"), (b.QuadratureAdjoint(abstol=1e-9, reltol=1e-9), "quadrature")):
        du0, dp = b.adjoint_sensitivities(sol, alg, t=t, dgdu_discrete=b.AffineCost(1.0, -2.0), sensealg=inner, abstol=1e-12, reltol=1e-12)
        cfg = O.make_cfg("lorenz", name, "tsit5_adaptive", N, t, 0.0, T, abstol=1e-12, reltol=1e-12, cost=("affine", 1.0, -2.0), quad_abstol=1e-9, quad_reltol=1e-9)
        ref = O.gradient(cfg, t, u0, p)
        assert _rel(du0, ref["du0"]) < 1e-6 and _rel(dp.ravel(), ref["dp"]) < 1e-6          # chaotic: rounding amplified ~1e4
        res[name] = dp.ravel()
    assert _rel(res["gauss"], res["interpolating"]) < 1e-6 and _rel(res["quadrature"], res["interpolating"]) < 1e-6


@pytest.mark.parametrize("family,stepper", [("lv", "tsit5_adaptive"), ("lorenz", "tsit5_adaptive"), ("robertson", "rosenbrock23")])
@pytest.mark.parametrize("shared_p", [True, False])
def test_gauss_kronrod_adjoint(family, stepper, shared_p):
    """GaussKronrodAdjoint (src/gauss_adjoint.jl:820-825): error-controlled G3/K7 (Tsit5) or G1/K3 (Rosenbrock23)
    quadrature of every accepted reverse step, explicit-stack bisection on the device == the oracle's recursion."""
    N = 37
    rng = np.random.default_rng(4)
    if family == "lv":
        u0 = 1.0 + 0.05 * rng.standard_normal((2, N)); p0 = np.array([1.5, 1.0, 3.0, 1.0]); T = 10.0; t = np.arange(0.0, 10.0001, 0.5)
    elif family == "lorenz":
        u0 = np.array([1.0, 0.0, 0.0])[:, None] + 0.01 * rng.standard_normal((3, N)); p0 = np.array([10.0, 28.0, 8.0 / 3.0]); T = 2.0; t = np.linspace(0.0, T, 21)
    else:
        u0 = np.repeat(np.array([[1.0], [0.0], [0.0]]), N, 1); p0 = np.array([0.04, 3e7, 1e4]); T = 100.0; t = np.logspace(-2, 2, 10); t[-1] = T
    p = p0 if shared_p else p0[:, None] * np.exp(0.02 * rng.standard_normal((len(p0), N)))
    tol = dict(abstol=1e-8, reltol=1e-8)
    eng = b.DeviceEnsemble(family, "gauss_kronrod", stepper, N, t, (0.0, T), 0.0, shared_p=shared_p, cost=b.AffineCost(1.0, -0.5), max_steps=8192, **tol)
    eng.forward(u0, p)
    du0, dp = eng.reverse()
    ref = O.gradient(O.make_cfg(family, "gauss_kronrod", stepper, N, t, 0.0, T, cost=("affine", 1.0, -0.5), shared_p=shared_p, **tol), t, u0, p)
    assert _rel(du0, ref["du0"]) < 1e-7
    assert _rel(dp, ref["dp"]) < 1e-6          # the bisection decisions sit on a 1e-7 threshold: a member may split differently
    # re-target the same forward pass: Gauss <-> GaussKronrod
    eng.set_reverse("gauss", cost=b.AffineCost(1.0, -0.5), t=t)
    _, dpg = eng.reverse()
    assert _rel(dpg, dp) < 1e-3
    eng.close()
    with pytest.raises(Exception):           # GaussKronrodAdjoint is an F64 path
        b.DeviceEnsemble("lv", "gauss_kronrod", "tsit5_fixed", 8, t, (0.0, T), 0.01, dtype="f32")


@pytest.mark.parametrize("family", ["lv", "lorenz"])
@pytest.mark.parametrize("shared_p", [True, False])
def test_gauss_kronrod_adjoint_fixed_step(family, shared_p):
    """GaussKronrodAdjoint with fixed-step Tsit5: the G3/K7 rule with bisection inside tsit5_reverse_kernel<SA_GK> (general-theta
    dense outputs of the adjoint and the forward step) vs the oracle; re-targeting the same handle to GaussAdjoint."""
    N = 150
    rng = np.random.default_rng(31)
    if family == "lv":
        T, dt, u0, p = 4.0, 0.05, np.exp(0.05 * rng.standard_normal((2, N))), np.array([1.5, 1.0, 3.0, 1.0])
    else:
        T, dt, u0, p = 1.0, 0.01, np.array([1.0, 0.0, 0.0])[:, None] + 0.1 * rng.standard_normal((3, N)), np.array([10.0, 28.0, 8.0 / 3.0])
    if not shared_p:
        p = p[:, None] * np.exp(0.02 * rng.standard_normal((len(p), N)))
    t = np.linspace(0.0, T, 11)
    eng = b.DeviceEnsemble(family, "gauss_kronrod", "tsit5_fixed", N, t, (0.0, T), dt, shared_p=shared_p, cost=b.AffineCost(1.0, -0.5))
    eng.forward(u0, p)
    du0, dp = eng.reverse()
    ref = O.gradient(O.make_cfg(family, "gauss_kronrod", "tsit5_fixed", N, t, 0.0, T, dt=dt, cost=("affine", 1.0, -0.5), shared_p=shared_p), t, u0, p)
    assert _rel(du0, ref["du0"]) < 1e-9 and _rel(dp, ref["dp"]) < 1e-6
    eng.set_reverse("gauss", cost=b.AffineCost(1.0, -0.5), t=t)
    _, dpg = eng.reverse()
    assert _rel(dpg, dp) < 1e-3
    eng.close()
