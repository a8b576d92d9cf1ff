"""a10 continuous cost functionals on the device (fixed-step Tsit5 path): dlam -= dgdu_continuous(y) at every adjoint
stage, alone or mixed with a discrete cost (test/Core7/mixed_costs.jl:19-110)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import scimlsensitivity_jl_b200 as b
from oracle import oracle as O


def _rel(a, ref):
    return np.abs(np.asarray(a) - ref).max() / max(np.abs(ref).max(), 1e-300)


@pytest.mark.parametrize("inner,name", [(b.InterpolatingAdjoint(), "interpolating"), (b.GaussAdjoint(), "gauss"),
                                        (b.BacksolveAdjoint(), "backsolve"), (b.QuadratureAdjoint(abstol=1e-10, reltol=1e-10), "quadrature")])
@pytest.mark.parametrize("mixed", [True, False])
def test_continuous_and_mixed_costs(inner, name, mixed):
    N, T, dt = 75, 2.0, 0.01
    rng = np.random.default_rng(4)
    u0 = np.ones((2, N)) * np.exp(0.1 * rng.standard_normal((2, N)))
    p = np.array([1.5, 1.0, 3.0, 1.0])
    t = np.linspace(0.0, T, 5)
    prob = b.EnsembleProblem(b.ODEProblem("lv", u0[:, 0], (0.0, T), p), u0s=u0)
    sol = b.solve(prob, b.Tsit5(dt=dt), saveat=t)
    du0, dp = b.adjoint_sensitivities(sol, b.Tsit5(dt=dt), t=t, sensealg=inner,
                                      dgdu_discrete=b.AffineCost(0.0, 1.0) if mixed else None,
                                      dgdu_continuous=b.QuadraticRunningCost(1.0, -0.3))
    ts = t if mixed else np.zeros(0)
    cfg = O.make_cfg("lv", name, "tsit5_fixed", N, ts, 0.0, T, dt=dt, cost=("affine", 0.0, 1.0), cont_cost=(1.0, -0.3),
                     quad_abstol=1e-10, quad_reltol=1e-10, ckpt_every_step=True)
    ref = O.gradient(cfg, ts, u0, p)
    assert _rel(du0, ref["du0"]) < 1e-8 and _rel(dp.ravel(), ref["dp"]) < (1e-7 if name == "quadrature" else 1e-8)
    # switching the continuous cost off again restores the purely discrete gradient
    if mixed:
        du0d, dpd = b.adjoint_sensitivities(sol, b.Tsit5(dt=dt), t=t, sensealg=inner, dgdu_discrete=b.AffineCost(0.0, 1.0))
        cfgd = O.make_cfg("lv", name, "tsit5_fixed", N, t, 0.0, T, dt=dt, cost=("affine", 0.0, 1.0), quad_abstol=1e-10, quad_reltol=1e-10, ckpt_every_step=True)
        refd = O.gradient(cfgd, t, u0, p)
        assert _rel(dpd.ravel(), refd["dp"]) < 1e-7


@pytest.mark.parametrize("inner,name", [(b.InterpolatingAdjoint(), "interpolating"), (b.GaussAdjoint(), "gauss"), (b.GaussKronrodAdjoint(), "gauss_kronrod"),
                                        (b.BacksolveAdjoint(), "backsolve"), (b.QuadratureAdjoint(abstol=1e-10, reltol=1e-10), "quadrature")])
def test_continuous_cost_adaptive_tsit5(inner, name):
    """The same functional on the error-controlled Tsit5 path (the solver the reference's mixed_costs tests use)."""
    N, T = 33, 2.0
    rng = np.random.default_rng(6)
    u0 = np.ones((2, N)) * np.exp(0.1 * rng.standard_normal((2, N)))
    p = np.array([1.5, 1.0, 3.0, 1.0])
    t = np.linspace(0.0, T, 5)
    tol = dict(abstol=1e-10, reltol=1e-10)
    prob = b.EnsembleProblem(b.ODEProblem("lv", u0[:, 0], (0.0, T), p), u0s=u0)
    sol = b.solve(prob, b.Tsit5(adaptive=True), saveat=t, **tol)
    du0, dp = b.adjoint_sensitivities(sol, b.Tsit5(adaptive=True), t=t, sensealg=inner, dgdu_discrete=b.AffineCost(0.0, 1.0),
                                      dgdu_continuous=b.QuadraticRunningCost(1.0, -0.3), **tol)
    cfg = O.make_cfg("lv", name, "tsit5_adaptive", N, t, 0.0, T, cost=("affine", 0.0, 1.0), cont_cost=(1.0, -0.3),
                     quad_abstol=1e-10, quad_reltol=1e-10, ckpt_every_step=True, **tol)
    ref = O.gradient(cfg, t, u0, p)
    assert _rel(du0, ref["du0"]) < 1e-7 and _rel(dp.ravel(), ref["dp"]) < 1e-6
    # and the oracle itself against differentiation of sum(l) + int g dt through the solver
    if name == "interpolating":
        lcfg = O.make_cfg("lv", "interpolating", "tsit5_adaptive", 1, t, 0.0, T, cost=("affine", 0.0, 1.0), cont_cost=(1.0, -0.3), abstol=1e-12, reltol=1e-12)
        e = 1e-6
        fd = np.array([(O.loss(lcfg, t, u0[:, :1], p + e * np.eye(4)[q])[0] - O.loss(lcfg, t, u0[:, :1], p - e * np.eye(4)[q])[0]) / (2 * e) for q in range(4)])
        r1 = O.gradient(O.make_cfg("lv", "interpolating", "tsit5_adaptive", 1, t, 0.0, T, cost=("affine", 0.0, 1.0), cont_cost=(1.0, -0.3), abstol=1e-12, reltol=1e-12), t, u0[:, :1], p)
        assert _rel(r1["dp"], fd) < 1e-6


# ---- the reference's mixed-cost family with a parameter part: g(u, p, t) = u1^2 + p1 (test/Core7/mixed_costs.jl:19-330) ----
MIXED_A, MIXED_E = [2.0, 0.0], [1.0, 0.0, 0.0, 0.0]          # dgdu = [2 u1, 0], dgdp = [1, 0, 0, 0]
ALGS = [(b.InterpolatingAdjoint(), "interpolating"), (b.GaussAdjoint(), "gauss"), (b.BacksolveAdjoint(), "backsolve"),
        (b.QuadratureAdjoint(abstol=1e-12, reltol=1e-12), "quadrature")]


@pytest.mark.parametrize("inner,name", ALGS)
@pytest.mark.parametrize("stepper", ["tsit5_adaptive", "tsit5_fixed"])
def test_reference_mixed_costs_discrete_with_dgdp(inner, name, stepper):
    """Discrete cost sum_k (u1(t_k)^2 + p1) at t = 1..9 with dgdu_discrete AND dgdp_discrete (mixed_costs.jl:199-330:
    save_start = save_end = false): du0, dp for every sensealg vs the oracle (itself pinned by finite differences of the
    cost, the reference's ForwardDiff comparison)."""
    N, T = 40, 10.0
    rng = np.random.default_rng(8)
    u0 = np.exp(0.05 * rng.standard_normal((2, N)))
    p = np.array([1.5, 1.0, 3.0, 1.0])
    t = np.arange(1.0, 10.0)
    adaptive = stepper == "tsit5_adaptive"
    alg = b.Tsit5(adaptive=True) if adaptive else b.Tsit5(dt=0.01)
    kw = dict(abstol=1e-10, reltol=1e-10) if adaptive else {}
    prob = b.EnsembleProblem(b.ODEProblem("lv", u0[:, 0], (0.0, T), p), u0s=u0)
    sol = b.solve(prob, alg, saveat=t, sensealg=b.B200Adjoint(inner), **kw)
    du0, dp = b.adjoint_sensitivities(sol, alg, t=t, sensealg=inner, dgdu_discrete=b.AffineCost(MIXED_A, 0.0),
                                      dgdp_discrete=b.ParamAffine(0.0, MIXED_E), checkpoints=t, **kw)
    cfg = O.make_cfg("lv", name, stepper, N, t, 0.0, T, dt=0.0 if adaptive else 0.01, cost_vec=(MIXED_A, 0.0, None, MIXED_E),
                     quad_abstol=1e-12, quad_reltol=1e-12, **kw)
    ref = O.gradient(cfg, t, u0, p)
    assert _rel(du0, ref["du0"]) < 1e-7 and _rel(dp.ravel(), ref["dp"]) < 1e-7
    # the parameter part is exactly 9 save times x N members in dp[0]
    du0b, dpb = b.adjoint_sensitivities(sol, alg, t=t, sensealg=inner, dgdu_discrete=b.AffineCost(MIXED_A, 0.0), checkpoints=t, **kw)
    assert abs((dp.ravel() - dpb.ravel())[0] - 9.0 * N) < 1e-6 and np.abs((dp.ravel() - dpb.ravel())[1:]).max() < 1e-9


@pytest.mark.parametrize("inner,name", ALGS)
@pytest.mark.parametrize("via_g", [False, True])
def test_reference_mixed_costs_continuous_with_dgdp(inner, name, via_g):
    """Continuous cost integral of (u1^2 + p1) over [0, 10] (mixed_costs.jl:19-196), given as dgdu_continuous + dgdp_continuous
    or as `g` alone; per-member parameters so that dp stays per member."""
    N, T = 24, 10.0
    rng = np.random.default_rng(9)
    u0 = np.exp(0.05 * rng.standard_normal((2, N)))
    p = np.array([1.5, 1.0, 3.0, 1.0])[:, None] * np.exp(0.02 * rng.standard_normal((4, N)))
    kw = dict(abstol=1e-10, reltol=1e-10)
    alg = b.Tsit5(adaptive=True)
    prob = b.EnsembleProblem(b.ODEProblem("lv", u0[:, 0], (0.0, T), p[:, 0]), u0s=u0, ps=p)
    sol = b.solve(prob, alg, saveat=[T], sensealg=b.B200Adjoint(inner), **kw)
    if via_g:
        du0, dp = b.adjoint_sensitivities(sol, alg, sensealg=inner, g=b.QuadraticRunningCost(MIXED_A, 0.0, None, MIXED_E), **kw)
    else:
        du0, dp = b.adjoint_sensitivities(sol, alg, sensealg=inner, dgdu_continuous=b.QuadraticRunningCost(MIXED_A, 0.0),
                                          dgdp_continuous=b.ParamAffine(0.0, MIXED_E), **kw)
    cfg = O.make_cfg("lv", name, "tsit5_adaptive", N, np.zeros(0), 0.0, T, cost=("affine", 0.0, 0.0), cont_vec=(MIXED_A, 0.0, None, MIXED_E),
                     shared_p=False, quad_abstol=1e-12, quad_reltol=1e-12, **kw)
    ref = O.gradient(cfg, np.zeros(0), u0, p)
    assert _rel(du0, ref["du0"]) < 1e-7 and _rel(dp, ref["dp"]) < 1e-7
