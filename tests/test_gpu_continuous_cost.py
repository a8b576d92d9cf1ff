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
