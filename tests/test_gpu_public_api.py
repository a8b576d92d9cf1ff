"""The reference-facing API on the device: solve(EnsembleProblem, ...), adjoint_sensitivities(...),
_concrete_solve_adjoint(...) -> (out, pullback).  Mirrors how the reference's tests drive the path
(test/Core3/adjoint.jl:53-57, test/Core1/concrete_solve_derivatives.jl:149-275, test/Core4/ensembles.jl:16-56)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import scimlsensitivity_jl_b200 as b
from oracle import oracle as O


def _rel(a, ref):
    return np.abs(np.asarray(a) - ref).max() / max(np.abs(ref).max(), 1e-300)


def _lorenz(N, seed=0):
    rng = np.random.default_rng(seed)
    return np.array([1.0, 0.0, 0.0])[:, None] + 0.1 * rng.standard_normal((3, N)), np.array([10.0, 28.0, 8.0 / 3.0])


@pytest.mark.parametrize("inner", [b.InterpolatingAdjoint(), b.GaussAdjoint(), b.BacksolveAdjoint()])
def test_adjoint_sensitivities_direct_interface(inner):
    """adjoint_sensitivities(sol, Tsit5(); t, dgdu_discrete = dg, sensealg) with dg = u - 2 (test/Core3/adjoint.jl:1169-1186)."""
    N, T, dt = 150, 2.0, 0.01
    u0, p = _lorenz(N)
    t = np.linspace(0, T, 21)
    prob = b.EnsembleProblem(b.ODEProblem("lorenz", u0[:, 0], (0.0, T), p), prob_func=lambda pr, i: u0[:, i])
    sol = b.solve(prob, b.Tsit5(dt=dt), b.EnsembleB200(), trajectories=N, saveat=t)
    assert sol.u.shape == (21, 3, N) and (sol.retcode == 0).all()
    du0, dp = b.adjoint_sensitivities(sol, b.Tsit5(dt=dt), t=t, dgdu_discrete=b.AffineCost(1.0, -2.0), sensealg=inner)
    assert dp.shape == (1, 3)                                   # dp comes back as a row (sensitivity_interface.jl:503-507)
    name = {b.InterpolatingAdjoint: "interpolating", b.GaussAdjoint: "gauss", b.BacksolveAdjoint: "backsolve"}[type(inner)]
    cfg = O.make_cfg("lorenz", name, "tsit5_fixed", N, t, 0.0, T, dt=dt, cost=("affine", 1.0, -2.0), ckpt_every_step=True)
    ref = O.gradient(cfg, t, u0, p)
    assert _rel(du0, ref["du0"]) < 1e-8 and _rel(dp.ravel(), ref["dp"]) < 1e-8
    # explicit cotangent array instead of the cost family
    du0e, dpe = b.adjoint_sensitivities(sol, b.Tsit5(dt=dt), t=t, dgdu_discrete=sol.u - 2.0, sensealg=inner)
    assert _rel(du0e, ref["du0"]) < 1e-8 and _rel(dpe.ravel(), ref["dp"]) < 1e-8
    with pytest.raises(ValueError):
        b.adjoint_sensitivities(sol, b.Tsit5(dt=dt), t=t, sensealg=inner)       # no cost given (interpolating_adjoint.jl:321-326)


@pytest.mark.parametrize("originator,arity", [(b.ChainRulesOriginator(), 6), (b.TrackerOriginator(), 5)])
def test_concrete_solve_adjoint_rrule(originator, arity):
    """(out, pullback) seam: loss = sum(sol) on LV, Delta = ones (test/Core1/concrete_solve_derivatives.jl:149-157)."""
    N, T, dt = 64, 10.0, 0.05
    rng = np.random.default_rng(1)
    u0 = np.ones((2, N)) * np.exp(0.1 * rng.standard_normal((2, N)))
    p = np.array([1.5, 1.0, 3.0, 1.0])
    prob = b.ODEProblem("lv", u0[:, 0], (0.0, T), p)
    out, pullback = b._concrete_solve_adjoint(prob, b.Tsit5(dt=dt), b.B200Adjoint(b.InterpolatingAdjoint()), u0, p, originator, saveat=0.1)
    assert out.u.shape == (101, 2, N)
    tang = pullback(np.ones_like(out.u))
    assert len(tang) == arity and all(isinstance(x, b.NoTangent) for x in tang[: arity - 3])
    du0, dp = tang[arity - 3], tang[arity - 2]
    assert du0.shape == u0.shape and dp.shape == p.shape
    saveat = np.linspace(0, T, 101)
    cfg = O.make_cfg("lv", "interpolating", "tsit5_fixed", N, saveat, 0.0, T, dt=dt, cost=("affine", 0.0, 1.0))
    ref = O.gradient(cfg, saveat, u0, p)
    assert np.abs(out.u - ref["saved"]).max() < 1e-10
    assert _rel(du0, ref["du0"]) < 1e-8 and _rel(dp, ref["dp"]) < 1e-8


def test_concrete_solve_adjoint_save_idxs_and_no_start():
    N, T, dt = 32, 1.0, 0.01
    u0, p = _lorenz(N, 3)
    prob = b.ODEProblem("lorenz", u0[:, 0], (0.0, T), p)
    out, pullback = b._concrete_solve_adjoint(prob, b.Tsit5(dt=dt), b.B200Adjoint(b.GaussAdjoint()), u0, p, None,
                                              saveat=0.1, save_idxs=[0, 2], save_start=False)
    # saveat::Number: the rrule's output keeps t0 (src/concrete_solve.jl:718-735); save_start = false => no_start (:962)
    assert out.u.shape == (11, 2, N) and out.t[0] == 0.0
    tang = pullback(np.ones_like(out.u))
    saveat = np.linspace(0.0, T, 11)
    dL = np.zeros((11, 3, N)); dL[:, [0, 2], :] = 1.0
    cfg = O.make_cfg("lorenz", "gauss", "tsit5_fixed", N, saveat, 0.0, T, dt=dt, no_start=True)
    ref = O.gradient(cfg, saveat, u0, p, dLdu=dL)
    assert _rel(tang[3], ref["du0"]) < 1e-8 and _rel(tang[4], ref["dp"]) < 1e-8


def test_sde_public_api_backsolve():
    """test/Core1/concrete_solve_derivatives.jl:736-787: SDE-LV diag noise, p=[1.5,1,3,1,0.1,0.1], dt=0.01, BacksolveAdjoint."""
    N, T, dt = 96, 1.0, 0.01
    u0 = np.ones((2, N)); p = np.array([1.5, 1.0, 3.0, 1.0, 0.1, 0.1])
    for alg, st in ((b.EulerHeun(dt=dt), "euler_heun"), (b.EM(dt=dt), "em")):
        prob = b.EnsembleProblem(b.SDEProblem("sde_lv", u0[:, 0], (0.0, T), p, seed=100), u0s=u0)
        sol = b.solve(prob, alg, saveat=0.01, sensealg=b.B200Adjoint(b.BacksolveAdjoint()))
        dW = sol.engine.noise()
        du0, dp = b.adjoint_sensitivities(sol, alg, t=sol.t, dgdu_discrete=b.AffineCost(0.0, 1.0), sensealg=b.BacksolveAdjoint(), checkpoints=sol.t)
        cfg = O.make_cfg("sde_lv", "backsolve", st, N, sol.t, 0.0, T, dt=dt, cost=("affine", 0.0, 1.0))
        ref = O.gradient(cfg, sol.t, u0, p, dW=dW)
        assert _rel(du0, ref["du0"]) < 1e-9 and _rel(dp.ravel(), ref["dp"]) < 1e-9


def test_unsupported_configs_fail_loudly():
    u0, p = _lorenz(8)
    prob = b.EnsembleProblem(b.ODEProblem("lorenz", u0[:, 0], (0.0, 1.0), p), u0s=u0)
    sol = b.solve(prob, b.Tsit5(dt=0.01), saveat=[0.005, 0.5])   # off-grid save times: the dense per-member framework takes over
    assert sol.u.shape == (2, 3, 8) and int(np.asarray(sol.retcode).sum()) == 0
    with pytest.raises(b.B200AdjError) as ei:                   # ... which is an F64 path
        b.DeviceEnsemble("lorenz", "gauss", "tsit5_fixed", 8, [0.005, 0.5], (0.0, 1.0), 0.01, dtype="f32")
    assert ei.value.code == -2
    with pytest.raises(b.B200AdjError):                          # horizon not a whole number of steps
        b.solve(prob, b.Tsit5(dt=0.03), saveat=[0.3])
    mprob = b.EnsembleProblem(b.ODEProblem("mlp", np.zeros(2), (0.0, 1.0), np.zeros(4482)), u0s=np.zeros((2, 8)))
    msol = b.solve(mprob, b.Tsit5(dt=0.05), saveat=0.5)
    with pytest.raises(b.B200AdjError):                          # QuadratureAdjoint is not built for the MLP family
        b.adjoint_sensitivities(msol, b.Tsit5(dt=0.05), t=msol.t, dgdu_discrete=b.AffineCost(1.0, 0.0), sensealg=b.QuadratureAdjoint())
