"""State-dependent events (ContinuousCallback) on the adaptive Tsit5 path: every ensemble member finds its own event times on
the device (root finding on the dense output), the reverse kernel applies the implicit event-time correction
(src/callback_tracking.jl:232-480).  Device vs oracle, the closed form of one bounce, and finite differences of the device's
own forward pass (the reference's checks: docs/src/examples/hybrid_jump/bouncing_ball.md,
test/Callbacks1/continuous_callbacks.jl -- every sensealg against differentiation through the solver)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import scimlsensitivity_jl_b200 as b
from oracle import oracle as O

pytestmark = pytest.mark.gpu

BALL = b.ContinuousCallback(idx=0, level=0.0, direction=-1, p_comp=1, p_param=1, p_sign=-1.0, max_events=32)
BALL_ORACLE = dict(idx=0, level=0.0, direction=-1, pcomp=1, pparam=1, psign=-1.0)
TOL = dict(abstol=1e-10, reltol=1e-10)


def _rel(a, ref):
    return float(np.max(np.abs(np.asarray(a) - ref)) / (np.max(np.abs(ref)) + 1e-300))


def _ball_inputs(N, shared_p, seed=3):
    rng = np.random.default_rng(seed)
    u0 = np.stack([50.0 + 5.0 * rng.standard_normal(N), 0.5 * rng.standard_normal(N)])      # heights 35..65: different bounce times
    p = np.array([9.8, 0.8]) if shared_p else np.stack([9.8 + 0.3 * rng.standard_normal(N), 0.8 + 0.03 * rng.standard_normal(N)])
    return u0, p


@pytest.mark.parametrize("shared_p", [True, False])
@pytest.mark.parametrize("sa,every", [("interpolating", False), ("gauss", False), ("gauss_kronrod", False), ("backsolve", True), ("backsolve", False)])
def test_bouncing_ball_device_vs_oracle(sa, every, shared_p):
    N = 48
    u0, p = _ball_inputs(N, shared_p)
    t = np.linspace(0.5, 15.0, 30)
    eng = b.DeviceEnsemble("ball", sa, "tsit5_adaptive", N, t, (0.0, 15.0), 0.0, cost=b.AffineCost(1.0, 0.0), shared_p=shared_p,
                           ckpt_every_step=every, **TOL)
    eng.set_continuous_callback(BALL)
    saved, status = eng.forward(u0, p)
    du0, dp = eng.reverse()
    counts, times = eng.event_times()
    cfg = O.make_cfg("ball", sa, "tsit5_adaptive", N, t, 0.0, 15.0, cost=("affine", 1.0, 0.0), shared_p=shared_p, ckpt_every_step=every,
                     crossing=BALL_ORACLE, **TOL)
    ref = O.gradient(cfg, t, u0, p)
    assert (np.asarray(status) == 0).all()
    assert counts.min() >= 3 and counts.max() <= 8 and len(set(counts.tolist())) > 1       # members bounce at their own times
    for i in range(N):
        assert np.all(np.diff(times[:counts[i], i]) > 0)
    assert _rel(saved, ref["saved"]) < 1e-9
    assert _rel(du0, ref["du0"]) < 1e-7 and _rel(dp, ref["dp"]) < 1e-7


def test_one_bounce_closed_form():
    """x0 dropped from rest: impact at t* = sqrt(2 x0 / g) with speed w = g t*, then x(T) = e w s - g s^2 / 2, s = T - t*.
    L = x(T):  dL/dx0 = e g s / w - (e w - g s) / w ... evaluated below by differentiating the closed form."""
    N = 8
    x0 = np.linspace(8.0, 15.0, N)
    g, e, T = 9.8, 0.8, 2.5

    def xT(x0, g, e):
        ts = np.sqrt(2 * x0 / g); w = g * ts; s = T - ts
        return e * w * s - 0.5 * g * s * s

    h = 1e-6
    d_x0 = (xT(x0 + h, g, e) - xT(x0 - h, g, e)) / (2 * h)
    d_g = (xT(x0, g + h, e) - xT(x0, g - h, e)) / (2 * h)
    d_e = (xT(x0, g, e + h) - xT(x0, g, e - h)) / (2 * h)
    u0 = np.stack([x0, np.zeros(N)])
    for sa in ("interpolating", "gauss", "gauss_kronrod", "backsolve"):
        eng = b.DeviceEnsemble("ball", sa, "tsit5_adaptive", N, [T], (0.0, T), 0.0, shared_p=False, **TOL)
        eng.set_continuous_callback(BALL)
        saved, status = eng.forward(u0, np.tile(np.array([[g], [e]]), (1, N)))
        dL = np.zeros((1, 2, N)); dL[0, 0, :] = 1.0
        du0, dp = eng.reverse(dL)
        counts, times = eng.event_times()
        assert (counts == 1).all()
        assert np.allclose(times[0], np.sqrt(2 * x0 / g), rtol=1e-9)
        assert np.allclose(np.asarray(saved)[0, 0], xT(x0, g, e), rtol=1e-8)
        assert np.allclose(np.asarray(du0)[0], d_x0, rtol=1e-6), sa
        assert np.allclose(np.asarray(dp)[0], d_g, rtol=1e-6) and np.allclose(np.asarray(dp)[1], d_e, rtol=1e-6), sa


def test_level_crossing_affine_affect_lv_vs_oracle():
    """a different condition / affect of the family: prey above the level 1.6 is harvested (u[0] <- 0.6 u[0] + 0.05), both
    directions of the Lotka-Volterra cycle cross the level, only the upward one fires."""
    N = 32
    rng = np.random.default_rng(5)
    u0 = 1.0 + 0.05 * rng.standard_normal((2, N))
    p = np.array([1.5, 1.0, 3.0, 1.0])
    t = np.arange(0.5, 10.0001, 0.5)
    cb = b.ContinuousCallback(idx=0, level=1.6, direction=+1, scale=[0.6, 1.0], shift=[0.05, 0.0], max_events=32)
    cr = dict(idx=0, level=1.6, direction=+1, scale=[0.6, 1.0], shift=[0.05, 0.0])
    for sa in ("interpolating", "gauss", "backsolve"):
        eng = b.DeviceEnsemble("lv", sa, "tsit5_adaptive", N, t, (0.0, 10.0), 0.0, cost=b.AffineCost(1.0, -1.0), **TOL)
        eng.set_continuous_callback(cb)
        saved, status = eng.forward(u0, p)
        du0, dp = eng.reverse()
        counts, _ = eng.event_times()
        cfg = O.make_cfg("lv", sa, "tsit5_adaptive", N, t, 0.0, 10.0, cost=("affine", 1.0, -1.0), crossing=cr, **TOL)
        ref = O.gradient(cfg, t, u0, p)
        assert (np.asarray(status) == 0).all() and counts.min() >= 1
        assert _rel(saved, ref["saved"]) < 1e-9
        assert _rel(du0, ref["du0"]) < 1e-6 and _rel(dp, ref["dp"]) < 1e-6, sa


@pytest.mark.parametrize("shared_p", [True, False])
@pytest.mark.parametrize("sa", ["interpolating", "gauss", "gauss_kronrod", "backsolve"])
def test_non_linear_affect_of_the_reference_tests_vs_oracle(sa, shared_p):
    """"u[1] += 3; u[2] = u[2]^2" at the impact, MSE loss sum((1 - u)^2) / 2 at saveat 0.5 (test/Callbacks2/continuous_callbacks.jl:
    240-251; tolerances 1e-12, :6-8): the Jacobian of the quadratic affect takes the place of the affine scale in the reverse
    kernel.  Members start from their own heights, so the impacts happen at their own times."""
    N = 40
    rng = np.random.default_rng(21)
    u0 = np.stack([5.0 + rng.random(N), 0.2 * rng.standard_normal(N)])
    p = np.array([9.8, 0.8]) if shared_p else np.stack([9.8 + 0.2 * rng.standard_normal(N), np.full(N, 0.8)])
    ts = np.arange(0.0, 2.5 + 1e-9, 0.5)
    kw = dict(abstol=1e-12, reltol=1e-12)
    cb = b.ContinuousCallback(idx=0, direction=-1, shift=[3.0, 0.0], sq_comp=1, sq_coef=1.0, max_events=8)
    eng = b.DeviceEnsemble("ball", sa, "tsit5_adaptive", N, ts, (0.0, 2.5), 0.0, cost=b.AffineCost(1.0, -1.0), shared_p=shared_p,
                           ckpt_every_step=True, **kw)
    eng.set_continuous_callback(cb)
    saved, status = eng.forward(u0, p)
    du0, dp = eng.reverse()
    counts, times = eng.event_times()
    cfg = O.make_cfg("ball", sa, "tsit5_adaptive", N, ts, 0.0, 2.5, cost=("affine", 1.0, -1.0), shared_p=shared_p, ckpt_every_step=True,
                     crossing=dict(idx=0, direction=-1, shift=[3.0, 0.0], qcomp=1, qcoef=1.0), **kw)
    ref = O.gradient(cfg, ts, u0, p)
    assert (np.asarray(status) == 0).all() and (counts == 1).all()
    assert _rel(saved, ref["saved"]) < 1e-9
    assert _rel(du0, ref["du0"]) < 1e-7 and _rel(dp, ref["dp"]) < 1e-7


def test_public_api_bouncing_ball_vs_finite_differences():
    """solve(EnsembleProblem(ODEProblem(ball; callback = ContinuousCallback(...))), Tsit5(), EnsembleB200(); sensealg) and its
    pullback against central differences of the device's own loss (bouncing_ball.md differentiates the final position)."""
    N = 6
    u0, p = _ball_inputs(N, True, seed=9)
    t = np.linspace(1.0, 12.0, 12)
    prob = b.ODEProblem("ball", u0[:, 0], (0.0, 12.0), p, callback=BALL)
    ens = b.EnsembleProblem(prob, u0s=u0)

    def loss(pp):
        sol = b.solve(b.EnsembleProblem(b.ODEProblem("ball", u0[:, 0], (0.0, 12.0), pp, callback=BALL), u0s=u0), b.Tsit5(adaptive=True),
                      b.EnsembleB200(), trajectories=N, saveat=t, abstol=1e-11, reltol=1e-11)
        return float(np.sum(np.asarray(sol.u) ** 2) / 2)

    out, pullback = b._concrete_solve_adjoint(ens, b.Tsit5(adaptive=True), b.B200Adjoint(b.GaussAdjoint()), u0, p, None, saveat=t,
                                              abstol=1e-11, reltol=1e-11)
    tang = pullback(np.asarray(out.u))                         # L = sum(u^2) / 2
    dp = np.asarray(tang[4]).reshape(-1)
    fd = np.zeros(2)
    for q in range(2):
        h = 1e-5 * p[q]
        e = np.zeros(2); e[q] = h
        fd[q] = (loss(p + e) - loss(p - e)) / (2 * h)
    assert np.allclose(dp, fd, rtol=2e-5), (dp, fd)


def test_refusals():
    t = np.linspace(0.5, 5.0, 10)
    eng = b.DeviceEnsemble("ball", "quadrature", "tsit5_adaptive", 4, t, (0.0, 5.0), 0.0, **TOL)
    with pytest.raises(b.B200AdjError) as ei:
        eng.set_continuous_callback(BALL)
    assert ei.value.code == -2                      # UNSUPPORTED: the reference's QuadratureAdjoint has no callback support either
    eng = b.DeviceEnsemble("lv", "gauss", "tsit5_fixed", 4, t, (0.0, 5.0), 0.01)
    with pytest.raises(b.B200AdjError):
        eng.set_continuous_callback(b.ContinuousCallback(idx=0, level=1.5))
    eng = b.DeviceEnsemble("ball", "gauss", "tsit5_adaptive", 4, t, (0.0, 5.0), 0.0, **TOL)
    with pytest.raises(b.B200AdjError):
        eng.set_continuous_callback(b.ContinuousCallback(idx=2))
    # capacity of the event list: status 3
    eng.set_continuous_callback(b.ContinuousCallback(idx=0, direction=-1, p_comp=1, p_param=1, p_sign=-1.0, max_events=1))
    u0 = np.tile(np.array([[5.0], [0.0]]), (1, 4))
    _, status = eng.forward(u0, np.array([9.8, 0.8]))
    assert (np.asarray(status) == 3).all()
