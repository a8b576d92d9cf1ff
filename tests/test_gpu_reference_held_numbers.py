"""Device-side twins of tests/test_reference_held_numbers.py: numbers the REFERENCE holds (its documentation's printed optimum,
its tests' literal fixtures), reproduced by the CUDA path through the C ABI -- with the oracle nowhere in the comparison.

* docs/src/examples/hybrid_jump/bouncing_ball.md:60  res.u = [0.866554105436901]: the stationary point of (x(15) - 20)^2 over
  the restitution coefficient.  Every ensemble member runs its own secant iteration on the DEVICE gradient (per-member p) from
  its own bracket; all of them land on the printed optimum.
* test/Callbacks2/continuous_vs_discrete.jl:19-21  tstop / vbefore / vafter of the first impact.
* test/Callbacks2/continuous_callbacks.jl:343  gND = [0.9999546000702386, 0.00018159971904994378] (condition u - 3/4 p[1],
  affect u += p[2]); the reference demands rtol 1e-10 of every sensealg (:344-359)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import scimlsensitivity_jl_b200 as b
from oracle import oracle as O

pytestmark = pytest.mark.gpu

SENSEALGS = ("interpolating", "gauss", "gauss_kronrod", "backsolve")
BALL = b.ContinuousCallback(idx=0, level=0.0, direction=-1, p_comp=1, p_param=1, p_sign=-1.0, max_events=16)
DOCS_OPTIMUM = 0.866554105436901
TSTOP, VBEFORE, VAFTER = 3.1943828249997, -31.30495168499705, 25.04396134799764
GND = np.array([0.9999546000702386, 0.00018159971904994378])
RELAX_CB = b.ContinuousCallback(idx=0, level=0.0, direction=0, level_param=0, level_coef=0.75, add_comp=0, add_param=1, add_coef=1.0,
                                max_events=4)
RELAX_ORACLE = dict(idx=0, level=0.0, direction=0, lparam=0, lcoef=0.75, acomp=0, aparam=1, acoef=1.0)


@pytest.mark.parametrize("tol", [(1e-6, 1e-3), (1e-10, 1e-10)], ids=["docs_default_tolerances", "tight"])
@pytest.mark.parametrize("sa", SENSEALGS)
def test_bouncing_ball_docs_optimum_from_the_device_gradient(sa, tol):
    N = 64
    ts = np.array([15.0])
    u0 = np.tile(np.array([[50.0], [0.0]]), (1, N))
    eng = b.DeviceEnsemble("ball", sa, "tsit5_adaptive", N, ts, (0.0, 15.0), 0.0, shared_p=False, ckpt_every_step=True,
                           abstol=tol[0], reltol=tol[1])
    eng.set_continuous_callback(BALL)

    def grad(theta):
        saved, status = eng.forward(u0, np.stack([np.full(N, 9.8), theta]))
        assert (np.asarray(status) == 0).all()
        x = np.asarray(saved)[0, 0]
        dL = np.zeros((1, 2, N)); dL[0, 0] = 2.0 * (x - 20.0)
        du0, dp = eng.reverse(dL)
        return x, np.asarray(dp)[1].copy()

    x, g = grad(np.where(np.arange(N) % 2 == 0, DOCS_OPTIMUM, 0.8))
    assert np.max(np.abs(x[0::2] - 19.367902815)) < 1e-8
    assert np.max(np.abs(g[0::2])) < 1e-9 * np.min(np.abs(g[1::2]))
    a = 0.862 + 0.004 * np.arange(N) / N
    bb = a + 0.003
    ga, gb = grad(a)[1], grad(bb)[1]
    for _ in range(12):
        den = np.where(gb == ga, 1.0, gb - ga)
        c = np.where(gb == ga, bb, bb - gb * (bb - a) / den)
        a, ga, bb, gb = bb, gb, c, grad(c)[1]
    assert np.max(np.abs(bb - DOCS_OPTIMUM)) < 1e-11, bb - DOCS_OPTIMUM


@pytest.mark.parametrize("tol", [(1e-12, 1e-12), (1e-6, 1e-3)])
def test_first_impact_event_location_on_the_device(tol):
    N = 32
    eps = 1e-3                                     # free flight either side of the impact: v(tstop -+ eps) = v-+ +- g eps
    ts = np.array([TSTOP - eps, TSTOP + eps, 5.0])
    eng = b.DeviceEnsemble("ball", "backsolve", "tsit5_adaptive", N, ts, (0.0, 5.0), 0.0, shared_p=True, abstol=tol[0], reltol=tol[1])
    eng.set_continuous_callback(BALL)
    saved, status = eng.forward(np.tile(np.array([[50.0], [0.0]]), (1, N)), np.array([9.8, 0.8]))
    counts, times = eng.event_times()
    saved = np.asarray(saved)
    assert (np.asarray(status) == 0).all() and (counts == 1).all()
    assert np.max(np.abs(times[0] - TSTOP)) < 1e-12
    assert np.max(np.abs(saved[0, 1] - (VBEFORE + 9.8 * eps))) < 1e-10
    assert np.max(np.abs(saved[1, 1] - (VAFTER - 9.8 * eps))) < 1e-10


@pytest.mark.parametrize("shared_p", [True, False])
@pytest.mark.parametrize("sa", SENSEALGS)
def test_parameter_dependent_condition_reproduces_gND_on_the_device(sa, shared_p):
    """u(10) = p1 + (p2 - p1/4) 4 e^-10 for u0 = 0 whatever p1, p2 > 0 are: every member's gradient is the printed gND."""
    N = 40
    ts = np.array([10.0])
    rng = np.random.default_rng(5)
    p = np.array([100.0, 50.0]) if shared_p else np.stack([100.0 + 20.0 * rng.random(N), 50.0 + 10.0 * rng.random(N)])
    p1 = p[0] if not shared_p else np.full(N, 100.0)
    eng = b.DeviceEnsemble("relax", sa, "tsit5_adaptive", N, ts, (0.0, 10.0), 0.0, shared_p=shared_p, ckpt_every_step=True,
                           abstol=1e-14, reltol=1e-14)
    eng.set_continuous_callback(RELAX_CB)
    saved, status = eng.forward(np.zeros((1, N)), p)
    du0, dp = eng.reverse(np.ones((1, 1, N)))
    counts, times = eng.event_times()
    dp = np.asarray(dp)
    assert (np.asarray(status) == 0).all() and (counts == 1).all()
    assert np.max(np.abs(times[0] - np.log(4.0))) < 1e-12
    rtol = 1e-6 if sa == "gauss_kronrod" else 1e-10            # GK: bisection threshold 1e-7 of IntegratingGKSumCallback
    if shared_p:
        assert np.allclose(dp.ravel() / N, GND, rtol=rtol, atol=0), dp.ravel() / N - GND
    else:
        assert np.allclose(dp, GND[:, None], rtol=rtol, atol=0), np.abs(dp - GND[:, None]).max(axis=1)
    assert np.allclose(np.asarray(du0)[0], -(p[1] - p1 / 4) * np.exp(-10.0) / (p1 / 4), rtol=1e-8)


@pytest.mark.parametrize("sa", SENSEALGS)
def test_parameter_dependent_condition_device_vs_oracle(sa):
    """Random starts below the level, per-member parameters, several save times: device vs oracle."""
    N = 48
    rng = np.random.default_rng(11)
    u0 = 40.0 * rng.random((1, N))
    p = np.stack([100.0 + 20.0 * rng.random(N), 30.0 + 10.0 * rng.random(N)])
    ts = np.linspace(0.5, 6.0, 12)
    kw = dict(abstol=1e-10, reltol=1e-10)
    eng = b.DeviceEnsemble("relax", sa, "tsit5_adaptive", N, ts, (0.0, 6.0), 0.0, cost=b.AffineCost(1.0, -2.0), shared_p=False,
                           ckpt_every_step=True, **kw)
    eng.set_continuous_callback(RELAX_CB)
    saved, status = eng.forward(u0, p)
    du0, dp = eng.reverse()
    cfg = O.make_cfg("relax", sa, "tsit5_adaptive", N, ts, 0.0, 6.0, cost=("affine", 1.0, -2.0), shared_p=False, ckpt_every_step=True,
                     crossing=RELAX_ORACLE, **kw)
    ref = O.gradient(cfg, ts, u0, p)
    rel = lambda a, r: float(np.max(np.abs(np.asarray(a) - r)) / np.max(np.abs(r)))
    assert (np.asarray(status) == 0).all()
    assert rel(saved, ref["saved"]) < 1e-9 and rel(du0, ref["du0"]) < 1e-7 and rel(dp, ref["dp"]) < 1e-7


def test_continuous_callback_params_are_validated():
    eng = b.DeviceEnsemble("relax", "gauss", "tsit5_adaptive", 4, [1.0], (0.0, 1.0), 0.0, abstol=1e-8, reltol=1e-8)
    with pytest.raises(Exception):
        eng.handle.set_continuous_callback_params(lparam=0, lcoef=0.75)          # before set_continuous_callback
    eng.set_continuous_callback(b.ContinuousCallback(idx=0))
    with pytest.raises(Exception):
        eng.handle.set_continuous_callback_params(lparam=2, lcoef=1.0)           # P = 2
    with pytest.raises(Exception):
        eng.handle.set_continuous_callback_params(acomp=1, aparam=0, acoef=1.0)  # d = 1
    eng.handle.set_continuous_callback_params(lparam=0, lcoef=0.75, acomp=0, aparam=1, acoef=1.0)


@pytest.mark.parametrize("shared_p", [True, False])
@pytest.mark.parametrize("sa", SENSEALGS)
def test_dosing_example_on_the_device(sa, shared_p):
    """"Dosing example" (test/Callbacks1/discrete_callbacks.jl:401-427): f = p[1] - u, at t = 8 the affect u[1] += p[2], loss u(10).
    Closed form u(10) = p1 (1 - e^-10) + p2 e^-2 for u0 = 0, so every member's gradient is [1 - e^-10, e^-2]; through the public
    API (PresetTimeCallback with an AffineAffect that adds a parameter) and against the oracle for random starts."""
    N = 24
    ts = np.array([10.0])
    rng = np.random.default_rng(3)
    p = np.array([100.0, 50.0]) if shared_p else np.stack([100.0 + 20.0 * rng.random(N), 50.0 + 10.0 * rng.random(N)])
    exact = np.array([1.0 - np.exp(-10.0), np.exp(-2.0)])
    kw = dict(abstol=1e-14, reltol=1e-14)
    eng = b.DeviceEnsemble("relax", sa, "tsit5_adaptive", N, ts, (0.0, 10.0), 0.0, shared_p=shared_p, ckpt_every_step=True, **kw)
    eng.set_events([8.0], [[1.0]], [[0.0]])
    eng.set_event_param_shift([0], [1], [1.0])
    saved, status = eng.forward(np.zeros((1, N)), p)
    du0, dp = eng.reverse(np.ones((1, 1, N)))
    dp = np.asarray(dp)
    assert (np.asarray(status) == 0).all()
    rtol = 1e-6 if sa == "gauss_kronrod" else 1e-10
    if shared_p:
        assert np.allclose(dp.ravel() / N, exact, rtol=rtol, atol=0), dp.ravel() / N - exact
    else:
        assert np.allclose(dp, exact[:, None], rtol=rtol, atol=0)
    # random starts, several save times, a second dose: device vs oracle
    u0 = 30.0 * rng.random((1, N))
    t2 = np.linspace(1.0, 10.0, 10)
    ev = ([3.0, 8.0], [[1.0], [0.5]], [[0.0], [1.0]])
    eng = b.DeviceEnsemble("relax", sa, "tsit5_adaptive", N, t2, (0.0, 10.0), 0.0, cost=b.AffineCost(1.0, -1.0), shared_p=shared_p,
                           ckpt_every_step=True, abstol=1e-10, reltol=1e-10)
    eng.set_events(*ev)
    eng.set_event_param_shift([0, 0], [1, 0], [1.0, -0.1])
    saved, status = eng.forward(u0, p)
    du0, dp = eng.reverse()
    cfg = O.make_cfg("relax", sa, "tsit5_adaptive", N, t2, 0.0, 10.0, cost=("affine", 1.0, -1.0), shared_p=shared_p, ckpt_every_step=True,
                     events=ev, event_padd=([0, 0], [1, 0], [1.0, -0.1]), abstol=1e-10, reltol=1e-10)
    ref = O.gradient(cfg, t2, u0, p)
    rel = lambda a, r: float(np.max(np.abs(np.asarray(a) - r)) / np.max(np.abs(r)))
    assert rel(saved, ref["saved"]) < 1e-9 and rel(du0, ref["du0"]) < 1e-7 and rel(dp, ref["dp"]) < 1e-7


def test_dosing_through_the_public_api():
    N = 8
    ts = np.array([10.0])
    cb = b.PresetTimeCallback([8.0], b.AffineAffect(1.0, 0.0, add_comp=0, add_param=1, add_coef=1.0))
    prob = b.EnsembleProblem(b.ODEProblem("relax", [0.0], (0.0, 10.0), [100.0, 50.0], callback=cb), u0s=np.zeros((1, N)))
    out, pullback = b._concrete_solve_adjoint(prob, b.Tsit5(adaptive=True), b.B200Adjoint(b.BacksolveAdjoint()), np.zeros((1, N)),
                                              np.array([100.0, 50.0]), None, saveat=ts, abstol=1e-12, reltol=1e-12)
    tang = pullback(np.ones_like(np.asarray(out.u)))
    dp = np.asarray(tang[4]).reshape(-1) / N
    assert np.allclose(dp, [1.0 - np.exp(-10.0), np.exp(-2.0)], rtol=1e-9), dp
    with pytest.raises(b.B200AdjError):        # the dt-grid kernels do not carry it: the dense framework does
        eng = b.DeviceEnsemble("lv", "gauss", "tsit5_fixed", 4, [1.0], (0.0, 1.0), 0.01)
        eng.set_events([0.5], [[1.0, 1.0]], [[0.0, 0.0]])
        eng.set_event_param_shift([0], [1], [1.0])
