"""Numbers the REFERENCE itself holds for this path (its tests' fixtures and its documentation's printed results), reproduced by
the oracle.  These are the pins that are not self-referential: none of them was produced by code of this repository.

| reference-held number | where | what it pins |
|---|---|---|
| res.u = [0.866554105436901]                                   | docs/src/examples/hybrid_jump/bouncing_ball.md:60 | the minimiser of (x(15) - 20)^2 over the restitution coefficient: x(15) peaks BELOW 20 (19.3679), so the optimum is the point where the GRADIENT through three bounces vanishes -- a root of the adjoint with the implicit event-time correction |
| tstop = 3.1943828249997, vbefore = -31.30495168499705, vafter = 25.04396134799764 | test/Callbacks2/continuous_vs_discrete.jl:19-21 | event location of the bouncing ball's first impact, left and right limits |
| gND = [0.9999546000702386, 0.00018159971904994378]            | test/Callbacks2/continuous_callbacks.jl:343 (rtol 1e-10 against every sensealg, :344-359) | condition that depends on a parameter, additive parameter affect: dtau/dp enters the gradient |
| 8.305557728239275 / 8.305305252400714 / 8.305266428305409     | test/Core6/forward_prob_kwargs.jl:28-30 | tests/test_oracle_relations.py::test_lv_reference_held_number_adaptive_tsit5_tol_1e12 (the three prints disagree at 3.5e-5; the oracle sits inside their spread and on the DOP853 value) |

The device-side twins are tests/test_gpu_reference_held_numbers.py."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O

SENSEALGS = ("interpolating", "gauss", "gauss_kronrod", "backsolve")
BALL = dict(idx=0, level=0.0, direction=-1, pcomp=1, pparam=1, psign=-1.0)          # condition u[1], affect u[2] = -p[2] u[2]
DOCS_OPTIMUM = 0.866554105436901                                                     # bouncing_ball.md:60
TSTOP, VBEFORE, VAFTER = 3.1943828249997, -31.30495168499705, 25.04396134799764      # continuous_vs_discrete.jl:19-21
GND = np.array([0.9999546000702386, 0.00018159971904994378])                         # continuous_callbacks.jl:343


def ball_loss_gradient(theta, sa, abstol, reltol):
    """d/dtheta of (x(15) - 20)^2 for u0 = [50, 0], p = [9.8, theta] (bouncing_ball.md:33-54), theta[N] -> (x[N], g[N])."""
    theta = np.atleast_1d(np.asarray(theta, dtype=np.float64))
    N = theta.size
    ts = np.array([15.0])
    u0 = np.tile(np.array([[50.0], [0.0]]), (1, N))
    p = np.stack([np.full(N, 9.8), theta])
    cfg = O.make_cfg("ball", sa, "tsit5_adaptive", N, ts, 0.0, 15.0, abstol=abstol, reltol=reltol, crossing=BALL, shared_p=False,
                     ckpt_every_step=True)
    x = O.forward(cfg, ts, u0, p)[0, 0]
    dL = np.zeros((1, 2, N)); dL[0, 0] = 2.0 * (x - 20.0)
    return x, O.gradient(cfg, ts, u0, p, dLdu=dL)["dp"][1]


def secant_roots(grad, a, b, iters=12):
    """Member-wise secant iteration on grad(theta)[i] = 0 from the brackets (a[i], b[i])."""
    ga, gb = grad(a), grad(b)
    for _ in range(iters):
        den = np.where(gb == ga, 1.0, gb - ga)
        c = np.where(gb == ga, b, b - gb * (b - a) / den)
        a, ga, b, gb = b, gb, c, grad(c)
    return b


@pytest.mark.parametrize("tol", [(1e-6, 1e-3), (1e-10, 1e-10)], ids=["docs_default_tolerances", "tight"])
@pytest.mark.parametrize("sa", SENSEALGS)
def test_bouncing_ball_docs_optimum_is_where_the_adjoint_gradient_vanishes(sa, tol):
    x, g = ball_loss_gradient([DOCS_OPTIMUM, 0.8], sa, *tol)
    assert abs(x[0] - 19.367902815) < 1e-8            # the end height never reaches the target 20: the optimum is a stationary point
    assert abs(g[0]) < 1e-9 * abs(g[1])               # observed 1e-14 of the gradient at the starting point 0.8
    # every start converges to the reference's printed optimum (observed |root - printed| = 1e-14)
    N = 8
    a = 0.862 + 0.004 * np.arange(N) / N
    root = secant_roots(lambda th: ball_loss_gradient(th, sa, *tol)[1], a, a + 0.003)
    assert np.max(np.abs(root - DOCS_OPTIMUM)) < 1e-11, root - DOCS_OPTIMUM


@pytest.mark.parametrize("tol", [(1e-12, 1e-12), (1e-6, 1e-3)])
def test_first_impact_event_location_matches_the_reference_fixture(tol):
    """continuous_vs_discrete.jl prescribes tstop / vbefore / vafter of the bouncing ball's first impact as literals so that a
    DiscreteCallback can replay the ContinuousCallback: the event the oracle's root finder locates is that event."""
    cfg = O.make_cfg("ball", "backsolve", "tsit5_adaptive", 1, np.array([5.0]), 0.0, 5.0, abstol=tol[0], reltol=tol[1], crossing=BALL)
    t, um, up = O.event_list(cfg, [50.0, 0.0], [9.8, 0.8])
    assert len(t) == 1
    assert abs(t[0] - TSTOP) < 1e-12 and abs(um[0, 1] - VBEFORE) < 1e-11 and abs(up[0, 1] - VAFTER) < 1e-11
    assert abs(um[0, 0]) < 1e-11


@pytest.mark.parametrize("sa", SENSEALGS)
def test_parameter_dependent_condition_reproduces_gND(sa):
    """continuous_callbacks.jl:317-359 ("Re-compile tape"): f = p[1] - u, condition u - 3/4 p[1], affect u += p[2], loss u(10),
    abstol = reltol = 1e-14; every sensealg is required to match the printed gND at rtol 1e-10."""
    cr = dict(idx=0, level=0.0, direction=0, lparam=0, lcoef=0.75, acomp=0, aparam=1, acoef=1.0)
    ts = np.array([10.0])
    cfg = O.make_cfg("relax", sa, "tsit5_adaptive", 1, ts, 0.0, 10.0, abstol=1e-14, reltol=1e-14, crossing=cr, ckpt_every_step=True)
    r = O.gradient(cfg, ts, np.array([[0.0]]), np.array([100.0, 50.0]), dLdu=np.ones((1, 1, 1)))
    # GaussKronrodAdjoint stops bisecting at |K - G| < 1e-7 (IntegratingGKSumCallback's default), 6e-8 here
    assert np.allclose(r["dp"], GND, rtol=1e-6 if sa == "gauss_kronrod" else 1e-10, atol=0), r["dp"] - GND
    assert abs(r["saved"][0, 0, 0] - (100.0 + 25.0 * 4.0 * np.exp(-10.0))) < 1e-10
    t, um, up = O.event_list(cfg, [0.0], [100.0, 50.0])
    assert len(t) == 1 and abs(t[0] - np.log(4.0)) < 1e-12 and abs(um[0, 0] - 75.0) < 1e-10 and abs(up[0, 0] - 125.0) < 1e-10


@pytest.mark.parametrize("stepper,kw", [("tsit5_adaptive", dict(abstol=1e-14, reltol=1e-14)), ("tsit5_fixed", dict(dt=0.01))])
@pytest.mark.parametrize("sa", SENSEALGS)
def test_dosing_example_closed_form(sa, stepper, kw):
    """"Dosing example" of test/Callbacks1/discrete_callbacks.jl:401-427 (f = p[1] - u, at t = 8 the affect u[1] += p[2], loss
    u(10), abstol = reltol = 1e-14; the reference asserts ForwardDiff == Zygote through BacksolveAdjoint).  No literal is printed
    there; the closed form is u(10) = p1 (1 - e^-10) + p2 e^-2, gradient [1 - e^-10, e^-2]."""
    ts = np.array([10.0])
    cfg = O.make_cfg("relax", sa, stepper, 1, ts, 0.0, 10.0, events=([8.0], [[1.0]], [[0.0]]), event_padd=([0], [1], [1.0]),
                     ckpt_every_step=True, **kw)
    r = O.gradient(cfg, ts, np.array([[0.0]]), np.array([100.0, 50.0]), dLdu=np.ones((1, 1, 1)))
    exact = np.array([1.0 - np.exp(-10.0), np.exp(-2.0)])
    assert abs(r["saved"][0, 0, 0] - (100.0 * exact[0] + 50.0 * exact[1])) < 1e-10
    assert np.allclose(r["dp"], exact, rtol=1e-6 if sa == "gauss_kronrod" else 1e-10, atol=0), r["dp"] - exact
