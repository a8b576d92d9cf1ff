"""Preset-time events (DiscreteCallback / PresetTimeCallback) on the Tsit5 paths (adaptive, and fixed step): device vs oracle, and the
reference's own relations (test/Callbacks1/discrete_callbacks.jl:200-231): every sensealg agrees with differentiation
through the solver, Backsolve == Interpolating == Gauss."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import scimlsensitivity_jl_b200 as b
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _rel(a, ref):
    return float(np.max(np.abs(np.asarray(a) - ref)) / (np.max(np.abs(ref)) + 1e-300))


CASES = {
    # affect!(integrator) = integrator.u[1] += 2.0 at t == 5            (discrete_callbacks.jl:263-269)
    "add_single": ([5.0], [[1.0, 1.0]], [[2.0, 0.0]]),
    # affecttimes = [2.03, 4.0, 8.0], u[1] += 2.0                        (:270-276)
    "add_multi": ([2.03, 4.0, 8.0], [[1.0, 1.0]] * 3, [[2.0, 0.0]] * 3),
    # integrator.u[1] = 2.0 at t == 5                                    (:286-292)
    "set_single": ([5.0], [[0.0, 1.0]], [[2.0, 0.0]]),
    # callbacks with no effect                                           (:249-262)
    "no_effect": ([5.0], [[1.0, 1.0]], [[0.0, 0.0]]),
    # parameter changing callback: integrator.p .= 2 .* integrator.p .- 0.5 at t == 5.1     (:294-303)
    "param_change": ([5.1], [[1.0, 1.0]], [[0.0, 0.0]], [[2.0] * 4], [[-0.5] * 4]),
    # a state kick and a (different) parameter change at different times
    "mixed": ([2.03, 5.1], [[1.0, 1.0], [1.0, 1.0]], [[2.0, 0.0], [0.0, 0.0]], [[1.0] * 4, [2.0, 1.0, 0.5, 1.0]], [[0.0] * 4, [-0.5, 0.0, 0.1, 0.0]]),
}


@pytest.mark.parametrize("case", sorted(CASES))
@pytest.mark.parametrize("sa,every", [("interpolating", False), ("gauss", False), ("backsolve", True), ("backsolve", False)])
def test_events_device_vs_oracle(case, sa, every):
    ev = CASES[case]
    N = 40
    rng = np.random.default_rng(11)
    u0 = 1.0 + 0.05 * rng.standard_normal((2, N))
    p = np.array([1.5, 1.0, 3.0, 1.0])
    t = np.arange(0.0, 10.0001, 0.5)                     # savingtimes = 0.5; t = 5 and 4, 8 coincide with events
    tol = dict(abstol=1e-10, reltol=1e-10)
    eng = b.DeviceEnsemble("lv", sa, "tsit5_adaptive", N, t, (0.0, 10.0), 0.0, cost=b.AffineCost(0.0, 1.0), ckpt_every_step=every,
                           max_steps=8192, **tol)
    eng.set_events(*ev)
    saved, status = eng.forward(u0, p)
    du0, dp = eng.reverse()
    cfg = O.make_cfg("lv", sa, "tsit5_adaptive", N, t, 0.0, 10.0, cost=("affine", 0.0, 1.0), ckpt_every_step=every, events=ev, **tol)
    ref = O.gradient(cfg, t, u0, p)
    assert (np.asarray(status) == 0).all()
    assert _rel(saved, ref["saved"]) < 1e-9
    assert _rel(du0, ref["du0"]) < 1e-7 and _rel(dp, ref["dp"]) < 1e-7
    # the saved state AT an event time is the post-event state
    if case == "set_single":
        assert np.allclose(np.asarray(saved)[10, 0, :], 2.0, atol=1e-12)
    eng.close()


@pytest.mark.parametrize("case", ["add_multi", "param_change", "mixed"])
def test_events_match_differentiation_through_the_solver(case):
    """g(sol) = sum(sol): gradient == finite differences of the loss through the (oracle) solver, all sensealgs equal."""
    ev = CASES[case]
    t = np.arange(0.0, 10.0001, 0.5)
    u0 = np.ones((2, 1)); p = np.array([1.5, 1.0, 3.0, 1.0])
    tol = dict(abstol=1e-12, reltol=1e-12)
    lcfg = O.make_cfg("lv", "interpolating", "tsit5_adaptive", 1, t, 0.0, 10.0, cost=("affine", 0.0, 1.0), events=ev, **tol)
    e = 1e-6
    fd_p = np.array([(O.loss(lcfg, t, u0, p + e * np.eye(4)[q])[0] - O.loss(lcfg, t, u0, p - e * np.eye(4)[q])[0]) / (2 * e) for q in range(4)])
    fd_u = np.array([(O.loss(lcfg, t, u0 + e * np.eye(2)[j][:, None], p)[0] - O.loss(lcfg, t, u0 - e * np.eye(2)[j][:, None], p)[0]) / (2 * e) for j in range(2)])
    res = {}
    for sa in ("interpolating", "gauss", "backsolve", "gauss_kronrod"):
        eng = b.DeviceEnsemble("lv", sa, "tsit5_adaptive", 1, t, (0.0, 10.0), 0.0, cost=b.AffineCost(0.0, 1.0), ckpt_every_step=True,
                               max_steps=16384, **tol)
        eng.set_events(*ev)
        eng.forward(u0, p)
        du0, dp = eng.reverse()
        assert _rel(dp, fd_p) < 1e-6 and _rel(np.asarray(du0)[:, 0], fd_u) < 1e-6, sa
        res[sa] = (np.asarray(du0).copy(), np.asarray(dp).copy())
        eng.close()
    assert _rel(res["backsolve"][1], res["interpolating"][1]) < 1e-7          # du01 ~ du03 rtol 1e-7 (:222-232)
    assert _rel(res["gauss"][1], res["interpolating"][1]) < 1e-7


def test_events_public_api_and_rejections():
    t = np.arange(0.0, 10.0001, 0.5)
    prob = b.ODEProblem("lv", np.ones(2), (0.0, 10.0), np.array([1.5, 1.0, 3.0, 1.0]))
    cbp = b.PresetTimeCallback([5.1], b.AffineAffect(1.0, 0.0, p_scale=2.0, p_shift=-0.5))
    assert [x.shape for x in cbp.tables(2, 4)] == [(1,), (1, 2), (1, 2), (1, 4), (1, 4)]
    solp = b.solve(b.EnsembleProblem(prob), b.Tsit5(adaptive=True), b.EnsembleB200(), trajectories=2, saveat=t, callback=cbp, abstol=1e-10, reltol=1e-10)
    _, dpp = b.adjoint_sensitivities(solp, b.Tsit5(adaptive=True), sensealg=b.GaussAdjoint(), t=t, dgdu_discrete=b.AffineCost(0.0, 1.0), abstol=1e-10, reltol=1e-10)
    refp = O.gradient(O.make_cfg("lv", "gauss", "tsit5_adaptive", 2, t, 0.0, 10.0, cost=("affine", 0.0, 1.0), abstol=1e-10, reltol=1e-10,
                                 events=CASES["param_change"]), t, np.ones((2, 2)), np.array([1.5, 1.0, 3.0, 1.0]))
    assert _rel(np.asarray(dpp).reshape(-1), refp["dp"]) < 1e-7
    cb = b.PresetTimeCallback([5.0], b.AffineAffect([1.0, 1.0], [2.0, 0.0]))
    sol = b.solve(b.EnsembleProblem(prob), b.Tsit5(adaptive=True), b.EnsembleB200(), trajectories=3, saveat=t, callback=cb, abstol=1e-10, reltol=1e-10)
    du0, dp = b.adjoint_sensitivities(sol, b.Tsit5(adaptive=True), sensealg=b.InterpolatingAdjoint(), t=t, dgdu_discrete=b.AffineCost(0.0, 1.0),
                                      abstol=1e-10, reltol=1e-10)
    cfg = O.make_cfg("lv", "interpolating", "tsit5_adaptive", 3, t, 0.0, 10.0, cost=("affine", 0.0, 1.0), abstol=1e-10, reltol=1e-10,
                     events=([5.0], [[1.0, 1.0]], [[2.0, 0.0]]))
    ref = O.gradient(cfg, t, np.ones((2, 3)), np.array([1.5, 1.0, 3.0, 1.0]))
    assert _rel(np.asarray(dp).reshape(-1), ref["dp"]) < 1e-7
    with pytest.raises(NotImplementedError):
        b.adjoint_sensitivities(sol, b.Tsit5(adaptive=True), sensealg=b.QuadratureAdjoint(), t=t, dgdu_discrete=b.AffineCost(0.0, 1.0))
    with pytest.raises(b.B200AdjError):      # fixed step: event times must lie on the dt grid
        b.solve(b.EnsembleProblem(prob), b.Tsit5(adaptive=False, dt=0.01), b.EnsembleB200(), trajectories=3, saveat=t,
                callback=b.PresetTimeCallback([5.005], b.AffineAffect([1.0, 1.0], [2.0, 0.0])))
    with pytest.raises(Exception):       # event time outside (t0, t1)
        eng = b.DeviceEnsemble("lv", "gauss", "tsit5_adaptive", 4, t, (0.0, 10.0), 0.0)
        eng.set_events([10.0], [[1.0, 1.0]], [[0.0, 0.0]])


@pytest.mark.parametrize("case", sorted(CASES))
@pytest.mark.parametrize("sa,every", [("interpolating", False), ("gauss", False), ("gauss_kronrod", False), ("backsolve", True), ("backsolve", False)])
def test_events_on_the_fixed_step_grid(case, sa, every):
    """The same event cases on the FIXED-step Tsit5 path (event times are dt-grid points: 2.03 = 203 dt): EV instantiations of
    tsit5_forward_kernel / tsit5_reverse_kernel against the oracle (test_oracle_relations.py pins the oracle's fixed-grid events
    by finite differences through the hybrid solve)."""
    ev = CASES[case]
    N, dt = 70, 0.01
    rng = np.random.default_rng(21)
    u0 = 1.0 + 0.05 * rng.standard_normal((2, N))
    p = np.array([1.5, 1.0, 3.0, 1.0])
    t = np.arange(0.0, 10.0001, 0.5)
    eng = b.DeviceEnsemble("lv", sa, "tsit5_fixed", N, t, (0.0, 10.0), dt, cost=b.AffineCost(0.0, 1.0), ckpt_every_step=every)
    eng.set_events(*ev)
    saved, status = eng.forward(u0, p)
    du0, dp = eng.reverse()
    cfg = O.make_cfg("lv", sa, "tsit5_fixed", N, t, 0.0, 10.0, dt=dt, cost=("affine", 0.0, 1.0), ckpt_every_step=every, events=ev)
    ref = O.gradient(cfg, t, u0, p)
    assert (np.asarray(status) == 0).all()
    assert _rel(saved, ref["saved"]) < 1e-10
    tol = 1e-6 if sa == "gauss_kronrod" else 1e-8
    assert _rel(du0, ref["du0"]) < tol and _rel(dp, ref["dp"]) < tol, (_rel(du0, ref["du0"]), _rel(dp, ref["dp"]))
    if case == "set_single":
        assert np.allclose(np.asarray(saved)[10, 0, :], 2.0, atol=1e-12)
    eng.close()


def test_fixed_step_events_through_the_public_api_per_member_parameters():
    t = np.arange(0.0, 10.0001, 0.5)
    N = 12
    rng = np.random.default_rng(5)
    u0 = 1.0 + 0.05 * rng.standard_normal((2, N))
    p = np.array([1.5, 1.0, 3.0, 1.0])[:, None] * np.exp(0.02 * rng.standard_normal((4, N)))
    cb = b.PresetTimeCallback([2.03, 5.1], [b.AffineAffect([1.0, 1.0], [2.0, 0.0]), b.AffineAffect(1.0, 0.0, p_scale=[2.0, 1.0, 0.5, 1.0], p_shift=[-0.5, 0.0, 0.1, 0.0])])
    prob = b.EnsembleProblem(b.ODEProblem("lv", u0[:, 0], (0.0, 10.0), p[:, 0]), u0s=u0, ps=p)
    sol = b.solve(prob, b.Tsit5(dt=0.01), b.EnsembleB200(), saveat=t, callback=cb)
    du0, dp = b.adjoint_sensitivities(sol, b.Tsit5(dt=0.01), sensealg=b.InterpolatingAdjoint(), t=t, dgdu_discrete=b.AffineCost(0.0, 1.0))
    cfg = O.make_cfg("lv", "interpolating", "tsit5_fixed", N, t, 0.0, 10.0, dt=0.01, cost=("affine", 0.0, 1.0), shared_p=False, events=CASES["mixed"])
    ref = O.gradient(cfg, t, u0, p)
    assert _rel(du0, ref["du0"]) < 1e-8 and _rel(dp, ref["dp"]) < 1e-8


@pytest.mark.parametrize("dtype,tol", [("f64", 1e-8), ("f32", 5e-4)])
@pytest.mark.parametrize("sa", ["interpolating", "gauss"])
def test_hybrid_neural_ode_events_on_the_mlp_family(sa, dtype, tol):
    """test/Core5/HybridNODE.jl:9-24: a neural RHS with an external kick "u[1] += 0.2 * cbinput[k]" at preset times.  The MLP
    family's CUDA-core kernels (F64 / F32) carry the events on the dt grid; device vs oracle (fp64), which
    tests/test_oracle_relations.py pins by finite differences through the hybrid solve."""
    H, N = 64, 100
    rng = np.random.default_rng(7)
    p = np.concatenate([(rng.standard_normal((H, 2)) / np.sqrt(2)).ravel(order="F"), 0.1 * rng.standard_normal(H),
                        (rng.standard_normal((H, H)) / np.sqrt(H)).ravel(order="F"), 0.1 * rng.standard_normal(H),
                        (rng.standard_normal((2, H)) / np.sqrt(H)).ravel(order="F"), 0.1 * rng.standard_normal(2)])
    T, dt = 3.0, 0.05
    ts = np.arange(0.25, T + 1e-9, 0.25)
    et = np.arange(0.5, T - 1e-9, 0.5)
    ev = (et, np.ones((len(et), 2)), np.stack([0.2 * rng.random(len(et)), np.zeros(len(et))], 1))
    u0 = rng.uniform(-1, 1, (2, N))
    eng = b.DeviceEnsemble("mlp", sa, "tsit5_fixed", N, ts, (0.0, T), dt, cost=b.AffineCost(1.0, -0.5), dtype=dtype)
    eng.set_events(*ev)
    saved, status = eng.forward(u0, p)
    du0, dp = eng.reverse()
    cfg = O.make_cfg("mlp", sa, "tsit5_fixed", N, ts, 0.0, T, dt=dt, cost=("affine", 1.0, -0.5), mlp_hidden=H, events=ev)
    ref = O.gradient(cfg, ts, u0, p)
    assert (np.asarray(status) == 0).all()
    assert _rel(saved, ref["saved"]) < (1e-10 if dtype == "f64" else 5e-5)
    assert _rel(du0, ref["du0"]) < tol and _rel(dp, ref["dp"]) < tol, (_rel(du0, ref["du0"]), _rel(dp, ref["dp"]))
    # the kick is visible in the primal: without events the same handle gives a different trajectory
    eng.set_events([], np.zeros((0, 2)), np.zeros((0, 2)))
    saved0, _ = eng.forward(u0, p)
    assert _rel(saved0, ref["saved"]) > 1e-3
    eng.close()


def test_mlp_events_refusals():
    ts = np.array([0.5, 1.0])
    eng = b.DeviceEnsemble("mlp", "interpolating", "tsit5_fixed", 64, ts, (0.0, 1.0), 0.05, dtype="bf16_f32acc")
    with pytest.raises(b.B200AdjError) as ei:        # the tensor-core path does not carry events
        eng.set_events([0.5], [[1.0, 1.0]], [[0.1, 0.0]])
    assert ei.value.code == -2
    eng = b.DeviceEnsemble("mlp", "interpolating", "tsit5_fixed", 64, ts, (0.0, 1.0), 0.05)
    with pytest.raises(b.B200AdjError):              # parameter-changing affects: not for the MLP family
        eng.set_events([0.5], [[1.0, 1.0]], [[0.1, 0.0]], np.ones((1, 4482)), np.zeros((1, 4482)))
    with pytest.raises(b.B200AdjError):              # off the dt grid
        eng.set_events([0.52], [[1.0, 1.0]], [[0.1, 0.0]])
