"""Edge cases of the reference's seam on the device: a single member, a single save time at t1
(`only_end`, vector cotangent), no save times at all with a continuous cost, save_start / save_end dropping the end points,
many tiny steps, a ragged ensemble just past a block boundary, repeated forward/reverse on one handle, matrix-shaped u0."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import scimlsensitivity_jl_b200 as b
from oracle import oracle as O


def _rel(a, ref):
    return np.abs(np.asarray(a) - ref).max() / max(np.abs(ref).max(), 1e-300)


P_LV = np.array([1.5, 1.0, 3.0, 1.0])


def test_single_member_and_only_end_vector_cotangent():
    """only_end: one output at t1, Delta may be a vector (src/concrete_solve.jl:716, 783-814)."""
    u0 = np.array([[1.0], [1.0]])
    prob = b.ODEProblem("lv", u0[:, 0], (0.0, 3.0), P_LV)
    out, pullback = b._concrete_solve_adjoint(prob, b.Tsit5(dt=0.01), b.B200Adjoint(b.InterpolatingAdjoint()), u0, P_LV, None,
                                              saveat=[3.0], save_start=False)
    assert out.u.shape == (1, 2, 1)
    tang = pullback(np.ones((2, 1)))                               # vector-shaped cotangent for the single output
    cfg = O.make_cfg("lv", "interpolating", "tsit5_fixed", 1, [3.0], 0.0, 3.0, dt=0.01, cost=("affine", 0.0, 1.0))
    ref = O.gradient(cfg, [3.0], u0, P_LV)
    assert _rel(tang[3], ref["du0"]) < 1e-9 and _rel(tang[4], ref["dp"]) < 1e-9


def test_save_start_save_end_false_and_saveat_number():
    """The rrule keeps t0 and t1 in the output when saveat is a number or an array (src/concrete_solve.jl:718-735, 752-769);
    save_start = false only makes the pullback ignore the cotangent at t0 (`no_start`, :962); with an EMPTY saveat the end
    points are dropped (:740-750); BacksolveAdjoint keeps the plain solver's saving behaviour (:713-717)."""
    N = 7
    rng = np.random.default_rng(0)
    u0 = np.exp(0.1 * rng.standard_normal((2, N)))
    prob = b.ODEProblem("lv", u0[:, 0], (0.0, 1.0), P_LV)
    out, pullback = b._concrete_solve_adjoint(prob, b.Tsit5(dt=0.01), b.B200Adjoint(b.GaussAdjoint()), u0, P_LV, None,
                                              saveat=0.25, save_start=False, save_end=False)
    assert np.allclose(out.t, [0.0, 0.25, 0.5, 0.75, 1.0]) and out.u.shape == (5, 2, N)
    tang = pullback(2.0 * out.u)
    cfg = O.make_cfg("lv", "gauss", "tsit5_fixed", N, out.t, 0.0, 1.0, dt=0.01, cost=("affine", 2.0, 0.0), no_start=True)
    ref = O.gradient(cfg, out.t, u0, P_LV)
    assert _rel(tang[3], ref["du0"]) < 1e-9 and _rel(tang[4], ref["dp"]) < 1e-9
    # an explicit array that contains t0: the same (sorted, t0 kept, its cotangent ignored)
    out2, pullback2 = b._concrete_solve_adjoint(prob, b.Tsit5(dt=0.01), b.B200Adjoint(b.InterpolatingAdjoint()), u0, P_LV, None,
                                                saveat=[0.5, 0.0, 1.0], save_start=False)
    assert np.allclose(out2.t, [0.0, 0.5, 1.0])
    tang2 = pullback2(np.ones_like(out2.u))
    ref2 = O.gradient(O.make_cfg("lv", "interpolating", "tsit5_fixed", N, out2.t, 0.0, 1.0, dt=0.01, cost=("affine", 0.0, 1.0), no_start=True), out2.t, u0, P_LV)
    assert _rel(tang2[3], ref2["du0"]) < 1e-9 and _rel(tang2[4], ref2["dp"]) < 1e-9
    # Backsolve: the forward solve's own saving behaviour (end points dropped)
    out3, _ = b._concrete_solve_adjoint(prob, b.Tsit5(dt=0.01), b.B200Adjoint(b.BacksolveAdjoint()), u0, P_LV, None,
                                        saveat=0.25, save_start=False, save_end=False)
    assert np.allclose(out3.t, [0.25, 0.5, 0.75])
    # empty saveat: every step is an output, end points dropped
    out4, _ = b._concrete_solve_adjoint(prob, b.Tsit5(dt=0.25), b.B200Adjoint(b.GaussAdjoint()), u0, P_LV, None, save_start=False, save_end=False)
    assert np.allclose(out4.t, [0.25, 0.5, 0.75])


def test_ragged_block_boundaries_and_matrix_u0():
    """N = 1, 31, 33, 149 (one past the SM count), 449 (one past the default block), with u0 given as d x N matrix."""
    saveat = np.linspace(0.0, 1.0, 11)
    for N in (1, 31, 33, 149, 449):
        rng = np.random.default_rng(N)
        u0 = np.array([1.0, 0.0, 0.0])[:, None] + 0.1 * rng.standard_normal((3, N))
        p = np.array([10.0, 28.0, 8.0 / 3.0])
        eng = b.DeviceEnsemble("lorenz", "gauss", "tsit5_fixed", N, saveat, (0.0, 1.0), 0.01, cost=b.AffineCost(1.0, -2.0))
        eng.forward(u0, p)
        du0, dp = eng.reverse()
        cfg = O.make_cfg("lorenz", "gauss", "tsit5_fixed", N, saveat, 0.0, 1.0, dt=0.01, cost=("affine", 1.0, -2.0))
        ref = O.gradient(cfg, saveat, u0, p, want_saved=False)
        assert _rel(du0, ref["du0"]) < 1e-9 and _rel(dp, ref["dp"]) < 1e-9, N
        eng.close()


def test_many_small_steps_and_handle_reuse():
    """S = 20000 steps (dt = 1e-4); the same handle serves several forward/reverse pairs with different inputs."""
    N = 40
    saveat = np.array([0.5, 1.0, 2.0])
    eng = b.DeviceEnsemble("lv", "interpolating", "tsit5_fixed", N, saveat, (0.0, 2.0), 1e-4, cost=b.AffineCost(0.0, 1.0))
    for seed in (0, 1):
        rng = np.random.default_rng(seed)
        u0 = np.exp(0.1 * rng.standard_normal((2, N)))
        p = P_LV * (1.0 + 0.05 * seed)
        saved, status = eng.forward(u0, p)
        du0, dp = eng.reverse()
        cfg = O.make_cfg("lv", "interpolating", "tsit5_fixed", N, saveat, 0.0, 2.0, dt=1e-4, cost=("affine", 0.0, 1.0))
        ref = O.gradient(cfg, saveat, u0, p)
        assert np.abs(saved - ref["saved"]).max() < 1e-10
        assert _rel(du0, ref["du0"]) < 1e-8 and _rel(dp, ref["dp"]) < 1e-8
    eng.close()


def test_state_errors():
    saveat = np.linspace(0.0, 1.0, 3)
    eng = b.DeviceEnsemble("lv", "gauss", "tsit5_fixed", 4, saveat, (0.0, 1.0), 0.01)
    with pytest.raises(b.B200AdjError) as ei:                       # reverse before forward
        eng.reverse(np.zeros((3, 2, 4)))
    assert ei.value.code == -5
    eng.forward(np.ones((2, 4)), P_LV)
    with pytest.raises(b.B200AdjError) as ei:                       # explicit cost but no cotangent array
        eng.handle.reverse(None, np.zeros((2, 4)), np.zeros(4))
    assert ei.value.code == -1
    eng.close()
    with pytest.raises(b.B200AdjError):                             # bad block size
        b.DeviceEnsemble("lv", "gauss", "tsit5_fixed", 4, saveat, (0.0, 1.0), 0.01, block_threads=48)
    with pytest.raises(b.B200AdjError):                             # descending save times
        b.DeviceEnsemble("lv", "gauss", "tsit5_fixed", 4, saveat[::-1].copy(), (0.0, 1.0), 0.01)


@pytest.mark.parametrize("per_sm", [160, 192, 224, 448])
@pytest.mark.parametrize("sa", ["gauss", "interpolating", "quadrature"])
def test_travelling_warp_groups_all_remainders(per_sm, sa):
    """One block per SM with 5, 6, 7 or 14 warp groups: 1, 2 or 3 groups travel round the four sub-partitions
    (csrc/ode_tsit5.cuh).  Per-member parameters, so every member's own dp and du0 are compared with the oracle (sampled),
    and the shared-parameter gradient equals the sum of the per-member ones."""
    import torch
    nsm = torch.cuda.get_device_properties(0).multi_processor_count
    N = nsm * per_sm - 37                      # ragged last block
    T, dt = 0.5, 0.01
    rng = np.random.default_rng(per_sm)
    u0 = np.array([1.0, 0.0, 0.0])[:, None] + 0.1 * rng.standard_normal((3, N))
    p = np.array([10.0, 28.0, 8.0 / 3.0])[:, None] * (1.0 + 0.01 * rng.standard_normal((3, N)))
    t = np.linspace(0.0, T, 6)
    kw = dict(quad_abstol=1e-10, quad_reltol=1e-10) if sa == "quadrature" else {}
    eng = b.DeviceEnsemble("lorenz", sa, "tsit5_fixed", N, t, (0.0, T), dt, shared_p=False, cost=b.AffineCost(1.0, -2.0), **kw)
    eng.forward(u0, p)
    du0, dp = eng.reverse()
    idx = np.r_[0:40, N - 40:N, rng.choice(N, 80, replace=False)]           # first block, ragged last block, random members
    ref = O.gradient(O.make_cfg("lorenz", sa, "tsit5_fixed", len(idx), t, 0.0, T, dt=dt, cost=("affine", 1.0, -2.0), shared_p=False, **kw),
                     t, u0[:, idx], p[:, idx])
    assert _rel(np.asarray(du0)[:, idx], ref["du0"]) < 1e-9
    assert _rel(np.asarray(dp)[:, idx], ref["dp"]) < (1e-7 if sa == "quadrature" else 1e-9)
    eng.close()
    if sa != "quadrature":
        p0 = np.array([10.0, 28.0, 8.0 / 3.0])
        e1 = b.DeviceEnsemble("lorenz", sa, "tsit5_fixed", N, t, (0.0, T), dt, shared_p=True, cost=b.AffineCost(1.0, -2.0))
        e2 = b.DeviceEnsemble("lorenz", sa, "tsit5_fixed", N, t, (0.0, T), dt, shared_p=False, cost=b.AffineCost(1.0, -2.0))
        e1.forward(u0, p0); e2.forward(u0, np.repeat(p0[:, None], N, 1))
        _, dps = e1.reverse(); _, dpm = e2.reverse()
        assert _rel(np.asarray(dps), np.asarray(dpm).sum(axis=1)) < 1e-11
        e1.close(); e2.close()


def test_pinned_outputs_do_not_alias_when_shapes_coincide():
    """pin_outputs=True with per-member parameters and P == d (Lorenz): du0 and dp are distinct page-locked buffers
    (the cache is keyed by role), and so are the two step-count arrays of an adaptive handle."""
    N = 40
    saveat = np.linspace(0.0, 1.0, 11)
    rng = np.random.default_rng(3)
    u0 = np.array([1.0, 0.0, 0.0])[:, None] + 0.1 * rng.standard_normal((3, N))
    p = np.array([10.0, 28.0, 8.0 / 3.0])[:, None] * np.exp(0.01 * rng.standard_normal((3, N)))
    eng = b.DeviceEnsemble("lorenz", "gauss", "tsit5_fixed", N, saveat, (0.0, 1.0), 0.01, shared_p=False,
                           cost=b.AffineCost(1.0, -2.0), pin_outputs=True)
    eng.forward(u0, p)
    du0, dp = eng.reverse()
    assert du0.ctypes.data != dp.ctypes.data
    cfg = O.make_cfg("lorenz", "gauss", "tsit5_fixed", N, saveat, 0.0, 1.0, dt=0.01, cost=("affine", 1.0, -2.0), shared_p=False)
    ref = O.gradient(cfg, saveat, u0, p, want_saved=False)
    assert _rel(du0, ref["du0"]) < 1e-9 and _rel(dp, ref["dp"]) < 1e-9
    eng.close()
    eng = b.DeviceEnsemble("lv", "gauss", "tsit5_adaptive", 8, saveat, (0.0, 1.0), 0.0, cost=b.AffineCost(0.0, 1.0),
                           abstol=1e-8, reltol=1e-8, pin_outputs=True)
    eng.forward(np.ones((2, 8)), P_LV)
    eng.reverse()
    f, r = eng.step_counts()
    assert f.ctypes.data != r.ctypes.data and (f > 0).all()      # (the reverse count is kept for QuadratureAdjoint only)
    eng.close()


def test_finer_reverse_times_do_not_disturb_the_forward_save_table():
    """adjoint_sensitivities(sol, t = finer grid) re-targets the reverse pass only: a later forward pass on the same handle
    still writes the create-time K save points (no overflow of the staging buffer, sol.t matches sol.u)."""
    N = 50
    coarse, fine = np.linspace(0.0, 1.0, 3), np.linspace(0.0, 1.0, 21)
    rng = np.random.default_rng(5)
    u0 = np.exp(0.1 * rng.standard_normal((2, N)))
    for stepper, dt, kw in (("tsit5_fixed", 0.01, {}), ("tsit5_adaptive", 0.0, dict(abstol=1e-9, reltol=1e-9))):
        eng = b.DeviceEnsemble("lv", "gauss", stepper, N, coarse, (0.0, 1.0), dt, cost=b.AffineCost(1.0, 0.0), **kw)
        s1, _ = eng.forward(u0, P_LV)
        eng.set_reverse("gauss", cost=b.AffineCost(1.0, 0.0), t=fine)
        du0, dp = eng.reverse()
        cfg = O.make_cfg("lv", "gauss", stepper, N, fine, 0.0, 1.0, dt=dt, cost=("affine", 1.0, 0.0), **kw)
        ref = O.gradient(cfg, fine, u0, P_LV, want_saved=False)
        assert _rel(du0, ref["du0"]) < 1e-7 and _rel(dp, ref["dp"]) < 1e-7
        s2, _ = eng.forward(u0, P_LV)
        assert s2.shape == (3, 2, N) and np.array_equal(np.asarray(s1), np.asarray(s2))
        eng.close()
