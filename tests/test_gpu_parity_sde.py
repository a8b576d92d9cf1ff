"""Parity of the SDE BacksolveAdjoint kernels against the oracle, with IDENTICAL Wiener increments on both sides
(the increments are generated on the device by the Philox counter, exported through b200adj_get_noise, and fed to
the oracle).  Tolerance 1e-9 relative (fp64; BASELINE C5 asks 1e-8)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import scimlsensitivity_jl_b200 as b
from oracle import oracle as O


def _rel(a, ref):
    return np.abs(np.asarray(a) - ref).max() / max(np.abs(ref).max(), 1e-300)


@pytest.mark.parametrize("family,p", [("sde_lv", [1.5, 1.0, 3.0, 1.0, 0.1, 0.1]), ("sde_linear", [1.01, 0.87])])
@pytest.mark.parametrize("stepper", ["em", "euler_heun"])
@pytest.mark.parametrize("mode", ["regen", "stored", "explicit_noise"])
@pytest.mark.parametrize("shared_p", [True, False])
def test_sde_backsolve_parity(family, p, stepper, mode, shared_p):
    N, T, dt = 300, 1.0, 0.01
    S = 100
    saveat = np.linspace(0.0, T, 101)
    rng = np.random.default_rng(3)
    u0 = np.ones((2, N)) * np.exp(0.05 * rng.standard_normal((2, N)))
    p = np.array(p)
    if not shared_p:
        p = p[:, None] * np.exp(0.02 * rng.standard_normal((len(p), N)))
    eng = b.DeviceEnsemble(family, "backsolve", stepper, N, saveat, (0.0, T), dt, shared_p=shared_p,
                           cost=b.AffineCost(1.0, 0.0), seed=100, traj_offset=7, stored_noise=(mode == "stored"))
    dW_in = None
    if mode == "explicit_noise":
        dW_in = np.sqrt(dt) * rng.standard_normal((S, 2, N))
    saved, status = eng.forward(u0, p, dW=dW_in)
    dW = eng.noise()
    if dW_in is not None:
        assert np.array_equal(dW, dW_in)
    else:
        # the Philox/Box-Muller stream is N(0, dt): loose moment check
        assert abs(dW.mean()) < 5 * np.sqrt(dt / dW.size) and abs(dW.var() / dt - 1) < 0.02
    du0, dp = eng.reverse()
    cfg = O.make_cfg(family, "backsolve", stepper, N, saveat, 0.0, T, dt=dt, cost=("affine", 1.0, 0.0), shared_p=shared_p, d=2)
    ref = O.gradient(cfg, saveat, u0, p, dW=dW)
    assert np.abs(saved - ref["saved"]).max() < 1e-11
    assert _rel(du0, ref["du0"]) < 1e-9
    assert _rel(dp, ref["dp"]) < 1e-9
    eng.close()


@pytest.mark.parametrize("stepper", ["em", "euler_heun"])
@pytest.mark.parametrize("shared_p", [True, False])
def test_sde_interpolating_parity(stepper, shared_p):
    """SDEAdjointProblem for InterpolatingAdjoint (src/interpolating_adjoint.jl:453-613): z = [lam; mu], y from the saved
    forward solution, untransformed drift; same Wiener increments on both sides."""
    N, T, dt = 130, 1.0, 0.01
    saveat = np.linspace(0.0, T, 21)
    rng = np.random.default_rng(8)
    u0 = np.ones((2, N)) * np.exp(0.05 * rng.standard_normal((2, N)))
    p = np.array([1.5, 1.0, 3.0, 1.0, 0.1, 0.1])
    if not shared_p:
        p = p[:, None] * np.exp(0.02 * rng.standard_normal((6, N)))
    eng = b.DeviceEnsemble("sde_lv", "interpolating", stepper, N, saveat, (0.0, T), dt, shared_p=shared_p, cost=b.AffineCost(1.0, -1.0), seed=3)
    saved, _ = eng.forward(u0, p)
    dW = eng.noise()
    du0, dp = eng.reverse()
    cfg = O.make_cfg("sde_lv", "interpolating", stepper, N, saveat, 0.0, T, dt=dt, cost=("affine", 1.0, -1.0), shared_p=shared_p)
    ref = O.gradient(cfg, saveat, u0, p, dW=dW)
    assert _rel(du0, ref["du0"]) < 1e-9 and _rel(dp, ref["dp"]) < 1e-9
    # retarget the same forward pass to BacksolveAdjoint: a different algorithm (transformed drift for EM), different numbers
    eng.set_reverse("backsolve", cost=b.AffineCost(1.0, -1.0))
    du0b, dpb = eng.reverse()
    cfgb = O.make_cfg("sde_lv", "backsolve", stepper, N, saveat, 0.0, T, dt=dt, cost=("affine", 1.0, -1.0), shared_p=shared_p)
    refb = O.gradient(cfgb, saveat, u0, p, dW=dW)
    assert _rel(dpb, refb["dp"]) < 1e-9
    eng.close()


def test_sde_noise_is_shard_independent():
    """Philox streams are keyed by the GLOBAL member index: two shards reproduce the unsharded increments."""
    N, T, dt = 128, 0.1, 0.01
    saveat = np.array([0.0, T])
    u0 = np.ones((2, N)); p = np.array([1.5, 1.0, 3.0, 1.0, 0.1, 0.1])
    full = b.DeviceEnsemble("sde_lv", "backsolve", "em", N, saveat, (0.0, T), dt, seed=9)
    full.forward(u0, p); W = full.noise()
    h = N // 2
    for lo in (0, h):
        sh = b.DeviceEnsemble("sde_lv", "backsolve", "em", h, saveat, (0.0, T), dt, seed=9, traj_offset=lo)
        sh.forward(u0[:, lo:lo + h], p)
        assert np.array_equal(sh.noise(), W[:, :, lo:lo + h])
        sh.close()
    full.close()


def test_sde_closed_form_stratonovich():
    """Linear SDE, EulerHeun: gradients match the closed form sum_k t_k u_k^2, sum_k W_k u_k^2
    (test/SDE1/sde_stratonovich.jl:105-113, rtol 1e-3 there)."""
    N, T, dt = 64, 1.0, 1e-3
    S = 1000
    saveat = np.linspace(0.0, T, 11)
    u0 = np.ones((2, N)) * np.array([[1.0], [0.5]]); p = np.array([1.01, 0.87])
    eng = b.DeviceEnsemble("sde_linear", "backsolve", "euler_heun", N, saveat, (0.0, T), dt, shared_p=False,
                           cost=b.AffineCost(1.0, 0.0), seed=1)
    saved, _ = eng.forward(u0, np.repeat(p[:, None], N, 1))
    dW = eng.noise()
    du0, dp = eng.reverse()
    W = np.concatenate([np.zeros((1, 2, N)), np.cumsum(dW, 0)], 0)
    idx = np.round(saveat / dt).astype(int)
    uex = u0[None] * np.exp(p[0] * saveat[:, None, None] + p[1] * W[idx])
    dp0 = (saveat[:, None, None] * uex ** 2).sum((0, 1))
    dp1 = (W[idx] * uex ** 2).sum((0, 1))
    assert _rel(dp[0], dp0) < 5e-3 and _rel(dp[1], dp1) < 1e-2
    assert _rel(du0, (uex ** 2 / u0[None]).sum(0)) < 5e-3
    eng.close()


@pytest.mark.parametrize("stepper", ["em", "euler_heun"])
def test_sde_interpolating_no_start(stepper):
    """no_start skips the cotangent of the first save time for every sensealg but Backsolve (src/adjoint_common.jl:761): the
    SDE InterpolatingAdjoint path honours the flag, BacksolveAdjoint ignores it."""
    N, T, dt = 120, 1.0, 0.01
    saveat = np.linspace(0.0, T, 11)
    p = np.array([1.5, 1.0, 3.0, 1.0, 0.1, 0.1])
    u0 = np.ones((2, N))
    out = {}
    for sa in ("interpolating", "backsolve"):
        for ns in (False, True):
            eng = b.DeviceEnsemble("sde_lv", sa, stepper, N, saveat, (0.0, T), dt, cost=b.AffineCost(1.0, 0.5), seed=5, no_start=ns, stored_noise=True)
            eng.forward(u0, p)
            dW = eng.noise()
            du0, dp = eng.reverse()
            ref = O.gradient(O.make_cfg("sde_lv", sa, stepper, N, saveat, 0.0, T, dt=dt, cost=("affine", 1.0, 0.5), no_start=ns), saveat, u0, p, dW=np.asarray(dW))
            assert _rel(du0, ref["du0"]) < 1e-9 and _rel(dp, ref["dp"]) < 1e-9
            out[(sa, ns)] = np.array(du0)
            eng.close()
    # the skipped jump is exactly the cotangent at t0: a u0 + b = 1.5
    assert np.allclose(out[("interpolating", False)] - out[("interpolating", True)], 1.5, atol=1e-12)
    assert np.array_equal(out[("backsolve", False)], out[("backsolve", True)])
