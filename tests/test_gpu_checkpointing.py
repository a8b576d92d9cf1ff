"""Interval checkpointing (CheckpointSolution machinery of src/interpolating_adjoint.jl:20-27, 54-112,
206-278 and src/gauss_adjoint.jl:40-46, 57-95, 167-212): cfg.checkpoint_every = C keeps the forward state every C steps only
and the reverse kernel re-solves each segment from its left checkpoint.  The gradient must not depend on C."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import scimlsensitivity_jl_b200 as b
from oracle import oracle as O


def _rel(a, ref):
    return np.abs(np.asarray(a) - ref).max() / max(np.abs(ref).max(), 1e-300)


@pytest.mark.parametrize("sensealg", ["interpolating", "gauss"])
@pytest.mark.parametrize("C", [2, 7, 8, 16, 1000])
def test_checkpoint_every_matches_oracle_and_every_step(sensealg, C):
    """S = 200 steps: C = 8 divides S, C = 7 leaves a short last segment, C = 1000 > S is one segment (clamped to S)."""
    N, T, dt = 1000, 2.0, 0.01
    if C == 1000:
        N = 96                                   # one segment of 200 states per slot: small blocks keep it inside shared memory
    saveat = np.linspace(0.0, T, 21)
    rng = np.random.default_rng(11)
    u0 = np.array([1.0, 0.0, 0.0])[:, None] + 0.1 * rng.standard_normal((3, N))
    p = np.array([10.0, 28.0, 8.0 / 3.0])
    ref = O.gradient(O.make_cfg("lorenz", sensealg, "tsit5_fixed", N, saveat, 0.0, T, dt=dt, cost=("affine", 1.0, -2.0)), saveat, u0, p)
    out = {}
    for c in (1, C):
        eng = b.DeviceEnsemble("lorenz", sensealg, "tsit5_fixed", N, saveat, (0.0, T), dt, cost=b.AffineCost(1.0, -2.0),
                               checkpoint_every=c, block_threads=32 if C == 1000 else 0)
        saved, status = eng.forward(u0, p)
        du0, dp = eng.reverse()
        assert int(np.asarray(status).sum()) == 0
        assert np.abs(np.asarray(saved) - ref["saved"]).max() < 1e-10
        assert _rel(du0, ref["du0"]) < 1e-8 and _rel(dp, ref["dp"]) < 1e-8
        du0b, dpb = eng.reverse()                # a second reverse pass on the same checkpoints
        assert np.array_equal(np.asarray(du0), np.asarray(du0b)) and np.array_equal(np.asarray(dp), np.asarray(dpb))
        out[c] = (np.array(du0), np.array(dp))
        eng.close()
    assert _rel(out[C][0], out[1][0]) < 1e-11 and _rel(out[C][1], out[1][1]) < 1e-11


def test_checkpoint_every_explicit_cotangent_per_member_p_and_fp32():
    N, T, dt, C = 333, 1.0, 0.01, 8
    saveat = np.linspace(0.0, T, 11)
    rng = np.random.default_rng(2)
    u0 = np.exp(0.1 * rng.standard_normal((2, N)))
    p = np.array([1.5, 1.0, 3.0, 1.0])[:, None] * np.exp(0.02 * rng.standard_normal((4, N)))
    dL = rng.standard_normal((11, 2, N))
    ref = O.gradient(O.make_cfg("lv", "interpolating", "tsit5_fixed", N, saveat, 0.0, T, dt=dt, shared_p=False), saveat, u0, p, dLdu=dL)
    eng = b.DeviceEnsemble("lv", "interpolating", "tsit5_fixed", N, saveat, (0.0, T), dt, shared_p=False, checkpoint_every=C)
    eng.forward(u0, p)
    du0, dp = eng.reverse(dL)
    assert _rel(du0, ref["du0"]) < 1e-8 and _rel(dp, ref["dp"]) < 1e-8
    eng.close()
    eng = b.DeviceEnsemble("lv", "gauss", "tsit5_fixed", N, saveat, (0.0, T), dt, shared_p=False, checkpoint_every=C, dtype="f32")
    eng.forward(u0, p)
    du0, dp = eng.reverse(dL)
    refg = O.gradient(O.make_cfg("lv", "gauss", "tsit5_fixed", N, saveat, 0.0, T, dt=dt, shared_p=False), saveat, u0, p, dLdu=dL)
    assert _rel(du0, refg["du0"]) < 2e-4 and _rel(dp, refg["dp"]) < 2e-4
    eng.close()


def test_checkpoint_every_rejected_where_the_reference_has_no_checkpointing():
    saveat = np.linspace(0.0, 1.0, 11)
    for sa in ("backsolve", "quadrature"):
        with pytest.raises(b.B200AdjError) as ei:
            b.DeviceEnsemble("lorenz", sa, "tsit5_fixed", 64, saveat, (0.0, 1.0), 0.01, checkpoint_every=8)
        assert ei.value.code == -2
    eng = b.DeviceEnsemble("lorenz", "gauss", "tsit5_fixed", 64, saveat, (0.0, 1.0), 0.01, checkpoint_every=8, cost=b.AffineCost(1.0, 0.0))
    with pytest.raises(b.B200AdjError):
        eng.set_reverse("backsolve", cost=b.AffineCost(1.0, 0.0))
    eng.close()
    with pytest.raises(b.B200AdjError) as ei:       # 200 states x 3 x 448 slots x 8 B does not fit in shared memory
        b.DeviceEnsemble("lorenz", "gauss", "tsit5_fixed", 65536, saveat, (0.0, 1.0), 0.005, checkpoint_every=200)
    assert ei.value.code == -1
