"""Parity of the CUDA path (through the C ABI) against the CPU oracle: ODE families, fixed-step Tsit5.

Tolerances (fp64): dp and du0 within 1e-8 relative of the oracle (north star asks 1e-6; observed ~1e-11), the
primal within 1e-10 absolute.  The oracle follows the reference's time-based interpolant lookup while the kernel
uses the step-aligned constant thetas, so agreement is to rounding amplified by the dynamics, not bitwise.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import scimlsensitivity_jl_b200 as b
from oracle import oracle as O

RTOL = 1e-8


def _rel(a, ref):
    return np.abs(np.asarray(a) - ref).max() / max(np.abs(ref).max(), 1e-300)


def _ensemble(family, N, seed=0):
    rng = np.random.default_rng(seed)
    if family == "lorenz":
        u0 = np.array([1.0, 0.0, 0.0])[:, None] + 0.1 * rng.standard_normal((3, N))
        p = np.array([10.0, 28.0, 8.0 / 3.0])
    elif family == "lv":
        u0 = np.array([1.0, 1.0])[:, None] * np.exp(0.1 * rng.standard_normal((2, N)))
        p = np.array([1.5, 1.0, 3.0, 1.0])
    else:
        u0 = np.array([1.0, 0.0, 0.0])[:, None] + 0.0 * rng.standard_normal((3, N))
        p = np.array([0.04, 3e2, 1e1])
    return u0, p


CASES = [
    # family, sensealg, T, dt, nsave, oracle kwargs / engine kwargs
    ("lorenz", "gauss", 10.0, 0.01, 101, {}),
    ("lorenz", "interpolating", 10.0, 0.01, 101, {}),
    ("lorenz", "backsolve", 2.0, 0.01, 21, {"ckpt_every_step": True}),
    ("lorenz", "backsolve", 2.0, 0.01, 21, {"ckpt_every_step": False}),
    ("lv", "gauss", 10.0, 0.05, 101, {}),
    ("lv", "interpolating", 10.0, 0.05, 101, {}),
    ("lv", "backsolve", 10.0, 0.05, 101, {"ckpt_every_step": False}),
    ("lorenz", "quadrature", 2.0, 0.01, 21, {}),
    ("lv", "quadrature", 10.0, 0.05, 101, {}),
    ("robertson", "gauss", 1.0, 0.001, 11, {}),
    ("robertson", "interpolating", 1.0, 0.001, 11, {}),
]


@pytest.mark.parametrize("family,sensealg,T,dt,nsave,kw", CASES)
@pytest.mark.parametrize("cost", ["affine", "explicit"])
@pytest.mark.parametrize("shared_p", [True, False])
def test_ode_parity(family, sensealg, T, dt, nsave, kw, cost, shared_p):
    N = 200   # not a multiple of the block size: exercises the ragged tail
    u0, p = _ensemble(family, N)
    saveat = np.linspace(0.0, T, nsave)
    if not shared_p:
        rng = np.random.default_rng(5)
        p = p[:, None] * np.exp(0.02 * rng.standard_normal((len(p), N)))
    every = kw.get("ckpt_every_step", False)
    ocost = ("affine", 1.0, -2.0)
    cfg = O.make_cfg(family, sensealg, "tsit5_fixed", N, saveat, 0.0, T, dt=dt, cost=ocost, shared_p=shared_p,
                     ckpt_every_step=every, quad_abstol=1e-9, quad_reltol=1e-9)
    ref = O.gradient(cfg, saveat, u0, p)
    eng = b.DeviceEnsemble(family, sensealg, "tsit5_fixed", N, saveat, (0.0, T), dt, shared_p=shared_p,
                           cost=b.AffineCost(1.0, -2.0) if cost == "affine" else None, ckpt_every_step=every,
                           quad_abstol=1e-9, quad_reltol=1e-9)
    saved, status = eng.forward(u0, p)
    assert (status == 0).all()
    assert np.abs(saved - ref["saved"]).max() <= 1e-10 * max(1.0, np.abs(ref["saved"]).max())
    dL = None if cost == "affine" else (saved - 2.0)
    du0, dp = eng.reverse(dL)
    assert _rel(du0, ref["du0"]) < RTOL
    # QuadratureAdjoint: adaptive Gauss-Kronrod is path dependent at the level of its own error estimate (1e-9 here)
    assert _rel(dp, ref["dp"]) < (1e-7 if sensealg == "quadrature" else RTOL)
    eng.close()


def test_device_buffers_and_reverse_retarget():
    """torch CUDA tensors in place of host arrays; one forward pass, several sensealgs on the same checkpoints."""
    import torch
    N, T, dt = 333, 5.0, 0.01
    u0, p = _ensemble("lorenz", N)
    saveat = np.linspace(0.0, T, 51)
    eng = b.DeviceEnsemble("lorenz", "gauss", "tsit5_fixed", N, saveat, (0.0, T), dt, on_device=True,
                           cost=b.AffineCost(1.0, -2.0))
    saved, status = eng.forward(torch.tensor(u0, device="cuda"), torch.tensor(p, device="cuda"))
    res = {}
    for sa in ["gauss", "interpolating", "backsolve", "quadrature"]:
        eng.set_reverse(sa, cost=b.AffineCost(1.0, -2.0), ckpt_every_step=True)
        eng.handle.set_tolerances(0.0, 0.0, 1e-9, 1e-9)
        du0, dp = eng.reverse()
        res[sa] = (du0.cpu().numpy(), dp.cpu().numpy())
        cfg = O.make_cfg("lorenz", sa, "tsit5_fixed", N, saveat, 0.0, T, dt=dt, cost=("affine", 1.0, -2.0), ckpt_every_step=True,
                         quad_abstol=1e-9, quad_reltol=1e-9)
        ref = O.gradient(cfg, saveat, u0, p, want_saved=False)
        assert _rel(res[sa][0], ref["du0"]) < RTOL and _rel(res[sa][1], ref["dp"]) < (1e-7 if sa == "quadrature" else RTOL)
    # the reference's own relation: all sensealgs agree (test/Core3/adjoint.jl:366-404), here at the dt=0.01 truncation level
    assert _rel(res["interpolating"][1], res["gauss"][1]) < 1e-5
    assert _rel(res["quadrature"][1], res["gauss"][1]) < 1e-5
    # fewer save times on the same forward pass
    t2 = saveat[::5]
    eng.set_reverse("gauss", cost=b.AffineCost(1.0, -2.0), t=t2)
    du0, dp = eng.reverse()
    cfg = O.make_cfg("lorenz", "gauss", "tsit5_fixed", N, t2, 0.0, T, dt=dt, cost=("affine", 1.0, -2.0))
    ref = O.gradient(cfg, t2, u0, p, want_saved=False)
    assert _rel(dp.cpu().numpy(), ref["dp"]) < RTOL
    eng.close()


def test_no_start_and_determinism():
    N, T, dt = 100, 1.0, 0.01
    u0, p = _ensemble("lv", N)
    saveat = np.linspace(0.0, T, 11)
    cfg = O.make_cfg("lv", "interpolating", "tsit5_fixed", N, saveat, 0.0, T, dt=dt, cost=("affine", 0.0, 1.0), no_start=True)
    ref = O.gradient(cfg, saveat, u0, p)
    eng = b.DeviceEnsemble("lv", "interpolating", "tsit5_fixed", N, saveat, (0.0, T), dt, cost=b.AffineCost(0.0, 1.0), no_start=True)
    eng.forward(u0, p)
    du0, dp = eng.reverse()
    assert _rel(du0, ref["du0"]) < RTOL and _rel(dp, ref["dp"]) < RTOL
    du0b, dpb = eng.reverse()
    assert np.array_equal(du0, du0b) and np.array_equal(dp, dpb)      # bitwise reproducible
    eng.close()


def test_full_size_properties():
    """BASELINE config C2 at full size (N=65536, T=10, dt=0.01): size-independent properties + a sampled oracle check."""
    import torch
    N, T, dt = 65536, 10.0, 0.01
    rng = np.random.default_rng(20260923)
    u0 = np.array([1.0, 0.0, 0.0])[:, None] + 0.1 * rng.standard_normal((3, N))
    p = np.array([10.0, 28.0, 8.0 / 3.0])
    saveat = np.linspace(0.0, T, 101)
    eng = b.DeviceEnsemble("lorenz", "gauss", "tsit5_fixed", N, saveat, (0.0, T), dt, on_device=True, cost=b.AffineCost(1.0, -2.0))
    saved, status = eng.forward(torch.tensor(u0, device="cuda"), torch.tensor(p, device="cuda"))
    assert int(status.sum()) == 0
    du0, dp = eng.reverse()
    du0, dp = du0.cpu().numpy(), dp.cpu().numpy()
    # (1) sampled members against the oracle
    idx = rng.choice(N, 128, replace=False)
    cfg = O.make_cfg("lorenz", "gauss", "tsit5_fixed", len(idx), saveat, 0.0, T, dt=dt, cost=("affine", 1.0, -2.0), shared_p=False)
    ref = O.gradient(cfg, saveat, u0[:, idx], np.repeat(p[:, None], len(idx), 1))
    assert _rel(du0[:, idx], ref["du0"]) < RTOL
    assert np.abs(saved[:, :, torch.tensor(idx, device="cuda")].cpu().numpy() - ref["saved"]).max() < 1e-9
    # (2) shared-p gradient == sum of per-member gradients (linearity of the reduction)
    eng2 = b.DeviceEnsemble("lorenz", "gauss", "tsit5_fixed", N, saveat, (0.0, T), dt, on_device=True, shared_p=False, cost=b.AffineCost(1.0, -2.0))
    eng2.forward(torch.tensor(u0, device="cuda"), torch.tensor(np.repeat(p[:, None], N, 1), device="cuda"))
    du0m, dpm = eng2.reverse()
    dpm = dpm.cpu().numpy()
    assert _rel(dpm[:, idx], ref["dp"]) < RTOL
    import math
    tot = np.array([math.fsum(dpm[q]) for q in range(3)])
    assert _rel(dp, tot) < 1e-12
    assert _rel(du0m.cpu().numpy(), du0) < 1e-12      # different template instantiation: same arithmetic, FMA contraction may differ
    # (3) explicit cotangent path == in-kernel affine cost
    eng.set_reverse("gauss", cost=None)
    du0e, dpe = eng.reverse(saved - 2.0)
    assert _rel(du0e.cpu().numpy(), du0) < 1e-12 and _rel(dpe.cpu().numpy(), dp) < 1e-12
    eng.close(); eng2.close()


# ---- fp32 variant of the fixed-step path ("C2 fp32, T = 1") -----------------------------------------
@pytest.mark.parametrize("family", ["lorenz", "lv"])
@pytest.mark.parametrize("sa", ["interpolating", "gauss", "backsolve"])
@pytest.mark.parametrize("shared_p", [True, False])
def test_fp32_variant_vs_fp64_oracle(family, sa, shared_p):
    """fp32 state / tables / tile, fp64 block reduction.  Tolerance: 2e-4 relative to the fp64 oracle at T = 1 (100 steps of
    fp32 rounding, eps = 6e-8, amplified by the adjoint's growth over the horizon; measured 1e-6 .. 3e-5)."""
    N, T, dt = 1000, 1.0, 0.01
    rng = np.random.default_rng(5)
    if family == "lorenz":
        u0 = np.array([1.0, 0.0, 0.0])[:, None] + 0.1 * rng.standard_normal((3, N)); p0 = np.array([10.0, 28.0, 8.0 / 3.0])
    else:
        u0 = 1.0 + 0.1 * rng.standard_normal((2, N)); p0 = np.array([1.5, 1.0, 3.0, 1.0])
    p = p0 if shared_p else p0[:, None] * (1.0 + 0.01 * rng.standard_normal((len(p0), N)))
    t = np.linspace(0.0, T, 11)
    for cost in (("affine", 1.0, -2.0), ("explicit",)):
        dL = None if cost[0] == "affine" else rng.standard_normal((len(t), u0.shape[0], N))
        eng = b.DeviceEnsemble(family, sa, "tsit5_fixed", N, t, (0.0, T), dt, shared_p=shared_p, dtype="f32",
                               cost=b.AffineCost(1.0, -2.0) if cost[0] == "affine" else None)
        saved, _ = eng.forward(u0.astype(np.float32), p.astype(np.float32))
        du0, dp = eng.reverse(None if dL is None else dL.astype(np.float32))
        assert saved.dtype == np.float32 and du0.dtype == np.float32 and dp.dtype == np.float32
        cfg = O.make_cfg(family, sa, "tsit5_fixed", N, t, 0.0, T, dt=dt, cost=cost, shared_p=shared_p)
        ref = O.gradient(cfg, t, u0.astype(np.float32).astype(np.float64), np.asarray(p, dtype=np.float32).astype(np.float64),
                         dLdu=None if dL is None else dL.astype(np.float32).astype(np.float64))
        assert _rel(saved, ref["saved"]) < 2e-5
        assert _rel(du0, ref["du0"]) < 2e-4, (cost[0], _rel(du0, ref["du0"]))
        assert _rel(dp, ref["dp"]) < 2e-4, (cost[0], _rel(dp, ref["dp"]))
        eng.close()


def test_fp32_rejects_quadrature_and_continuous_cost():
    t = np.linspace(0.0, 1.0, 11)
    with pytest.raises(Exception):
        b.DeviceEnsemble("lorenz", "quadrature", "tsit5_fixed", 64, t, (0.0, 1.0), 0.01, dtype="f32")
    with pytest.raises(Exception):
        b.DeviceEnsemble("robertson", "gauss", "tsit5_fixed", 64, t, (0.0, 1.0), 0.01, dtype="f32")


@pytest.mark.parametrize("sensealg,every", [("interpolating", False), ("gauss", False), ("gauss_kronrod", False), ("quadrature", False),
                                            ("backsolve", False), ("backsolve", True)])
def test_fixed_step_off_grid_save_times(sensealg, every):
    """Save / jump times OFF the dt grid with fixed-step Tsit5 (src/concrete_solve.jl:752-769: the primal is the dense forward
    solution interpolated at saveat; the jump times are tstops of the fixed-dt reverse solve, whose grid then shifts): the
    handle is routed to the dense per-member framework run with a constant step.  Device vs oracle, which is pinned against
    finite differences through the solve for exactly this case (test_oracle_relations.py)."""
    N, T, dt = 130, 3.0, 0.01
    rng = np.random.default_rng(41)
    u0 = np.exp(0.05 * rng.standard_normal((2, N)))
    p = np.array([1.5, 1.0, 3.0, 1.0])
    t = np.array([0.013, 0.5, 1.2345, 2.0, 2.999])
    kw = dict(quad_abstol=1e-12, quad_reltol=1e-12)
    eng = b.DeviceEnsemble("lv", sensealg, "tsit5_fixed", N, t, (0.0, T), dt, cost=b.AffineCost(1.0, -0.5), ckpt_every_step=every, **kw)
    saved, status = eng.forward(u0, p)
    du0, dp = eng.reverse()
    ref = O.gradient(O.make_cfg("lv", sensealg, "tsit5_fixed", N, t, 0.0, T, dt=dt, cost=("affine", 1.0, -0.5), ckpt_every_step=every, **kw), t, u0, p)
    assert int(np.asarray(status).sum()) == 0 and np.abs(np.asarray(saved) - ref["saved"]).max() < 1e-11
    tol = 1e-6 if sensealg == "gauss_kronrod" else 1e-8
    assert _rel(du0, ref["du0"]) < tol and _rel(dp, ref["dp"]) < tol, (_rel(du0, ref["du0"]), _rel(dp, ref["dp"]))
    eng.close()


def test_dense_forward_flag_allows_retargeting_to_off_grid_times():
    """adjoint_sensitivities(sol, t = off-grid times) on a grid-aligned handle needs the dense forward solution: without
    B200ADJ_FLAG_DENSE_FORWARD the C ABI says UNSUPPORTED, with it the reverse pass stops at the requested times."""
    N, T, dt = 33, 2.0, 0.01
    rng = np.random.default_rng(42)
    u0 = np.exp(0.05 * rng.standard_normal((2, N)))
    p = np.array([1.5, 1.0, 3.0, 1.0])
    grid, off = np.linspace(0.0, T, 5), np.array([0.333, 1.0, 1.777])
    eng = b.DeviceEnsemble("lv", "gauss", "tsit5_fixed", N, grid, (0.0, T), dt, cost=b.AffineCost(1.0, 0.0))
    eng.forward(u0, p)
    with pytest.raises(b.B200AdjError) as ei:
        eng.set_reverse("gauss", cost=b.AffineCost(1.0, 0.0), t=off)
    assert ei.value.code == -2
    eng.close()
    eng = b.DeviceEnsemble("lv", "gauss", "tsit5_fixed", N, grid, (0.0, T), dt, cost=b.AffineCost(1.0, 0.0), dense_forward=True)
    eng.forward(u0, p)
    eng.set_reverse("interpolating", cost=b.AffineCost(1.0, 0.0), t=off)
    du0, dp = eng.reverse()
    ref = O.gradient(O.make_cfg("lv", "interpolating", "tsit5_fixed", N, off, 0.0, T, dt=dt, cost=("affine", 1.0, 0.0)), off, u0, p)
    assert _rel(du0, ref["du0"]) < 1e-8 and _rel(dp, ref["dp"]) < 1e-8
    eng.close()
