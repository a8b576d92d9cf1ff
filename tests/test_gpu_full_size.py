"""BASELINE configs C3, C4, C5 at their FULL sizes: size-independent properties (shard invariance, linearity of the
gradient in the cotangent, additivity of the shared-parameter gradient over shards, bitwise reproducibility) plus a
sampled comparison with the oracle.  (C2 at full size: test_gpu_parity_ode.py::test_full_size_properties.)"""
import math
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


def test_c3_robertson_full_size():
    """C3: Robertson N = 16384, per-member k, Rosenbrock23 (1e-8), QuadratureAdjoint (1e-10)."""
    N, T = 16384, 100.0
    rng = np.random.default_rng(20260923)
    t = np.logspace(-2, 2, 10); t[-1] = T
    u0 = np.repeat(np.array([[1.0], [0.0], [0.0]]), N, 1)
    k = np.array([0.04, 3e7, 1e4])[:, None] * np.exp(0.05 * rng.standard_normal((3, N)))
    kw = dict(abstol=1e-8, reltol=1e-8, quad_abstol=1e-10, quad_reltol=1e-10)
    eng = b.DeviceEnsemble("robertson", "quadrature", "rosenbrock23", N, t, (0.0, T), 0.0, shared_p=False, cost=b.AffineCost(1.0, 0.0),
                           max_steps=8192, **kw)
    saved, status = eng.forward(u0, k)
    assert (np.asarray(status) == 0).all()
    du0, dp = eng.reverse()
    du0, dp = np.asarray(du0).copy(), np.asarray(dp).copy()
    assert np.isfinite(du0).all() and np.isfinite(dp).all()
    # conservation: y1 + y2 + y3 = 1 along every trajectory (d/dt sum = 0 for Robertson)
    assert np.abs(np.asarray(saved).sum(axis=1) - 1.0).max() < 1e-6
    # bitwise reproducible
    du0b, dpb = eng.reverse()
    assert np.array_equal(du0, np.asarray(du0b)) and np.array_equal(dp, np.asarray(dpb))
    # shard invariance: members are independent => a shard solved alone gives the same numbers
    lo, hi = 5000, 5000 + 1024
    sh = b.DeviceEnsemble("robertson", "quadrature", "rosenbrock23", hi - lo, t, (0.0, T), 0.0, shared_p=False, cost=b.AffineCost(1.0, 0.0),
                          max_steps=8192, **kw)
    sh.forward(u0[:, lo:hi], k[:, lo:hi])
    du0s, dps = sh.reverse()
    assert np.array_equal(np.asarray(du0s), du0[:, lo:hi]) and np.array_equal(np.asarray(dps), dp[:, lo:hi])
    sh.close()
    # sampled members against the oracle
    idx = rng.choice(N, 48, replace=False)
    ref = O.gradient(O.make_cfg("robertson", "quadrature", "rosenbrock23", len(idx), t, 0.0, T, cost=("affine", 1.0, 0.0), shared_p=False, **kw),
                     t, u0[:, idx], k[:, idx])
    assert _rel(du0[:, idx], ref["du0"]) < 1e-7
    err = np.abs(dp[:, idx] - ref["dp"]) / (np.abs(ref["dp"]).max(axis=1, keepdims=True))
    # BASELINE bound for C3: 1e-5.  Observed 3e-9 median / 3e-8 worst row: the device bisects the same segments in the same
    # order as the oracle (arg-max ties resolve to the lowest segment index on both sides)
    assert np.median(err) < 1e-7 and err.max() < 1e-6
    nrm = np.linalg.norm(dp[:, idx] - ref["dp"], axis=0) / np.linalg.norm(ref["dp"], axis=0)      # the norm quadgk controls
    assert nrm.max() < 1e-6
    eng.close()


def test_c4_mlp_full_size():
    """C4: MLP 2->64->64->2, N = 4096, InterpolatingAdjoint, fp32: linearity in the cotangent, shard additivity of dp."""
    N, T, dt, H = 4096, 1.5, 0.05, 64
    rng = np.random.default_rng(1)
    t = np.linspace(0.05, T, 30)
    u0 = rng.uniform(-2, 2, (2, N)).astype(np.float32)
    p = np.concatenate([(rng.standard_normal((H, 2)) / np.sqrt(2)).ravel(order="F"), 0.1 * rng.standard_normal(H),
                        (rng.standard_normal((H, H)) / np.sqrt(H)).ravel(order="F"), 0.1 * rng.standard_normal(H),
                        (rng.standard_normal((2, H)) / np.sqrt(H)).ravel(order="F"), 0.1 * rng.standard_normal(2)]).astype(np.float32)
    eng = b.DeviceEnsemble("mlp", "interpolating", "tsit5_fixed", N, t, (0.0, T), dt, dtype="f32")
    saved, status = eng.forward(u0, p)
    assert (np.asarray(status) == 0).all()
    d1 = rng.standard_normal((30, 2, N)).astype(np.float32); d2 = rng.standard_normal((30, 2, N)).astype(np.float32)
    g1 = [np.asarray(x).copy() for x in eng.reverse(d1)]
    g2 = [np.asarray(x).copy() for x in eng.reverse(d2)]
    g12 = [np.asarray(x).copy() for x in eng.reverse((2.0 * d1 - 0.5 * d2).astype(np.float32))]
    for a1, a2, a12 in zip(g1, g2, g12):       # the adjoint is linear in the cotangent
        assert _rel(a12, 2.0 * a1 - 0.5 * a2) < 2e-4
    # shared-parameter gradient is additive over shards; du0 of a shard is the shard of du0
    acc = np.zeros_like(g1[1], dtype=np.float64)
    for lo in range(0, N, 1024):
        sh = b.DeviceEnsemble("mlp", "interpolating", "tsit5_fixed", 1024, t, (0.0, T), dt, dtype="f32")
        sh.forward(u0[:, lo:lo + 1024], p)
        du0s, dps = sh.reverse(d1[:, :, lo:lo + 1024])
        assert _rel(np.asarray(du0s), g1[0][:, lo:lo + 1024]) < 1e-6
        acc += np.asarray(dps, dtype=np.float64)
        sh.close()
    assert _rel(acc, g1[1]) < 1e-4
    # sampled members against the fp64 oracle (per-member du0)
    idx = np.sort(rng.choice(N, 64, replace=False))
    ref = O.gradient(O.make_cfg("mlp", "interpolating", "tsit5_fixed", len(idx), t, 0.0, T, dt=dt, mlp_hidden=H), t, u0[:, idx].astype(np.float64),
                     p.astype(np.float64), dLdu=d1[:, :, idx].astype(np.float64))
    assert _rel(g1[0][:, idx], ref["du0"]) < 1e-4
    eng.close()


def test_c5_sde_full_size():
    """C5: SDE Lotka-Volterra, diagonal noise, N = 131072, EM dt = 0.01, BacksolveAdjoint, Philox noise regenerated."""
    N, T, dt = 131072, 1.0, 0.01
    t = np.linspace(0.0, T, 101)
    u0 = np.ones((2, N)); p = np.array([1.5, 1.0, 3.0, 1.0, 0.1, 0.1])
    eng = b.DeviceEnsemble("sde_lv", "backsolve", "em", N, t, (0.0, T), dt, cost=b.AffineCost(0.0, 1.0), seed=20260923)
    saved, status = eng.forward(u0, p)
    assert (np.asarray(status) == 0).all()
    du0, dp = [np.asarray(x).copy() for x in eng.reverse()]
    du0b, dpb = eng.reverse()
    assert np.array_equal(du0, np.asarray(du0b)) and np.array_equal(dp, np.asarray(dpb))        # bitwise reproducible
    # ensemble statistics: E[u(T)] of the Ito SDE with linear multiplicative noise equals the noise-free drift mean only
    # approximately; the size-independent check is the shard structure below.  Shards (global member offsets) reproduce
    # the unsharded run: du0 bitwise, dp additive
    acc = np.zeros(6)
    nsh = 4
    for r in range(nsh):
        lo, hi = r * N // nsh, (r + 1) * N // nsh
        sh = b.DeviceEnsemble("sde_lv", "backsolve", "em", hi - lo, t, (0.0, T), dt, cost=b.AffineCost(0.0, 1.0), seed=20260923, traj_offset=lo)
        sh.forward(u0[:, lo:hi], p)
        du0s, dps = sh.reverse()
        assert np.array_equal(np.asarray(du0s), du0[:, lo:hi])
        acc += np.asarray(dps)
        sh.close()
    assert _rel(acc, dp) < 1e-11
    # sampled members against the oracle with the same Wiener increments
    lo = 77777
    sm = b.DeviceEnsemble("sde_lv", "backsolve", "em", 256, t, (0.0, T), dt, cost=b.AffineCost(0.0, 1.0), seed=20260923, traj_offset=lo,
                          stored_noise=True, shared_p=False)
    pm = np.repeat(p[:, None], 256, 1)
    sm.forward(u0[:, lo:lo + 256], pm)
    dW = sm.noise()
    du0m, dpm = sm.reverse()
    assert _rel(np.asarray(du0m), du0[:, lo:lo + 256]) < 1e-12
    ref = O.gradient(O.make_cfg("sde_lv", "backsolve", "em", 256, t, 0.0, T, dt=dt, cost=("affine", 0.0, 1.0), shared_p=False), t,
                     u0[:, lo:lo + 256], pm, dW=dW)
    assert _rel(du0[:, lo:lo + 256], ref["du0"]) < 1e-9 and _rel(np.asarray(dpm), ref["dp"]) < 1e-9
    sm.close(); eng.close()
