#!/usr/bin/env python
"""Generate tests/golden/golden_v3.npz from the CPU oracle: the callback-family cases added after golden_v2 -- state-dependent
events (bouncing ball, parameter-dependent level + additive parameter affect on the Relax family, non-linear affect), the
"Dosing example" (preset-time affect that adds a parameter) and the hybrid neural ODE (preset-time kicks on the MLP family).
Like v1 / v2 these are ORACLE outputs (the reference cannot run here); they freeze the oracle against accidental change.  What
pins these code paths to the reference is tests/test_reference_held_numbers.py.
Run from the repo root:  python tests/golden/make_golden_v3.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

N = 4
TOL = dict(abstol=1e-10, reltol=1e-10)
BALL = dict(idx=0, level=0.0, direction=-1, pcomp=1, pparam=1, psign=-1.0)
RELAX = dict(idx=0, level=0.0, direction=0, lparam=0, lcoef=0.75, acomp=0, aparam=1, acoef=1.0)
SQ = dict(idx=0, direction=-1, shift=[3.0, 0.0], qcomp=1, qcoef=1.0)
SAS = ("interpolating", "gauss", "gauss_kronrod", "backsolve")


def mlp_params(rng, H=64):
    return np.concatenate([(rng.standard_normal((H, 2)) / np.sqrt(2)).ravel(order="F"), 0.1 * rng.standard_normal(H),
                           (rng.standard_normal((H, H)) / np.sqrt(H)).ravel(order="F"), 0.1 * rng.standard_normal(H),
                           (rng.standard_normal((2, H)) / np.sqrt(H)).ravel(order="F"), 0.1 * rng.standard_normal(2)])


def cases():
    rng = np.random.default_rng(2026)
    out = {}
    ub = np.stack([50.0 + 5.0 * rng.standard_normal(N), 0.5 * rng.standard_normal(N)])
    for sa in SAS:
        out[f"ball_{sa}"] = (dict(family="ball", sensealg=sa, stepper="tsit5_adaptive", saveat=np.linspace(0.5, 15.0, 30), T=15.0,
                                  kw=dict(TOL, crossing=BALL, ckpt_every_step=True, cost=("affine", 1.0, 0.0))), ub, np.array([9.8, 0.8]))
    ur = 40.0 * rng.random((1, N))
    for sa in SAS:
        out[f"relax_level_{sa}"] = (dict(family="relax", sensealg=sa, stepper="tsit5_adaptive", saveat=np.linspace(0.5, 6.0, 12), T=6.0,
                                         kw=dict(TOL, crossing=RELAX, ckpt_every_step=True, cost=("affine", 1.0, -2.0))), ur, np.array([100.0, 35.0]))
        out[f"relax_dosing_{sa}"] = (dict(family="relax", sensealg=sa, stepper="tsit5_adaptive", saveat=np.linspace(1.0, 10.0, 10), T=10.0,
                                          kw=dict(TOL, events=([3.0, 8.0], [[1.0], [0.5]], [[0.0], [1.0]]), event_padd=([0, 0], [1, 0], [1.0, -0.1]),
                                                  ckpt_every_step=True, cost=("affine", 1.0, -1.0))), ur, np.array([100.0, 35.0]))
    us = np.stack([5.0 + rng.random(N), 0.2 * rng.standard_normal(N)])
    for sa in SAS:
        out[f"ball_square_{sa}"] = (dict(family="ball", sensealg=sa, stepper="tsit5_adaptive", saveat=np.arange(0.0, 2.51, 0.5), T=2.5,
                                         kw=dict(abstol=1e-12, reltol=1e-12, crossing=SQ, ckpt_every_step=True, cost=("affine", 1.0, -1.0))), us, np.array([9.8, 0.8]))
    et = np.arange(0.5, 2.99, 0.5)
    ev = (et, np.ones((len(et), 2)), np.stack([0.2 * rng.random(len(et)), np.zeros(len(et))], 1))
    um = rng.uniform(-1, 1, (2, N))
    pm = mlp_params(rng)
    for sa in ("interpolating", "gauss"):
        out[f"hybrid_node_{sa}"] = (dict(family="mlp", sensealg=sa, stepper="tsit5_fixed", saveat=np.arange(0.25, 3.01, 0.25), T=3.0,
                                         kw=dict(dt=0.05, mlp_hidden=64, events=ev, cost=("affine", 1.0, -0.5))), um, pm)
    return out


def build():
    out = {}
    for name, (c, u0, p) in cases().items():
        cfg = O.make_cfg(c["family"], c["sensealg"], c["stepper"], N, c["saveat"], 0.0, c["T"], **c["kw"])
        r = O.gradient(cfg, c["saveat"], u0, p)
        out[name + "/saved"] = r["saved"]
        out[name + "/du0"] = r["du0"]
        out[name + "/dp"] = r["dp"]
        out[name + "/steps"] = r["steps"]
    return out


if __name__ == "__main__":
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v3.npz")
    np.savez_compressed(path, **build())
    print("wrote", path, os.path.getsize(path), "bytes")
