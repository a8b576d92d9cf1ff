#!/usr/bin/env python
"""Generate tests/golden/golden_v2.npz from the CPU oracle: the adaptive steppers, GaussKronrodAdjoint, preset-time events
(state and parameter affects), the continuous cost and the stiff QuadratureAdjoint case.  Like golden_v1 these are ORACLE
outputs (the reference cannot run here); they freeze the oracle against accidental change.
Run from the repo root:  python tests/golden/make_golden_v2.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

N = 4
LV_P = [1.5, 1.0, 3.0, 1.0]
EVENTS = ([2.03, 5.1], [[1.0, 1.0], [1.0, 1.0]], [[2.0, 0.0], [0.0, 0.0]], [[1.0] * 4, [2.0, 1.0, 0.5, 1.0]], [[0.0] * 4, [-0.5, 0.0, 0.1, 0.0]])
TOL = dict(abstol=1e-9, reltol=1e-9)
CASES = {
    # name: (family, sensealg, stepper, T, save times, p, make_cfg kwargs)
    **{f"lv_adaptive_{sa}": ("lv", sa, "tsit5_adaptive", 10.0, np.arange(0.0, 10.01, 0.5), LV_P, dict(TOL, ckpt_every_step=True))
       for sa in ("interpolating", "gauss", "gauss_kronrod", "quadrature", "backsolve")},
    **{f"lv_events_{sa}": ("lv", sa, "tsit5_adaptive", 10.0, np.arange(0.0, 10.01, 0.5), LV_P, dict(TOL, ckpt_every_step=True, events=EVENTS))
       for sa in ("interpolating", "gauss", "gauss_kronrod", "backsolve")},
    "lv_adaptive_contcost": ("lv", "interpolating", "tsit5_adaptive", 2.0, np.linspace(0.0, 2.0, 5), LV_P, dict(TOL, cont_cost=(1.0, -0.3))),
    "lorenz_adaptive_backsolve": ("lorenz", "backsolve", "tsit5_adaptive", 2.0, np.linspace(0.0, 2.0, 21), [10.0, 28.0, 8.0 / 3.0], dict(TOL, ckpt_every_step=True)),
    **{f"robertson_ros23_{sa}": ("robertson", sa, "rosenbrock23", 100.0, np.r_[np.logspace(-2, 2, 10)[:-1], 100.0], [0.04, 3e7, 1e4],
                                 dict(abstol=1e-8, reltol=1e-8, quad_abstol=1e-10, quad_reltol=1e-10))
       for sa in ("gauss", "gauss_kronrod", "quadrature")},
}


def inputs(family, name):
    rng = np.random.default_rng(sum(map(ord, name)))
    if family == "lorenz":
        return np.array([1.0, 0.0, 0.0])[:, None] + 0.05 * rng.standard_normal((3, N))
    if family == "robertson":
        return np.repeat(np.array([[1.0], [0.0], [0.0]]), N, 1)
    return np.ones((2, N)) * np.exp(0.05 * rng.standard_normal((2, N)))


def build():
    out = {}
    for name, (fam, sa, st, T, ts, p, kw) in CASES.items():
        u0 = inputs(fam, name)
        cost = ("affine", 1.0, 0.0) if fam == "robertson" else ("affine", 1.0, -2.0)
        r = O.gradient(O.make_cfg(fam, sa, st, N, ts, 0.0, T, cost=cost, **kw), ts, u0, np.array(p))
        out[name + "/u0"] = u0
        out[name + "/saved"] = r["saved"]
        out[name + "/du0"] = r["du0"]
        out[name + "/dp"] = r["dp"]
        out[name + "/steps"] = r["steps"]
    return out


if __name__ == "__main__":
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v2.npz")
    np.savez_compressed(path, **build())
    print("wrote", path, os.path.getsize(path), "bytes")
