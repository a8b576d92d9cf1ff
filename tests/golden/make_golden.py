#!/usr/bin/env python
"""Generate tests/golden/golden_v1.npz from the CPU oracle.

The reference (Julia) cannot run in this image and ships no golden vectors for the continuous-adjoint path, so these
fixtures are ORACLE outputs (the oracle itself is pinned by tests/test_oracle_relations.py).  They freeze the oracle
against accidental change (CPU test) and give the device path a second, stored target (GPU test).
Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

CASES = {
    # name: (family, sensealg, stepper, T, dt, nsave, p, oracle kwargs)
    "lorenz_gauss": ("lorenz", "gauss", "tsit5_fixed", 2.0, 0.01, 21, [10.0, 28.0, 8.0 / 3.0], {}),
    "lorenz_interp": ("lorenz", "interpolating", "tsit5_fixed", 2.0, 0.01, 21, [10.0, 28.0, 8.0 / 3.0], {}),
    "lorenz_backsolve": ("lorenz", "backsolve", "tsit5_fixed", 2.0, 0.01, 21, [10.0, 28.0, 8.0 / 3.0], {"ckpt_every_step": True}),
    "lv_gauss": ("lv", "gauss", "tsit5_fixed", 10.0, 0.05, 101, [1.5, 1.0, 3.0, 1.0], {}),
    "lv_interp": ("lv", "interpolating", "tsit5_fixed", 10.0, 0.05, 101, [1.5, 1.0, 3.0, 1.0], {}),
    "sdelv_em": ("sde_lv", "backsolve", "em", 1.0, 0.01, 101, [1.5, 1.0, 3.0, 1.0, 0.1, 0.1], {}),
    "sdelv_eh": ("sde_lv", "backsolve", "euler_heun", 1.0, 0.01, 101, [1.5, 1.0, 3.0, 1.0, 0.1, 0.1], {}),
}
N = 8


def inputs(family, name):
    rng = np.random.default_rng(abs(hash(name)) % 2 ** 31 if False else sum(map(ord, name)))
    if family == "lorenz":
        u0 = np.array([1.0, 0.0, 0.0])[:, None] + 0.1 * rng.standard_normal((3, N))
    else:
        u0 = np.ones((2, N)) * np.exp(0.1 * rng.standard_normal((2, N)))
    return u0, rng


def build():
    out = {}
    for name, (fam, sa, st, T, dt, nsave, p, kw) in CASES.items():
        u0, rng = inputs(fam, name)
        saveat = np.linspace(0.0, T, nsave)
        dW = None
        if st in ("em", "euler_heun"):
            dW = np.sqrt(dt) * rng.standard_normal((int(round(T / dt)), 2, N))
        cfg = O.make_cfg(fam, sa, st, N, saveat, 0.0, T, dt=dt, cost=("affine", 1.0, -2.0), **kw)
        r = O.gradient(cfg, saveat, u0, np.array(p), dW=dW)
        out[name + "/u0"] = u0
        out[name + "/p"] = np.array(p)
        out[name + "/saveat"] = saveat
        if dW is not None:
            out[name + "/dW"] = dW
        out[name + "/saved"] = r["saved"]
        out[name + "/du0"] = r["du0"]
        out[name + "/dp"] = r["dp"]
    return out


if __name__ == "__main__":
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v1.npz")
    np.savez_compressed(path, **build())
    print("wrote", path, os.path.getsize(path), "bytes")
