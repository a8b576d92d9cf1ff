"""User RHS families as plug-ins (the reference's user `vjp` / `vjp_p` seam, src/derivative_wrappers.jl:
284-359, test/Core3/user_vjp.jl:14-38 -- "user VJP vs ForwardDiff, rtol 1e-5").  The plug-in is built HERE by nvcc from a
header outside the library (examples/), registered, and driven through the same public API as the built-in families."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import scimlsensitivity_jl_b200 as b
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _rel(a, ref):
    return np.abs(np.asarray(a) - ref).max() / max(np.abs(ref).max(), 1e-300)


def _plugin(header, struct, name, has_jac):
    """The plug-ins are prebuilt in-tree by __graft_entry__.build() (examples/libb200fam_*.so travel with the snapshot); a
    stale one is refused by the library's ABI tag and rebuilt here."""
    so = os.path.join(ROOT, "examples", f"libb200fam_{name}.so")
    if os.path.exists(so):
        try:
            return b.register_family(so)
        except b.B200AdjError:
            os.remove(so)
    return b.register_family(b.build_family_plugin(os.path.join(ROOT, "examples", header), struct, name, out=so, has_jac=has_jac))


@pytest.fixture(scope="module")
def plugins():
    lv = _plugin("lv_clone_family.cuh", "LvClone", "lv_clone", True)
    vdp = _plugin("vanderpol_family.cuh", "VanDerPol", "vanderpol", False)
    assert lv[1:] == (2, 4) and vdp[1:] == (2, 2)
    return lv, vdp


@pytest.mark.parametrize("stepper,sensealg", [("tsit5_fixed", "gauss"), ("tsit5_fixed", "interpolating"), ("tsit5_fixed", "backsolve"),
                                              ("tsit5_fixed", "quadrature"), ("tsit5_adaptive", "interpolating"), ("tsit5_adaptive", "gauss_kronrod"),
                                              ("rosenbrock23", "gauss"), ("rosenbrock23", "interpolating")])
def test_plugin_family_reproduces_the_builtin_family(plugins, stepper, sensealg):
    """LV written as a plug-in: same kernels, same arithmetic => the same numbers as the built-in family, and the oracle's."""
    N, T = 77, 2.0
    rng = np.random.default_rng(3)
    u0 = np.exp(0.05 * rng.standard_normal((2, N)))
    p = np.array([1.5, 1.0, 3.0, 1.0])
    t = np.linspace(0.0, T, 9)
    kw = dict(abstol=1e-9, reltol=1e-9, quad_abstol=1e-11, quad_reltol=1e-11)
    dt = 0.01 if stepper == "tsit5_fixed" else 0.0
    res = {}
    for fam in ("lv", "lv_clone"):
        eng = b.DeviceEnsemble(fam, sensealg, stepper, N, t, (0.0, T), dt, cost=b.AffineCost(1.0, -0.5), **kw)
        saved, _ = eng.forward(u0, p)
        du0, dp = eng.reverse()
        res[fam] = (np.array(saved), np.array(du0), np.array(dp))
        eng.close()
    for x, y in zip(res["lv"], res["lv_clone"]):
        assert np.array_equal(x, y)
    ref = O.gradient(O.make_cfg("lv", sensealg, stepper, N, t, 0.0, T, dt=dt, cost=("affine", 1.0, -0.5), **kw), t, u0, p)
    assert _rel(res["lv_clone"][1], ref["du0"]) < 1e-6 and _rel(res["lv_clone"][2], ref["dp"]) < 1e-6


@pytest.mark.parametrize("sensealg", [b.InterpolatingAdjoint(), b.GaussAdjoint(), b.BacksolveAdjoint(), b.QuadratureAdjoint(abstol=1e-12, reltol=1e-12)])
def test_new_user_family_matches_finite_differences_through_the_solve(plugins, sensealg):
    """A family the library has never seen (van der Pol, d = 2, P = 2), through the public API: the adjoint gradient of
    L = sum_k sum_j (u_j(t_k)^2 / 2 - 0.3 u_j(t_k)) vs central differences of the device's own forward solve (the
    reference's user-VJP test compares with ForwardDiff at rtol 1e-5)."""
    N, T, dt = 5, 3.0, 0.005
    rng = np.random.default_rng(4)
    u0 = np.array([2.0, 0.0])[:, None] + 0.1 * rng.standard_normal((2, N))
    p = np.array([0.8, 1.2])[:, None] * np.exp(0.05 * rng.standard_normal((2, N)))
    t = np.linspace(0.0, T, 7)
    alg = b.Tsit5(dt=dt)

    def loss(u0_, p_):
        prob = b.EnsembleProblem(b.ODEProblem("vanderpol", u0_[:, 0], (0.0, T), p_[:, 0]), u0s=u0_, ps=p_)
        sol = b.solve(prob, alg, saveat=t)
        u = np.asarray(sol.u)
        return (0.5 * u ** 2 - 0.3 * u).sum(axis=(0, 1)), sol              # per member

    _, sol = loss(u0, p)
    du0, dp = b.adjoint_sensitivities(sol, alg, t=t, dgdu_discrete=b.AffineCost(1.0, -0.3), sensealg=sensealg, checkpoints=t)
    h = 1e-6
    for j in range(2):
        e = np.zeros((2, 1)); e[j] = h
        fd_u = (loss(u0 + e, p)[0] - loss(u0 - e, p)[0]) / (2 * h)
        fd_p = (loss(u0, p + e)[0] - loss(u0, p - e)[0]) / (2 * h)
        assert np.allclose(np.asarray(du0)[j], fd_u, rtol=2e-6, atol=1e-7)
        assert np.allclose(np.asarray(dp)[j], fd_p, rtol=2e-6, atol=1e-7)


def test_plugin_errors_are_reported():
    with pytest.raises(b.B200AdjError):
        b.register_family("/nonexistent/libnope.so")
    with pytest.raises(KeyError):
        b.solve(b.EnsembleProblem(b.ODEProblem("not_registered", np.ones(2), (0.0, 1.0), np.ones(2))), b.Tsit5(dt=0.1), trajectories=2, saveat=0.5)
