"""Small end-to-end runs for compute-sanitizer (memcheck / racecheck): WHAT = ode | mlp | adaptive | sde."""
import sys, os, numpy as np
sys.path.insert(0, ".")
import scimlsensitivity_jl_b200 as b
what = sys.argv[1] if len(sys.argv) > 1 else "ode"
rng = np.random.default_rng(0)
if what == "ode":          # travelling warp groups active (160 slots per SM), Gauss + Interpolating, shared and per-member p
    N = 148 * 160 - 37; T, dt = 0.1, 0.01; t = np.linspace(0, T, 3)
    u0 = np.array([1.0, 0, 0])[:, None] + 0.1 * rng.standard_normal((3, N))
    for sa in ("gauss", "interpolating"):
        for sp in (True, False):
            p = np.array([10.0, 28.0, 8 / 3]); p = p if sp else np.repeat(p[:, None], N, 1)
            e = b.DeviceEnsemble("lorenz", sa, "tsit5_fixed", N, t, (0, T), dt, shared_p=sp, cost=b.AffineCost(1.0, -2.0))
            e.forward(u0, p); du0, dp = e.reverse(); print(what, sa, sp, np.asarray(dp).ravel()[:3]); e.close()
elif what == "mlp":
    H = 64; N = 130; T, dt = 0.15, 0.05; t = np.linspace(0.05, T, 3)
    p = 0.3 * rng.standard_normal(4482); u0 = rng.uniform(-2, 2, (2, N))
    for dt_ in ("bf16_f32acc", "f32"):
        e = b.DeviceEnsemble("mlp", "interpolating", "tsit5_fixed", N, t, (0, T), dt, dtype=dt_, cost=b.AffineCost(1.0, -0.5))
        e.forward(u0, p); du0, dp = e.reverse(); print(what, dt_, np.asarray(dp)[:3]); e.close()
elif what == "adaptive":
    N = 70; t = np.arange(0, 2.01, 0.5); u0 = 1 + 0.05 * rng.standard_normal((2, N)); p = np.array([1.5, 1.0, 3.0, 1.0])
    for sa in ("interpolating", "gauss", "gauss_kronrod", "backsolve", "quadrature"):
        e = b.DeviceEnsemble("lv", sa, "tsit5_adaptive", N, t, (0, 2.0), 0.0, cost=b.AffineCost(0.0, 1.0), abstol=1e-8, reltol=1e-8, ckpt_every_step=True)
        if sa != "quadrature": e.set_events([0.7, 1.0], [[1, 1], [1, 1]], [[0.5, 0], [0, 0]], [[1.0] * 4, [2.0, 1, 1, 1]], [[0.0] * 4, [-0.5, 0, 0, 0]])
        e.forward(u0, p); du0, dp = e.reverse(); print(what, sa, np.asarray(dp)); e.close()
    ts = np.logspace(-2, 1, 4); u0r = np.repeat(np.array([[1.0], [0], [0]]), N, 1); k = np.array([0.04, 3e7, 1e4])
    for sa in ("gauss", "gauss_kronrod", "quadrature"):
        e = b.DeviceEnsemble("robertson", sa, "rosenbrock23", N, ts, (0, 10.0), 0.0, cost=b.AffineCost(1.0, 0.0), abstol=1e-6, reltol=1e-6)
        e.forward(u0r, k); du0, dp = e.reverse(); print(what, "ros", sa, np.asarray(dp)); e.close()
else:
    N = 200; t = np.linspace(0, 0.1, 11); u0 = np.ones((2, N)); p = np.array([1.5, 1.0, 3.0, 1.0, 0.1, 0.1])
    for st in ("em", "euler_heun"):
        for sa in ("backsolve", "interpolating"):
            e = b.DeviceEnsemble("sde_lv", sa, st, N, t, (0, 0.1), 0.01, cost=b.AffineCost(0.0, 1.0), seed=3)
            e.forward(u0, p); du0, dp = e.reverse(); print(what, st, sa, np.asarray(dp)[:3]); e.close()
