"""Small end-to-end runs for compute-sanitizer (memcheck / racecheck): WHAT = ode | mlp | adaptive | sde | r2 | r3."""
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
elif what == "r2":         # round-2 kernels: SEG (checkpoint_every), EV (events on the dt grid), GK, augmented Rosenbrock23, MLP Gauss, CC, dgdp
    N = 300; T, dt = 0.4, 0.01; t = np.linspace(0, T, 5)
    u0 = np.array([1.0, 0, 0])[:, None] + 0.1 * rng.standard_normal((3, N)); p = np.array([10.0, 28.0, 8 / 3])
    for sa, C in (("gauss", 4), ("interpolating", 8), ("gauss_kronrod", 1)):
        e = b.DeviceEnsemble("lorenz", sa, "tsit5_fixed", N, t, (0, T), dt, cost=b.AffineCost(1.0, -2.0), checkpoint_every=C)
        e.forward(u0, p); du0, dp = e.reverse(); print(what, "seg", sa, C, np.asarray(dp)); e.close()
    ul = 1 + 0.05 * rng.standard_normal((2, N)); pl = np.array([1.5, 1.0, 3.0, 1.0])
    for sa in ("gauss", "backsolve"):
        e = b.DeviceEnsemble("lv", sa, "tsit5_fixed", N, np.linspace(0, 1.0, 5), (0, 1.0), 0.01, cost=b.AffineCost(0.0, 1.0))
        e.set_events([0.3, 0.5], [[1, 1], [1, 1]], [[0.5, 0], [0, 0]], [[1.0] * 4, [2.0, 1, 1, 1]], [[0.0] * 4, [-0.5, 0, 0, 0]])
        e.forward(ul, pl); du0, dp = e.reverse(); print(what, "ev", sa, np.asarray(dp)); e.close()
    e = b.DeviceEnsemble("lv", "gauss", "tsit5_fixed", N, [0.013, 0.5, 0.97], (0, 1.0), 0.01, cost=b.AffineCost([2.0, 0.0], 0.0))
    e.set_reverse("gauss", cost=b.AffineCost([2.0, 0.0], 0.0), dgdp=b.ParamAffine([0.0] * 4, [1.0, 0, 0, 0]))
    e.forward(ul, pl); du0, dp = e.reverse(); print(what, "offgrid+dgdp", np.asarray(dp)); e.close()
    ts = np.logspace(-2, 1, 4); u0r = np.repeat(np.array([[1.0], [0], [0]]), 70, 1); k = np.array([0.04, 3e7, 1e4])
    for sa in ("interpolating", "backsolve"):
        e = b.DeviceEnsemble("robertson", sa, "rosenbrock23", 70, ts, (0, 10.0), 0.0, cost=b.AffineCost(1.0, 0.0), abstol=1e-6, reltol=1e-6)
        e.forward(u0r, k); du0, dp = e.reverse(); print(what, "ros-aug", sa, np.asarray(dp)); e.close()
    pm = 0.3 * rng.standard_normal(4482); um = rng.uniform(-2, 2, (2, 130)); tm = np.linspace(0.05, 0.15, 3)
    for dt_ in ("bf16_f32acc", "f32", "f64"):
        e = b.DeviceEnsemble("mlp", "gauss", "tsit5_fixed", 130, tm, (0, 0.15), 0.05, dtype=dt_, cost=b.AffineCost(1.0, -0.5))
        e.forward(um, pm); du0, dp = e.reverse(); print(what, "mlp-gauss", dt_, np.asarray(dp)[:3]); e.close()
    ub = np.stack([10.0 + rng.standard_normal(70), np.zeros(70)]); tb = np.linspace(0.5, 5.0, 10)
    for sa in ("interpolating", "gauss", "gauss_kronrod", "backsolve"):
        e = b.DeviceEnsemble("ball", sa, "tsit5_adaptive", 70, tb, (0, 5.0), 0.0, cost=b.AffineCost(1.0, 0.0), abstol=1e-8, reltol=1e-8)
        e.set_continuous_callback(b.ContinuousCallback(idx=0, direction=-1, p_comp=1, p_param=1, p_sign=-1.0, max_events=16))
        e.forward(ub, np.array([9.8, 0.8])); du0, dp = e.reverse(); print(what, "cc", sa, np.asarray(dp)); e.close()
elif what == "r3":         # re-entry kernels: Relax (d = 1 TMA records) with a parameter-dependent condition, quadratic affect, MLP events
    N = 70; tr = np.linspace(0.5, 4.0, 8)
    cb = b.ContinuousCallback(idx=0, direction=0, level_param=0, level_coef=0.75, add_comp=0, add_param=1, add_coef=1.0, max_events=4)
    for sa in ("interpolating", "gauss", "gauss_kronrod", "backsolve"):
        for sp in (True, False):
            pr = np.array([100.0, 50.0]) if sp else np.stack([100.0 + 5 * rng.random(N), 50.0 + rng.random(N)])
            e = b.DeviceEnsemble("relax", sa, "tsit5_adaptive", N, tr, (0, 4.0), 0.0, cost=b.AffineCost(1.0, 0.0), shared_p=sp, abstol=1e-8, reltol=1e-8, ckpt_every_step=True)
            e.set_continuous_callback(cb)
            e.forward(40.0 * rng.random((1, N)), pr); du0, dp = e.reverse(); print(what, "relax", sa, sp, np.asarray(dp).ravel()[:2]); e.close()
    ub = np.stack([5.0 + rng.random(N), np.zeros(N)]); tb = np.arange(0.0, 2.51, 0.5)
    for sa in ("interpolating", "backsolve"):
        e = b.DeviceEnsemble("ball", sa, "tsit5_adaptive", N, tb, (0, 2.5), 0.0, cost=b.AffineCost(1.0, -1.0), abstol=1e-8, reltol=1e-8)
        e.set_continuous_callback(b.ContinuousCallback(idx=0, direction=-1, shift=[3.0, 0.0], sq_comp=1, sq_coef=1.0, max_events=8))
        e.forward(ub, np.array([9.8, 0.8])); du0, dp = e.reverse(); print(what, "sq", sa, np.asarray(dp)); e.close()
    pm = 0.3 * rng.standard_normal(4482); um = rng.uniform(-2, 2, (2, 130)); tm = np.linspace(0.05, 0.3, 6)
    for dt_ in ("f32", "f64"):
        for sa in ("interpolating", "gauss"):
            e = b.DeviceEnsemble("mlp", sa, "tsit5_fixed", 130, tm, (0, 0.3), 0.05, dtype=dt_, cost=b.AffineCost(1.0, -0.5))
            e.set_events([0.1, 0.2], [[1, 1], [1, 1]], [[0.1, 0], [0.05, 0]])
            e.forward(um, pm); du0, dp = e.reverse(); print(what, "mlp-ev", dt_, sa, np.asarray(dp)[:3]); e.close()
else:
    N = 200; t = np.linspace(0, 0.1, 11); u0 = np.ones((2, N)); p = np.array([1.5, 1.0, 3.0, 1.0, 0.1, 0.1])
    for st in ("em", "euler_heun"):
        for sa in ("backsolve", "interpolating"):
            e = b.DeviceEnsemble("sde_lv", sa, st, N, t, (0, 0.1), 0.01, cost=b.AffineCost(0.0, 1.0), seed=3)
            e.forward(u0, p); du0, dp = e.reverse(); print(what, st, sa, np.asarray(dp)[:3]); e.close()
