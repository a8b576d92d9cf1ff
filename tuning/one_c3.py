import sys, os, numpy as np, torch, time
sys.path.insert(0, ".")
import scimlsensitivity_jl_b200 as b
N = int(os.environ.get("NMEM", "16384")); T = 100.0
rng = np.random.default_rng(20260923)
t = np.logspace(-2, 2, 10); t[-1] = T
u0 = np.repeat(np.array([[1.0], [0.0], [0.0]]), N, 1)
k = np.array([0.04, 3e7, 1e4])[:, None] * np.exp(0.05 * rng.standard_normal((3, N)))
kw = dict(abstol=1e-8, reltol=1e-8, quad_abstol=1e-10, quad_reltol=1e-10)
eng = b.DeviceEnsemble("robertson", os.environ.get("SA", "quadrature"), "rosenbrock23", N, t, (0.0, T), 0.0, shared_p=False, on_device=True, cost=b.AffineCost(1.0, 0.0), max_steps=8192, **kw)
u0d = torch.tensor(u0, device="cuda"); kd = torch.tensor(k, device="cuda")
for _ in range(2):
    eng.forward(u0d, kd, want_saved=False, want_status=False); du0, dp = eng.reverse()
torch.cuda.synchronize()
f, r = eng.step_counts()
print("fwd steps mean", float(np.mean(np.asarray(f.cpu() if hasattr(f, 'cpu') else f))), "rev steps mean", float(np.mean(np.asarray(r.cpu() if hasattr(r, 'cpu') else r))))
print(dp[:, :2])
