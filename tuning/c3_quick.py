"""C3 (Robertson / Rosenbrock23 / QuadratureAdjoint) at N members: phase times and parity of a sample against the oracle."""
import sys, os, time, numpy as np, torch
sys.path.insert(0, ".")
import scimlsensitivity_jl_b200 as b
from oracle import oracle as O
N = int(os.environ.get("NMEM", "16384")); T = 100.0
rng = np.random.default_rng(20260923)
saveat = np.logspace(-2, 2, 10); saveat[-1] = T
u0 = np.repeat(np.array([[1.0], [0.0], [0.0]]), N, 1)
p = np.array([0.04, 3e7, 1e4])[:, None] * np.exp(0.05 * rng.standard_normal((3, N)))
kw = dict(abstol=1e-8, reltol=1e-8, quad_abstol=1e-10, quad_reltol=1e-10)
eng = b.DeviceEnsemble("robertson", "quadrature", "rosenbrock23", N, saveat, (0.0, T), 0.0, shared_p=False, on_device=True,
                       cost=b.AffineCost(1.0, 0.0), max_steps=8192, block_threads=int(os.environ.get("BLOCK", "0")), **kw)
u0d = torch.tensor(u0, device="cuda"); pd = torch.tensor(p, device="cuda")
du0 = torch.empty_like(u0d); dp = torch.empty_like(pd)
for _ in range(2):
    eng.handle.forward(u0d, pd, None, None, None); eng.handle.reverse(None, du0, dp)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
R = 5; tf = tr = 0.0
for _ in range(R):
    ev[0].record(); eng.handle.forward(u0d, pd, None, None, None); ev[1].record(); eng.handle.reverse(None, du0, dp); ev[2].record()
    torch.cuda.synchronize(); tf += ev[0].elapsed_time(ev[1]); tr += ev[1].elapsed_time(ev[2])
print(f"C3 N={N}: forward {tf/R:.2f} ms, reverse (adjoint solve + quadgk) {tr/R:.2f} ms -> {N/((tf+tr)/R*1e-3):.3e} members/s")
idx = np.sort(rng.choice(N, min(N, 64), replace=False))
ref = O.gradient(O.make_cfg("robertson", "quadrature", "rosenbrock23", len(idx), saveat, 0.0, T, cost=("affine", 1.0, 0.0), shared_p=False, **kw),
                 saveat, u0[:, idx], p[:, idx], want_saved=False)
d = dp.cpu().numpy()[:, idx]; r = ref["dp"]
nrm = np.linalg.norm(d - r, axis=0) / np.linalg.norm(r, axis=0)
row = np.abs(d - r) / np.abs(r).max(axis=1, keepdims=True)
print(f"dp per-member 2-norm rel err: median {np.median(nrm):.2e} max {nrm.max():.2e}; per-row rel err: median {np.median(row):.2e} max {row.max():.2e}; du0 {np.abs(du0.cpu().numpy()[:, idx]-ref['du0']).max()/np.abs(ref['du0']).max():.2e}")
