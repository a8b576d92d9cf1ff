import sys, numpy as np
sys.path.insert(0, ".")
import scimlsensitivity_jl_b200 as b
from oracle import oracle as O
T=100.0; fam="robertson"
saveat = np.logspace(-2, 2, 10); saveat[-1] = T
N=100
rng=np.random.default_rng(0)
u0=np.repeat(np.array([1.0,0,0])[:,None],N,1); k=np.array([0.04,3e7,1e4])
p = k[:,None]*np.exp(0.05*rng.standard_normal((3,N)))
for sa in ["gauss", "quadrature"]:
    cfg = O.make_cfg(fam, sa, "rosenbrock23", N, saveat, 0.0, T, abstol=1e-8, reltol=1e-8, cost=("affine",1.0,0.0), shared_p=False, quad_abstol=1e-10, quad_reltol=1e-10)
    ref = O.gradient(cfg, saveat, u0, p)
    eng = b.DeviceEnsemble(fam, sa, "rosenbrock23", N, saveat, (0.0,T), 0.0, shared_p=False, cost=b.AffineCost(1.0,0.0), abstol=1e-8, reltol=1e-8, quad_abstol=1e-10, quad_reltol=1e-10, max_steps=8192)
    saved, st = eng.forward(u0, p); du0, dp = eng.reverse()
    f, r = eng.step_counts()
    e_u = np.abs(du0-ref["du0"]).max(0)/np.abs(ref["du0"]).max(0)
    e_p = np.abs(dp-ref["dp"])/np.abs(ref["dp"])
    print(sa, "rev steps min/max", r.min(), r.max(), "du0 err max", e_u.max(), "dp rel err per param max", e_p.max(1), "worst members", np.argsort(-e_p[0])[:5], np.sort(-e_p[0])[:5])
    eng.close()
