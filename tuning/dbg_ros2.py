import sys, numpy as np
sys.path.insert(0, ".")
import scimlsensitivity_jl_b200 as b
from oracle import oracle as O
T=100.0; fam="robertson"
saveat = np.logspace(-2, 2, 10); saveat[-1] = T
for N, shared in [(2, True), (2, False), (100, False)]:
  for sa in ["gauss", "quadrature"]:
    rng=np.random.default_rng(0)
    u0=np.repeat(np.array([1.0,0,0])[:,None],N,1); k=np.array([0.04,3e7,1e4])
    p = k if shared else k[:,None]*np.exp(0.05*rng.standard_normal((3,N)))
    cfg = O.make_cfg(fam, sa, "rosenbrock23", N, saveat, 0.0, T, abstol=1e-8, reltol=1e-8, cost=("affine",1.0,0.0), shared_p=shared, quad_abstol=1e-10, quad_reltol=1e-10)
    ref = O.gradient(cfg, saveat, u0, p)
    eng = b.DeviceEnsemble(fam, sa, "rosenbrock23", N, saveat, (0.0,T), 0.0, shared_p=shared, cost=b.AffineCost(1.0,0.0), abstol=1e-8, reltol=1e-8, quad_abstol=1e-10, quad_reltol=1e-10)
    saved, st = eng.forward(u0, p); du0, dp = eng.reverse()
    f, r = eng.step_counts()
    print(N, shared, sa, "fwd", f[:3], ref["steps"][:3], "rev", r[:3], "du0 gpu", du0[:,0], "ref", ref["du0"][:,0], "dp", np.asarray(dp).reshape(3,-1)[:,0], np.asarray(ref["dp"]).reshape(3,-1)[:,0])
    eng.close()
