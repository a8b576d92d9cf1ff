import sys, numpy as np
sys.path.insert(0, ".")
import scimlsensitivity_jl_b200 as b
from oracle import oracle as O
for fam, u0v, pv, T in [("robertson",[1.0,0,0],[0.04,3e7,1e4],100.0), ("lorenz",[1.0,0,0],[10.0,28.0,8/3],1.0), ("lv",[1.0,1.0],[1.5,1.0,3.0,1.0],2.0)]:
    for saveat in [np.array([T]), np.array([0.5*T, T]), np.linspace(0.1*T, T, 10)]:
        N=2
        u0=np.repeat(np.array(u0v)[:,None],N,1); p=np.array(pv)
        cfg = O.make_cfg(fam, "gauss", "rosenbrock23", N, saveat, 0.0, T, abstol=1e-8, reltol=1e-8, cost=("affine",1.0,0.0))
        ref = O.gradient(cfg, saveat, u0, p)
        eng = b.DeviceEnsemble(fam, "gauss", "rosenbrock23", N, saveat, (0.0,T), 0.0, cost=b.AffineCost(1.0,0.0), abstol=1e-8, reltol=1e-8)
        saved, st = eng.forward(u0, p); du0, dp = eng.reverse()
        f, r = eng.step_counts()
        print(fam, len(saveat), "fwd", f[0], ref["steps"][0], "du0 gpu", du0[:,0], "ref", ref["du0"][:,0], "dp", dp, ref["dp"])
        eng.close()
