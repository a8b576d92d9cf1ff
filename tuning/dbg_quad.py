import sys, numpy as np, torch
sys.path.insert(0, ".")
import scimlsensitivity_jl_b200 as b
from oracle import oracle as O
def run(N,T,nsave,on_device,first=None):
    dt=0.01; rng=np.random.default_rng(0)
    u0=np.array([1.0,0,0])[:,None]+0.1*rng.standard_normal((3,N)); p=np.array([10.0,28.0,8/3])
    saveat=np.linspace(0,T,nsave)
    eng=b.DeviceEnsemble("lorenz","gauss","tsit5_fixed",N,saveat,(0.0,T),dt,on_device=on_device,cost=b.AffineCost(1.0,-2.0))
    if on_device: eng.forward(torch.tensor(u0,device="cuda"),torch.tensor(p,device="cuda"))
    else: eng.forward(u0,p)
    out={}
    for sa in (first or [])+["quadrature"]:
        eng.set_reverse(sa,cost=b.AffineCost(1.0,-2.0),ckpt_every_step=True); eng.handle.set_tolerances(0,0,1e-9,1e-9)
        du0,dp=eng.reverse(); out[sa]=np.asarray(dp.cpu().numpy() if on_device else dp)
    cfg=O.make_cfg("lorenz","quadrature","tsit5_fixed",N,saveat,0.0,T,dt=dt,cost=("affine",1.0,-2.0),quad_abstol=1e-9,quad_reltol=1e-9)
    ref=O.gradient(cfg,saveat,u0,p,want_saved=False)
    print(N,T,nsave,on_device,first,"rel err",np.abs(out["quadrature"]-ref["dp"]).max()/np.abs(ref["dp"]).max())
    eng.close()
run(200,2.0,21,False); run(200,5.0,51,False); run(333,5.0,51,False); run(333,5.0,51,True); run(333,5.0,51,True,["gauss","interpolating","backsolve"]); run(333,2.0,21,True,["backsolve"])
