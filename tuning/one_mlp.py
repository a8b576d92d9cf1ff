import sys, numpy as np, torch
sys.path.insert(0, ".")
import scimlsensitivity_jl_b200 as b
H=64; N=4096; T,dt=1.5,0.05; saveat=np.linspace(0.05,T,30)
rng=np.random.default_rng(0)
u0=rng.uniform(-2,2,(2,N))
p=np.concatenate([(rng.standard_normal((H,2))/np.sqrt(2)).ravel(order="F"),0.1*rng.standard_normal(H),(rng.standard_normal((H,H))/np.sqrt(H)).ravel(order="F"),0.1*rng.standard_normal(H),(rng.standard_normal((2,H))/np.sqrt(H)).ravel(order="F"),0.1*rng.standard_normal(2)])
eng=b.DeviceEnsemble("mlp","interpolating","tsit5_fixed",N,saveat,(0.0,T),dt,on_device=True,dtype="bf16_f32acc",cost=b.AffineCost(1.0,-0.5))
u0d=torch.tensor(u0,device="cuda",dtype=torch.float32); pd=torch.tensor(p,device="cuda",dtype=torch.float32)
for _ in range(2):
    eng.forward(u0d,pd,want_saved=False,want_status=False); du0,dp=eng.reverse()
torch.cuda.synchronize(); print(dp[:4])
