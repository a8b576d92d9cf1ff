import sys, os, numpy as np, torch
sys.path.insert(0, ".")
import scimlsensitivity_jl_b200 as b, bench
N=65536; saveat=np.linspace(0,10,101)
u0,p=bench.make_inputs(N)
blk=int(os.environ.get("BLK","0"))
eng=b.DeviceEnsemble("lorenz","gauss","tsit5_fixed",N,saveat,(0.0,10.0),0.01,on_device=True,cost=b.AffineCost(1.0,-2.0),block_threads=blk)
u0d=torch.tensor(u0,device="cuda"); pd=torch.tensor(p,device="cuda")
for _ in range(2):
    eng.forward(u0d,pd,want_saved=False,want_status=False); du0,dp=eng.reverse()
torch.cuda.synchronize(); print(dp)
