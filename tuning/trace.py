import sys, os, numpy as np, torch
sys.path.insert(0, ".")
import scimlsensitivity_jl_b200 as b, bench
N=65536; saveat=np.linspace(0,10,101)
u0,p=bench.make_inputs(N)
for blk in (0,448,512,256,128):
    eng=b.DeviceEnsemble("lorenz","gauss","tsit5_fixed",N,saveat,(0.0,10.0),0.01,on_device=True,cost=b.AffineCost(1.0,-2.0),block_threads=blk,trace=True)
    u0d=torch.tensor(u0,device="cuda"); pd=torch.tensor(p,device="cuda")
    for _ in range(3):
        eng.forward(u0d,pd,want_saved=False,want_status=False); du0,dp=eng.reverse()
    torch.cuda.synchronize()
    tr=eng.handle.block_trace().astype(np.int64)
    t0=tr[:,1].min(); st=(tr[:,1]-t0)/1e3; en=(tr[:,2]-t0)/1e3
    sm=tr[:,0]; cnt=np.bincount(sm,minlength=148)
    print(f"block={blk} grid={len(tr)} SMs used={np.count_nonzero(cnt)} blocks/SM hist={np.bincount(cnt)}; start us: min {st.min():.1f} max {st.max():.1f}; end us: min {en.min():.1f} max {en.max():.1f}; dur us: min {(en-st).min():.1f} mean {(en-st).mean():.1f} max {(en-st).max():.1f}")
    late=(st>50).sum(); print("  blocks starting >50us late:", late)
    for c in sorted(set(cnt)):
        m=np.isin(sm,np.where(cnt==c)[0]); 
        if m.any(): print(f"  SMs with {c} blocks: mean dur {(en-st)[m].mean():.1f} us, last end {en[m].max():.1f}")
    eng.close()
