"""C2 with interval checkpointing (cfg.checkpoint_every = C): time and checkpoint memory per gradient, parity vs C = 1."""
import sys, numpy as np, torch
sys.path.insert(0, ".")
import scimlsensitivity_jl_b200 as b, bench
N = 65536; saveat = np.linspace(0, 10, 101)
u0, p = bench.make_inputs(N)
u0d = torch.tensor(u0, device="cuda"); pd = torch.tensor(p, device="cuda")
ref = None
for C in (1, 4, 8):
    torch.cuda.synchronize(); free0 = torch.cuda.mem_get_info()[0]
    eng = b.DeviceEnsemble("lorenz", "gauss", "tsit5_fixed", N, saveat, (0.0, 10.0), 0.01, on_device=True, cost=b.AffineCost(1.0, -2.0), checkpoint_every=C)
    mem = (free0 - torch.cuda.mem_get_info()[0]) / 1e9
    for _ in range(3):
        eng.forward(u0d, pd, want_saved=False, want_status=False); du0, dp = eng.reverse()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    tf = tr = 0.0; R = 20
    for _ in range(R):
        ev[0].record(); eng.forward(u0d, pd, want_saved=False, want_status=False); ev[1].record(); du0, dp = eng.reverse(); ev[2].record()
        torch.cuda.synchronize(); tf += ev[0].elapsed_time(ev[1]); tr += ev[1].elapsed_time(ev[2])
    dpc = dp.cpu().numpy()
    if ref is None:
        ref = dpc
    print(f"checkpoint_every={C:2d}: device memory {mem:.3f} GB, forward {tf/R:.3f} ms, reverse {tr/R:.3f} ms, total {(tf+tr)/R:.3f} ms, "
          f"dp rel diff vs C=1 {np.abs(dpc-ref).max()/np.abs(ref).max():.1e}")
    eng.close()
