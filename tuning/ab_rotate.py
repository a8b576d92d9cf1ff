"""A/B timing of the reverse kernel (C2): forward/reverse ms via CUDA events, 20 reps each."""
import sys, os, numpy as np, torch
sys.path.insert(0, ".")
import scimlsensitivity_jl_b200 as b, bench
N = int(os.environ.get("NMEM", "65536")); saveat = np.linspace(0, 10, 101)
u0, p = bench.make_inputs(N)
sa = os.environ.get("SA", "gauss")
eng = b.DeviceEnsemble("lorenz", sa, "tsit5_fixed", N, saveat, (0.0, 10.0), 0.01, on_device=True, cost=b.AffineCost(1.0, -2.0), no_rotate=bool(int(os.environ.get("NO_ROTATE", "0"))))
u0d = torch.tensor(u0, device="cuda"); pd = torch.tensor(p, device="cuda")
for _ in range(3):
    eng.forward(u0d, pd, want_saved=False, want_status=False); du0, dp = eng.reverse()
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
tf = tr = 0.0
R = 20
for _ in range(R):
    ev[0].record(); eng.forward(u0d, pd, want_saved=False, want_status=False); ev[1].record(); du0, dp = eng.reverse(); ev[2].record()
    torch.cuda.synchronize(); tf += ev[0].elapsed_time(ev[1]); tr += ev[1].elapsed_time(ev[2])
print(f"NO_ROTATE={os.environ.get("NO_ROTATE","0")} SA={sa} N={N} fwd {tf/R:.4f} ms rev {tr/R:.4f} ms dp {dp.cpu().numpy()}")
