"""Error metrics and timing of the bf16 MLP path (B200ADJ_MLP_TC=1: in-loop tensor cores, =0: tapes + dW2 GEMM)."""
import os, sys, numpy as np, torch
sys.path.insert(0, ".")
import scimlsensitivity_jl_b200 as b
from oracle import oracle as O
H = 64; P = 4482
def weights():
    rng = np.random.default_rng(1)
    return np.concatenate([(rng.standard_normal((H, 2)) / np.sqrt(2)).ravel(order="F"), 0.1 * rng.standard_normal(H),
                           (rng.standard_normal((H, H)) / np.sqrt(H)).ravel(order="F"), 0.1 * rng.standard_normal(H),
                           (rng.standard_normal((2, H)) / np.sqrt(H)).ravel(order="F"), 0.1 * rng.standard_normal(2)])
rel = lambda a, r: float(np.max(np.abs(np.asarray(a) - r)) / np.max(np.abs(r)))
T, dt = 1.5, 0.05; saveat = np.linspace(0.05, T, 30)
for N in (100, 4096):
    rng = np.random.default_rng(0)
    u0 = rng.uniform(-2, 2, (2, N)); p = weights()
    for cost in (("affine", 1.0, -0.5), ("explicit",)):
        dL = None if cost[0] == "affine" else rng.standard_normal((30, 2, N))
        cfg = O.make_cfg("mlp", "interpolating", "tsit5_fixed", N, saveat, 0.0, T, dt=dt, cost=cost, mlp_hidden=H)
        ref = O.gradient(cfg, saveat, u0, p, dLdu=dL)
        eng = b.DeviceEnsemble("mlp", "interpolating", "tsit5_fixed", N, saveat, (0.0, T), dt, dtype="bf16_f32acc",
                               cost=b.AffineCost(1.0, -0.5) if cost[0] == "affine" else None)
        saved, status = eng.forward(u0, p)
        du0, dp = eng.reverse(dL)
        blocks = {"W1": slice(0, 128), "b1": slice(128, 192), "W2": slice(192, 192 + 4096), "b2": slice(4288, 4352), "W3": slice(4352, 4480), "b3": slice(4480, 4482)}
        print(f"TC={os.environ.get('B200ADJ_MLP_TC','1')} N={N} cost={cost[0]} status={int(np.asarray(status).sum())} saved {rel(saved, ref['saved']):.2e} du0 {rel(du0, ref['du0']):.2e} dp {rel(dp, ref['dp']):.2e} | " +
              " ".join(f"{k} {rel(np.asarray(dp)[v], ref['dp'][v]):.2e}" for k, v in blocks.items()) +
              f" | rms W2 {np.sqrt(np.mean(((np.asarray(dp)[blocks['W2']] - ref['dp'][blocks['W2']]) / np.abs(ref['dp'][blocks['W2']]).max()) ** 2)):.2e}")
        eng.close()
# timing at N = 4096 on device buffers
N = 4096
rng = np.random.default_rng(0)
u0 = torch.tensor(rng.uniform(-2, 2, (2, N)), device="cuda", dtype=torch.float32); p = torch.tensor(weights(), device="cuda", dtype=torch.float32)
eng = b.DeviceEnsemble("mlp", "interpolating", "tsit5_fixed", N, saveat, (0.0, T), dt, on_device=True, dtype="bf16_f32acc", cost=b.AffineCost(1.0, -0.5))
for _ in range(3):
    eng.forward(u0, p, want_saved=False, want_status=False); eng.reverse()
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
tf = tr = 0.0
for _ in range(20):
    ev[0].record(); eng.forward(u0, p, want_saved=False, want_status=False); ev[1].record(); eng.reverse(); ev[2].record()
    torch.cuda.synchronize(); tf += ev[0].elapsed_time(ev[1]); tr += ev[1].elapsed_time(ev[2])
print(f"TC={os.environ.get('B200ADJ_MLP_TC','1')} N=4096 forward {tf/20:.3f} ms reverse(+reduce) {tr/20:.3f} ms -> {N/((tf+tr)/20*1e-3):.3e} traj/s")
