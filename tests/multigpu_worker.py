"""Run under torchrun (one rank per GPU): the public API shards the ensemble over ranks and all-reduces dp; rank 0
checks the result against the unsharded oracle.  Used by tests/test_gpu_multigpu.py."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import scimlsensitivity_jl_b200 as b  # noqa: E402
from oracle import oracle as O  # noqa: E402


def main():
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    rank, world = dist.get_rank(), dist.get_world_size()
    N, T, dt = 1000, 2.0, 0.01
    rng = np.random.default_rng(0)
    u0 = np.array([1.0, 0.0, 0.0])[:, None] + 0.1 * rng.standard_normal((3, N))
    p = np.array([10.0, 28.0, 8.0 / 3.0])
    t = np.linspace(0, T, 21)
    prob = b.EnsembleProblem(b.ODEProblem("lorenz", u0[:, 0], (0.0, T), p), u0s=u0)
    sol = b.solve(prob, b.Tsit5(dt=dt), b.EnsembleB200(device=local), saveat=t)
    lo, hi = b.shard_bounds(N)
    assert sol.u.shape == (21, 3, hi - lo)
    du0, dp = b.adjoint_sensitivities(sol, b.Tsit5(dt=dt), t=t, dgdu_discrete=b.AffineCost(1.0, -2.0), sensealg=b.GaussAdjoint())
    # the all-reduce ran behind the C ABI (b200adj_comm_init + ncclAllReduce inside b200adj_reverse), not in torch.distributed
    assert sol.engine.comm_attached and sol.engine.handle.comm_size == (world, rank)
    fused = sol.engine.handle.comm_is_fused       # peer-memory mailboxes mapped: the reverse kernel reduced dp itself
    # the same gradient through ncclAllReduce (flag) must agree to the last bits of a 2-term sum
    eng_n = b.DeviceEnsemble("lorenz", "gauss", "tsit5_fixed", hi - lo, t, (0.0, T), dt, cost=b.AffineCost(1.0, -2.0), device=local, nccl_allreduce=True)
    b.distributed.attach_comm(eng_n)
    assert not eng_n.handle.comm_is_fused
    eng_n.forward(u0[:, lo:hi], p)
    _, dp_n = eng_n.reverse()
    assert np.allclose(np.asarray(dp_n).ravel(), np.asarray(dp).ravel(), rtol=1e-13, atol=0)
    for _ in range(3):                             # repeated gradients on one handle: epochs / parities of the mailboxes
        _, dp_again = b.adjoint_sensitivities(sol, b.Tsit5(dt=dt), t=t, dgdu_discrete=b.AffineCost(1.0, -2.0), sensealg=b.GaussAdjoint())
        assert np.array_equal(np.asarray(dp_again), np.asarray(dp))
    eng_n.close()
    cfg = O.make_cfg("lorenz", "gauss", "tsit5_fixed", N, t, 0.0, T, dt=dt, cost=("affine", 1.0, -2.0))
    ref = O.gradient(cfg, t, u0, p)
    e_dp = np.abs(dp.ravel() - ref["dp"]).max() / np.abs(ref["dp"]).max()
    e_u = np.abs(du0 - ref["du0"][:, lo:hi]).max() / np.abs(ref["du0"]).max()
    # SDE: Philox streams keyed by the global member index => shards reproduce the unsharded noise
    sprob = b.EnsembleProblem(b.SDEProblem("sde_lv", np.ones(2), (0.0, 1.0), np.array([1.5, 1.0, 3.0, 1.0, 0.1, 0.1]), seed=7), u0s=np.ones((2, N)))
    ssol = b.solve(sprob, b.EM(dt=0.01), b.EnsembleB200(device=local), saveat=0.1, sensealg=b.B200Adjoint(b.BacksolveAdjoint()))
    dW = ssol.engine.noise()
    sdu0, sdp = b.adjoint_sensitivities(ssol, b.EM(dt=0.01), t=ssol.t, dgdu_discrete=b.AffineCost(0.0, 1.0), sensealg=b.BacksolveAdjoint(), checkpoints=ssol.t)
    gW = [None] * world
    dist.all_gather_object(gW, dW)
    ok = True
    if rank == 0:
        Wfull = np.concatenate(gW, axis=2)
        scfg = O.make_cfg("sde_lv", "backsolve", "em", N, ssol.t, 0.0, 1.0, dt=0.01, cost=("affine", 0.0, 1.0))
        sref = O.gradient(scfg, ssol.t, np.ones((2, N)), np.array([1.5, 1.0, 3.0, 1.0, 0.1, 0.1]), dW=Wfull)
        e_s = np.abs(sdp.ravel() - sref["dp"]).max() / np.abs(sref["dp"]).max()
        print(f"MULTIGPU world={world} fused_allreduce={fused} dp_err={e_dp:.2e} du0_err={e_u:.2e} sde_dp_err={e_s:.2e}")
        ok = e_dp < 1e-8 and e_u < 1e-8 and e_s < 1e-9
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
