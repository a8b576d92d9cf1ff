"""world_size-2 gloo test of the multi-GPU host logic on CPU: contiguous member sharding + the single all-reduce of
dG/dp.  The per-rank device pass is stood in for by the oracle (tests may call it); what is checked is that
shard -> local gradient -> all-reduce reproduces the unsharded gradient and that du0 stays sharded consistently."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, N, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import scimlsensitivity_jl_b200 as b
    from scimlsensitivity_jl_b200 import distributed as D
    from oracle import oracle as O
    assert D.world() == (rank, world)
    lo, hi = D.shard_bounds(N)
    rng = np.random.default_rng(0)
    u0 = np.array([1.0, 0.0, 0.0])[:, None] + 0.1 * rng.standard_normal((3, N))
    p = np.array([10.0, 28.0, 8 / 3]); saveat = np.linspace(0, 1, 11)
    cfg = O.make_cfg("lorenz", "gauss", "tsit5_fixed", hi - lo, saveat, 0.0, 1.0, dt=0.01, cost=("affine", 1.0, -2.0))
    r = O.gradient(cfg, saveat, u0[:, lo:hi], p, nthreads=1)

    class Eng:          # what allreduce_dp needs to know about the engine
        shared_p = True
    dp = D.allreduce_dp(r["dp"].copy(), Eng())
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), dp=dp, du0=r["du0"], lo=lo, hi=hi)
    # per-member parameters: no collective at all
    Eng.shared_p = False
    x = np.full(3, float(rank))
    assert np.array_equal(D.allreduce_dp(x, Eng()), x)
    dist.destroy_process_group()


def test_two_rank_shard_and_allreduce(tmp_path):
    N, world = 37, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, N, str(tmp_path)), nprocs=world, join=True)
    sys.path.insert(0, ROOT)
    from oracle import oracle as O
    rng = np.random.default_rng(0)
    u0 = np.array([1.0, 0.0, 0.0])[:, None] + 0.1 * rng.standard_normal((3, N))
    p = np.array([10.0, 28.0, 8 / 3]); saveat = np.linspace(0, 1, 11)
    cfg = O.make_cfg("lorenz", "gauss", "tsit5_fixed", N, saveat, 0.0, 1.0, dt=0.01, cost=("affine", 1.0, -2.0))
    full = O.gradient(cfg, saveat, u0, p, nthreads=1)
    parts = [np.load(os.path.join(str(tmp_path), f"r{r}.npz")) for r in range(world)]
    assert np.allclose(parts[0]["dp"], parts[1]["dp"], rtol=0, atol=0)          # every rank holds the same reduced dp
    assert np.allclose(parts[0]["dp"], full["dp"], rtol=1e-13)
    du0 = np.concatenate([q["du0"] for q in parts], axis=1)
    assert np.array_equal(du0, full["du0"])
    assert int(parts[0]["hi"]) == int(parts[1]["lo"])
