"""Multi-GPU: one process per GPU (torch.distributed), ensemble members sharded contiguously over ranks, exactly one
all-reduce(sum) of dG/dp per gradient when the parameters are shared.  No collective runs during
time stepping: members are independent (that is what EnsembleProblem means, test/Core4/ensembles.jl:22-31).
"""
import numpy as np


def _dist():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist
    except Exception:
        pass
    return None


def world():
    d = _dist()
    return (d.get_rank(), d.get_world_size()) if d else (0, 1)


def shard_bounds(N, rank=None, world_size=None):
    """Contiguous block [lo, hi) of members owned by `rank`: i in [g*N/G, (g+1)*N/G)."""
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    return (rank * N) // world_size, ((rank + 1) * N) // world_size


def allreduce_dp(dp, eng=None):
    """Sum the shared-parameter gradient over ranks (NCCL for CUDA tensors, gloo for host arrays)."""
    d = _dist()
    if d is None or d.get_world_size() == 1 or (eng is not None and not eng.shared_p):
        return dp
    import torch
    if hasattr(dp, "data_ptr"):
        d.all_reduce(dp, op=d.ReduceOp.SUM)
        return dp
    t = torch.from_numpy(np.ascontiguousarray(dp))
    if d.get_backend() == "nccl":
        t = t.cuda()
        d.all_reduce(t, op=d.ReduceOp.SUM)
        return t.cpu().numpy()
    d.all_reduce(t, op=d.ReduceOp.SUM)
    return t.numpy()
