"""Multi-GPU: one process per GPU, ensemble members sharded contiguously over ranks, exactly one all-reduce(sum) of
dG/dp per gradient when the parameters are shared.  No collective runs during time stepping: members are
independent (that is what EnsembleProblem means, test/Core4/ensembles.jl:22-31).

The collective itself lives BEHIND the C ABI (b200adj_comm_init / b200adj_reverse, csrc/comm.cu: ncclAllReduce on the
handle's stream); torch.distributed is only the host channel that carries the 128-byte NCCL id from rank 0 to the
others (what Distributed.jl / MPI would do for a Julia host).  `allreduce_dp` remains as the host-side fallback for
handles without a communicator (the gloo CPU tests).
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


def attach_comm(eng):
    """Give the engine's handle an NCCL communicator spanning all ranks (id broadcast through torch.distributed's store);
    afterwards b200adj_reverse reduces dp itself.  No-op for world size 1."""
    d = _dist()
    if d is None or d.get_world_size() == 1 or getattr(eng, "comm_attached", False):
        return eng
    from . import _lib
    box = [_lib.comm_unique_id() if d.get_rank() == 0 else None]
    d.broadcast_object_list(box, src=0)
    eng.handle.comm_init(d.get_world_size(), d.get_rank(), box[0])
    eng.comm_attached = True
    return eng


def allreduce_dp(dp, eng=None):
    """Sum the shared-parameter gradient over ranks on the host side (NCCL for CUDA tensors, gloo for host arrays).
    Skipped when the handle owns a communicator: b200adj_reverse already returned the reduced gradient."""
    d = _dist()
    if d is None or d.get_world_size() == 1 or (eng is not None and (not eng.shared_p or getattr(eng, "comm_attached", False))):
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
