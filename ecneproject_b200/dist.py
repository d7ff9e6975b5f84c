"""Multi-GPU plumbing (one process per GPU): bootstrap the engine's exchange channel from an already
initialised torch.distributed process group.  The data path itself is inside libecne_b200.so:
row ranges balanced by stored terms (ecne_shard_rows), wire state replicated, and every Jacobi round's
update records pulled by the peers over NVLink inside the persistent sweep kernel."""
import ctypes as C
import os

import numpy as np

from . import _abi


def shard_rows(problem_handle, rank, world):
    """[lo, hi) of the rows rank `rank` sweeps (pure host code, no GPU needed)."""
    lib = _abi.engine_lib()
    lo, hi = C.c_uint64(0), C.c_uint64(0)
    st = lib.ecne_shard_rows(C.byref(problem_handle.c), rank, world, C.byref(lo), C.byref(hi))
    if st != 0:
        raise ValueError(lib.ecne_last_error().decode())
    return int(lo.value), int(hi.value)


def broadcast_unique_id(make_id, rank, device=None):
    """Rank 0 makes the 128-byte id, everybody receives it through torch.distributed."""
    import torch
    import torch.distributed as dist
    buf = torch.zeros(128, dtype=torch.uint8, device=device or "cpu")
    if rank == 0:
        buf.copy_(torch.from_numpy(np.frombuffer(make_id(), dtype=np.uint8).copy()))
    dist.broadcast(buf, src=0)
    return bytes(buf.cpu().numpy().tobytes())


def init_from_torch(local_rank=None):
    """Bind this process to its GPU and join the engine's communicator.  torch.distributed must be
    initialised (nccl or gloo).  Returns (rank, world)."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    if local_rank is None:
        local_rank = int(os.environ.get("LOCAL_RANK", rank))
    lib = _abi.engine_lib()
    st = lib.ecne_init(local_rank)
    if st != 0:
        raise RuntimeError(lib.ecne_last_error().decode())

    def make_id():
        out = (C.c_uint8 * 128)()
        if lib.ecne_dist_unique_id(out) != 0:
            raise RuntimeError(lib.ecne_last_error().decode())
        return bytes(out)
    dev = torch.device("cuda", local_rank) if dist.get_backend() == "nccl" else None
    uid = broadcast_unique_id(make_id, rank, dev)
    arr = (C.c_uint8 * 128).from_buffer_copy(uid)
    st = lib.ecne_dist_init(rank, world, arr)
    if st != 0:
        raise RuntimeError(lib.ecne_last_error().decode())
    from . import api
    api._initialised = True
    return rank, world
