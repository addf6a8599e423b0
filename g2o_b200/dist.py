"""Multi-GPU plumbing: one process per GPU, ``torch.distributed`` carries the collectives of the landmark-sharded
solve (the reduced camera system and its right-hand side once per LM trial, three scalars per chi2 / scale evaluation,
the replicated camera step).  The C ABI only sees a function pointer (``g2ocu_allreduce_fn``)."""
from __future__ import annotations

import ctypes

import numpy as np


class _DevicePointer:
    """Exposes a raw device pointer through ``__cuda_array_interface__`` so torch can wrap it without a copy."""

    def __init__(self, ptr: int, count: int):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f8", "data": (ptr, False), "version": 2, "strides": None}


def make_allreduce(backend_is_cuda: bool = True, group=None):
    """Returns ``fn(ptr, count, op, stream) -> int`` for :meth:`CudaSolver.set_shard`.  ``op`` 0 = sum, 1 = max, 2 = in-place
    reduce-scatter of ``world * count`` doubles.
    With a CUDA backend the collective is enqueued in stream order on the solver's stream (no host synchronisation);
    with gloo (CPU tests) ``ptr`` is a host pointer."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group); rank = dist.get_rank(group)

    def run(t, op):
        if op == 2:     # in-place reduce-scatter: rank r keeps the sum of range r (g2ocu.h: G2OCU_OP_REDUCE_SCATTER_SUM)
            count = t.numel() // world
            if backend_is_cuda:
                dist.reduce_scatter_tensor(t[rank * count:(rank + 1) * count], t, op=dist.ReduceOp.SUM, group=group)
            else:       # gloo has no reduce_scatter: all-reduce is a superset of the contract
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        else:
            dist.all_reduce(t, op=dist.ReduceOp.SUM if op == 0 else dist.ReduceOp.MAX, group=group)

    def fn(ptr, count, op, stream):
        total = count * world if op == 2 else count
        if backend_is_cuda:
            t = torch.as_tensor(_DevicePointer(ptr, total), device="cuda")
            ext = torch.cuda.ExternalStream(stream) if stream else torch.cuda.current_stream()
            with torch.cuda.stream(ext):
                run(t, op)
        else:
            buf = (ctypes.c_double * total).from_address(ptr)
            run(torch.from_numpy(np.frombuffer(buf, dtype=np.float64)), op)
        return 0
    return fn


def install_torch_allreduce(solver, rank: int, world: int, group=None):
    import torch.distributed as dist
    cuda = dist.get_backend(group) == "nccl"
    solver.set_shard(rank, world, make_allreduce(cuda, group))


def nccl_library_path() -> str:
    """libnccl.so.2 of the running torch build (the nvidia-nccl wheel); falls back to the soname for a system NCCL."""
    import glob
    import os
    import sys
    for root in sys.path:
        hits = glob.glob(os.path.join(root, "nvidia", "nccl", "lib", "libnccl.so.2"))
        if hits:
            return hits[0]
    return "libnccl.so.2"


def install_nccl(solver, rank: int, world: int, group=None):
    """Direct NCCL: rank 0 creates an ncclUniqueId through the C ABI, torch.distributed only carries its 128 bytes to the other
    ranks; afterwards no collective of the solve touches Python."""
    import torch.distributed as dist
    from . import _lib
    path = nccl_library_path()
    box = [None]
    if rank == 0:
        buf = ctypes.create_string_buffer(128)
        rc = _lib.lib().g2ocu_nccl_unique_id(path.encode(), buf)
        if rc != 0:
            raise RuntimeError("g2ocu_nccl_unique_id failed: " + _lib.lib().g2ocu_last_error(None).decode())
        box[0] = buf.raw
    dist.broadcast_object_list(box, src=0, group=group)
    solver.set_shard_nccl(rank, world, path, box[0])


def install_p2p(solver, rank: int, world: int, group=None, schur: bool = True):
    """NVLink peer-memory exchange for the slab PCG (after ``build_structure``): CUDA-IPC handles travel over torch.distributed once."""
    import torch.distributed as dist
    mine = solver.p2p_export()
    handles = [None] * world
    dist.all_gather_object(handles, mine, group=group)
    solver.p2p_import(b"".join(handles))
    if schur:   # the reduction of the reduced camera system through peer memory as well (PCG solver)
        mine = solver.p2p_export_schur()
        handles = [None] * world
        dist.all_gather_object(handles, mine, group=group)
        solver.p2p_import_schur(b"".join(handles))
    dist.barrier(group=group)
