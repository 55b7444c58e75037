"""Landmark sharding over torch.distributed (one process per GPU, NCCL over NVLink/NVSwitch).

The path shards by landmark: every rank packs the FULL graph, keeps a contiguous,
edge-balanced landmark range (vio_set_shard) and accumulates its partial reduced system.  The only
exchange step is an all-reduce (sum) of [S values | b_S correction | b_p | diag(H_pp)] per linearisation
plus two tiny scalar all-reduces per trial step; the reduced solve then runs redundantly and
deterministically on every rank, so no broadcast of the pose update is needed.
"""
import numpy as np


class _CudaPtr:
    """Zero-copy view of device memory for torch.as_tensor via __cuda_array_interface__."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f8", "data": (int(ptr), False), "version": 2,
                                         "strides": None}


def make_allreduce_hook():
    import torch
    import torch.distributed as dist

    def hook(ptr, n, stream):
        # the C side orders its kernels on `stream` (the handle's stream, possibly one torch has never seen): run the
        # collective on that stream so NCCL waits for the linearise kernels and the mirror/finalise kernels wait for NCCL
        t = torch.as_tensor(_CudaPtr(ptr, n), device="cuda")
        if stream:
            with torch.cuda.stream(torch.cuda.ExternalStream(int(stream))):
                dist.all_reduce(t, op=dist.ReduceOp.SUM)
        else:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return 0

    return hook


def init_native_nccl(problem, rank=None, world=None):
    """Native collectives: libvio_b200.so creates its own NCCL communicator (vio_nccl_init) and issues every all-reduce of
    the solve itself on the handle's stream - no callback into Python.  torch.distributed (any backend) is only used once,
    to ship rank 0's 128-byte ncclUniqueId to the other ranks.  Call on every rank before set_graph."""
    import torch.distributed as dist
    from . import capi
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    box = [capi.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    problem.nccl_init(rank, world, box[0])
    return problem


def shard_ranges(edge_ptr, world):
    """Landmark cut points used by vio_set_shard: contiguous ranges balanced by edge count.

    edge_ptr: int array (L+1) CSR offsets of the landmark-sorted edges.  Returns world+1 cut indices.
    (Host logic mirrored here so it can be tested on CPU with gloo.)
    """
    edge_ptr = np.asarray(edge_ptr)
    L = edge_ptr.shape[0] - 1
    E = int(edge_ptr[-1])
    cuts = [0]
    for r in range(1, world):
        target = E * r // world
        cuts.append(int(min(np.searchsorted(edge_ptr, target, side="left"), L)))
    cuts.append(L)
    for i in range(1, len(cuts)):
        cuts[i] = max(cuts[i], cuts[i - 1])
    return cuts
