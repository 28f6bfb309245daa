"""Host-side plumbing for the row-sharded store: one process per GPU, `torch.distributed` for the control plane only.

The data path has no collective: kernels reach a remote shard with plain loads/stores through CUDA-IPC peer mappings
(NVLink). What the ranks exchange here, once, are the 64-byte IPC handles of their shards.
"""
from __future__ import annotations

import numpy as np


def exchange_handles(handle: bytes, group=None) -> bytes:
    """all_gather of one 64-byte handle per rank -> world x 64 bytes in rank order (works with gloo and nccl)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    mine = torch.tensor(list(handle), dtype=torch.uint8, device=dev)
    out = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(out, mine, group=group)
    return b"".join(bytes(t.cpu().numpy().tobytes()) for t in out)


def connect_peers(model, group=None) -> None:
    """Connects every table of `model` (a capi.Model on a sharded graph) to the shards of all other ranks."""
    for t in range(model.n_tables):
        model.open_peers(t, exchange_handles(model.ipc_handle(t), group))


def init_exchange(group=None) -> None:
    """Bootstraps the library's own NCCL communicator (bulk-exchange mode): rank 0's ncclUniqueId is broadcast over the
    host's process group, then every rank joins."""
    import torch
    import torch.distributed as dist

    from . import capi

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    uid = capi.nccl_unique_id() if rank == 0 else bytes(128)
    t = torch.tensor(list(uid), dtype=torch.uint8, device=dev)
    dist.broadcast(t, src=0, group=group)
    capi.nccl_init(bytes(t.cpu().numpy().tobytes()), rank, world)


def owned_rows(V: int, rank: int, world: int) -> np.ndarray:
    """Global vertex ids of the rows rank holds, in local-row order."""
    return np.arange(rank, V, world)


def gather_table(model, table: int, V: int, group=None) -> np.ndarray | None:
    """Rank 0 gets the full [V x dim] table (float32) assembled from every shard; other ranks get None."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    local = model.get_rows(table, dtype=np.float32)
    objs = [None] * world if rank == 0 else None
    dist.gather_object(local, objs, dst=0, group=group)
    if rank != 0:
        return None
    full = np.zeros((V, model.dim), dtype=np.float32)
    for r, part in enumerate(objs):
        full[owned_rows(V, r, world)] = part
    return full
