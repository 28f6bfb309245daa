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


# ---- rotating shards (include/smore_b200.h "rotating shards") ---------------------------------------------------------
def connect_rotation(model, group=None) -> None:
    """One process per GPU: every rank opens the slot buffers of the NEXT rank of the ring (CUDA IPC)."""
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    handles = exchange_handles(model.rot_ipc_handles(), group)
    nxt = (rank + 1) % world
    model.rot_open_next(handles[192 * nxt:192 * (nxt + 1)])


def connect_rotation_local(models) -> None:
    """All shards in this process (several shards on one device; tests): plain device pointers."""
    n = len(models)
    ptrs = [m.rot_slot_ptrs() for m in models]
    for r, m in enumerate(models):
        m.rot_set_next_ptrs(ptrs[(r + 1) % n])


def train_line_rotating(models, p, episodes, first_episode=0, barrier=None, world=None, rank0=0, trainer="line"):
    """Runs `episodes` episodes of the block-cyclic schedule. `models`: the shard(s) driven by this process -- [model] with
    one process per GPU (then `barrier` must synchronise the ranks, e.g. torch.distributed.barrier), or all shards of the
    ring in rank order (tests; no barrier needed). p.total = samples of ALL ranks per episode; p.sched_total (if set) is the
    length of the whole LR schedule in samples and the offset advances by p.total per episode; sub-streams are
    p.stream_base + ((episode * world + rank) << 20) + warp. trainer: "line" or "bpr" (the Go BPR). Returns (samples,
    kernel milliseconds) per model."""
    import copy

    world = world or len(models)
    done = [0] * len(models)
    ms = [0.0] * len(models)
    for e in range(first_episode, first_episode + episodes):
        for m in models:
            m.rot_send_begin(e)
        for i, m in enumerate(models):
            q = copy.copy(p)
            q.stream_base = p.stream_base + ((e * world + rank0 + i) << 20)
            if p.sched_total:
                q.sched_offset = p.sched_offset + (e - first_episode) * p.total
            st = m.train_bpr_episode(q, e) if trainer == "bpr" else m.train_line_episode(q, e)
            done[i] += st["samples"]
            ms[i] += st["kernel_ms"]
        for m in models:
            m.rot_send_end(e)
        if barrier is not None:
            barrier()
    return done, ms
