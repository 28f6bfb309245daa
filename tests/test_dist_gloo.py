"""CPU (gloo, world_size 2): the host-side control plane of the row-sharded store -- handle exchange in rank order and
the ownership arithmetic. The data path has no collective (peer loads/stores), so this is all the N>1 host logic."""
import os

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from smore_b200 import dist as sdist


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    handle = bytes([rank * 16 + (i % 16) for i in range(64)])
    allh = sdist.exchange_handles(handle)
    ok = len(allh) == 64 * world and all(allh[64 * r:64 * (r + 1)] == bytes([r * 16 + (i % 16) for i in range(64)])
                                         for r in range(world))

    class FakeModel:  # duck-typed capi.Model: the gather path only needs get_rows / dim
        dim = 4

        def get_rows(self, table, dtype=np.float32):
            rows = sdist.owned_rows(11, rank, world)
            return (rows[:, None] * 10 + np.arange(4)[None, :]).astype(dtype)

    full = sdist.gather_table(FakeModel(), 0, 11)
    if rank == 0:
        ok = ok and np.array_equal(full, (np.arange(11)[:, None] * 10 + np.arange(4)[None, :]).astype(np.float32))
    else:
        ok = ok and full is None
    out[rank] = ok
    dist.destroy_process_group()


def test_handle_exchange_and_gather_world2():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, 29653, out), nprocs=world, join=True)
    assert all(out[r] for r in range(world))


def test_owned_rows_partition():
    for V, world in ((10, 2), (11, 4), (7, 8), (5, 1)):
        seen = np.concatenate([sdist.owned_rows(V, r, world) for r in range(world)])
        assert sorted(seen.tolist()) == list(range(V))
        for r in range(world):
            assert all(v % world == r for v in sdist.owned_rows(V, r, world))
