"""CPU (gloo, world_size 2): the host-side control plane of the row-sharded store -- handle exchange in rank order and
the ownership arithmetic -- and the episode protocol of the rotating shards (send_begin / train / send_end / barrier, slot
roles, sub-stream and schedule bookkeeping) with a fake model whose "copy engine" is a gloo isend / irecv pair. The data
path has no collective (peer loads/stores, copy-engine pushes), so this is all the N>1 host logic."""
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


# ---- rotating shards: the episode protocol of smore_b200/dist.py over a real process group --------------------------------
def _ring_q(rank, episode, nsub):  # rotation.cu ring_q: the sub-part `rank` trains in `episode`
    return (2 * rank - episode) % nsub


class _FakeRotModel:
    """Duck-typed capi.Model on a rotating graph. The three slot buffers hold sub-part NUMBERS instead of rows and the
    copy-engine transfer of smore_rot_send_begin is a gloo isend / irecv pair, with the library's slot roles
    (host_common.h smore_rotation_s): T(e) = slot[e % 3], O(e) = slot[(e + 2) % 3] -> the next rank's I(e) = slot[(e + 1) % 3]."""

    def __init__(self, rank, world):
        import torch

        self.rank, self.world, self.nsub = rank, world, 2 * world
        self.slot = [torch.tensor([2 * rank]), torch.tensor([-1]), torch.tensor([2 * rank + 1])]  # at home, episode 0
        self.pending, self.log, self.opened = [], [], None

    def rot_ipc_handles(self):
        return bytes([self.rank * 3 + k for k in range(3) for _ in range(64)])

    def rot_open_next(self, handles):
        self.opened = handles

    def rot_send_begin(self, e):
        assert not self.pending
        nxt, prv = (self.rank + 1) % self.world, (self.rank - 1) % self.world
        self.pending = [dist.isend(self.slot[(e + 2) % 3], nxt, tag=e), dist.irecv(self.slot[(e + 1) % 3], prv, tag=e)]

    def train_line_episode(self, q, e):
        self.log.append((e, int(self.slot[e % 3]), int(q.stream_base), int(q.sched_offset)))
        return {"samples": 10 + self.rank, "kernel_ms": 1.0}

    def rot_send_end(self, e):
        for w in self.pending:
            w.wait()
        self.pending = []


def _rot_worker(rank, world, port, out):
    import types

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m = _FakeRotModel(rank, world)
    sdist.connect_rotation(m)
    nxt = (rank + 1) % world
    ok = m.opened == bytes([nxt * 3 + k for k in range(3) for _ in range(64)])  # the NEXT rank's three slot handles
    p = types.SimpleNamespace(stream_base=5, sched_total=1000, sched_offset=100, total=7)
    episodes = 2 * world  # one cycle
    done, ms = sdist.train_line_rotating([m], p, episodes, first_episode=0, barrier=dist.barrier, world=world, rank0=rank)
    ok = ok and done == [episodes * (10 + rank)] and ms == [float(episodes)]
    for e, q, sb, so in m.log:
        ok = ok and q == _ring_q(rank, e, 2 * world)            # the resident sub-part is the one the schedule names
        ok = ok and sb == 5 + ((e * world + rank) << 20)        # sub-streams: unique per (episode, rank)
        ok = ok and so == 100 + e * 7                           # the LR schedule advances by p.total per episode
    # after a cycle every sub-part is home again
    e = episodes
    ok = ok and int(m.slot[e % 3]) == 2 * rank and int(m.slot[(e + 2) % 3]) == 2 * rank + 1
    logs = [None] * world
    dist.all_gather_object(logs, [(q, rank) for _, q, _, _ in m.log])
    met = sorted(x for l in logs for x in l)
    ok = ok and met == sorted((q, r) for q in range(2 * world) for r in range(world))  # every block exactly once per cycle
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_rotating_shards_protocol_world2():
    world = 2
    out = mp.Manager().dict()
    mp.spawn(_rot_worker, args=(world, 29654, out), nprocs=world, join=True)
    assert all(out[r] for r in range(world))


def test_rotating_shards_protocol_world4():
    world = 4
    out = mp.Manager().dict()
    mp.spawn(_rot_worker, args=(world, 29655, out), nprocs=world, join=True)
    assert all(out[r] for r in range(world))
