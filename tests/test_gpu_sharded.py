"""GPU: the row-sharded store. Shards are emulated on ONE device: world graphs + models in one process
connected by raw pointers (same kernels, same address arithmetic as the CUDA-IPC path), and -- second test -- two
PROCESSES on one device connected through real CUDA-IPC handles exchanged over gloo."""
import os

import numpy as np
import pytest
import torch.multiprocessing as mp

from oracle import bindings as B
from smore_b200 import capi, synth
from smore_b200 import dist as sdist
from tests import graphs
from tests.quality import evaluate_sampled as evaluate, sbm_graph

pytestmark = pytest.mark.gpu
SEED = 20261018


def test_world1_shard_is_identity():
    src, dst, w = graphs.random_graph(300, 4000, seed=61)
    off, col, ww, _ = B.edges_to_csr(src, dst, w, 1)
    res = []
    for shard in (False, True):
        g = capi.Graph.from_csr(off, col, ww)
        if shard:
            info = g.set_shard(0, 1)
            assert info["n_local"] == g.V and info["source_mass_fraction"] == 1.0
        m = capi.Model(g, 16, 2, capi.F64)
        m.init(0, True, SEED), m.init(1, False, SEED)
        p = capi.default_params()
        p.mode, p.seed, p.total = capi.MODE_DETERMINISTIC, SEED, 20000
        m.train_line(p)
        res.append((m.get_rows(0), m.get_rows(1)))
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1])


def test_sharded_init_and_mass_fractions():
    src, dst, w = graphs.random_graph(301, 4000, seed=63)
    off, col, ww, _ = B.edges_to_csr(src, dst, w, 1)
    g0 = capi.Graph.from_csr(off, col, ww)
    full = capi.Model(g0, 8, 1, capi.F64)
    full.init(0, True, SEED)
    W = full.get_rows(0)
    world, fr, pairs = 4, [], []
    for r in range(world):
        g = capi.Graph.from_csr(off, col, ww)
        info = g.set_shard(r, world)
        fr.append(info["source_mass_fraction"])
        rows = sdist.owned_rows(g.V, r, world)
        assert info["n_local"] == len(rows)
        m = capi.Model(g, 8, 1, capi.F64)
        m.init(0, True, SEED)
        assert np.array_equal(m.get_rows(0), W[rows])  # init does not depend on the sharding
        # the rank samples edges whose TARGET it owns, and negatives among its own vertices
        st, _ = g.sample(capi.SAMPLE_SOURCE_TARGET, SEED, r, 20000)
        n, _ = g.sample(capi.SAMPLE_NEGATIVE, SEED, r, 5000)
        assert (st[1::2] % world == r).all() and (n % world == r).all()
        pairs.append(st.reshape(-1, 2))
    assert abs(sum(fr) - 1.0) < 1e-12
    # union of the ranks' edge samples (weighted by their mass) ~ the unsharded source->target distribution:
    # compare per-source-vertex frequencies on the hottest sources
    ref, _ = g0.sample(capi.SAMPLE_SOURCE_TARGET, SEED, 99, 80000)
    ref_cnt = np.bincount(ref[0::2], minlength=g0.V) / 80000
    mix = np.zeros(g0.V)
    for f, pr in zip(fr, pairs):
        mix += f * np.bincount(pr[:, 0], minlength=g0.V) / len(pr)
    hot = np.argsort(-ref_cnt)[:20]
    assert np.allclose(mix[hot], ref_cnt[hot], rtol=0.15, atol=2e-3)


def _sbm():
    src, dst, w = sbm_graph(n_comm=150, comm_size=80, deg=24, p_in=0.85, seed=5)
    (ts, td, tw), (hs, hd, _) = synth.split_edges(src, dst, w, 0.10, seed=6)
    off, col, ww, labels = synth.csr_from_edges(ts, td, tw, True)
    lab2id = {int(l): i for i, l in enumerate(labels)}
    ok = np.array([(int(a) in lab2id) and (int(b) in lab2id) for a, b in zip(hs, hd)])
    test_s = np.array([lab2id[int(a)] for a in hs[ok]])
    test_d = np.array([lab2id[int(b)] for b in hd[ok]])
    train_adj = {v: set(col[off[v]:off[v + 1]].tolist()) for v in range(len(labels))}
    return off, col, ww, test_s, test_d, train_adj


def _params(total, seed):
    p = capi.default_params()
    p.semantics, p.mode, p.seed, p.alpha, p.total, p.negative_samples = capi.SEM_CPP, capi.MODE_HOGWILD, seed, 0.025, total, 5
    p.max_warps = 512
    return p


def _ipc_worker(rank, world, port, out):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    off, col, ww, test_s, test_d, train_adj = _sbm()
    V, dim, total = len(off) - 1, 32, 14_000_000  # LINE-2 from zero contexts needs > 6M updates to take off here
    capi.check(capi.lib().smore_init(0))  # both ranks share device 0: IPC mappings work within one device too
    g = capi.Graph.from_csr(off, col, ww)
    g.set_shard(rank, world)
    m = capi.Model(g, dim, 2, capi.F32)
    m.init(0, True, 5), m.init(1, False, 5)
    sdist.connect_peers(m)
    dist.barrier()
    p = _params(total, 17)
    p.stream_base = rank * (1 << 20)
    st = m.train_line(p)
    dist.barrier()
    Wv, Wc = sdist.gather_table(m, 0, V), sdist.gather_table(m, 1, V)
    if rank == 0:
        a, r = evaluate(Wv.astype(np.float64), Wc.astype(np.float64), test_s, test_d, train_adj, np.random.default_rng(2))
        out["auc"], out["recall"] = a, r
    out[f"samples{rank}"] = st["samples"]
    dist.barrier()
    m.close()
    dist.destroy_process_group()


def test_two_processes_cuda_ipc_on_one_device():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_ipc_worker, args=(world, 29657, out), nprocs=world, join=True)
    print(dict(out))
    assert out["samples0"] > 0 and out["samples1"] > 0
    assert out["auc"] > 0.85  # both shards learned and every cross-shard context row was reachable


# ---- bulk-exchange mode (smore_model_enable_exchange): remote vertex rows move in per-super-batch all-to-alls ---------
def _exchange_shards(off, col, ww, V, dim, world, init_v, init_c, superbatch, dtype=capi.F32, hot=-1.0):
    ms = []
    for r in range(world):
        gr = capi.Graph.from_csr(off, col, ww)
        gr.set_shard(r, world)
        mr = capi.Model(gr, dim, 2, dtype)
        rows = sdist.owned_rows(V, r, world)
        mr.set_rows(0, init_v[rows]), mr.set_rows(1, init_c[rows])
        mr.enable_exchange(superbatch, hot)
        ms.append(mr)
    if hot >= 0:  # hot rows are reached through the peer mappings
        for t in range(2):
            ptrs = [mr.device_ptr(t) for mr in ms]
            for mr in ms:
                mr.set_peer_ptrs(t, ptrs)
    return ms


def test_exchange_routing_is_exact():
    """alpha so small that no update can change a row in fp64: every table must come back bit-identical. Any mis-routed
    request, row or delta (owner adds `returned - sent`) would show up as a changed row."""
    src, dst, w = graphs.random_graph(2000, 30000, seed=71)
    off, col, ww, _ = B.edges_to_csr(src, dst, w, 1)
    V, dim, world = len(off) - 1, 24, 4
    rng = np.random.default_rng(3)
    init_v, init_c = (rng.random((V, dim)) - 0.5) / dim, (rng.random((V, dim)) - 0.5) / dim
    ms = _exchange_shards(off, col, ww, V, dim, world, init_v, init_c, superbatch=5000, dtype=capi.F64)
    p = _params(200_000, 5)
    p.alpha = 1e-30
    stats = capi.train_line_group(ms, p)
    assert sum(s["samples"] for s in stats) > 150_000
    moved = 0
    for r, mr in enumerate(ms):
        rows = sdist.owned_rows(V, r, world)
        assert np.array_equal(mr.get_rows(0), init_v[rows])
        assert np.array_equal(mr.get_rows(1), init_c[rows])
        xs = mr.exchange_stats()
        assert xs["superbatches"] == 10  # ceil(200000 / (5000 * 4))
        moved += xs["rows_requested"]
    # ~3/4 of the sources are remote; dedup inside a super-batch can only lower the count
    assert 0.1 * 200_000 < moved <= 0.8 * 200_000


def test_exchange_stages_the_right_rows():
    """K = 0, zero contexts, tiny alpha (sigmoid LUT stays in its 0.5 bin), vertex row v = [(v+1), 1, (v+1)^2, ...] * scale:
    a context row then holds sum_i a_i * row(source_i). Contexts whose second moment equals the squared mean were fed by
    ONE source vertex -- which must be a neighbour, whether its row was local or came through the exchange."""
    off, col, ww, *_ = _sbm()
    V, dim = len(off) - 1, 32
    adj = {v: set(col[off[v]:off[v + 1]].tolist()) for v in range(V)}
    ids = np.arange(V) + 1.0
    init_v = np.full((V, dim), 1e-4)
    init_v[:, 0], init_v[:, 2] = ids * 1e-4, ids * ids * 1e-8
    for world, sb, total in ((2, 1 << 15, 20000), (4, 1 << 15, 20000), (8, 1 << 11, 30000)):
        ms = _exchange_shards(off, col, ww, V, dim, world, init_v, np.zeros((V, dim)), superbatch=sb, dtype=capi.F64)
        p = _params(total, 9)
        p.negative_samples, p.alpha = 0, 1e-6
        capi.train_line_group(ms, p)
        Wc = np.zeros((V, dim))
        for r, mr in enumerate(ms):
            Wc[sdist.owned_rows(V, r, world)] = mr.get_rows(1)
        hit = np.flatnonzero(Wc[:, 1] > 0)
        mean = Wc[hit, 0] / Wc[hit, 1]
        var = Wc[hit, 2] / Wc[hit, 1] * 1e4 - mean * mean
        single = np.abs(var) < 1e-3  # (two sources one id apart with weights 1 : 99 already give 0.0099)
        x = np.round(mean[single] - 1).astype(int)
        good = np.array([xi in adj[c] for c, xi in zip(hit[single], x)])
        remote = (x % world) != (hit[single] % world)
        assert remote.sum() > 150 and good[remote].mean() > 0.995, (world, remote.sum(), good[remote].mean())
        assert good.mean() > 0.995


def _partition_worker(rank, out):
    import ctypes as C

    capi.check(capi.lib().smore_init(0))
    t, p, s = C.c_int(), C.c_int(), C.c_int()
    capi.check(capi.lib().smore_debug_sm_partition(32, C.byref(t), C.byref(p), C.byref(s)))
    out.update(total=t.value, part=p.value, seen=s.value)


def test_update_stream_runs_on_its_sm_partition():
    """The exchange mode keeps SMs free for the communication kernels by running the persistent update kernel in a green
    context: a probe kernel launched on that stream must stay inside the partition. (Own process: the first call decides
    the partition for the process.)"""
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_partition_worker, args=(out,), nprocs=1, join=True)
    assert out["total"] >= 100
    assert 0 < out["part"] < out["total"] and out["seen"] == out["part"], dict(out)
