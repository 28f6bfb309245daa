import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from smore_b200 import capi, dist as sdist
from tests.test_gpu_sharded import _sbm, _params, _exchange_shards
from tests.test_gpu_quality import evaluate

off, col, ww, test_s, test_d, train_adj = _sbm()
V, dim, total = len(off) - 1, 32, 12_000_000
init = ((np.random.default_rng(1).random((V, dim)) - 0.5) / dim)

def collect(ms, world):
    Wv, Wc = np.zeros((V, dim)), np.zeros((V, dim))
    for r, mr in enumerate(ms):
        rows = sdist.owned_rows(V, r, world)
        Wv[rows], Wc[rows] = mr.get_rows(0), mr.get_rows(1)
    return Wv, Wc

def report(tag, Wv, Wc):
    a, r = evaluate(Wv, Wc, test_s, test_d, train_adj, np.random.default_rng(2))
    print(f"{tag}: AUC {a:.4f} rec {r:.4f} |dWv| {np.linalg.norm(Wv - init):.3f} |Wc| {np.linalg.norm(Wc):.3f} max|Wv| {np.abs(Wv).max():.3f}", flush=True)

def peer_shards(world):
    ms = []
    for r in range(world):
        gr = capi.Graph.from_csr(off, col, ww)
        gr.set_shard(r, world)
        mr = capi.Model(gr, dim, 2, capi.F32)
        rows = sdist.owned_rows(V, r, world)
        mr.set_rows(0, init[rows]), mr.set_rows(1, np.zeros((len(rows), dim)))
        ms.append(mr)
    for t in range(2):
        ptrs = [mr.device_ptr(t) for mr in ms]
        for mr in ms:
            mr.set_peer_ptrs(t, ptrs)
    return ms

world = 4
for rounds, reseed in ((20, True), (92, True), (92, False), (366, True)):
    ms = peer_shards(world)
    for k in range(rounds):
        for r, mr in enumerate(ms):
            p = _params(total // rounds, 100 + k if reseed else 100)
            p.stream_base = r * (1 << 20) + (0 if reseed else k * (1 << 24))
            p.sched_total, p.sched_offset = total, k * (total // rounds)
            mr.train_line(p)
    report(f"peer-mode world={world} rounds={rounds} reseed={reseed}", *collect(ms, world))
    del ms
for tot in (24_000_000,):
    ms = _exchange_shards(off, col, ww, V, dim, world, init, np.zeros((V, dim)), superbatch=1 << 15, hot=0.0)
    capi.train_line_group(ms, _params(tot, 100))
    report(f"exchange all-hot world={world} sb=32768 total={tot}", *collect(ms, world))
