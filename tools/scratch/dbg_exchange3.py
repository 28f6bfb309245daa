import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from smore_b200 import capi, dist as sdist
from tests.test_gpu_sharded import _sbm, _params, _exchange_shards

off, col, ww, test_s, test_d, train_adj = _sbm()
V, dim = len(off) - 1, 32
adj = {v: set(col[off[v]:off[v + 1]].tolist()) for v in range(V)}
ids = np.arange(V) + 1.0
init_v = np.full((V, dim), 1e-4)
init_v[:, 0] = ids * 1e-4
init_v[:, 2] = ids * ids * 1e-8
for world, sb, total in ((2, 1 << 15, 20000), (4, 1 << 15, 20000), (4, 1 << 15, 100000), (8, 1 << 13, 100000)):
    ms = _exchange_shards(off, col, ww, V, dim, world, init_v, np.zeros((V, dim)), superbatch=sb, dtype=capi.F64)
    p = _params(total, 9)
    p.negative_samples, p.alpha = 0, 1e-6
    st = capi.train_line_group(ms, p)
    Wc = np.zeros((V, dim))
    for r, mr in enumerate(ms):
        rows = sdist.owned_rows(V, r, world)
        Wc[rows] = mr.get_rows(1)
    hit = np.flatnonzero(Wc[:, 1] > 0)
    mean = Wc[hit, 0] / Wc[hit, 1]
    var = Wc[hit, 2] / Wc[hit, 1] * 1e4 - mean * mean
    single = np.abs(var) < 1e-6 * mean * mean + 1e-9
    x = np.round(mean[single] - 1).astype(int)
    good = np.array([xi in adj[c] for c, xi in zip(hit[single], x)])
    remote = (x % world) != (hit[single] % world)
    print(f"world={world} sb={sb} total={total}: hit {len(hit)} single-source {single.sum()} neighbours {good.sum()} | remote-source singles {remote.sum()} of which neighbours {(good & remote).sum()} {ms[0].exchange_stats()}", flush=True)
