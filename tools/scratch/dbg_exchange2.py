import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from smore_b200 import capi, dist as sdist
from tests.test_gpu_sharded import _sbm, _params, _exchange_shards

off, col, ww, test_s, test_d, train_adj = _sbm()
V, dim = len(off) - 1, 32
adj = {v: set(col[off[v]:off[v + 1]].tolist()) for v in range(V)}
init_v = np.full((V, dim), 1e-4)
init_v[:, 0] = (np.arange(V) + 1) * 1e-4
for world, sb, total in ((2, 1 << 15, 60000), (4, 1 << 15, 60000), (4, 1 << 15, 240000), (4, 1 << 12, 60000), (8, 1 << 14, 200000)):
    ms = _exchange_shards(off, col, ww, V, dim, world, init_v, np.zeros((V, dim)), superbatch=sb, dtype=capi.F64)
    p = _params(total, 9)
    p.negative_samples, p.alpha = 0, 1e-6
    st = capi.train_line_group(ms, p)
    Wc = np.zeros((V, dim)); Wv = np.zeros((V, dim))
    for r, mr in enumerate(ms):
        rows = sdist.owned_rows(V, r, world)
        Wc[rows] = mr.get_rows(1); Wv[rows] = mr.get_rows(0)
    hit = np.flatnonzero(Wc[:, 1] > 0)
    ratio = Wc[hit, 0] / Wc[hit, 1] - 1
    single = np.abs(ratio - np.round(ratio)) < 1e-6
    ok = sum(int(round(x)) in adj[c] for c, x in zip(hit[single], ratio[single]))
    print(f"world={world} sb={sb} total={total}: samples {sum(s['samples'] for s in st)} contexts hit {len(hit)} single-source {single.sum()} of which neighbours {ok}; max|dWv| {np.abs(Wv-init_v).max():.2e} stats {ms[0].exchange_stats()}", flush=True)
