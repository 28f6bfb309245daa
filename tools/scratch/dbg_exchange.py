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

for world, sb, hot in ((4, 1 << 15, -1.0), (4, 1 << 17, -1.0), (8, 1 << 15, -1.0), (8, 1 << 15, 0.25)):
    ms = _exchange_shards(off, col, ww, V, dim, world, init, np.zeros((V, dim)), superbatch=sb, hot=hot)
    capi.train_line_group(ms, _params(total, 100))
    report(f"exchange world={world} sb={sb} hot={hot} {ms[0].exchange_stats()}", *collect(ms, world))
    del ms
