#!/usr/bin/env python
"""Reference-path quality numbers for the §8f models (HPE, MF, Skew-OPT) on the planted-partition graph of the LINE
quality gate (tests/test_zz_gpu_quality_gates.py): the oracle restatement -- pinned bit-exactly against the compiled reference --
trained on one stream, then held-out AUC / recall@10. Written to tests/golden/quality_baselines_v1.json so that the
Hogwild GPU gates of these models ("within 0.5 % of the reference path") have their CPU side precomputed.

    python tests/golden/make_quality_baselines.py        (a few minutes of CPU)
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import bindings as B  # noqa: E402
from smore_b200 import synth  # noqa: E402
from tests.quality import evaluate_sampled as evaluate, sbm_graph  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "quality_baselines_v1.json")
SEED = 20261018


def main():
    src, dst, w = sbm_graph(n_comm=150, comm_size=80, deg=24, p_in=0.85, seed=5)
    (ts, td, tw), (hs, hd, _) = synth.split_edges(src, dst, w, 0.10, seed=6)
    off, col, ww, labels = synth.csr_from_edges(ts, td, tw, True)
    lab2id = {int(l): i for i, l in enumerate(labels)}
    ok = np.array([(int(a) in lab2id) and (int(b) in lab2id) for a, b in zip(hs, hd)])
    test_s = np.array([lab2id[int(a)] for a in hs[ok]])
    test_d = np.array([lab2id[int(b)] for b in hd[ok]])
    V = len(labels)
    train_adj = {v: set(col[off[v]:off[v + 1]].tolist()) for v in range(V)}
    dim = 32
    init_v = (np.random.default_rng(1).random((V, dim)) - 0.5) / dim
    init_c = (np.random.default_rng(3).random((V, dim)) - 0.5) / dim
    res = {"graph": "sbm_graph(150, 80, 24, 0.85, seed=5), 10 % held out (split seed 6), undirected CSR", "dim": dim,
           "init": "vertex: default_rng(1), context: default_rng(3), (U-0.5)/dim", "seed": SEED, "models": {}}

    def run(name, fn, Wv, Wc, **params):
        t0 = time.time()
        fn()
        a, r = evaluate(Wv, Wc, test_s, test_d, train_adj, np.random.default_rng(2))
        res["models"][name] = dict(params, auc=float(a), recall_at_10=float(r), seconds=round(time.time() - t0, 1))
        print(name, res["models"][name], flush=True)

    g = B.OracleGraph(B.SEM_CPP, off, col, ww)
    a, c = init_v.copy(), init_c.copy()
    run("hpe", lambda: g.train_hpe_cpp(a, c, 5, 5, 0.01, 0.025, 3_000_000, SEED, 0), a, c,
        walk_steps=5, negative_samples=5, reg=0.01, alpha=0.025, total=3_000_000)
    gn = B.OracleGraph(B.SEM_CPP, off, col, ww, neg_method=B.NEG_NO_DEGREES)
    m = init_v.copy()
    run("mf", lambda: gn.train_mf_cpp(m, 5, 0.01, 0.025, 12_000_000, SEED, 0), m, m,
        negative_samples=5, reg=0.01, alpha=0.025, total=12_000_000, negative_method="no_degrees")
    s = init_v.copy() + 0.01
    run("skewopt", lambda: gn.train_skewopt_cpp(s, 10.0, 3.0, 3, 0.025, 3_000_000, SEED, 0), s, s,
        xi=10.0, omega=3.0, eta=3, alpha=0.025, total=3_000_000, negative_method="no_degrees", init_offset=0.01)
    json.dump(res, open(OUT, "w"), indent=1)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
