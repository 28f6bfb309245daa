"""CPU: the precomputed reference-path quality numbers (tests/golden/quality_baselines_v1.json) are reproducible -- the
oracle is deterministic on one stream, so re-running the cheapest entry must give the stored AUC exactly. (The Hogwild
GPU gates of HPE / MF / Skew-OPT compare against these numbers.)"""
import json
import os

import numpy as np

from oracle import bindings as B
from smore_b200 import synth
from tests.quality import evaluate_sampled as evaluate, sbm_graph

Q = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "quality_baselines_v1.json")))


def test_baselines_are_sane():
    for name in ("hpe", "mf", "skewopt"):
        assert Q["models"][name]["auc"] > 0.9 and Q["models"][name]["recall_at_10"] > 0.09


def test_skewopt_baseline_reproduces():
    src, dst, w = sbm_graph(n_comm=150, comm_size=80, deg=24, p_in=0.85, seed=5)
    (ts, td, tw), (hs, hd, _) = synth.split_edges(src, dst, w, 0.10, seed=6)
    off, col, ww, labels = synth.csr_from_edges(ts, td, tw, True)
    lab2id = {int(l): i for i, l in enumerate(labels)}
    ok = np.array([(int(a) in lab2id) and (int(b) in lab2id) for a, b in zip(hs, hd)])
    test_s = np.array([lab2id[int(a)] for a in hs[ok]])
    test_d = np.array([lab2id[int(b)] for b in hd[ok]])
    V = len(labels)
    train_adj = {v: set(col[off[v]:off[v + 1]].tolist()) for v in range(V)}
    m = Q["models"]["skewopt"]
    g = B.OracleGraph(B.SEM_CPP, off, col, ww, neg_method=B.NEG_NO_DEGREES)
    W = (np.random.default_rng(1).random((V, Q["dim"])) - 0.5) / Q["dim"] + m["init_offset"]
    g.train_skewopt_cpp(W, m["xi"], m["omega"], m["eta"], m["alpha"], m["total"], Q["seed"], 0)
    a, r = evaluate(W, W, test_s, test_d, train_adj, np.random.default_rng(2))
    assert abs(a - m["auc"]) < 1e-6 and abs(r - m["recall_at_10"]) < 1e-6  # (libm / BLAS may differ in the last bits between machines)
