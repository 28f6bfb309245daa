#!/usr/bin/env python
"""Generates tests/golden/golden_go_v1.npz from the independent plain-Python restatement of the Go tree
(tests/golden/go_restatement.py). Pure CPU, no compiled code involved:

    python tests/golden/make_golden_go.py        (about a minute)

The fixtures pin the Go-semantics side of oracle/smore_oracle.cpp (tests/test_go_pin.py) and of the CUDA kernels
(tests/test_gpu_golden_go.py): alias tables, sampler replays, and embeddings after deterministic LINE-1 / LINE-2 / BPR /
DeepWalk runs, all under the replayed Philox stream."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests import graphs  # noqa: E402
from tests.golden import go_restatement as GO  # noqa: E402

SEED = 20261018
OUT = os.path.join(ROOT, "tests", "golden", "golden_go_v1.npz")
SHUFFLE_STREAM = 1 << 63


def tables(pn, G, tag):
    G[f"{tag}_vertex_prob"], G[f"{tag}_vertex_alias"] = np.array(pn.vertex_at[0]), np.array(pn.vertex_at[1], dtype=np.int64)
    G[f"{tag}_negative_prob"], G[f"{tag}_negative_alias"] = np.array(pn.negative_at[0]), np.array(pn.negative_at[1], dtype=np.int64)
    G[f"{tag}_out_degree"], G[f"{tag}_in_degree"] = np.array(pn.out_degree), np.array(pn.in_degree)
    rng = GO.Words(SEED, 1)
    G[f"{tag}_source"] = np.array([pn.source_sample(rng) for _ in range(2000)], dtype=np.int64)
    rng = GO.Words(SEED, 2)
    G[f"{tag}_negative"] = np.array([pn.negative_sample(rng) for _ in range(2000)], dtype=np.int64)
    rng = GO.Words(SEED, 3)
    st = []
    for _ in range(2000):
        s = pn.source_sample(rng)
        st += [s, pn.target_sample(s, rng)]
    G[f"{tag}_source_target"] = np.array(st, dtype=np.int64)
    G[f"{tag}_source_target_words"] = rng.pos


def init(V, dim, seed):
    Wv, Wc = graphs.init_tables(V, dim, seed=seed)
    return Wv, Wc


def main():
    G = {}
    dim = 8
    # ---- README graph, both loading modes ----
    src, dst, w = graphs.readme_graph()
    for und in (0, 1):
        pn = GO.ProNet(src.tolist(), dst.tolist(), w.tolist(), bool(und))
        tables(pn, G, f"readme{und}")
    G["sigmoid"] = np.array(GO.ProNet(src.tolist(), dst.tolist(), w.tolist(), False).sigmoid)

    # ---- a 60-vertex graph with parallel edges and hubs: LINE order 2 / order 1, DeepWalk ----
    src, dst, w = graphs.random_graph(60, 400, seed=71)
    G["g60_src"], G["g60_dst"], G["g60_w"] = src, dst, w
    pn = GO.ProNet(src.tolist(), dst.tolist(), w.tolist(), True)
    tables(pn, G, "g60")
    V = pn.max_vid
    Wv, Wc = init(V, dim, 5)
    G["g60_init_v"], G["g60_init_c"] = Wv, Wc
    for order, total in ((2, 25000), (1, 25000)):  # > 2 LR ticks
        a, c = Wv.tolist(), Wc.tolist()
        rng = GO.Words(SEED, 0)
        pos = GO.train_line(pn, a, c, order, dim, total, total, 5, 0.025, rng)
        G[f"g60_line{order}_v"], G[f"g60_line{order}_c"], G[f"g60_line{order}_words"] = np.array(a), np.array(c), pos
        G[f"g60_line{order}_total"] = total
    a, c = Wv.tolist(), Wc.tolist()
    rng, sh = GO.Words(SEED, 0), GO.Words(SEED, SHUFFLE_STREAM)
    pos, pairs = GO.train_deepwalk(pn, a, c, dim, 2, 12, 3, 5, 0.025, rng, sh)
    G["g60_dw_v"], G["g60_dw_c"], G["g60_dw_words"], G["g60_dw_pairs"] = np.array(a), np.array(c), pos, pairs
    G["g60_dw_args"] = np.array([2, 12, 3, 5])
    # node2vec (Go tree only): biased second-order walks, p = 0.5 (returns likely), q = 2 (BFS-like)
    a, c = Wv.tolist(), Wc.tolist()
    rng, sh = GO.Words(SEED, 0), GO.Words(SEED, SHUFFLE_STREAM)
    pos, pairs = GO.train_node2vec(pn, a, c, dim, 2, 12, 3, 5, 0.025, 0.5, 2.0, rng, sh)
    G["g60_n2v_v"], G["g60_n2v_c"], G["g60_n2v_words"], G["g60_n2v_pairs"] = np.array(a), np.array(c), pos, pairs
    G["g60_n2v_args"] = np.array([2, 12, 3, 5, 0.5, 2.0])
    walks = []
    for start in range(0, pn.max_vid, 5):
        rng = GO.Words(SEED, 100 + start)
        wk = GO.biased_random_walk(pn, start, 20, 4.0, 0.25, rng)
        walks.append(np.array([start, rng.pos, len(wk)] + wk + [-1] * (21 - len(wk)), dtype=np.int64))
    G["g60_n2v_walks"] = np.stack(walks)  # rows: start, words consumed, length, the walk (p = 4, q = 0.25, 20 steps, stream 100 + start)

    # ---- a 40 x 25 bipartite graph, directed (cmd/bpr default): BPR ----
    src, dst, w = graphs.bipartite_graph(40, 25, 500, seed=73)
    G["bip_src"], G["bip_dst"], G["bip_w"] = src, dst, w
    pn = GO.ProNet(src.tolist(), dst.tolist(), w.tolist(), False)
    tables(pn, G, "bip")
    V = pn.max_vid
    Wv, Wc = init(V, dim, 6)
    G["bip_init_v"], G["bip_init_c"] = Wv, Wc
    a, c = Wv.tolist(), Wc.tolist()
    rng = GO.Words(SEED, 0)
    total = 30000
    pos = GO.train_bpr(pn, a, c, dim, total, total, 0.025, 0.001, rng)
    G["bip_bpr_v"], G["bip_bpr_c"], G["bip_bpr_words"], G["bip_bpr_total"] = np.array(a), np.array(c), pos, total
    # node2vec on the DIRECTED bipartite graph: every item is a sink, walks stop after one step (node2vec.go:100-104)
    a, c = Wv.tolist(), Wc.tolist()
    rng, sh = GO.Words(SEED, 0), GO.Words(SEED, SHUFFLE_STREAM)
    pos, pairs = GO.train_node2vec(pn, a, c, dim, 3, 10, 4, 5, 0.025, 2.0, 0.5, rng, sh)
    G["bip_n2v_v"], G["bip_n2v_c"], G["bip_n2v_words"], G["bip_n2v_pairs"] = np.array(a), np.array(c), pos, pairs
    G["bip_n2v_args"] = np.array([3, 10, 4, 5, 2.0, 0.5])
    np.savez_compressed(OUT, **G)
    print("wrote", OUT, {k: (v.shape if hasattr(v, "shape") else v) for k, v in G.items() if "words" in k or "pairs" in k})


if __name__ == "__main__":
    main()
