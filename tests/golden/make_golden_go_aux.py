#!/usr/bin/env python
"""Generates tests/golden/golden_go_aux_v1.npz: the two-graph / three-table Go models (CPR: cpr.go, TPR: tpr.go) from the
independent plain-Python restatement (tests/golden/go_restatement.py). Pure CPU, no compiled code:

    python tests/golden/make_golden_go_aux.py        (about a minute)

Both models look a vertex of the FIRST graph up in the SECOND graph by its vid (cpr.go:148-169, tpr.go:108), so the edge
lists are written such that the shared entities (CPR: users, TPR: items) are the first names to appear in both files and
therefore get the same vids; the remaining vids of the second graph are its own items / words."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests.golden import go_restatement as GO  # noqa: E402

SEED = 20261018
OUT = os.path.join(ROOT, "tests", "golden", "golden_go_aux_v1.npz")
DIM = 8


def two_domain_edges(n_shared, n_a, n_b, e_a, e_b, seed):
    """Graph A: shared x A-items, graph B: shared x B-items (labels: shared 0.., A-items 1000.., B-items 2000..). Each list
    starts with one line per shared entity in label order, each naming a NEW item, so shared entity k has vid 2k in BOTH
    graphs (vids go by first appearance, source before destination); B knows fewer shared entities than A (vids past B's
    range) and an odd vid is an A-item in A but a B-item in B -- the reference looks it up all the same."""
    rng = np.random.RandomState(seed)
    pa = 1.0 / np.arange(1, n_a + 1)
    pa /= pa.sum()
    a_src = np.concatenate([np.arange(n_shared), rng.randint(0, n_shared, size=e_a)])
    a_dst = 1000 + np.concatenate([np.arange(n_shared), rng.choice(n_a, size=e_a, p=pa)])
    nb_shared = n_shared - 7
    b_src = np.concatenate([np.arange(nb_shared), rng.randint(0, nb_shared, size=e_b)])
    b_dst = 2000 + np.concatenate([np.arange(nb_shared), rng.randint(0, n_b, size=e_b)])
    a_w = rng.randint(1, 6, size=len(a_src)).astype(np.float64)
    b_w = rng.randint(1, 6, size=len(b_src)).astype(np.float64)
    return (a_src, a_dst, a_w), (b_src, b_dst, b_w)


def rows(n, seed):
    return (np.random.RandomState(seed).random_sample((n, DIM)) - 0.5) / DIM


def main():
    G = {}
    # ---- CPR: target domain 40 users x 45 items, source domain 33 of those users x 40 items, undirected (cmd/cpr default) ----
    (ts, td, tw), (ss, sd, sw) = two_domain_edges(40, 45, 40, 400, 250, seed=91)
    for k, v in (("cpr_t_src", ts), ("cpr_t_dst", td), ("cpr_t_w", tw), ("cpr_s_src", ss), ("cpr_s_dst", sd), ("cpr_s_w", sw)):
        G[k] = v
    tg = GO.ProNet(ts.tolist(), td.tolist(), tw.tolist(), True)
    sg = GO.ProNet(ss.tolist(), sd.tolist(), sw.tolist(), True)
    assert tg.keys[:80:2] == list(range(40)) and sg.keys[:66:2] == list(range(33))
    U0, T0, S0 = rows(max(tg.max_vid, sg.max_vid), 1), rows(tg.max_vid, 2), rows(sg.max_vid, 3)
    G["cpr_init_u"], G["cpr_init_t"], G["cpr_init_s"] = U0, T0, S0
    for tag, margin, total in (("cpr", 8.0, 25000), ("cpr_m0", 0.05, 12000)):  # default margin; a margin that gates updates
        U, T, S = U0.tolist(), T0.tolist(), S0.tolist()
        rng = GO.Words(SEED, 0)
        pos = GO.train_cpr(tg, sg, U, T, S, DIM, total, total, 0.1, 0.01, 0.01, margin, rng)
        assert np.array_equal(np.array(S), S0)
        G[f"{tag}_u"], G[f"{tag}_t"], G[f"{tag}_words"] = np.array(U), np.array(T), pos
        G[f"{tag}_args"] = np.array([0.1, 0.01, 0.01, margin, total])
    # ---- TPR: 30 users x 20 items DIRECTED item -> user lines first?  no: tpr.go samples users from the user-item graph,
    # items must share vids with the item-word graph, so the user-item list starts with one line per ITEM (item -> user,
    # undirected, cmd/tpr default) and the item-word list with one line per item (item -> word) ----
    (us, ud, uw), (ws, wd, ww) = two_domain_edges(27, 30, 40, 350, 160, seed=93)  # shared = items
    for k, v in (("tpr_ui_src", us), ("tpr_ui_dst", ud), ("tpr_ui_w", uw), ("tpr_iw_src", ws), ("tpr_iw_dst", wd), ("tpr_iw_w", ww)):
        G[k] = v
    ui = GO.ProNet(us.tolist(), ud.tolist(), uw.tolist(), True)
    iw = GO.ProNet(ws.tolist(), wd.tolist(), ww.tolist(), False)  # words are sinks: only items list words
    U0, I0, W0 = rows(ui.max_vid, 4), rows(ui.max_vid, 5), rows(iw.max_vid, 6)
    G["tpr_init_u"], G["tpr_init_i"], G["tpr_init_w"] = U0, I0, W0
    total = 20000
    U, I, W = U0.tolist(), I0.tolist(), W0.tolist()
    rng = GO.Words(SEED, 0)
    pos = GO.train_tpr(ui, iw, U, I, W, DIM, total, total, 0.025, 0.001, 0.5, rng)
    G["tpr_u"], G["tpr_i"], G["tpr_w"], G["tpr_words"] = np.array(U), np.array(I), np.array(W), pos
    G["tpr_args"] = np.array([0.025, 0.001, 0.5, total])
    np.savez_compressed(OUT, **G)
    print("wrote", OUT, {k: int(v) for k, v in G.items() if k.endswith("_words")})


if __name__ == "__main__":
    main()
