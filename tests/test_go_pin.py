"""CPU: the Go-semantics half of the oracle (oracle/smore_oracle.cpp) against a SECOND, independently written restatement
of the Go sources (tests/golden/go_restatement.py -> tests/golden/golden_go_v1.npz), bit for bit, plus hand-computed
known answers on the README graph. There is no Go toolchain in this image (parity with a Go BUILD stays unpinned); two
restatements written separately from the same sources agreeing exactly is the evidence available without one."""
import os

import numpy as np

from oracle import bindings as B
from tests import graphs
from tests.golden import go_restatement as GO

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_go_v1.npz"))
SEED = 20261018


def _oracle(src, dst, w, undirected):
    off, col, ww, ids = B.edges_to_csr(src, dst, w, undirected)
    return B.OracleGraph(B.SEM_GO, off, col, ww, max_line=len(src)), ids


def _check_tables(og, tag):
    p, a = og.alias(0)
    assert np.array_equal(p, G[f"{tag}_vertex_prob"]) and np.array_equal(a, G[f"{tag}_vertex_alias"])
    p, a = og.alias(1)
    assert np.array_equal(p, G[f"{tag}_negative_prob"]) and np.array_equal(a, G[f"{tag}_negative_alias"])
    assert np.array_equal(og.sample(0, SEED, 1, 2000)[0], G[f"{tag}_source"])
    assert np.array_equal(og.sample(1, SEED, 2, 2000)[0], G[f"{tag}_negative"])
    st, words = og.sample(3, SEED, 3, 2000)
    assert np.array_equal(st, G[f"{tag}_source_target"]) and words == int(G[f"{tag}_source_target_words"])


def test_readme_graph_known_answers_by_hand():
    """Directed README graph: out-degrees [8,0,0,6,0,4] (userA, itemA, itemC, userB, itemB, userC). BuildAliasMethod(power 1):
    norm = d * 6 / 18 = [8/3, 0, 0, 2, 0, 4/3]; small = [1,2,4], large = [0,3,5]; popping from the back pairs
    4<-5, 5<-3, 2<-3, 3<-0, 1<-0 and leaves 0: prob = [1, 0, 0, 1/3, 0, 1/3], alias = [0, 0, 3, 0, 5, 3]."""
    src, dst, w = graphs.readme_graph()
    pn = GO.ProNet(src.tolist(), dst.tolist(), w.tolist(), False)
    assert pn.out_degree == [8.0, 0.0, 0.0, 6.0, 0.0, 4.0] and pn.in_degree == [0.0, 8.0, 5.0, 0.0, 5.0, 0.0]
    prob, alias = pn.vertex_at
    assert alias == [0, 0, 3, 0, 5, 3]
    assert np.allclose(prob, [1, 0, 0, 1 / 3, 0, 1 / 3], rtol=0, atol=1e-15)
    # the table encodes d / 18 exactly up to rounding
    p = np.array(prob) / 6
    for i, a in enumerate(alias):
        p[a] += (1 - prob[i]) / 6
    assert np.allclose(p, np.array(pn.out_degree) / 18, atol=1e-15)
    # sigmoid table: 1001 entries, exact end points and centre (pronet.go:90-95)
    assert len(pn.sigmoid) == 1001 and pn.sigmoid[500] == 0.5 and pn.fast_sigmoid(8.0) == pn.sigmoid[1000]
    assert pn.fast_sigmoid(-8.0000001) == 0.0 and pn.fast_sigmoid(8.0000001) == 1.0
    # the oracle agrees with the hand computation too
    og, _ = _oracle(src, dst, w, 0)
    p2, a2 = og.alias(0)
    assert a2.tolist() == alias and np.array_equal(p2, np.array(prob))
    assert np.array_equal(og.sigmoid_table(), G["sigmoid"])


def test_tables_and_samplers_match_the_python_restatement():
    src, dst, w = graphs.readme_graph()
    for und in (0, 1):
        og, _ = _oracle(src, dst, w, und)
        _check_tables(og, f"readme{und}")
        do, di = og.degrees()
        assert np.array_equal(do, G[f"readme{und}_out_degree"]) and np.array_equal(di, G[f"readme{und}_in_degree"])
    og, _ = _oracle(G["g60_src"], G["g60_dst"], G["g60_w"], 1)
    _check_tables(og, "g60")
    og, _ = _oracle(G["bip_src"], G["bip_dst"], G["bip_w"], 0)
    _check_tables(og, "bip")


def test_line_bpr_deepwalk_embeddings_match_the_python_restatement():
    og, _ = _oracle(G["g60_src"], G["g60_dst"], G["g60_w"], 1)
    for order in (2, 1):
        a, c = G["g60_init_v"].copy(), G["g60_init_c"].copy()
        total = int(G[f"g60_line{order}_total"])
        pos = og.train_line_go(a, c, order, 5, 0.025, total, SEED, 0)
        assert pos == int(G[f"g60_line{order}_words"])
        assert np.array_equal(a, G[f"g60_line{order}_v"]) and np.array_equal(c, G[f"g60_line{order}_c"])
    a, c = G["g60_init_v"].copy(), G["g60_init_c"].copy()
    wt, ws, win, K = (int(x) for x in G["g60_dw_args"])
    pos, pairs = og.train_deepwalk_go(a, c, wt, ws, win, K, 0.025, SEED, 0)
    assert pos == int(G["g60_dw_words"]) and pairs == int(G["g60_dw_pairs"])
    assert np.array_equal(a, G["g60_dw_v"]) and np.array_equal(c, G["g60_dw_c"])
    og, _ = _oracle(G["bip_src"], G["bip_dst"], G["bip_w"], 0)
    a, c = G["bip_init_v"].copy(), G["bip_init_c"].copy()
    total = int(G["bip_bpr_total"])
    pos = og.train_bpr_go(a, c, 0.025, 0.001, total, SEED, 0)
    assert pos == int(G["bip_bpr_words"])
    assert np.array_equal(a, G["bip_bpr_v"]) and np.array_equal(c, G["bip_bpr_c"])


def test_node2vec_matches_the_python_restatement():
    """Go-only model: biased second-order walks (node2vec.go:82-173) and the training loop, incl. a directed graph whose
    walks stop at sinks after one step."""
    og, _ = _oracle(G["g60_src"], G["g60_dst"], G["g60_w"], 1)
    for row in G["g60_n2v_walks"]:
        start, words, n = int(row[0]), int(row[1]), int(row[2])
        walk, used = og.biased_walk_go(start, 20, 4.0, 0.25, SEED, 100 + start)
        assert used == words and walk.tolist() == row[3:3 + n].tolist()
    for tag, und in (("g60", 1), ("bip", 0)):
        og, _ = _oracle(G[f"{tag}_src"], G[f"{tag}_dst"], G[f"{tag}_w"], und)
        wt, ws, win, K, p, q = G[f"{tag}_n2v_args"]
        a, c = G[f"{tag}_init_v"].copy(), G[f"{tag}_init_c"].copy()
        pos, pairs = og.train_node2vec_go(a, c, int(wt), int(ws), int(win), int(K), 0.025, float(p), float(q), SEED, 0)
        assert pos == int(G[f"{tag}_n2v_words"]) and pairs == int(G[f"{tag}_n2v_pairs"])
        assert np.array_equal(a, G[f"{tag}_n2v_v"]) and np.array_equal(c, G[f"{tag}_n2v_c"])


def test_fixture_is_reproducible_from_the_restatement():
    """Re-runs the cheapest entry: the committed .npz is what go_restatement.py produces."""
    pn = GO.ProNet(G["bip_src"].tolist(), G["bip_dst"].tolist(), G["bip_w"].tolist(), False)
    a, c = G["bip_init_v"].tolist(), G["bip_init_c"].tolist()
    pos = GO.train_bpr(pn, a, c, 8, 5000, 5000, 0.025, 0.001, GO.Words(SEED, 0))
    og, _ = _oracle(G["bip_src"], G["bip_dst"], G["bip_w"], 0)
    a2, c2 = G["bip_init_v"].copy(), G["bip_init_c"].copy()
    assert og.train_bpr_go(a2, c2, 0.025, 0.001, 5000, SEED, 0) == pos
    assert np.array_equal(np.array(a), a2) and np.array_equal(np.array(c), c2)


# ---- CPR / TPR (Go tree only; two graphs, three tables): tests/golden/golden_go_aux_v1.npz ---------------------------------
GA = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_go_aux_v1.npz"))


def test_cpr_matches_the_python_restatement():
    tg, _ = _oracle(GA["cpr_t_src"], GA["cpr_t_dst"], GA["cpr_t_w"], 1)
    sg, _ = _oracle(GA["cpr_s_src"], GA["cpr_s_dst"], GA["cpr_s_w"], 1)
    for tag in ("cpr", "cpr_m0"):
        alpha, ureg, ireg, margin, total = GA[f"{tag}_args"]
        U, T, S = GA["cpr_init_u"].copy(), GA["cpr_init_t"].copy(), GA["cpr_init_s"].copy()
        words = tg.train_cpr_go(sg, U, T, S, alpha, ureg, ireg, margin, int(total), int(total), SEED, 0)
        assert words == int(GA[f"{tag}_words"])
        assert np.array_equal(U, GA[f"{tag}_u"]) and np.array_equal(T, GA[f"{tag}_t"])
        assert np.array_equal(S, GA["cpr_init_s"])  # the source-domain rows are only read (cpr.go:160-167)
    assert not np.array_equal(GA["cpr_u"], GA["cpr_init_u"])
    # the gating margin really gates: fewer rows moved than with the default margin of 8
    moved = lambda a, b: int((np.abs(a - b).max(axis=1) > 0).sum())
    assert moved(GA["cpr_m0_t"], GA["cpr_init_t"]) <= moved(GA["cpr_t"], GA["cpr_init_t"])


def test_tpr_matches_the_python_restatement():
    ui, _ = _oracle(GA["tpr_ui_src"], GA["tpr_ui_dst"], GA["tpr_ui_w"], 1)
    iw, _ = _oracle(GA["tpr_iw_src"], GA["tpr_iw_dst"], GA["tpr_iw_w"], 0)
    alpha, lam, tw, total = GA["tpr_args"]
    U, I, W = GA["tpr_init_u"].copy(), GA["tpr_init_i"].copy(), GA["tpr_init_w"].copy()
    words = ui.train_tpr_go(iw, U, I, W, alpha, lam, tw, int(total), int(total), SEED, 0)
    assert words == int(GA["tpr_words"])
    assert np.array_equal(U, GA["tpr_u"]) and np.array_equal(I, GA["tpr_i"]) and np.array_equal(W, GA["tpr_w"])
    assert not np.array_equal(W, GA["tpr_init_w"])  # word rows are trained (tpr.go:216-232)
