"""CPU: the oracle restatement against the committed golden vectors (tests/golden/golden_v1.npz, generated from the
compiled reference by tests/golden/make_golden.py). This is what pins the oracle where /root/reference is absent."""
import os

import numpy as np
import pytest

from oracle import bindings as B
from tests import graphs

SEED = 20261018
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_v1.npz"))


def test_philox_known_answers():
    for c, k, o in zip(G["philox_kat_ctr"], G["philox_kat_key"], G["philox_kat_out"]):
        assert np.array_equal(B.philox_block(c, k), o)
    # stream layout: word n = lane n&3 of block n>>2, counter = {blk_lo, blk_hi, stream_lo, stream_hi}, key = seed
    w = B.stream_words(0xa4093822_00000000 >> 32 | (0x299f31d0 << 32), 0, 0, 8)
    assert np.array_equal(w[:4], B.philox_block([0, 0, 0, 0], [0xa4093822, 0x299f31d0]))
    assert np.array_equal(w[4:], B.philox_block([1, 0, 0, 0], [0xa4093822, 0x299f31d0]))


@pytest.mark.parametrize("und", [0, 1])
def test_readme_graph(und):
    src, dst, w = graphs.readme_graph()
    off, col, ww, ids = B.edges_to_csr(src, dst, w, und)
    assert np.array_equal(off, G[f"readme{und}_off"]) and np.array_equal(col, G[f"readme{und}_col"])
    assert np.array_equal(ww, G[f"readme{und}_w"])
    g = B.OracleGraph(B.SEM_CPP, off, col, ww)
    for which, nm in ((0, "vertex"), (1, "negative"), (2, "context")):
        p, a = g.alias(which)
        assert np.array_equal(p, G[f"readme{und}_{nm}_prob"]) and np.array_equal(a, G[f"readme{und}_{nm}_alias"])
    assert np.array_equal(g.sample(0, SEED, 1, 2000)[0], G[f"readme{und}_source"])
    assert np.array_equal(g.sample(1, SEED, 2, 2000)[0], G[f"readme{und}_negative"])
    assert np.array_equal(g.sample(3, SEED, 3, 2000)[0], G[f"readme{und}_source_target"])
    assert np.array_equal(g.sigmoid_table(), G["sigmoid_table"])


def g300(neg=B.NEG_DEGREES):
    off, col, ww, _ = B.edges_to_csr(G["g300_src"], G["g300_dst"], G["g300_w"], 1)
    return B.OracleGraph(B.SEM_CPP, off, col, ww, neg_method=neg)


def test_g300_tables_samplers_walks():
    g = g300()
    for which, nm in ((0, "vertex"), (1, "negative"), (2, "context")):
        p, a = g.alias(which)
        assert np.array_equal(p, G[f"g300_{nm}_prob"]) and np.array_equal(a, G[f"g300_{nm}_alias"])
    assert np.array_equal(g.sample(3, SEED, 4, 20000)[0], G["g300_source_target"])
    k, off = 0, 0
    for start in range(0, 300, 7):
        for mode, w0, w1 in ((0, 5, 0), (1, 2, 5)):
            wk, a, b = g.walk_pairs(SEED, 1000 + start, start, 40, mode, w0, w1)
            n = G["g300_walk_len"][k]
            assert np.array_equal(wk, G["g300_walks"][k][:n])
            m = G["g300_walk_npairs"][k]
            assert np.array_equal(a, G["g300_pair_v"][off:off + m]) and np.array_equal(b, G["g300_pair_c"][off:off + m])
            off += m
            k += 1


@pytest.mark.parametrize("order", [1, 2])
def test_line_train(order):
    g = g300()
    a, c = G["g300_init_v"].copy(), G["g300_init_c"].copy()
    pos = g.train_line_cpp(a, a if order == 1 else c, 5, 0.025, 1000000, SEED, 0)
    assert pos == int(G[f"line{order}_words"])
    assert np.array_equal(a, G[f"line{order}_v"])
    if order == 2:
        assert np.array_equal(c, G["line2_c"])


@pytest.mark.parametrize("walklets", [0, 1])
def test_walk_models_train(walklets):
    g = g300()
    a, c = G["g300_init_v2"].copy(), G["g300_init_c2"].copy()
    nm = "walklets" if walklets else "deepwalk"
    w0, w1 = (2, 4) if walklets else (5, 0)
    pos, pairs = g.train_walk_cpp(walklets, a, c, 3, 20, w0, w1, 5, 0.025, SEED, 0)
    assert pos == int(G[f"{nm}_words"]) and pairs > 0
    assert np.array_equal(a, G[f"{nm}_v"]) and np.array_equal(c, G[f"{nm}_c"])


def bip(und):
    off, col, ww, ids = B.edges_to_csr(G["bip_src"], G["bip_dst"], G["bip_w"], und)
    return B.OracleGraph(B.SEM_CPP, off, col, ww, neg_method=B.NEG_NO_DEGREES), ids


def test_bpr_warp_hoprec_train():
    g, _ = bip(0)
    a = G["bpr_init"].copy()
    assert g.train_bpr_cpp(a, 0.025, 1000000, SEED, 0) == int(G["bpr_words"])
    assert np.array_equal(a, G["bpr_v"])
    a = G["warp_init"].copy()
    pos, tries = g.train_warp_cpp(a, 0.025, 1000000, SEED, 0)
    assert pos == int(G["warp_words"]) and tries > 1000000
    assert np.array_equal(a, G["warp_v"])
    g, ids = bip(1)
    g.set_field(G["hoprec_field"])
    a = G["hoprec_init"].copy()
    assert g.train_hoprec_cpp(a, 3, 0.025, 1000000, SEED, 0) == int(G["hoprec_words"])
    assert np.array_equal(a, G["hoprec_v"])


# ---- HPE (tests/golden/golden_hpe_v1.npz, tests/golden/make_golden_hpe.py) --------------------------------------------
GH = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_hpe_v1.npz"))


@pytest.mark.parametrize("tag", ["a", "b"])
def test_hpe_train_golden(tag):
    off, col, ww, _ = B.edges_to_csr(GH["g300_src"], GH["g300_dst"], GH["g300_w"], 1)
    g = B.OracleGraph(B.SEM_CPP, off, col, ww)
    steps, K, reg = GH[f"hpe_{tag}_args"]
    a, c = GH["init_v"].copy(), GH["init_c"].copy()
    pos = g.train_hpe_cpp(a, c, int(steps), int(K), float(reg), 0.025, 1000000, SEED, 0)
    assert pos == int(GH[f"hpe_{tag}_words"])
    assert np.array_equal(a, GH[f"hpe_{tag}_v"]) and np.array_equal(c, GH[f"hpe_{tag}_c"])


# ---- MF (tests/golden/golden_mf_v1.npz, tests/golden/make_golden_mf.py) -----------------------------------------------
GM = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_mf_v1.npz"))


@pytest.mark.parametrize("tag", ["bip", "small"])
def test_mf_train_golden(tag):
    off, col, ww, _ = B.edges_to_csr(GM[f"{tag}_src"], GM[f"{tag}_dst"], GM[f"{tag}_w"], 0)
    g = B.OracleGraph(B.SEM_CPP, off, col, ww, neg_method=B.NEG_NO_DEGREES)
    K, reg = GM[f"{tag}_args"]
    a = GM[f"{tag}_init"].copy()
    pos = g.train_mf_cpp(a, int(K), float(reg), 0.025, 1000000, SEED, 0)
    assert pos == int(GM[f"{tag}_words"])
    assert np.array_equal(a, GM[f"{tag}_v"])


# ---- Skew-OPT (tests/golden/golden_skewopt_v1.npz): oracle groundwork, the kernel is the next row -----------------------
GS = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_skewopt_v1.npz"))


@pytest.mark.parametrize("tag", ["a", "b"])
def test_skewopt_train_golden(tag):
    off, col, ww, _ = B.edges_to_csr(GS["bip_src"], GS["bip_dst"], GS["bip_w"], 0)
    g = B.OracleGraph(B.SEM_CPP, off, col, ww, neg_method=B.NEG_NO_DEGREES)
    xi, omega, eta = GS[f"{tag}_args"]
    a = GS[f"{tag}_init"].copy()
    pos = g.train_skewopt_cpp(a, float(xi), float(omega), int(eta), 0.025, 1000000, SEED, 0)
    assert pos == int(GS[f"{tag}_words"])
    assert np.array_equal(a, GS[f"{tag}_v"])
    assert not np.array_equal(a, GS[f"{tag}_init"])
