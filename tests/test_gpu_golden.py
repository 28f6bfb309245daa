"""GPU: the CUDA path against the committed golden vectors generated from the compiled reference
(tests/golden/make_golden.py): tables and sampler/walk output bit-exact, deterministic-mode embeddings after the
reference's own 1M-sample Train() within 1e-5 relative."""
import os

import numpy as np
import pytest

from oracle import bindings as B
from smore_b200 import capi
from tests import graphs

pytestmark = pytest.mark.gpu
SEED = 20261018
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_v1.npz"))


def rel_err(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


def params(**kw):
    p = capi.default_params()
    p.semantics, p.mode, p.seed, p.stream_base = capi.SEM_CPP, capi.MODE_DETERMINISTIC, SEED, 0
    for k, v in kw.items():
        setattr(p, k, v)
    return p


@pytest.mark.parametrize("und", [0, 1])
def test_readme_graph_ingest_tables_samplers(tmp_path, und):
    # through the text ingest path, with the README's own vertex names
    src, dst, w = graphs.readme_graph()
    names = ["userA", "itemA", "itemC", "userB", "itemB", "userC"]
    path = str(tmp_path / "net.txt")
    B.write_edge_list(path, src, dst, w, names=names)
    g = capi.Graph.from_edge_list(path, und, capi.SEM_CPP)
    off, col, ww = g.csr()
    assert np.array_equal(off, G[f"readme{und}_off"]) and np.array_equal(col, G[f"readme{und}_col"])
    assert np.array_equal(ww, G[f"readme{und}_w"])
    assert g.names() == list(G[f"readme{und}_names"])
    for which, nm in ((0, "vertex"), (1, "negative"), (2, "context")):
        p, a = g.alias(which)
        assert np.array_equal(p, G[f"readme{und}_{nm}_prob"]) and np.array_equal(a, G[f"readme{und}_{nm}_alias"])
    assert np.array_equal(g.sample(0, SEED, 1, 2000)[0], G[f"readme{und}_source"])
    assert np.array_equal(g.sample(1, SEED, 2, 2000)[0], G[f"readme{und}_negative"])
    assert np.array_equal(g.sample(3, SEED, 3, 2000)[0], G[f"readme{und}_source_target"])


def g300(neg=capi.NEG_DEGREES):
    off, col, ww, _ = B.edges_to_csr(G["g300_src"], G["g300_dst"], G["g300_w"], 1)
    return capi.Graph.from_csr(off, col, ww, semantics=capi.SEM_CPP, negative_method=neg)


def test_g300_samplers_walks():
    g = g300()
    assert np.array_equal(g.sample(3, SEED, 4, 20000)[0], G["g300_source_target"])
    k, off = 0, 0
    for start in range(0, 300, 7):
        for mode, w0, w1 in ((0, 5, 0), (1, 2, 5)):
            wk, a, b = g.walk_pairs(SEED, 1000 + start, start, 40, mode, w0, w1)
            n, m = G["g300_walk_len"][k], G["g300_walk_npairs"][k]
            assert np.array_equal(wk, G["g300_walks"][k][:n])
            assert np.array_equal(a, G["g300_pair_v"][off:off + m]) and np.array_equal(b, G["g300_pair_c"][off:off + m])
            off += m
            k += 1


@pytest.mark.parametrize("order", [1, 2])
def test_line_train_1m(order):
    g = g300()
    m = capi.Model(g, 8, 1 if order == 1 else 2, capi.F64)
    m.set_rows(0, G["g300_init_v"])
    if order == 2:
        m.set_rows(1, G["g300_init_c"])
    st = m.train_line(params(total=1000000, order=order))
    assert st["words_stream0"] == int(G[f"line{order}_words"])
    assert rel_err(m.get_rows(0), G[f"line{order}_v"]) == 0.0
    if order == 2:
        assert rel_err(m.get_rows(1), G["line2_c"]) == 0.0


@pytest.mark.parametrize("walklets", [0, 1])
def test_walk_models(walklets):
    g = g300()
    m = capi.Model(g, 8, 2, capi.F64)
    m.set_rows(0, G["g300_init_v2"])
    m.set_rows(1, G["g300_init_c2"])
    nm = "walklets" if walklets else "deepwalk"
    p = params(walk_times=3, walk_steps=20, window_min=2 if walklets else 1, window_max=4 if walklets else 5)
    st = m.train_walklets(p) if walklets else m.train_deepwalk(p)
    assert st["words_stream0"] == int(G[f"{nm}_words"])
    assert rel_err(m.get_rows(0), G[f"{nm}_v"]) == 0.0 and rel_err(m.get_rows(1), G[f"{nm}_c"]) == 0.0


def test_ranking_models():
    off, col, ww, _ = B.edges_to_csr(G["bip_src"], G["bip_dst"], G["bip_w"], 0)
    g = capi.Graph.from_csr(off, col, ww, semantics=capi.SEM_CPP, negative_method=capi.NEG_NO_DEGREES)
    for nm in ("bpr", "warp"):
        m = capi.Model(g, 8, 1, capi.F64)
        m.set_rows(0, G[f"{nm}_init"])
        st = (m.train_bpr if nm == "bpr" else m.train_warp)(params(total=1000000))
        assert st["words_stream0"] == int(G[f"{nm}_words"])
        assert rel_err(m.get_rows(0), G[f"{nm}_v"]) == 0.0
    off, col, ww, _ = B.edges_to_csr(G["bip_src"], G["bip_dst"], G["bip_w"], 1)
    g = capi.Graph.from_csr(off, col, ww, semantics=capi.SEM_CPP, negative_method=capi.NEG_NO_DEGREES)
    g.set_field(G["hoprec_field"])
    m = capi.Model(g, 8, 1, capi.F64)
    m.set_rows(0, G["hoprec_init"])
    st = m.train_hoprec(params(total=1000000, walk_steps=3))
    assert st["words_stream0"] == int(G["hoprec_words"])
    assert rel_err(m.get_rows(0), G["hoprec_v"]) == 0.0
