"""GPU: the Go-semantics kernels against the fixtures of the independent plain-Python restatement of the Go tree
(tests/golden/go_restatement.py -> golden_go_v1.npz): alias tables and sampler replays bit-exact, deterministic fp64
embeddings of LINE order 1 / 2, BPR and DeepWalk bit-identical (north-star bar: 1e-5 relative)."""
import os

import numpy as np
import pytest

from oracle import bindings as B
from smore_b200 import capi
from tests import graphs

pytestmark = pytest.mark.gpu
SEED = 20261018
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_go_v1.npz"))


def _graph(src, dst, w, undirected):
    off, col, ww, _ = B.edges_to_csr(src, dst, w, undirected)
    return capi.Graph.from_csr(off, col, ww, semantics=capi.SEM_GO, n_lines=len(src))


def _params(**kw):
    p = capi.default_params()
    p.semantics, p.mode, p.seed, p.stream_base = capi.SEM_GO, capi.MODE_DETERMINISTIC, SEED, 0
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def _check_tables(g, tag):
    p, a = g.alias(capi.AT_VERTEX)
    assert np.array_equal(p, G[f"{tag}_vertex_prob"]) and np.array_equal(a, G[f"{tag}_vertex_alias"])
    p, a = g.alias(capi.AT_NEGATIVE)
    assert np.array_equal(p, G[f"{tag}_negative_prob"]) and np.array_equal(a, G[f"{tag}_negative_alias"])
    assert np.array_equal(g.sample(capi.SAMPLE_SOURCE, SEED, 1, 2000)[0], G[f"{tag}_source"])
    assert np.array_equal(g.sample(capi.SAMPLE_NEGATIVE, SEED, 2, 2000)[0], G[f"{tag}_negative"])
    st, words = g.sample(capi.SAMPLE_SOURCE_TARGET, SEED, 3, 2000)
    assert np.array_equal(st, G[f"{tag}_source_target"]) and words == int(G[f"{tag}_source_target_words"])


def test_tables_and_samplers():
    src, dst, w = graphs.readme_graph()
    for und in (0, 1):
        _check_tables(_graph(src, dst, w, und), f"readme{und}")
    _check_tables(_graph(G["g60_src"], G["g60_dst"], G["g60_w"], 1), "g60")
    _check_tables(_graph(G["bip_src"], G["bip_dst"], G["bip_w"], 0), "bip")


@pytest.mark.parametrize("order", [2, 1])
def test_line(order):
    g = _graph(G["g60_src"], G["g60_dst"], G["g60_w"], 1)
    m = capi.Model(g, 8, 1 if order == 1 else 2, capi.F64)
    m.set_rows(0, G["g60_init_v"])
    if order == 2:
        m.set_rows(1, G["g60_init_c"])
    st = m.train_line(_params(total=int(G[f"g60_line{order}_total"]), order=order, negative_samples=5))
    assert st["words_stream0"] == int(G[f"g60_line{order}_words"])
    assert np.array_equal(m.get_rows(0), G[f"g60_line{order}_v"])
    if order == 2:
        assert np.array_equal(m.get_rows(1), G[f"g60_line{order}_c"])


def test_deepwalk():
    g = _graph(G["g60_src"], G["g60_dst"], G["g60_w"], 1)
    wt, ws, win, K = (int(x) for x in G["g60_dw_args"])
    m = capi.Model(g, 8, 2, capi.F64)
    m.set_rows(0, G["g60_init_v"]), m.set_rows(1, G["g60_init_c"])
    st = m.train_deepwalk(_params(walk_times=wt, walk_steps=ws, window_min=1, window_max=win, negative_samples=K))
    assert st["words_stream0"] == int(G["g60_dw_words"]) and st["pair_updates"] == int(G["g60_dw_pairs"])
    assert np.array_equal(m.get_rows(0), G["g60_dw_v"]) and np.array_equal(m.get_rows(1), G["g60_dw_c"])


def test_bpr():
    g = _graph(G["bip_src"], G["bip_dst"], G["bip_w"], 0)
    m = capi.Model(g, 8, 2, capi.F64)
    m.set_rows(0, G["bip_init_v"]), m.set_rows(1, G["bip_init_c"])
    st = m.train_bpr(_params(total=int(G["bip_bpr_total"]), lambda_=0.001))
    assert st["words_stream0"] == int(G["bip_bpr_words"])
    assert np.array_equal(m.get_rows(0), G["bip_bpr_v"]) and np.array_equal(m.get_rows(1), G["bip_bpr_c"])
