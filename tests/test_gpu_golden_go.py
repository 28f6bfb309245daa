"""GPU: the Go-semantics kernels against the fixtures of the independent plain-Python restatement of the Go tree
(tests/golden/go_restatement.py -> golden_go_v1.npz): alias tables and sampler replays bit-exact, deterministic fp64
embeddings of LINE order 1 / 2, BPR and DeepWalk bit-identical (north-star bar: 1e-5 relative)."""
import os
import subprocess

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


@pytest.mark.parametrize("tag,und", [("g60", 1), ("bip", 0)])
def test_node2vec(tag, und):
    """Go-only model (internal/models/node2vec): deterministic fp64 embeddings bit-identical to the Python restatement, on
    an undirected graph with hubs and parallel edges and on a directed one whose walks stop at sinks."""
    g = _graph(G[f"{tag}_src"], G[f"{tag}_dst"], G[f"{tag}_w"], und)
    wt, ws, win, K, p, q = G[f"{tag}_n2v_args"]
    m = capi.Model(g, 8, 2, capi.F64)
    m.set_rows(0, G[f"{tag}_init_v"]), m.set_rows(1, G[f"{tag}_init_c"])
    st = m.train_node2vec(_params(walk_times=int(wt), walk_steps=int(ws), window_min=1, window_max=int(win), negative_samples=int(K),
                                  n2v_p=float(p), n2v_q=float(q)))
    assert st["words_stream0"] == int(G[f"{tag}_n2v_words"]) and st["pair_updates"] == int(G[f"{tag}_n2v_pairs"])
    assert np.array_equal(m.get_rows(0), G[f"{tag}_n2v_v"]) and np.array_equal(m.get_rows(1), G[f"{tag}_n2v_c"])
    # C++ semantics have no node2vec
    gc = capi.Graph.from_csr(*B.edges_to_csr(G[f"{tag}_src"], G[f"{tag}_dst"], G[f"{tag}_w"], und)[:3])
    mc = capi.Model(gc, 8, 2, capi.F64)
    pc = capi.default_params()
    with pytest.raises(capi.SmoreError):
        mc.train_node2vec(pc)


def test_node2vec_cli(tmp_path):
    """cmd/node2vec/main.go flags (-p -q, window 10, walk_steps 80) and the Go "%.6f" writer."""
    net, rep = str(tmp_path / "net.txt"), str(tmp_path / "rep.txt")
    B.write_edge_list(net, G["g60_src"], G["g60_dst"], G["g60_w"])
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([os.path.join(root, "smore_b200", "bin", "node2vec"), "-train", net, "-save", rep, "-dimensions", "16",
                        "-walk_times", "2", "-p", "0.5", "-q", "2", "-threads", "4"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lines = open(rep).read().split("\n")
    n, dim = map(int, lines[0].split())
    assert dim == 16 and len(lines) == n + 2 and len(lines[1].split(" ")) == 17 and len(lines[1].split(" ")[1].split(".")[1]) == 6
    r = subprocess.run([os.path.join(root, "smore_b200", "bin", "node2vec"), "-train", net, "-save", rep, "-semantics", "cpp"],
                       capture_output=True, text=True)
    assert r.returncode != 0 and "only in the Go tree" in r.stderr
