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


# ---- CPR / TPR (Go tree only): two graphs, three tables -- tests/golden/golden_go_aux_v1.npz ------------------------------
GA = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_go_aux_v1.npz"))


def _aux_csr(src, dst, w, undirected):
    off, col, _, _ = B.edges_to_csr(src, dst, w, undirected)
    return off, col


@pytest.mark.parametrize("tag", ["cpr", "cpr_m0"])
def test_cpr(tag):
    """internal/models/cpr: deterministic fp64 rows bit-identical to the Python restatement (default margin 8 and a margin
    that gates updates); the source-domain rows are only read."""
    g = _graph(GA["cpr_t_src"], GA["cpr_t_dst"], GA["cpr_t_w"], 1)
    V = GA["cpr_init_t"].shape[0]
    alpha, ureg, ireg, margin, total = GA[f"{tag}_args"]
    m = capi.Model(g, 8, 2, capi.F64)
    m.set_rows(0, GA["cpr_init_u"][:V]), m.set_rows(1, GA["cpr_init_t"])
    m.attach_aux(*_aux_csr(GA["cpr_s_src"], GA["cpr_s_dst"], GA["cpr_s_w"], 1), rows=GA["cpr_init_s"])
    st = m.train_cpr(_params(total=int(total), alpha=float(alpha), lambda_=float(ureg), item_reg=float(ireg), margin=float(margin)))
    assert st["words_stream0"] == int(GA[f"{tag}_words"])
    assert np.array_equal(m.get_rows(0), GA[f"{tag}_u"][:V]) and np.array_equal(m.get_rows(1), GA[f"{tag}_t"])
    assert np.array_equal(GA[f"{tag}_u"][V:], GA["cpr_init_u"][V:])  # user rows past the target graph are never trained
    assert np.array_equal(m.get_aux_rows(), GA["cpr_init_s"])


def test_tpr():
    """internal/models/tpr: user, item AND word rows bit-identical to the Python restatement."""
    g = _graph(GA["tpr_ui_src"], GA["tpr_ui_dst"], GA["tpr_ui_w"], 1)
    alpha, lam, tw, total = GA["tpr_args"]
    m = capi.Model(g, 8, 2, capi.F64)
    m.set_rows(0, GA["tpr_init_u"]), m.set_rows(1, GA["tpr_init_i"])
    m.attach_aux(*_aux_csr(GA["tpr_iw_src"], GA["tpr_iw_dst"], GA["tpr_iw_w"], 0), rows=GA["tpr_init_w"])
    st = m.train_tpr(_params(total=int(total), alpha=float(alpha), lambda_=float(lam), text_weight=float(tw)))
    assert st["words_stream0"] == int(GA["tpr_words"])
    assert np.array_equal(m.get_rows(0), GA["tpr_u"]) and np.array_equal(m.get_rows(1), GA["tpr_i"])
    assert np.array_equal(m.get_aux_rows(), GA["tpr_w"])


@pytest.mark.parametrize("model", ["cpr", "tpr"])
def test_cpr_tpr_fp32_and_hogwild(model):
    """The fp32 path that Hogwild runs (red.global.add row deltas): one stream against the oracle within 2e-3 relative to
    the row scale, then a Hogwild run at the library's occupancy moves every table without producing a NaN; a model without
    an attached second graph, and C++ semantics, are refused."""
    a, b = ("cpr_t", "cpr_s") if model == "cpr" else ("tpr_ui", "tpr_iw")
    und_b = 1 if model == "cpr" else 0
    g = _graph(GA[f"{a}_src"], GA[f"{a}_dst"], GA[f"{a}_w"], 1)
    og = B.OracleGraph(B.SEM_GO, *B.edges_to_csr(GA[f"{a}_src"], GA[f"{a}_dst"], GA[f"{a}_w"], 1)[:3], max_line=len(GA[f"{a}_src"]))
    ob = B.OracleGraph(B.SEM_GO, *B.edges_to_csr(GA[f"{b}_src"], GA[f"{b}_dst"], GA[f"{b}_w"], und_b)[:3], max_line=len(GA[f"{b}_src"]))
    if model == "cpr":
        V = GA["cpr_init_t"].shape[0]
        t0, t1, t2 = GA["cpr_init_u"][:V].copy(), GA["cpr_init_t"].copy(), GA["cpr_init_s"].copy()
    else:
        t0, t1, t2 = GA["tpr_init_u"].copy(), GA["tpr_init_i"].copy(), GA["tpr_init_w"].copy()
    total = 20000
    m = capi.Model(g, 8, 2, capi.F32)
    with pytest.raises(capi.SmoreError):
        (m.train_cpr if model == "cpr" else m.train_tpr)(_params(total=100))
    m.set_rows(0, t0), m.set_rows(1, t1)
    m.attach_aux(*_aux_csr(GA[f"{b}_src"], GA[f"{b}_dst"], GA[f"{b}_w"], und_b), rows=t2)
    p = _params(total=total, alpha=0.05, lambda_=0.01, item_reg=0.01, margin=8.0, text_weight=0.5)
    r0, r1, r2 = t0.copy(), t1.copy(), t2.copy()
    if model == "cpr":
        m.train_cpr(p)
        og.train_cpr_go(ob, r0, r1, r2, 0.05, 0.01, 0.01, 8.0, total, total, SEED, 0)
    else:
        m.train_tpr(p)
        og.train_tpr_go(ob, r0, r1, r2, 0.05, 0.01, 0.5, total, total, SEED, 0)
    for got, want in ((m.get_rows(0), r0), (m.get_rows(1), r1), (m.get_aux_rows(), r2)):
        assert np.abs(got - want).max() < 2e-3 * max(np.abs(want).max(), 1e-3)
    # Hogwild
    h = capi.Model(g, 8, 2, capi.F32)
    h.set_rows(0, t0), h.set_rows(1, t1)
    h.attach_aux(*_aux_csr(GA[f"{b}_src"], GA[f"{b}_dst"], GA[f"{b}_w"], und_b), rows=t2)
    p.mode, p.total = capi.MODE_HOGWILD, 200_000
    st = (h.train_cpr if model == "cpr" else h.train_tpr)(p)
    assert st["samples"] > 0.9 * p.total
    for got, init in ((h.get_rows(0), t0), (h.get_rows(1), t1)):
        assert np.isfinite(got).all() and not np.array_equal(got, init)
    aux = h.get_aux_rows()
    t2f = t2.astype(np.float32).astype(np.float64)  # (CPR only reads its third table)
    assert np.isfinite(aux).all() and (np.array_equal(aux, t2f) if model == "cpr" else not np.array_equal(aux, t2f))
    # random third table when no rows are given; C++ semantics have neither model
    h.attach_aux(*_aux_csr(GA[f"{b}_src"], GA[f"{b}_dst"], GA[f"{b}_w"], und_b), rows=None, seed=5)
    r = h.get_aux_rows()
    assert np.abs(r).max() <= 0.5 / 8 and r.std() > 0.01
    gc = capi.Graph.from_csr(*B.edges_to_csr(GA[f"{a}_src"], GA[f"{a}_dst"], GA[f"{a}_w"], 1)[:3])
    mc = capi.Model(gc, 8, 2, capi.F32)
    mc.attach_aux(*_aux_csr(GA[f"{b}_src"], GA[f"{b}_dst"], GA[f"{b}_w"], und_b))
    with pytest.raises(capi.SmoreError):
        (mc.train_cpr if model == "cpr" else mc.train_tpr)(capi.default_params())


def test_cpr_and_tpr_cli(tmp_path):
    """cmd/cpr/main.go and cmd/tpr/main.go: two edge lists in, three "%.6f" files out, rows named by the graph they belong to."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    t, s = str(tmp_path / "t.txt"), str(tmp_path / "s.txt")
    B.write_edge_list(t, GA["cpr_t_src"], GA["cpr_t_dst"], GA["cpr_t_w"])
    B.write_edge_list(s, GA["cpr_s_src"], GA["cpr_s_dst"], GA["cpr_s_w"])
    outs = [str(tmp_path / f"{k}.rep") for k in ("user", "target", "source")]
    r = subprocess.run([os.path.join(root, "smore_b200", "bin", "cpr"), "-train_target", t, "-train_source", s, "-save_user", outs[0],
                        "-save_target", outs[1], "-save_source", outs[2], "-dimensions", "16", "-update_times", "1", "-threads", "2"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "[CPR: Cross-Domain Preference Ranking]" in r.stdout and "Progress: 100.00 %" in r.stdout
    Vt, Vs = GA["cpr_init_t"].shape[0], GA["cpr_init_s"].shape[0]
    for path, rows in zip(outs, (Vt, Vt, Vs)):
        lines = open(path).read().split("\n")
        assert lines[0] == f"{rows} 16" and len(lines) == rows + 2
        assert len(lines[1].split(" ")) == 17 and len(lines[1].split(" ")[1].split(".")[1]) == 6
    assert open(outs[0]).read().split("\n")[1].split(" ")[0] == f"v{GA['cpr_t_src'][0]}"  # first name of the target graph
    assert open(outs[2]).read().split("\n")[2].split(" ")[0] == f"v{GA['cpr_s_dst'][0]}"  # second name of the source graph
    ui, iw = str(tmp_path / "ui.txt"), str(tmp_path / "iw.txt")
    B.write_edge_list(ui, GA["tpr_ui_src"], GA["tpr_ui_dst"], GA["tpr_ui_w"])
    B.write_edge_list(iw, GA["tpr_iw_src"], GA["tpr_iw_dst"], GA["tpr_iw_w"])
    outs = [str(tmp_path / f"{k}.rep") for k in ("u", "i", "w")]
    r = subprocess.run([os.path.join(root, "smore_b200", "bin", "tpr"), "-train_ui", ui, "-train_iw", iw, "-save_user", outs[0],
                        "-save_item", outs[1], "-save_word", outs[2], "-dimensions", "16", "-sample_times", "20"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "[TPR: Text-aware Preference Ranking]" in r.stdout
    assert open(outs[2]).read().split("\n")[0] == f"{GA['tpr_init_w'].shape[0]} 16"
    r = subprocess.run([os.path.join(root, "smore_b200", "bin", "tpr"), "-train_ui", ui, "-save_user", outs[0]], capture_output=True, text=True)
    assert r.returncode != 0 and "-train_iw" in r.stderr
