"""GPU parity of the fp32 THROUGHPUT path (the instantiation bench.py times) against the fp64 CPU oracle: every trainer,
deterministic single stream, same Philox words, short runs.

Why a tolerance and short runs: fp32 tables are updated with red.global.add of fp32 deltas, dot products are butterfly sums
and products contract into FMAs, so a score differs from the oracle's by ~1e-7 relative; once such a score lands in the
neighbouring bin of the 1000-entry sigmoid table (or flips a margin gate of WARP / HOP-Rec / Skew-OPT) the two runs diverge
for good. On the runs below that does not happen (the consumed word count, which depends on every gate decision, is
asserted equal) and the embeddings agree to REL_TOL = 2e-3 of the largest entry -- the bar tests/test_gpu_parity.py already
sets for LINE. The bit-exact checks of the same kernels live in the fp64 tests."""
import numpy as np
import pytest

from oracle import bindings as B
from smore_b200 import capi
from tests import graphs
from tests.test_gpu_parity import make, params, rel_err, SEED

pytestmark = pytest.mark.gpu
REL_TOL = 2e-3


def _model(dg, dim, tables, *rows):
    m = capi.Model(dg, dim, tables, capi.F32)
    for t, W in enumerate(rows):
        m.set_rows(t, W)
    return m


@pytest.mark.parametrize("order,dim", [(2, 128), (1, 64)])
def test_line_go_f32(order, dim):
    src, dst, w = graphs.random_graph(500, 6000, seed=13)
    og, dg, _ = make(src, dst, w, 1, capi.SEM_GO)
    total = 5000
    Wv, Wc = graphs.init_tables(og.V, dim, seed=1)
    a, c = Wv.copy(), Wc.copy()
    pos = og.train_line_go(a, c, order, 5, 0.025, total, SEED, 0)
    m = _model(dg, dim, 1 if order == 1 else 2, *((Wv,) if order == 1 else (Wv, Wc)))
    st = m.train_line(params(capi.SEM_GO, total=total, order=order))
    assert st["words_stream0"] == pos
    assert rel_err(m.get_rows(0), a) < REL_TOL
    if order == 2:
        assert rel_err(m.get_rows(1), c) < REL_TOL


@pytest.mark.parametrize("dim", [64, 128])
def test_bpr_go_f32(dim):
    src, dst, w = graphs.bipartite_graph(200, 150, 4000, seed=19)
    og, dg, _ = make(src, dst, w, 0, capi.SEM_GO)
    total = 8000
    Wv, Wc = graphs.init_tables(og.V, dim, seed=3)
    a, c = Wv.copy(), Wc.copy()
    pos = og.train_bpr_go(a, c, 0.025, 0.001, total, SEED, 0)
    m = _model(dg, dim, 2, Wv, Wc)
    st = m.train_bpr(params(capi.SEM_GO, total=total, lambda_=0.001))
    assert st["words_stream0"] == pos
    assert rel_err(m.get_rows(0), a) < REL_TOL and rel_err(m.get_rows(1), c) < REL_TOL


@pytest.mark.parametrize("dim", [64, 128])
def test_bpr_cpp_f32(dim):
    src, dst, w = graphs.bipartite_graph(200, 150, 4000, seed=23)
    og, dg, _ = make(src, dst, w, 0, capi.SEM_CPP, neg=capi.NEG_NO_DEGREES)
    total = 4000
    W, _ = graphs.init_tables(og.V, dim, seed=4)
    a = W.copy()
    pos = og.train_bpr_cpp(a, 0.025, total, SEED, 0)
    m = _model(dg, dim, 1, W)
    st = m.train_bpr(params(capi.SEM_CPP, total=total))
    assert st["words_stream0"] == pos
    assert rel_err(m.get_rows(0), a) < REL_TOL


@pytest.mark.parametrize("dim,scale", [(128, 1.0), (64, 40.0)])
def test_warp_f32(dim, scale):
    src, dst, w = graphs.bipartite_graph(200, 150, 4000, seed=29)
    og, dg, _ = make(src, dst, w, 0, capi.SEM_CPP, neg=capi.NEG_NO_DEGREES)
    total = 4000
    W, _ = graphs.init_tables(og.V, dim, seed=5)
    W *= scale
    a = W.copy()
    pos, tries = og.train_warp_cpp(a, 0.025, total, SEED, 0)
    m = _model(dg, dim, 1, W)
    st = m.train_warp(params(capi.SEM_CPP, total=total))
    assert st["words_stream0"] == pos  # every margin test (f < 1) went the oracle's way
    assert abs(st["mean_tries"] - tries / total) < 1e-12
    assert rel_err(m.get_rows(0), a) < REL_TOL


def test_hoprec_f32():
    nu, ni, dim, steps = 120, 80, 128, 3
    src, dst, w = graphs.bipartite_graph(nu, ni, 2500, seed=31)
    og, dg, ids = make(src, dst, w, 1, capi.SEM_CPP, neg=capi.NEG_NO_DEGREES)
    fl = np.zeros(og.V, dtype=np.int32)
    for l, vid in ids.items():
        fl[vid] = 0 if l < nu else 1
    og.set_field(fl)
    dg.set_field(fl)
    total = 1500
    W, _ = graphs.init_tables(og.V, dim, seed=6)
    W *= 8.0
    a = W.copy()
    pos = og.train_hoprec_cpp(a, steps, 0.025, total, SEED, 0)
    m = _model(dg, dim, 1, W)
    st = m.train_hoprec(params(capi.SEM_CPP, total=total, walk_steps=steps))
    assert st["words_stream0"] == pos  # every gate (f > margin) and rejection loop went the oracle's way
    assert rel_err(m.get_rows(0), a) < REL_TOL


@pytest.mark.parametrize("sem,walklets", [(capi.SEM_CPP, 0), (capi.SEM_CPP, 1), (capi.SEM_GO, 0)])
def test_walk_models_f32(sem, walklets):
    src, dst, w = graphs.random_graph(150, 700, seed=37)
    og, dg, _ = make(src, dst, w, 1, sem)
    dim = 128
    Wv, Wc = graphs.init_tables(og.V, dim, seed=7)
    a, c = Wv.copy(), Wc.copy()
    wt, ws, max_walks = 1, 20, 40
    if sem == capi.SEM_CPP:
        w0, w1 = (2, 4) if walklets else (5, 0)
        pos, pairs = og.train_walk_cpp(walklets, a, c, wt, ws, w0, w1, 5, 0.025, SEED, 0, max_walks)
    else:
        pos, pairs = og.train_deepwalk_go(a, c, wt, ws, 5, 5, 0.025, SEED, 0, max_walks)
    m = _model(dg, dim, 2, Wv, Wc)
    p = params(sem, walk_times=wt, walk_steps=ws, window_min=2 if walklets else 1, window_max=4 if walklets else 5,
               max_walks=max_walks)
    st = m.train_walklets(p) if walklets else m.train_deepwalk(p)
    assert st["words_stream0"] == pos and st["pair_updates"] == pairs
    assert rel_err(m.get_rows(0), a) < REL_TOL and rel_err(m.get_rows(1), c) < REL_TOL


def test_node2vec_f32():
    """Go tree only: the biased second-order walk (fp64 weights inside, whatever the table type) feeding fp32 pair updates."""
    src, dst, w = graphs.random_graph(150, 700, seed=37)
    og, dg, _ = make(src, dst, w, 1, capi.SEM_GO)
    dim = 128
    Wv, Wc = graphs.init_tables(og.V, dim, seed=7)
    a, c = Wv.copy(), Wc.copy()
    wt, ws, max_walks = 1, 20, 40
    pos, pairs = og.train_node2vec_go(a, c, wt, ws, 5, 5, 0.025, 0.5, 2.0, SEED, 0, max_walks)
    m = _model(dg, dim, 2, Wv, Wc)
    p = params(capi.SEM_GO, walk_times=wt, walk_steps=ws, window_min=1, window_max=5, max_walks=max_walks, n2v_p=0.5, n2v_q=2.0)
    st = m.train_node2vec(p)
    assert st["words_stream0"] == pos and st["pair_updates"] == pairs
    assert rel_err(m.get_rows(0), a) < REL_TOL and rel_err(m.get_rows(1), c) < REL_TOL


def test_hpe_f32():
    src, dst, w = graphs.random_graph(300, 4000, seed=47)
    og, dg, _ = make(src, dst, w, 1, capi.SEM_CPP)
    dim, total, steps, K, reg = 128, 1500, 5, 5, 0.01
    Wv, Wc = graphs.init_tables(og.V, dim, seed=8)
    a, c = Wv.copy(), Wc.copy()
    pos = og.train_hpe_cpp(a, c, steps, K, reg, 0.025, total, SEED, 0)
    m = _model(dg, dim, 2, Wv, Wc)
    st = m.train_hpe(params(capi.SEM_CPP, total=total, walk_steps=steps, negative_samples=K, lambda_=reg))
    assert st["words_stream0"] == pos
    assert rel_err(m.get_rows(0), a) < REL_TOL and rel_err(m.get_rows(1), c) < REL_TOL


def test_mf_f32():
    src, dst, w = graphs.bipartite_graph(200, 150, 4000, seed=53)
    og, dg, _ = make(src, dst, w, 0, capi.SEM_CPP, neg=capi.NEG_NO_DEGREES)
    dim, total, K, reg = 128, 5000, 5, 0.01
    W, _ = graphs.init_tables(og.V, dim, seed=9)
    a = W.copy()
    pos = og.train_mf_cpp(a, K, reg, 0.025, total, SEED, 0)
    m = _model(dg, dim, 1, W)
    st = m.train_mf(params(capi.SEM_CPP, total=total, negative_samples=K, lambda_=reg))
    assert st["words_stream0"] == pos
    assert rel_err(m.get_rows(0), a) < REL_TOL


def test_skewopt_f32():
    src, dst, w = graphs.bipartite_graph(200, 150, 4000, seed=59)
    og, dg, _ = make(src, dst, w, 0, capi.SEM_CPP, neg=capi.NEG_NO_DEGREES)
    dim, total = 128, 2000
    W, _ = graphs.init_tables(og.V, dim, seed=10)
    a = W.copy()
    pos = og.train_skewopt_cpp(a, 10.0, 3.0, 3, 0.025, total, SEED, 0)
    m = _model(dg, dim, 1, W)
    st = m.train_skewopt(params(capi.SEM_CPP, total=total, xi=10.0, omega=3.0, eta=3))
    assert st["words_stream0"] == pos
    assert rel_err(m.get_rows(0), a) < REL_TOL
