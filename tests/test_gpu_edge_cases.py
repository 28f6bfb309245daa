"""GPU: edge cases and error behaviour of the C ABI (ragged / degenerate inputs, parameter limits, bad arguments)."""
import numpy as np
import pytest

from oracle import bindings as B
from smore_b200 import capi
from tests import graphs

pytestmark = pytest.mark.gpu
SEED = 7


def both(src, dst, w, und, sem=capi.SEM_CPP, neg=capi.NEG_DEGREES):
    off, col, ww, ids = B.edges_to_csr(src, dst, w, und)
    n_lines = len(col) if sem == capi.SEM_CPP else len(src)
    return (B.OracleGraph(sem, off, col, ww, max_line=n_lines, neg_method=neg),
            capi.Graph.from_csr(off, col, ww, semantics=sem, negative_method=neg, n_lines=n_lines))


def det(sem, **kw):
    p = capi.default_params()
    p.semantics, p.mode, p.seed = sem, capi.MODE_DETERMINISTIC, SEED
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def test_single_edge_graph():
    og, dg = both(np.array([0]), np.array([1]), np.array([2.0]), 1)
    Wv, Wc = graphs.init_tables(2, 8, 1, context_zero=True)
    a, c = Wv.copy(), Wc.copy()
    pos = og.train_line_cpp(a, c, 5, 0.025, 500, SEED, 0)
    m = capi.Model(dg, 8, 2, capi.F64)
    m.set_rows(0, Wv), m.set_rows(1, Wc)
    st = m.train_line(det(capi.SEM_CPP, total=500))
    assert st["words_stream0"] == pos
    assert np.array_equal(m.get_rows(0), a) and np.array_equal(m.get_rows(1), c)


@pytest.mark.parametrize("K", [0, 1, 7, 31])
def test_negative_sample_counts(K):
    src, dst, w = graphs.random_graph(100, 900, seed=3)
    og, dg = both(src, dst, w, 1)
    Wv, Wc = graphs.init_tables(og.V, 12, 2, context_zero=True)
    a, c = Wv.copy(), Wc.copy()
    pos = og.train_line_cpp(a, c, K, 0.025, 3000, SEED, 0)
    m = capi.Model(dg, 12, 2, capi.F64)
    m.set_rows(0, Wv), m.set_rows(1, Wc)
    st = m.train_line(det(capi.SEM_CPP, total=3000, negative_samples=K))
    assert st["words_stream0"] == pos == (3000 - 1) * (4 + 2 * K)
    assert np.array_equal(m.get_rows(0), a) and np.array_equal(m.get_rows(1), c)


@pytest.mark.parametrize("dim", [1, 3, 33, 100, 256])
def test_odd_dimensions(dim):
    src, dst, w = graphs.random_graph(80, 700, seed=5)
    og, dg = both(src, dst, w, 1, sem=capi.SEM_GO)
    Wv, Wc = graphs.init_tables(og.V, dim, 3)
    a, c = Wv.copy(), Wc.copy()
    og.train_line_go(a, c, 2, 5, 0.025, 2000, SEED, 0)
    m = capi.Model(dg, dim, 2, capi.F64)
    m.set_rows(0, Wv), m.set_rows(1, Wc)
    m.train_line(det(capi.SEM_GO, total=2000))
    assert np.array_equal(m.get_rows(0), a) and np.array_equal(m.get_rows(1), c)


def test_isolated_sinks_and_duplicate_edges():
    # directed: vertices 3..5 are sinks, (0 -> 1) appears three times (kept as parallel entries), self loop on 2
    src = np.array([0, 0, 0, 1, 2, 2, 0, 1])
    dst = np.array([1, 1, 1, 3, 2, 4, 5, 0])
    w = np.array([1.0, 2.0, 3.0, 1.0, 5.0, 1.0, 1.0, 2.0])
    for sem in (capi.SEM_CPP, capi.SEM_GO):
        og, dg = both(src, dst, w, 0, sem=sem)
        assert np.array_equal(og.sample(3, SEED, 0, 3000)[0], dg.sample(capi.SAMPLE_SOURCE_TARGET, SEED, 0, 3000)[0])
        for start in range(og.V):
            a = og.walk_pairs(SEED, start, start, 12, 0, 3)
            b = dg.walk_pairs(SEED, start, start, 12, 0, 3)
            assert all(np.array_equal(x, y) for x, y in zip(a, b))
    # LINE-1 on the shared table exercises the self-loop (vertex row == context row) ordered path
    og, dg = both(src, dst, w, 0)
    W, _ = graphs.init_tables(og.V, 8, 4)
    W *= 30
    a = W.copy()
    og.train_line_cpp(a, a, 5, 0.05, 4000, SEED, 0)
    m = capi.Model(dg, 8, 1, capi.F64)
    m.set_rows(0, W)
    m.train_line(det(capi.SEM_CPP, total=4000, order=1, alpha=0.05))
    assert np.array_equal(m.get_rows(0), a)


def test_argument_errors():
    off = np.array([0, 1, 2], dtype=np.int64)
    col = np.array([1, 0], dtype=np.int32)
    w = np.ones(2)
    with pytest.raises(capi.SmoreError, match="out of range"):
        capi.Graph.from_csr(off, np.array([1, 7], dtype=np.int32), w)
    with pytest.raises(capi.SmoreError, match="row_off"):
        capi.Graph.from_csr(np.array([0, 2, 1], dtype=np.int64), col, w)
    g = capi.Graph.from_csr(off, col, w)
    with pytest.raises(capi.SmoreError, match="dim"):
        capi.Model(g, 0, 2)
    with pytest.raises(capi.SmoreError, match="unsupported"):
        m = capi.Model(g, 1000, 2)
        m.train_line(det(capi.SEM_CPP, total=10))
    m = capi.Model(g, 8, 2)
    with pytest.raises(capi.SmoreError, match="semantics"):
        m.train_line(det(capi.SEM_GO, total=10))
    with pytest.raises(capi.SmoreError, match="negative_samples"):
        m.train_line(det(capi.SEM_CPP, total=10, negative_samples=40))
    with pytest.raises(capi.SmoreError, match="alpha"):
        m.train_line(det(capi.SEM_CPP, total=10, alpha=0.0))
    with pytest.raises(capi.SmoreError, match="only in the C\\+\\+ tree"):
        gg = capi.Graph.from_csr(off, col, w, semantics=capi.SEM_GO)
        capi.Model(gg, 8, 2).train_warp(det(capi.SEM_GO, total=10))
    with pytest.raises(capi.SmoreError, match="field"):
        capi.Model(g, 8, 1).train_hoprec(det(capi.SEM_CPP, total=10, walk_steps=2))
    with pytest.raises(capi.SmoreError, match="bounds"):
        m.set_rows(0, np.zeros((5, 8)))
    with pytest.raises(capi.SmoreError, match="tables"):
        capi.Model(g, 8, 1).train_line(det(capi.SEM_CPP, total=10))  # LINE-2 needs two tables
    with pytest.raises(capi.SmoreError, match="cannot open"):
        capi.Graph.from_edge_list("/nonexistent/file.txt", True)


def test_ingest_skips_malformed_lines(tmp_path):
    path = tmp_path / "net.txt"
    path.write_text("a b 1\nonly_two_fields 3\n\nc a 2.5\nb c notanumber\n  d   a\t4  \n")
    for sem, lines in ((capi.SEM_CPP, 6), (capi.SEM_GO, 3)):
        g = capi.Graph.from_edge_list(str(path), True, semantics=sem)
        assert g.names() == ["a", "b", "c", "d"]
        assert g.V == 4 and g.E == 6 and g.n_lines == lines
        off, col, w = g.csr()
        assert col[off[0]:off[1]].tolist() == [1, 2, 3] and w[off[0]:off[1]].tolist() == [1.0, 2.5, 4.0]


def test_chunked_schedule_continues_the_learning_rate():
    """sched_total / sched_offset: a call can be one chunk of a longer LR schedule. With alpha at its floor for the whole
    second chunk (offset == total), rows move by at most alpha*1e-4-sized steps; with offset 0 they move normally."""
    src, dst, w = graphs.random_graph(150, 1500, seed=9)
    _, dg = both(src, dst, w, 1)
    Wv, Wc = graphs.init_tables(dg.V, 16, 5)
    moved = []
    for offset in (0, 10**9):
        m = capi.Model(dg, 16, 2, capi.F64)
        m.set_rows(0, Wv), m.set_rows(1, Wc)
        m.train_line(det(capi.SEM_CPP, total=5000, sched_total=10**9, sched_offset=offset))
        moved.append(np.abs(m.get_rows(0) - Wv).max())
    assert moved[0] > 1e-4 and moved[1] < moved[0] * 1e-3


def test_async_row_copies_are_ordered_on_the_device():
    """smore_model_{set,get}_rows_f32_async: an upload is ordered after the read-backs of the same TABLE issued before it (and a
    read-back after earlier uploads of that table) by events between the two copy streams -- the host does not have to wait in
    between, and copies of the other table are free to run next to them. bench.py's end-to-end leg relies on it to
    double-buffer two model instances."""
    import torch

    src, dst, w = graphs.random_graph(5000, 40000, seed=3)
    off, col, ww, _ = B.edges_to_csr(src, dst, w, 1)
    g = capi.Graph.from_csr(off, col, ww)
    V, dim = len(off) - 1, 128
    m = capi.Model(g, dim, 2, capi.F32)
    a = torch.full((V, dim), 1.0, dtype=torch.float32).pin_memory()
    b = torch.full((V, dim), 2.0, dtype=torch.float32).pin_memory()
    out1 = torch.zeros((V, dim), dtype=torch.float32).pin_memory()
    out2 = torch.zeros((V, dim), dtype=torch.float32).pin_memory()
    c = torch.full((V, dim), 3.0, dtype=torch.float32).pin_memory()
    out3 = torch.zeros((V, dim), dtype=torch.float32).pin_memory()
    out4 = torch.zeros((V, dim), dtype=torch.float32).pin_memory()
    for _ in range(20):  # no host wait between the calls
        m.set_rows_async(0, a.numpy())
        m.set_rows_async(1, c.numpy())
        m.get_rows_async(0, out1.numpy())   # must see `a`
        m.set_rows_async(0, b.numpy())      # must not overtake the read-back above
        m.get_rows_async(1, out3.numpy())   # the other table, interleaved: must see `c`
        m.set_rows_async(1, a.numpy())      # must not overtake the read-back of table 1
        m.get_rows_async(0, out2.numpy())   # must see `b`
        m.get_rows_async(1, out4.numpy())   # must see `a`
        m.wait_copies()
        assert float(out1.min()) == 1.0 == float(out1.max())
        assert float(out2.min()) == 2.0 == float(out2.max())
        assert float(out3.min()) == 3.0 == float(out3.max())
        assert float(out4.min()) == 1.0 == float(out4.max())
        out1.zero_(), out2.zero_(), out3.zero_(), out4.zero_()
    with pytest.raises(capi.SmoreError):
        m.set_rows_async(0, a.numpy()[: V - 1], first=2)
