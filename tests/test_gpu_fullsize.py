"""GPU, BASELINE.json full sizes (configs[0] BPR 10k x 10k / 1M edges; configs[1] LINE V = 1M / 10M edges): properties that
do not need an oracle run at that scale -- alias tables reproduce their target distribution, every sampled pair is a real
edge and hot sources appear with the right frequency, training moves only rows it may move, repeated deterministic runs
are bit-identical, Hogwild bookkeeping adds up."""
import numpy as np
import pytest

from smore_b200 import capi, synth

pytestmark = pytest.mark.gpu
SEED = 20261018


def implied_distribution(prob, alias):
    n = len(prob)
    p = prob / n
    a = np.where(alias < 0, np.arange(n), alias)
    np.add.at(p, a, (1.0 - prob) / n)
    return p


@pytest.fixture(scope="module")
def c2():
    src, dst, w = synth.power_law_edges(1_000_000, 10_000_000, SEED)
    off, col, ww, _ = synth.csr_from_edges(src, dst, w, True)
    g = capi.Graph.from_csr(off, col, ww, semantics=capi.SEM_CPP)
    return g, off, col, ww


def test_c2_alias_tables_reproduce_their_distributions(c2):
    g, off, col, ww = c2
    out_deg = np.add.reduceat(ww, off[:-1])
    in_deg = np.bincount(col, weights=ww, minlength=g.V)
    for which, dist in ((capi.AT_VERTEX, out_deg), (capi.AT_NEGATIVE, in_deg + out_deg)):
        prob, alias = g.alias(which)
        want = dist ** 0.75
        want /= want.sum()
        # the reference's fp64 Vose construction itself drifts by ~3e-12 over 1M entries (same drift in the oracle)
        assert np.abs(implied_distribution(prob, alias) - want).max() < 1e-10
    # context table: per-vertex sub-tables; check the 1000 largest neighbourhoods
    prob, alias = g.alias(capi.AT_CONTEXT)
    for v in np.argsort(-np.diff(off))[:1000]:
        o, b = off[v], off[v + 1] - off[v]
        p = prob[o:o + b] / b
        tgt = np.zeros(g.V)
        np.add.at(tgt, col[o:o + b], p)
        a = alias[o:o + b]
        np.add.at(tgt, np.where(a < 0, col[o:o + b], a), (1.0 - prob[o:o + b]) / b)
        want = np.zeros(g.V)
        np.add.at(want, col[o:o + b], ww[o:o + b] ** 0.75)
        assert np.abs(tgt - want / want.sum()).max() < 1e-12


def test_c2_device_sampler_statistics(c2):
    g, off, col, ww = c2
    n = 4_000_000
    st, used = g.sample(capi.SAMPLE_SOURCE_TARGET, SEED, 3, n)
    assert used == 4 * n
    s, t = st[0::2], st[1::2]
    # every sampled (s, t) is an edge: t must occur in s' adjacency slice
    pick = np.random.default_rng(0).integers(0, n, 20000)
    for i in pick:
        sl = col[off[s[i]]:off[s[i] + 1]]
        assert t[i] in sl
    out_deg = np.add.reduceat(ww, off[:-1])
    p = out_deg ** 0.75
    p /= p.sum()
    hot = np.argsort(-p)[:200]
    cnt = np.bincount(s, minlength=g.V)[hot]
    sigma = np.sqrt(n * p[hot] * (1 - p[hot]))
    assert (np.abs(cnt - n * p[hot]) < 5 * sigma).all()


def test_c2_hogwild_step_properties(c2):
    g, off, col, ww = c2
    dim = 128
    m = capi.Model(g, dim, 2, capi.F32)
    m.init(0, True, 1), m.init(1, False, 1)
    p = capi.default_params()
    p.mode, p.seed, p.total = capi.MODE_HOGWILD, 5, 1 << 24
    st = m.train_line(p)
    assert st["pair_updates"] == st["samples"] and 0.99 * p.total <= st["samples"] <= p.total
    chk = np.random.default_rng(1).integers(0, g.V - 4096, 8)
    for first in chk:  # spot-check slices of both tables
        for t in (0, 1):
            W = m.get_rows(t, int(first), 4096, dtype=np.float32)
            assert np.isfinite(W).all() and np.abs(W).max() < 10.0
    # context rows only move when sampled: after 16M updates on a 1M-vertex graph (almost) every row moved, and the
    # positive pairs now score above random pairs
    Wv = m.get_rows(0, 0, 200000, dtype=np.float32)
    Wc = m.get_rows(1, 0, 200000, dtype=np.float32)
    assert (np.abs(Wc).max(axis=1) > 0).mean() > 0.9
    srcs = np.repeat(np.arange(200000), np.diff(off[:200001]))
    cols = col[: off[200000]]
    ok = cols < 200000
    rng = np.random.default_rng(2)
    pick = rng.integers(0, ok.sum(), 50000)
    pos = np.einsum("ij,ij->i", Wv[srcs[ok][pick]], Wc[cols[ok][pick]])
    neg = np.einsum("ij,ij->i", Wv[srcs[ok][pick]], Wc[rng.integers(0, 200000, 50000)])
    assert pos.mean() > neg.mean()


def test_c2_deterministic_mode_is_reproducible(c2):
    g, off, col, ww = c2
    res = []
    for _ in range(2):
        m = capi.Model(g, 128, 2, capi.F32)
        m.init(0, True, 9), m.init(1, False, 9)
        p = capi.default_params()
        p.mode, p.seed, p.total = capi.MODE_DETERMINISTIC, 11, 20000
        st = m.train_line(p)
        assert st["words_stream0"] == (20000 - 1) * 14
        res.append((m.get_rows(0, 0, 100000, dtype=np.float32), m.get_rows(1, 0, 100000, dtype=np.float32)))
        m.close()
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1])


def test_c1_bpr_full_size_properties():
    """configs[0]: 10k users x 10k items, 1M weighted edges, BPR dim 64 (Go semantics, the CPU-runnable case)."""
    nu = 10_000
    src, dst, w = synth.bipartite_edges(nu, 10_000, 1_000_000, 7)
    off, col, ww, labels = synth.csr_from_edges(src, dst, w, False)
    g = capi.Graph.from_csr(off, col, ww, semantics=capi.SEM_GO, n_lines=len(src))
    is_item = labels >= nu
    m = capi.Model(g, 64, 2, capi.F32)
    m.init(0, True, 1), m.init(1, True, 2)
    Wv0, Wc0 = m.get_rows(0), m.get_rows(1)
    p = capi.default_params()
    p.semantics, p.mode, p.seed, p.total, p.lambda_ = capi.SEM_GO, capi.MODE_HOGWILD, 3, 10 * len(src), 0.001
    st = m.train_bpr(p)
    assert 0.99 * p.total <= st["samples"] <= p.total
    Wv, Wc = m.get_rows(0), m.get_rows(1)
    assert np.isfinite(Wv).all() and np.isfinite(Wc).all()
    # users are the only sources (directed user -> item): item rows of the vertex table never move
    assert np.array_equal(Wv[is_item], Wv0[is_item])
    assert (np.abs(Wv[~is_item] - Wv0[~is_item]).max(axis=1) > 0).all()
    # ranking quality on the training interactions: a user's items outrank random items
    rng = np.random.default_rng(0)
    e = rng.integers(0, len(col), 100000)
    srcs = np.repeat(np.arange(g.V), np.diff(off))
    items = np.flatnonzero(is_item)
    pos = np.einsum("ij,ij->i", Wv[srcs[e]], Wc[col[e]])
    neg = np.einsum("ij,ij->i", Wv[srcs[e]], Wc[items[rng.integers(0, len(items), 100000)]])
    assert (pos > neg).mean() > 0.8
