"""Pins the restated oracle (oracle/smore_oracle.cpp, C++ semantics) against the UNMODIFIED compiled reference
(oracle/_ref/libsmore_ref.so = /root/reference/src compiled -O2 with the Philox shim instead of src/random.cpp).

The reference has no tests or golden vectors of its own, so these comparisons -- plus the golden files
generated from the same library (tests/golden/) -- are what pins parity. Skipped where the compiled reference is absent.
"""
import os

import numpy as np
import pytest

from oracle import bindings as B
from tests import graphs

pytestmark = pytest.mark.ref

SEED = 20261018


def _mk(tmp_path, src, dst, w, undirected, kind, dim, order=2, field=None):
    path = os.path.join(tmp_path, "g.txt")
    B.write_edge_list(path, src, dst, w)
    ffile = None
    if field is not None:
        ffile = os.path.join(tmp_path, "f.txt")
        with open(ffile, "w") as f:
            for name, fl in field:
                f.write(f"{name} {fl}\n")
    ref = B.Ref(kind, path, undirected, dim, order=order, field_file=ffile)
    off, col, ww, ids = B.edges_to_csr(src, dst, w, undirected)
    return ref, (off, col, ww), ids


@pytest.mark.parametrize("undirected", [0, 1])
def test_ingest_csr_and_alias_tables(tmp_path, undirected):
    src, dst, w = graphs.random_graph(300, 4000, seed=3)
    ref, (off, col, ww), ids = _mk(str(tmp_path), src, dst, w, undirected, B.K_LINE, 4)
    roff, rcol, rw = ref.csr()
    assert np.array_equal(off, roff) and np.array_equal(col, rcol) and np.array_equal(ww, rw)
    assert ref.names() == [f"v{k}" for k in sorted(ids, key=ids.get)]
    assert ref.max_line == len(col)  # doubled when undirected (src/proNet.cpp:230-231)
    g = B.OracleGraph(B.SEM_CPP, off, col, ww)
    for a, b in zip(g.degrees(), ref.degrees()):
        assert np.array_equal(a, b)
    for which in (0, 1, 2):
        p0, a0 = g.alias(which)
        p1, a1 = ref.alias(which)
        assert np.array_equal(p0, p1), which  # bit-exact fp64
        assert np.array_equal(a0, a1), which
    assert np.array_equal(g.sigmoid_table(), ref.sigmoid_table())
    for x in (-9.0, -8.0, -7.999, -0.3, 0.0, 1e-9, 0.016, 3.3, 7.9999, 8.0, 8.5):
        assert g.fast_sigmoid(x) == ref.fast_sigmoid(x)


def test_alias_method_known_distributions():
    rng = np.random.RandomState(0)
    cases = [np.array([1.0]), np.array([1.0, 1.0, 1.0, 1.0]), np.array([0.0, 2.0, 0.0, 5.0, 1.0]),
             np.array([1e-3, 1e3, 1.0, 7.0]), rng.randint(1, 100, size=1000).astype(float),
             rng.pareto(1.5, size=5000) + 1e-3]
    for dist in cases:
        p0, a0 = B.alias_build(B.SEM_CPP, dist)
        p1, a1 = B.Ref.alias_method(dist)
        assert np.array_equal(p0, p1) and np.array_equal(a0, a1)
        # Go alias (alias.go:10-90) at power 0.75 is the same Vose construction; it normalises as (w^p * n) / sum
        # instead of w^p * (n / sum), so the tables agree only to rounding: compare the implied distributions.
        p2, a2 = B.alias_build(B.SEM_GO, dist, 0.75)
        n = len(dist)
        for pp, aa in ((p1, np.where(a1 == -1, np.arange(n), a1)), (p2, a2)):
            implied = pp / n
            np.add.at(implied, aa, (1.0 - pp) / n)
            want = dist ** 0.75 / (dist ** 0.75).sum()
            assert np.allclose(implied, want, rtol=0, atol=1e-12)


def test_neg_method_no_degrees(tmp_path):
    src, dst, w = graphs.bipartite_graph(50, 40, 600, seed=5)
    ref, (off, col, ww), _ = _mk(str(tmp_path), src, dst, w, 0, B.K_BPR, 4)
    g = B.OracleGraph(B.SEM_CPP, off, col, ww, neg_method=B.NEG_NO_DEGREES)
    for which in (0, 1, 2):
        assert all(np.array_equal(a, b) for a, b in zip(g.alias(which), ref.alias(which)))


def test_sampler_replay(tmp_path):
    src, dst, w = graphs.random_graph(500, 6000, seed=7)
    ref, (off, col, ww), _ = _mk(str(tmp_path), src, dst, w, 1, B.K_LINE, 4)
    g = B.OracleGraph(B.SEM_CPP, off, col, ww)
    n = 200000
    for which in (0, 1, 3):
        ref.seed(SEED, 5)
        r = ref.sample(which, n)
        o, pos = g.sample(which, SEED, 5, n)
        assert np.array_equal(r, o)
        assert pos == ref.pos()
    arg = np.random.RandomState(1).randint(0, g.V, size=n)
    ref.seed(SEED, 9)
    assert np.array_equal(ref.sample(2, n, arg), g.sample(2, SEED, 9, n, arg)[0])


def test_walks_and_pairs(tmp_path):
    # directed graph with sinks: exercises the dead-end teleport (src/proNet.cpp:712-718)
    src, dst, w = graphs.random_graph(120, 260, seed=11, zipf=False)
    ref, (off, col, ww), _ = _mk(str(tmp_path), src, dst, w, 0, B.K_DEEPWALK, 4)
    g = B.OracleGraph(B.SEM_CPP, off, col, ww)
    for start in range(g.V):
        for mode, w0, w1 in ((0, 5, 0), (0, 1, 0), (1, 2, 5), (1, 1, 3)):
            ref.seed(SEED, start)
            a = ref.walk_pairs(start, 40, mode, w0, w1)
            b = g.walk_pairs(SEED, start, start, 40, mode, w0, w1)
            for x, y in zip(a, b):
                assert np.array_equal(x, y)


@pytest.mark.parametrize("order", [1, 2])
def test_line_train_matches_reference_train(tmp_path, order):
    """Unmodified LINE::Train (1M samples, the CLI's minimum) vs the restated loop: same stream, same tables."""
    src, dst, w = graphs.random_graph(400, 5000, seed=13)
    dim = 8
    ref, (off, col, ww), _ = _mk(str(tmp_path), src, dst, w, 1, B.K_LINE, dim, order=order)
    g = B.OracleGraph(B.SEM_CPP, off, col, ww)
    Wv, Wc = graphs.init_tables(g.V, dim, seed=1, context_zero=True)
    ref.set_rows(0, Wv)
    if order == 2:
        ref.set_rows(1, Wc)
    ref.seed(SEED, 0)
    ref.train(1, 5, alpha=0.025, workers=1)
    a, c = Wv.copy(), Wc.copy()
    pos = g.train_line_cpp(a, a if order == 1 else c, 5, 0.025, 1000000, SEED, 0)
    assert pos == ref.pos()
    assert np.array_equal(a, ref.get_rows(0))  # bit-exact: same fp64 operations in the same order
    if order == 2:
        assert np.array_equal(c, ref.get_rows(1))


def test_bpr_and_warp_train(tmp_path):
    src, dst, w = graphs.bipartite_graph(200, 150, 4000, seed=17)
    dim = 8
    for kind in (B.K_BPR, B.K_WARP):
        ref, (off, col, ww), _ = _mk(str(tmp_path), src, dst, w, 0, kind, dim)
        g = B.OracleGraph(B.SEM_CPP, off, col, ww, neg_method=B.NEG_NO_DEGREES)
        W, _ = graphs.init_tables(g.V, dim, seed=2)
        ref.set_rows(0, W)
        ref.seed(SEED, 0)
        ref.train(1, 5, alpha=0.025, workers=1)
        a = W.copy()
        if kind == B.K_BPR:
            pos = g.train_bpr_cpp(a, 0.025, 1000000, SEED, 0)
        else:
            pos, tries = g.train_warp_cpp(a, 0.025, 1000000, SEED, 0)
            assert tries >= 1000000
        assert pos == ref.pos()
        assert np.array_equal(a, ref.get_rows(0))


def test_hoprec_train(tmp_path):
    nu, ni = 120, 80
    src, dst, w = graphs.bipartite_graph(nu, ni, 2500, seed=19)
    dim = 8
    labels = sorted(set(src.tolist()) | set(dst.tolist()))
    # users first in the field file so that field id 0 == user (src/proNet.cpp:368-372, HBPR.cpp:98)
    field = [(f"v{l}", "u" if l < nu else "i") for l in labels]
    ref, (off, col, ww), ids = _mk(str(tmp_path), src, dst, w, 1, B.K_HOPREC, dim, field=field)
    g = B.OracleGraph(B.SEM_CPP, off, col, ww, neg_method=B.NEG_NO_DEGREES)
    fl = np.zeros(g.V, dtype=np.int32)
    for l, vid in ids.items():
        fl[vid] = 0 if l < nu else 1
    assert np.array_equal(fl, ref.fields())
    g.set_field(fl)
    W, _ = graphs.init_tables(g.V, dim, seed=3)
    ref.set_rows(0, W)
    ref.seed(SEED, 0)
    ref.train(1, 3, alpha=0.025, workers=1)
    a = W.copy()
    pos = g.train_hoprec_cpp(a, 3, 0.025, 1000000, SEED, 0)
    assert pos == ref.pos()
    assert np.array_equal(a, ref.get_rows(0))


@pytest.mark.parametrize("walklets", [0, 1])
def test_deepwalk_walklets_train(tmp_path, walklets):
    src, dst, w = graphs.random_graph(300, 1500, seed=23)
    dim = 8
    kind = B.K_WALKLETS if walklets else B.K_DEEPWALK
    ref, (off, col, ww), _ = _mk(str(tmp_path), src, dst, w, 1, kind, dim)
    g = B.OracleGraph(B.SEM_CPP, off, col, ww)
    Wv, Wc = graphs.init_tables(g.V, dim, seed=4)
    ref.set_rows(0, Wv)
    ref.set_rows(1, Wc)
    ref.seed(SEED, 0)
    if walklets:
        ref.train(2, 20, 2, 4, 5, alpha=0.025, workers=1)
        w0, w1 = 2, 4
    else:
        ref.train(2, 20, 5, 5, alpha=0.025, workers=1)
        w0, w1 = 5, 0
    a, c = Wv.copy(), Wc.copy()
    pos, pairs = g.train_walk_cpp(walklets, a, c, 2, 20, w0, w1, 5, 0.025, SEED, 0)
    assert pairs > 0
    assert pos == ref.pos()
    assert np.array_equal(a, ref.get_rows(0))
    assert np.array_equal(c, ref.get_rows(1))


def test_update_pair_aliasing_cases(tmp_path):
    """Rows that alias (LINE-1 self pairs, BPR pos==neg) must follow the reference's in-place order."""
    src, dst, w = graphs.random_graph(30, 200, seed=29)
    dim = 6
    ref, (off, col, ww), _ = _mk(str(tmp_path), src, dst, w, 1, B.K_LINE, dim, order=1)
    g = B.OracleGraph(B.SEM_CPP, off, col, ww)
    W, _ = graphs.init_tables(g.V, dim, seed=5)
    W *= 20
    for v, c in ((3, 3), (3, 4), (0, 1)):
        ref.set_rows(0, W)
        ref.seed(SEED, 2)
        ref.update_pair(v, c, 5, 0.05)
        a = W.copy()
        g.update_pair_cpp(a, a, v, c, 5, 0.05, SEED, 2)
        assert np.array_equal(a, ref.get_rows(0))
    refb, _, _ = _mk(str(tmp_path), src, dst, w, 1, B.K_BPR, dim)
    g = B.OracleGraph(B.SEM_CPP, off, col, ww, neg_method=B.NEG_NO_DEGREES)  # BPR ctor: "no_degrees" (BPR.cpp:4-7)
    for v, ci, cj in ((1, 2, 2), (1, 1, 2), (1, 2, 1), (4, 5, 6)):
        for fn_ref, fn_orc in ((refb.update_bpr_pair, g.update_bpr_pair_cpp), (refb.update_warp_pair, g.update_warp_pair_cpp)):
            refb.set_rows(0, W)
            refb.seed(SEED, 3)
            fn_ref(v, ci, cj, 0.05)
            a = W.copy()
            fn_orc(a, v, ci, cj, 0.05, SEED, 3)
            assert np.array_equal(a, refb.get_rows(0))


def test_hpe_train_and_update_community(tmp_path):
    """HPE: HPE::Train (src/model/HPE.cpp:93-147) = UpdateCommunity (src/proNet.cpp:3018-3054,
    Opt_SigmoidRegSGD :1332-1351) + UpdatePair with the roles swapped, against the compiled reference."""
    src, dst, w = graphs.random_graph(250, 3000, seed=23)
    dim = 12
    ref, (off, col, ww), _ = _mk(str(tmp_path), src, dst, w, 1, B.K_HPE, dim)
    g = B.OracleGraph(B.SEM_CPP, off, col, ww)
    Wv, Wc = graphs.init_tables(g.V, dim, seed=4)
    Wc = (np.random.RandomState(5).random_sample(Wc.shape) - 0.5) / dim  # HPE::Init: both tables random (HPE.cpp:38-56)
    # one UpdateCommunity call, incl. a sink-free 4-step walk and repeated negatives on a tiny K
    ref.set_rows(0, Wv), ref.set_rows(1, Wc)
    a, c = Wv.copy(), Wc.copy()
    ref.seed(SEED, 3)
    ref.update_community(7, int(col[off[7]]), 0.01, 4, 5, 0.025)
    pos = g.update_community_cpp(a, c, 7, int(col[off[7]]), 0.01, 4, 5, 0.025, SEED, 3)
    assert pos == ref.pos()
    assert np.array_equal(a, ref.get_rows(0)) and np.array_equal(c, ref.get_rows(1))
    # the whole Train() loop: 1M samples, walk_steps 3, K 5, reg 0.01
    ref.set_rows(0, Wv), ref.set_rows(1, Wc)
    a, c = Wv.copy(), Wc.copy()
    ref.seed(SEED, 0)
    ref.train_hpe(1, 3, 5, 0.01, alpha=0.025, workers=1)
    pos = g.train_hpe_cpp(a, c, 3, 5, 0.01, 0.025, 1000000, SEED, 0)
    assert pos == ref.pos()
    assert np.array_equal(a, ref.get_rows(0)) and np.array_equal(c, ref.get_rows(1))


def test_mf_train_and_update_factorized_pair(tmp_path):
    """MF: MF::Train (src/model/MF.cpp:50-98) = UpdateFactorizedPair (src/proNet.cpp:2591-2614,
    Opt_SGD :991-1012) on ONE table in both roles, negatives "no_degrees" (MF.cpp:4-7), against the compiled reference."""
    src, dst, w = graphs.bipartite_graph(160, 120, 3000, seed=29)
    dim = 12
    ref, (off, col, ww), _ = _mk(str(tmp_path), src, dst, w, 0, B.K_MF, dim)
    g = B.OracleGraph(B.SEM_CPP, off, col, ww, neg_method=B.NEG_NO_DEGREES)
    W, _ = graphs.init_tables(g.V, dim, seed=6)
    # single calls incl. the aliasing cases of a shared table: context == vertex, and an ordinary pair
    for v, c in ((5, 5), (5, int(col[off[5]]))):
        ref.set_rows(0, W)
        a = W.copy()
        ref.seed(SEED, 7)
        ref.update_factorized_pair(v, c, 0.01, 5, 0.025)
        pos = g.update_factorized_pair_cpp(a, v, c, 0.01, 5, 0.025, SEED, 7)
        assert pos == ref.pos()
        assert np.array_equal(a, ref.get_rows(0))
    ref.set_rows(0, W)
    a = W.copy()
    ref.seed(SEED, 0)
    ref.train_mf(1, 5, 0.01, alpha=0.025, workers=1)
    pos = g.train_mf_cpp(a, 5, 0.01, 0.025, 1000000, SEED, 0)
    assert pos == ref.pos()
    assert np.array_equal(a, ref.get_rows(0))


def test_skewopt_train_and_update_sbpr_pair(tmp_path):
    """Skew-OPT (oracle groundwork, the kernel is next): SPR::Train (src/model/SkewOPT.cpp) =
    UpdateSBPRPair (src/proNet.cpp:1517-1566: 16 margin-gated rounds, every negative drawn inside) with Opt_SBPRSGD
    (:1070-1098: (f - xi)/omega, gate at 2, clamp at -2, eta-th power chain), against the compiled reference."""
    src, dst, w = graphs.bipartite_graph(160, 120, 3000, seed=37)
    dim = 12
    ref, (off, col, ww), _ = _mk(str(tmp_path), src, dst, w, 0, B.K_SKEWOPT, dim)
    g = B.OracleGraph(B.SEM_CPP, off, col, ww, neg_method=B.NEG_NO_DEGREES)
    W, _ = graphs.init_tables(g.V, dim, seed=9)
    W = W * 40.0 + 0.01  # scores around xi/omega-scale so that rounds are gated, clamped and accepted (Init adds 0.01, SkewOPT.cpp)
    for xi, omega, eta in ((10.0, 3.0, 3), (0.5, 0.7, 2)):
        ref.set_rows(0, W)
        a = W.copy()
        ref.seed(SEED, 5)
        ref.update_sbpr_pair(3, int(col[off[3]]), xi, omega, eta, 0.025)
        pos = g.update_sbpr_pair_cpp(a, 3, int(col[off[3]]), xi, omega, eta, 0.025, SEED, 5)
        assert pos == ref.pos() == 32
        assert np.array_equal(a, ref.get_rows(0))
    for xi, omega, eta in ((10.0, 3.0, 3), (0.5, 0.7, 2)):
        ref.set_rows(0, W)
        a = W.copy()
        ref.seed(SEED, 0)
        ref.train_skewopt(1, xi, omega, eta, alpha=0.025, workers=1)
        pos = g.train_skewopt_cpp(a, xi, omega, eta, 0.025, 1000000, SEED, 0)
        assert pos == ref.pos()
        assert np.array_equal(a, ref.get_rows(0))
