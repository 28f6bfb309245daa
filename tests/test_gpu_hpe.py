"""GPU: HPE -- k_hpe against the golden vectors generated from the compiled reference
(tests/golden/golden_hpe_v1.npz) and against the oracle on cases built to repeat context rows inside a step."""
import os
import subprocess

import numpy as np
import pytest

from oracle import bindings as B
from smore_b200 import capi
from tests import graphs

pytestmark = pytest.mark.gpu
SEED = 20261018
GH = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_hpe_v1.npz"))
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _params(total, steps, K, reg, mode=capi.MODE_DETERMINISTIC, seed=SEED):
    p = capi.default_params()
    p.semantics, p.mode, p.seed, p.stream_base, p.alpha = capi.SEM_CPP, mode, seed, 0, 0.025
    p.total, p.walk_steps, p.negative_samples, p.lambda_ = total, steps, K, reg
    return p


@pytest.mark.parametrize("tag", ["a"])  # ("b" is checked against the oracle on the CPU: one warp takes ~35 s per 1 M samples)
def test_hpe_matches_the_compiled_reference(tag):
    off, col, ww, _ = B.edges_to_csr(GH["g300_src"], GH["g300_dst"], GH["g300_w"], 1)
    g = capi.Graph.from_csr(off, col, ww, semantics=capi.SEM_CPP)
    steps, K, reg = GH[f"hpe_{tag}_args"]
    m = capi.Model(g, 8, 2, capi.F64)
    m.set_rows(0, GH["init_v"]), m.set_rows(1, GH["init_c"])
    st = m.train_hpe(_params(1_000_000, int(steps), int(K), float(reg)))
    assert st["words_stream0"] == int(GH[f"hpe_{tag}_words"])
    assert np.array_equal(m.get_rows(0), GH[f"hpe_{tag}_v"]) and np.array_equal(m.get_rows(1), GH[f"hpe_{tag}_c"])


def test_hpe_repeated_contexts_and_sinks_match_the_oracle():
    """12 vertices and K = 20: almost every step repeats a context row (ORDERED path); directed edges into sinks end
    walks early. dim 20 exercises the masked row layout."""
    rng = np.random.RandomState(5)
    src = rng.randint(0, 8, 60)
    dst = rng.randint(0, 12, 60)          # vertices 8..11 only ever appear as targets: sinks
    w = rng.randint(1, 4, 60).astype(np.float64)
    off, col, ww, _ = B.edges_to_csr(src, dst, w, 0)
    og = B.OracleGraph(B.SEM_CPP, off, col, ww)
    dg = capi.Graph.from_csr(off, col, ww, semantics=capi.SEM_CPP)
    dim = 20
    Wv, Wc = graphs.init_tables(og.V, dim, seed=8)
    a, c = Wv.copy(), Wc.copy()
    pos = og.train_hpe_cpp(a, c, 6, 20, 0.02, 0.025, 30000, SEED, 0)
    m = capi.Model(dg, dim, 2, capi.F64)
    m.set_rows(0, Wv), m.set_rows(1, Wc)
    st = m.train_hpe(_params(30000, 6, 20, 0.02))
    assert st["words_stream0"] == pos
    assert np.array_equal(m.get_rows(0), a) and np.array_equal(m.get_rows(1), c)


def test_hpe_hogwild_fp32_and_errors():
    off, col, ww, _ = B.edges_to_csr(GH["g300_src"], GH["g300_dst"], GH["g300_w"], 1)
    g = capi.Graph.from_csr(off, col, ww, semantics=capi.SEM_CPP)
    m = capi.Model(g, 128, 2, capi.F32)
    m.init(0, True, 1), m.init(1, True, 2)
    st = m.train_hpe(_params(2_000_000, 5, 5, 0.01, mode=capi.MODE_HOGWILD))
    assert st["samples"] > 1_900_000 and st["pair_updates"] >= 5 * st["samples"]
    assert np.isfinite(m.get_rows(0)).all() and np.abs(m.get_rows(0)).max() < 50
    one = capi.Model(g, 16, 1, capi.F32)
    with pytest.raises(capi.SmoreError):
        one.train_hpe(_params(1000, 5, 5, 0.01))       # needs both tables
    gg = capi.Graph.from_csr(off, col, ww, semantics=capi.SEM_GO)
    mg = capi.Model(gg, 16, 2, capi.F32)
    p = _params(1000, 5, 5, 0.01)
    p.semantics = capi.SEM_GO
    with pytest.raises(capi.SmoreError):
        mg.train_hpe(p)                                 # the Go hpe model is LINE-2


def test_hpe_cli(tmp_path):
    src, dst, w = graphs.random_graph(150, 2000, seed=31)
    net, rep = str(tmp_path / "net.txt"), str(tmp_path / "rep.txt")
    B.write_edge_list(net, src, dst, w)
    r = subprocess.run([os.path.join(ROOT, "smore_b200", "bin", "hpe"), "-train", net, "-save", rep, "-dimensions", "16",
                        "-sample_times", "1", "-walk_steps", "3", "-reg", "0.01", "-threads", "2"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lines = open(rep).read().split("\n")
    n, dim = map(int, lines[0].split())
    assert dim == 16 and len(lines) == n + 2 and len(lines[1].split(" ")) == 17
