"""GPU: MF -- k_line<..., KIND = 1> (UpdateFactorizedPair, one table in both roles) against the
golden vectors generated from the compiled reference (tests/golden/golden_mf_v1.npz)."""
import os
import subprocess

import numpy as np
import pytest

from oracle import bindings as B
from smore_b200 import capi
from tests import graphs

pytestmark = pytest.mark.gpu
SEED = 20261018
GM = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_mf_v1.npz"))
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _params(total, K, reg, mode=capi.MODE_DETERMINISTIC):
    p = capi.default_params()
    p.semantics, p.mode, p.seed, p.stream_base, p.alpha = capi.SEM_CPP, mode, SEED, 0, 0.025
    p.total, p.negative_samples, p.lambda_ = total, K, reg
    return p


@pytest.mark.parametrize("tag", ["bip", "small"])
def test_mf_matches_the_compiled_reference(tag):
    """"small" is a 40-vertex general graph: negatives repeat and hit the vertex row itself (ORDERED path)."""
    off, col, ww, _ = B.edges_to_csr(GM[f"{tag}_src"], GM[f"{tag}_dst"], GM[f"{tag}_w"], 0)
    g = capi.Graph.from_csr(off, col, ww, semantics=capi.SEM_CPP, negative_method=capi.NEG_NO_DEGREES)
    K, reg = GM[f"{tag}_args"]
    m = capi.Model(g, 8, 1, capi.F64)
    m.set_rows(0, GM[f"{tag}_init"])
    st = m.train_mf(_params(1_000_000, int(K), float(reg)))
    assert st["words_stream0"] == int(GM[f"{tag}_words"])
    assert np.array_equal(m.get_rows(0), GM[f"{tag}_v"])


def test_mf_hogwild_fp32_and_cli(tmp_path):
    src, dst, w = graphs.bipartite_graph(400, 300, 9000, seed=41)
    off, col, ww, _ = B.edges_to_csr(src, dst, w, 0)
    g = capi.Graph.from_csr(off, col, ww, semantics=capi.SEM_CPP, negative_method=capi.NEG_NO_DEGREES)
    m = capi.Model(g, 128, 1, capi.F32)
    m.init(0, True, 1)
    st = m.train_mf(_params(2_000_000, 5, 0.01, mode=capi.MODE_HOGWILD))
    assert st["samples"] > 1_900_000
    W = m.get_rows(0)
    assert np.isfinite(W).all() and np.abs(W).max() < 50
    # observed pairs score higher than random ones after training (label +1 vs -1)
    nu = 400
    s = np.einsum("ij,ij->i", W[np.repeat(np.arange(len(off) - 1), np.diff(off))], W[col])
    rng = np.random.default_rng(1)
    items = np.unique(col)
    r = np.einsum("ij,ij->i", W[rng.integers(0, len(off) - 1, 20000)], W[rng.choice(items, 20000)])
    assert s.mean() > r.mean() + 0.05
    with pytest.raises(capi.SmoreError):
        p = _params(1000, 5, 0.01)
        p.semantics = capi.SEM_GO
        m.train_mf(p)
    net, rep = str(tmp_path / "net.txt"), str(tmp_path / "rep.txt")
    B.write_edge_list(net, src, dst, w)
    out = subprocess.run([os.path.join(ROOT, "smore_b200", "bin", "mf"), "-train", net, "-save", rep, "-dimensions", "16",
                          "-sample_times", "1", "-reg", "0.01"], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    lines = open(rep).read().split("\n")
    n, dim = map(int, lines[0].split())
    assert dim == 16 and len(lines) == n + 2
