"""GPU: Skew-OPT -- k_skewopt (UpdateSBPRPair: 16 margin-gated rounds on one shared table) against
the golden vectors generated from the compiled reference (tests/golden/golden_skewopt_v1.npz)."""
import os
import subprocess

import numpy as np
import pytest

from oracle import bindings as B
from smore_b200 import capi
from tests import graphs

pytestmark = pytest.mark.gpu
SEED = 20261018
GS = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_skewopt_v1.npz"))
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _params(total, xi, omega, eta, mode=capi.MODE_DETERMINISTIC):
    p = capi.default_params()
    p.semantics, p.mode, p.seed, p.stream_base, p.alpha = capi.SEM_CPP, mode, SEED, 0, 0.025
    p.total, p.xi, p.omega, p.eta = total, xi, omega, eta
    return p


def _graph():
    off, col, ww, _ = B.edges_to_csr(GS["bip_src"], GS["bip_dst"], GS["bip_w"], 0)
    return capi.Graph.from_csr(off, col, ww, semantics=capi.SEM_CPP, negative_method=capi.NEG_NO_DEGREES)


def test_skewopt_matches_the_compiled_reference():
    # ("b" -- xi 0.5, omega 0.7, eta 2 -- is checked against the oracle on the CPU; one warp needs ~25 s per 1 M samples)
    xi, omega, eta = GS["a_args"]
    m = capi.Model(_graph(), 8, 1, capi.F64)
    m.set_rows(0, GS["a_init"])
    st = m.train_skewopt(_params(1_000_000, float(xi), float(omega), int(eta)))
    assert st["words_stream0"] == int(GS["a_words"])
    assert np.array_equal(m.get_rows(0), GS["a_v"])


def test_skewopt_short_run_matches_the_oracle_with_other_parameters():
    off, col, ww, _ = B.edges_to_csr(GS["bip_src"], GS["bip_dst"], GS["bip_w"], 0)
    og = B.OracleGraph(B.SEM_CPP, off, col, ww, neg_method=B.NEG_NO_DEGREES)
    a = GS["b_init"].copy()
    pos = og.train_skewopt_cpp(a, 0.5, 0.7, 2, 0.025, 60000, SEED, 0)
    m = capi.Model(_graph(), 8, 1, capi.F64)
    m.set_rows(0, GS["b_init"])
    st = m.train_skewopt(_params(60000, 0.5, 0.7, 2))
    assert st["words_stream0"] == pos
    assert np.array_equal(m.get_rows(0), a)


def test_skewopt_hogwild_fp32_and_cli(tmp_path):
    src, dst, w = graphs.bipartite_graph(400, 300, 9000, seed=43)
    off, col, ww, _ = B.edges_to_csr(src, dst, w, 0)
    g = capi.Graph.from_csr(off, col, ww, semantics=capi.SEM_CPP, negative_method=capi.NEG_NO_DEGREES)
    m = capi.Model(g, 128, 1, capi.F32)
    m.init(0, True, 1)
    st = m.train_skewopt(_params(1_000_000, 10.0, 3.0, 3, mode=capi.MODE_HOGWILD))
    assert st["samples"] > 900_000 and st["pair_updates"] == 16 * st["samples"]
    assert np.isfinite(m.get_rows(0)).all()
    p = _params(1000, 10.0, 3.0, 3)
    p.eta = 0
    with pytest.raises(capi.SmoreError):
        m.train_skewopt(p)
    net, rep = str(tmp_path / "net.txt"), str(tmp_path / "rep.txt")
    B.write_edge_list(net, src, dst, w)
    out = subprocess.run([os.path.join(ROOT, "smore_b200", "bin", "skewopt"), "-train", net, "-save", rep, "-dimensions", "16",
                          "-sample_times", "1"], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    lines = open(rep).read().split("\n")
    n, dim = map(int, lines[0].split())
    assert dim == 16 and len(lines) == n + 2
