"""GPU: the C++ host CLI (smore_b200/bin/*, mirror of the reference CLIs) end to end: text ingest -> Init -> Train ->
SaveWeights, compared with the oracle driven by the same Philox-defined init and draw stream, in the reference's own
output formats (C++ iostream %g, src/model/LINE.cpp:35-38; Go %.6f, internal/models/line/line.go:219-228)."""
import os
import subprocess

import numpy as np
import pytest

from oracle import bindings as B
from tests import graphs

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "smore_b200", "bin")
SEED = 77


def philox_init(V, dim, seed, table):
    words = B.stream_words(seed, (1 << 62) + table, 0, V * dim).astype(np.float64)
    return ((words / 4294967296.0 - 0.5) / dim).reshape(V, dim)


def expected_file(names, W, fmt):
    lines = [f"{W.shape[0]} {W.shape[1]}"]
    for n, row in zip(names, W):
        lines.append(n + "".join(fmt % x for x in row))
    return "\n".join(lines) + "\n"


@pytest.mark.parametrize("sem", ["cpp", "go"])
def test_line_cli_matches_oracle(tmp_path, sem):
    src, dst, w = graphs.random_graph(200, 2500, seed=51)
    net, out = str(tmp_path / "net.txt"), str(tmp_path / "rep.txt")
    B.write_edge_list(net, src, dst, w)
    dim = 16
    # sample_times=1: cpp -> 1e6 samples, go -> 1 * MaxLine samples
    cmd = [os.path.join(BIN, "line"), "-train", net, "-save", out, "-dimensions", str(dim), "-sample_times", "1",
           "-negative_samples", "5", "-alpha", "0.025", "-threads", "4", "-semantics", sem, "-mode", "deterministic",
           "-dtype", "f64", "-seed", str(SEED)] + (["-undirected", "1"] if sem == "cpp" else ["-undirected=true"])
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    off, col, ww, ids = B.edges_to_csr(src, dst, w, True)
    names = [f"v{k}" for k in sorted(ids, key=ids.get)]
    V = len(names)
    Wv = philox_init(V, dim, SEED, 0)
    if sem == "cpp":
        g = B.OracleGraph(B.SEM_CPP, off, col, ww)
        Wc = np.zeros((V, dim))
        g.train_line_cpp(Wv, Wc, 5, 0.025, 1_000_000, SEED, 0)
        want = expected_file(names, Wv, " %g")
    else:
        g = B.OracleGraph(B.SEM_GO, off, col, ww, max_line=len(src))
        Wc = philox_init(V, dim, SEED, 1)
        g.train_line_go(Wv, Wc, 2, 5, 0.025, len(src), SEED, 0)
        want = expected_file(names, Wv, " %.6f")
    assert open(out).read() == want


def test_bpr_and_hoprec_cli_run(tmp_path):
    nu, ni = 150, 90
    src, dst, w = graphs.bipartite_graph(nu, ni, 2500, seed=53)
    net, out, fld = str(tmp_path / "net.txt"), str(tmp_path / "rep.txt"), str(tmp_path / "field.txt")
    B.write_edge_list(net, src, dst, w)
    with open(fld, "w") as f:
        for l in sorted(set(src.tolist()) | set(dst.tolist())):
            f.write(f"v{l} {'u' if l < nu else 'i'}\n")
    for cmd in ([os.path.join(BIN, "bpr"), "-train", net, "-save", out, "-dimensions", "64", "-sample_times", "2"],
                [os.path.join(BIN, "smore"), "hoprec", "-train", net, "-field", fld, "-save", out, "-dimensions", "32",
                 "-sample_times", "1", "-walk_steps", "3"],
                [os.path.join(BIN, "warp"), "-train", net, "-save", out, "-sample_times", "1"]):
        res = subprocess.run(cmd, capture_output=True, text=True)
        assert res.returncode == 0, res.stdout + res.stderr
        head = open(out).readline().split()
        assert int(head[0]) == len(set(src.tolist()) | set(dst.tolist()))
        body = np.loadtxt(out, skiprows=1, usecols=range(1, int(head[1]) + 1))
        assert np.isfinite(body).all()


def test_cli_errors(tmp_path):
    res = subprocess.run([os.path.join(BIN, "line"), "-train", str(tmp_path / "missing.txt"), "-save", "x"],
                         capture_output=True, text=True)
    assert res.returncode != 0 and "cannot open" in res.stderr
