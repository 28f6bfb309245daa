"""CPU: the multi-threaded text ingest (smore_edge_list_to_csr, host-only) against a line-by-line restatement of the
reference's rules (src/proNet.cpp:158-224; pronet.go:128-165): ids in first-appearance order (source before target),
lines with fewer than three fields or an unparsable weight skipped, duplicates kept, adjacency in file order with the
reverse entry right behind the forward one when undirected. The result must not depend on the thread count."""
import os
import subprocess
import sys

import numpy as np
import pytest

from smore_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _write(path, n_vertices, n_lines, seed):
    rng = np.random.default_rng(seed)
    lines = []
    for i in range(n_lines):
        a, b = rng.integers(0, n_vertices, 2)
        w = rng.choice(["1", "2.5", "3", "+4", "1e-2", "0.125", "7"])
        sep = rng.choice([" ", "\t", "  ", " \t "])
        lines.append(f"v{a}{sep}n{b}{sep}{w}" + ("\r" if i % 97 == 0 else ""))
        if i % 53 == 0:
            lines.append("")                      # empty line
        if i % 71 == 0:
            lines.append(f"v{a} n{b}")            # two fields: skipped
        if i % 89 == 0:
            lines.append(f"v{a} n{b} abc")        # bad weight: skipped
        if i % 101 == 0:
            lines.append(f"x{a} x{a} 2 trailing") # self loop, extra field ignored
    with open(path, "w") as f:
        f.write("\n".join(lines))                  # no newline at the end of the file


def _restate(path, undirected):
    ids, names, edges = {}, [], []
    for line in open(path, "rb").read().split(b"\n"):
        tok = line.replace(b"\r", b" ").split()
        if len(tok) < 3:
            continue
        try:
            w = float(tok[2])
        except ValueError:
            continue
        for t in tok[:2]:
            if t not in ids:
                ids[t] = len(names)
                names.append(t.decode())
        edges.append((ids[tok[0]], ids[tok[1]], w))
    adj = [[] for _ in names]
    for a, b, w in edges:
        adj[a].append((b, w))
        if undirected:
            adj[b].append((a, w))
    off = np.zeros(len(names) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(x) for x in adj])
    col = np.array([b for x in adj for b, _ in x], dtype=np.int32)
    w = np.array([w for x in adj for _, w in x])
    return off, col, w, names, len(edges)


@pytest.mark.parametrize("undirected", [False, True])
def test_ingest_matches_restatement(tmp_path, undirected):
    path = str(tmp_path / "g.txt")
    _write(path, 3000, 200_000, seed=5)  # ~3.4 MB: several chunks
    exp = _restate(path, undirected)
    got = capi.edge_list_to_csr(path, undirected)
    assert got[4] == exp[4] and got[3] == exp[3]
    assert np.array_equal(got[0], exp[0]) and np.array_equal(got[1], exp[1]) and np.array_equal(got[2], exp[2])


def test_ingest_independent_of_thread_count(tmp_path):
    path = str(tmp_path / "g.txt")
    _write(path, 5000, 150_000, seed=6)
    code = ("import sys, hashlib; sys.path.insert(0, %r); from smore_b200 import capi; r = capi.edge_list_to_csr(%r, True); "
            "h = hashlib.sha1(); [h.update(x.tobytes()) for x in r[:3]]; h.update('\\n'.join(r[3]).encode()); print(h.hexdigest(), r[4])"
            % (ROOT, path))
    outs = set()
    for t in ("1", "2", "7"):
        env = dict(os.environ, SMORE_HOST_THREADS=t)
        outs.add(subprocess.run([sys.executable, "-c", code], env=env, check=True, capture_output=True, text=True).stdout.strip())
    assert len(outs) == 1, outs


def test_ingest_errors_and_empty(tmp_path):
    with pytest.raises(capi.SmoreError):
        capi.edge_list_to_csr(str(tmp_path / "missing.txt"), False)
    p = tmp_path / "empty.txt"
    p.write_text("")
    off, col, w, names, n = capi.edge_list_to_csr(str(p), True)
    assert len(off) == 1 and len(col) == 0 and names == [] and n == 0
