"""GPU: warm start (-load_v / -load_c: proNet::LoadPreTrain, src/proNet.cpp:238-286) and the binary checkpoint."""
import os
import subprocess

import numpy as np
import pytest

from oracle import bindings as B
from smore_b200 import capi
from tests import graphs

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "smore_b200", "bin")


def _graph(tmp_path, n=300, e=4000, seed=81):
    src, dst, w = graphs.random_graph(n, e, seed=seed)
    net = str(tmp_path / "net.txt")
    B.write_edge_list(net, src, dst, w)
    return net, capi.Graph.from_edge_list(net, True)


def _read_text(path):
    lines = open(path).read().split("\n")
    n, dim = map(int, lines[0].split())
    names, rows = [], []
    for ln in lines[1:1 + n]:
        t = ln.split(" ")
        names.append(t[0])
        rows.append([float(x) for x in t[1:]])
    return names, np.array(rows).reshape(n, dim)


@pytest.mark.parametrize("fmt", [0, 1])
def test_pretrain_roundtrip(tmp_path, fmt):
    net, g = _graph(tmp_path)
    dim = 24
    a = capi.Model(g, dim, 2, capi.F64)
    a.init(0, True, 5), a.init(1, True, 6)
    rep = str(tmp_path / "rep.txt")
    a.save_weights(rep, table=0, fmt=fmt)
    names, W = _read_text(rep)
    assert names == g.names()
    # shuffle the lines, add an unknown vertex and a short line: order must not matter, strangers are skipped
    lines = open(rep).read().split("\n")
    body = lines[1:1 + g.V]
    np.random.default_rng(1).shuffle(body)
    body.insert(7, "stranger" + " 0.5" * dim)
    body.insert(11, names[3] + " 1 2 3")
    shuffled = str(tmp_path / "shuffled.txt")
    open(shuffled, "w").write("\n".join([lines[0]] + body) + "\n")
    b = capi.Model(g, dim, 2, capi.F64)
    b.init(0, False), b.init(1, False)
    assert b.load_pretrain(shuffled, table=1) == g.V
    assert np.array_equal(b.get_rows(1), W)          # exactly the numbers in the file
    assert not b.get_rows(0).any()                   # the other table is untouched
    assert b.load_pretrain(rep, table=0) == g.V and np.array_equal(b.get_rows(0), W)
    # dimension mismatch: the reference skips the file
    c = capi.Model(g, dim + 1, 1, capi.F32)
    c.init(0, False)
    assert c.load_pretrain(rep) == 0 and not c.get_rows(0).any()
    with pytest.raises(capi.SmoreError):
        b.load_pretrain(str(tmp_path / "missing.txt"))


def test_pretrain_on_a_sharded_model(tmp_path):
    net, g = _graph(tmp_path)
    dim = 8
    a = capi.Model(g, dim, 1, capi.F64)
    a.init(0, True, 9)
    rep = str(tmp_path / "rep.txt")
    a.save_weights(rep, fmt=0)
    _, W = _read_text(rep)
    world = 4
    for r in range(world):
        gr = capi.Graph.from_edge_list(net, True)
        gr.set_shard(r, world)
        m = capi.Model(gr, dim, 1, capi.F64)
        m.init(0, False)
        rows = np.arange(r, gr.V, world)
        assert m.load_pretrain(rep) == len(rows)
        assert np.array_equal(m.get_rows(0), W[rows])


@pytest.mark.parametrize("dtype", [capi.F32, capi.F64])
def test_checkpoint_roundtrip_and_resume(tmp_path, dtype):
    net, g = _graph(tmp_path)
    dim = 32
    p = capi.default_params()
    p.mode, p.seed, p.total = capi.MODE_DETERMINISTIC, 3, 30_000
    p.sched_total = 60_000
    # one run of 60 k samples in two chunks ...
    a = capi.Model(g, dim, 2, dtype)
    a.init(0, True, 1), a.init(1, False, 1)
    a.train_line(p)
    assert a.progress() == {"seed": 3, "next_stream": 1, "sched_total": 60_000, "sched_done": 30_000}
    ck = str(tmp_path / "half.ckpt")
    a.save_checkpoint(ck)
    p2 = capi.default_params()
    p2.mode, p2.seed, p2.total, p2.sched_total, p2.sched_offset, p2.stream_base = capi.MODE_DETERMINISTIC, 3, 30_000, 60_000, 30_000, 1
    a.train_line(p2)
    assert a.progress()["sched_done"] == 60_000 and a.progress()["next_stream"] == 2
    # ... equals: first chunk, checkpoint, NEW model, resume from the position the checkpoint carries, second chunk
    b = capi.Model(g, dim, 2, dtype)
    assert b.progress()["sched_total"] == 0
    b.load_checkpoint(ck)
    pos = b.progress()
    assert pos == {"seed": 3, "next_stream": 1, "sched_total": 60_000, "sched_done": 30_000}
    p3 = capi.default_params()
    p3.mode, p3.seed, p3.total = capi.MODE_DETERMINISTIC, pos["seed"], pos["sched_total"] - pos["sched_done"]
    p3.sched_total, p3.sched_offset, p3.stream_base = pos["sched_total"], pos["sched_done"], pos["next_stream"]
    b.train_line(p3)
    assert np.array_equal(a.get_rows(0), b.get_rows(0)) and np.array_equal(a.get_rows(1), b.get_rows(1))
    # a checkpoint only fits the model it was taken from
    with pytest.raises(capi.SmoreError):
        capi.Model(g, dim + 8, 2, dtype).load_checkpoint(ck)
    with pytest.raises(capi.SmoreError):
        capi.Model(g, dim, 1, dtype).load_checkpoint(ck)
    open(str(tmp_path / "junk.ckpt"), "wb").write(b"not a checkpoint")
    with pytest.raises(capi.SmoreError):
        b.load_checkpoint(str(tmp_path / "junk.ckpt"))
    trunc = str(tmp_path / "trunc.ckpt")
    open(trunc, "wb").write(open(ck, "rb").read()[:-100])
    with pytest.raises(capi.SmoreError):
        b.load_checkpoint(trunc)


def test_cli_load_v_and_checkpoint_flags(tmp_path):
    net, g = _graph(tmp_path, n=120, e=1500, seed=83)
    rep1, rep2, ck = str(tmp_path / "r1.txt"), str(tmp_path / "r2.txt"), str(tmp_path / "m.ckpt")
    base = [os.path.join(BIN, "deepwalk"), "-train", net, "-dimensions", "16", "-walk_times", "1", "-walk_steps", "5",
            "-semantics", "cpp", "-mode", "deterministic", "-dtype", "f64"]
    r = subprocess.run(base + ["-save", rep1, "-checkpoint", ck], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert os.path.getsize(ck) > 2 * g.V * 16 * 8
    r = subprocess.run(base + ["-save", rep2, "-load_v", rep1, "-load_c", rep1], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert r.stdout.count(f"# of Pre-train:\t\t{g.V}") == 2
    r = subprocess.run(base + ["-save", rep2, "-resume", ck], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run(base + ["-save", rep2, "-resume", rep1], capture_output=True, text=True)
    assert r.returncode != 0 and "not a smore_b200 checkpoint" in r.stderr
    # an interrupted Hogwild run continues where its checkpoint left it: schedule position and unused sampler streams
    line = [os.path.join(BIN, "line"), "-train", net, "-dimensions", "16", "-sample_times", "2", "-semantics", "cpp"]
    r = subprocess.run(line + ["-save", rep1, "-checkpoint", ck, "-max_chunks", "8"], capture_output=True, text=True)
    assert r.returncode == 0 and "stopped by -max_chunks" in r.stdout, r.stdout + r.stderr
    r = subprocess.run(line + ["-save", rep2, "-resume", ck], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "resume:\t\t\t800000 of 2000000 samples done" in r.stdout and "Progress: 100.00 %" in r.stdout, r.stdout
    done = [l for l in r.stdout.split("\n") if "samples," in l][0]
    assert 1_100_000 < int(done.split()[0]) <= 1_200_000, done  # only the remaining 12 of 20 chunks ran
