"""GPU: smore_progress -- the one entry point that may be called from another thread while a train call blocks
(it replaces the "Alpha / Progress" line the reference's workers print every MONITOR samples,
src/model/LINE.cpp:179-187, line.go:133-142)."""
import os
import subprocess
import threading
import time

import numpy as np
import pytest

from smore_b200 import capi, synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "smore_b200", "bin")


def _model(nv=200_000, dim=128):
    src, dst, w = synth.power_law_edges(nv, 10 * nv, 5)
    off, col, ww, _ = synth.csr_from_edges(src, dst, w, True)
    g = capi.Graph.from_csr(off, col, ww)
    m = capi.Model(g, dim, 2, capi.F32)
    m.init(0, True, 1), m.init(1, False, 2)
    return g, m


def test_progress_before_any_training_reads_zero():
    g, m = _model(2000, 16)
    assert m.live_progress() == {"done": 0, "total": 0, "alpha": 0.0, "running": False}


def test_progress_is_polled_while_train_blocks():
    g, m = _model()
    p = capi.default_params()
    p.mode, p.seed, p.total, p.alpha = capi.MODE_HOGWILD, 3, 400_000_000, 0.025  # ~0.5 s of device time
    seen, result = [], {}

    def train():
        result["stats"] = m.train_line(p)

    t = threading.Thread(target=train)
    t.start()
    while t.is_alive():
        seen.append(m.live_progress())  # (ctypes drops the GIL inside the blocking train call)
        time.sleep(0.01)
    t.join()
    running = [s for s in seen if s["running"]]
    assert len(running) >= 5, "the poll never saw the call in flight"
    assert all(s["total"] == p.total for s in running)
    done = [s["done"] for s in running]
    assert all(b >= a for a, b in zip(done, done[1:])) and done[-1] > done[0], "progress did not advance"
    assert max(done) <= p.total
    alphas = [s["alpha"] for s in running]
    assert all(0 < a <= p.alpha for a in alphas) and alphas[-1] < alphas[0]
    # the reference's schedule at the polled position: alpha * (1 - done / total), refreshed every MONITOR samples
    mid = running[len(running) // 2]
    assert abs(mid["alpha"] - p.alpha * (1 - mid["done"] / p.total)) < p.alpha * 0.02
    end = m.live_progress()
    assert not end["running"] and end["total"] == p.total
    assert end["done"] == result["stats"]["samples"] and end["done"] > 0.99 * p.total
    assert end["alpha"] < 0.02 * p.alpha


def test_progress_of_a_chunk_of_a_longer_schedule_and_of_walks():
    g, m = _model(20_000, 32)
    p = capi.default_params()
    p.mode, p.seed, p.total = capi.MODE_HOGWILD, 3, 1_000_000
    p.sched_total, p.sched_offset = 4_000_000, 2_000_000
    st = m.train_line(p)
    end = m.live_progress()
    assert end["total"] == 4_000_000 and end["done"] == 2_000_000 + st["samples"] and not end["running"]
    assert abs(end["alpha"] - p.alpha * (1 - 3_000_000 / 4_000_000)) < p.alpha * 0.01
    q = capi.default_params()
    q.mode, q.seed, q.walk_times, q.walk_steps, q.window_min, q.window_max = capi.MODE_HOGWILD, 3, 2, 10, 1, 3
    st = m.train_deepwalk(q)
    end = m.live_progress()
    assert end["total"] == end["done"] == st["samples"] and 39_000 < end["total"] <= 40_000  # units: walks (walk_times x V)


def test_cli_prints_the_reference_progress_line(tmp_path):
    rng = np.random.default_rng(4)
    net = tmp_path / "net.txt"
    with open(net, "w") as f:
        for a, b in zip(rng.integers(0, 3000, 60_000), rng.integers(0, 3000, 60_000)):
            f.write(f"v{a} v{b} 1\n")
    out = subprocess.run([os.path.join(BIN, "line"), "-train", str(net), "-save", str(tmp_path / "rep.txt"), "-dimensions", "64",
                          "-sample_times", "60"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "\tAlpha: " in out.stdout and "Progress: 100.00 %" in out.stdout
