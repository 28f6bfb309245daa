#!/usr/bin/env python
"""Throughput of ONE rank's block kernel of the rotating-shard mode at the full per-GPU footprint of BASELINE configs[4]
(12.5 M context rows + 6.25 M resident vertex rows + 500 M block-table entries: ~18 GB touched at random), on a single GPU:
the graph is built as rank 0 of a 2-rank ring and the rank's sub-parts "rotate" onto itself. This is the kernel the N > 1
lines of bench.py are bound by; the tool exists to tune it (SMORE_B200_LIB selects an experiment build of the library).

    python tools/bench_block.py [--scale 1.0] [--episodes 8] [--episode-batch 8388608] [--split]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from smore_b200 import capi  # noqa: E402
from smore_b200 import dist as sdist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--episodes", type=int, default=8)
    ap.add_argument("--episode-batch", type=int, default=1 << 23)
    ap.add_argument("--split", action="store_true")
    ap.add_argument("--dim", type=int, default=128)
    a = ap.parse_args()
    V, E = int(25_000_000 * a.scale), int(500_000_000 * a.scale)
    t0 = time.time()
    g = capi.Graph.synthetic_rotating(V, E, 20261018, 0, 2)
    m = capi.Model(g, a.dim, 2, capi.F32)
    m.init(0, True, 1), m.init(1, False, 1)
    m.enable_rotation()
    sdist.connect_rotation_local([m])  # a ring of one: the outgoing sub-part lands in this rank's own incoming slot
    build_s = time.time() - t0
    p = capi.default_params()
    p.semantics, p.mode, p.seed, p.alpha, p.negative_samples = capi.SEM_CPP, capi.MODE_HOGWILD, 1, 0.025, 5
    p.total = a.episode_batch * 2  # both ranks' samples per episode; this rank runs its ~half
    p.neg_mode = capi.PAIRING_SPLIT if a.split else capi.PAIRING_COUPLED
    sdist.train_line_rotating([m], p, 4, world=2)  # warm-up: one full cycle
    done, ms = sdist.train_line_rotating([m], p, a.episodes, first_episode=4, world=2)
    rows = 8 if a.split else 7
    byt = 2 * rows * a.dim * 4 + (84 if a.split else 76)
    peak = 6542.7
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = float(json.load(open(pk))["hbm_gbs"])
    rate = done[0] / (ms[0] * 1e-3)
    print(json.dumps({"lib": os.environ.get("SMORE_B200_LIB", "default"), "V": V, "E_lines": E, "build_s": round(build_s, 1),
                      "updates_per_s": rate, "ms_per_episode": ms[0] / a.episodes, "algorithmic_bytes_per_update": byt,
                      "frac_of_measured_hbm": rate * byt / 1e9 / peak, "split": a.split}), flush=True)


if __name__ == "__main__":
    main()
