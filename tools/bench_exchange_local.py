"""Single-GPU cost model of the bulk-exchange mode: `world` shards of configs[1] on ONE device, driven by
smore_train_line_group (device copies instead of NCCL). All phases run back to back on one GPU, so the number is the
HBM cost of the whole protocol (requests, gather, staging copies, updates, apply) per update -- what a rank pays locally
when NVLink is not the limit."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from smore_b200 import capi, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--world", type=int, default=2)
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--superbatch", type=int, default=1 << 20)
    ap.add_argument("--total", type=int, default=1 << 25)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--hot-threshold", type=float, default=-1.0)
    a = ap.parse_args()
    capi.check(capi.lib().smore_init(0))
    nv = int(1_000_000 * a.scale)
    src, dst, w = synth.power_law_edges(nv, 10 * nv, 20261018)
    off, col, ww, _ = synth.csr_from_edges(src, dst, w, True)
    ms = []
    for r in range(a.world):
        g = capi.Graph.from_csr(off, col, ww)
        g.set_shard(r, a.world)
        m = capi.Model(g, 128, 2, capi.F32)
        m.init(0, True, 1), m.init(1, False, 1)
        m.enable_exchange(a.superbatch, a.hot_threshold)
        ms.append(m)
    if a.hot_threshold >= 0:
        for t in range(2):
            ptrs = [m.device_ptr(t) for m in ms]
            for m in ms:
                m.set_peer_ptrs(t, ptrs)
    p = capi.default_params()
    p.mode, p.seed, p.total, p.negative_samples = capi.MODE_HOGWILD, 1, a.total, 5
    for step in range(a.steps + 1):
        p.stream_base = step << 30
        t0 = time.time()
        st = capi.train_line_group(ms, p)
        dt = time.time() - t0
        n = sum(s["samples"] for s in st)
        xs = [m.exchange_stats() for m in ms]
        print(json.dumps({"world": a.world, "superbatch": a.superbatch, "samples": n, "wall_s": dt,
                          "device_ms": st[0]["kernel_ms"], "updates_per_s": n / (st[0]["kernel_ms"] * 1e-3),
                          "rows_requested": sum(x["rows_requested"] for x in xs), "superbatches": xs[0]["superbatches"],
                          "hot_vertices": xs[0]["hot_vertices"], "hot_threshold": a.hot_threshold}), flush=True)


if __name__ == "__main__":
    main()
