#!/usr/bin/env python
"""Diagnostic (CPU): does the device's W-stream work split of the Go DeepWalk loop -- run SEQUENTIALLY by the oracle, so
without any concurrency -- move recall@10 on the quality gate's problem? (tools/go_walk_probe.py is the GPU side.)"""
import sys, os, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import bindings as B  # noqa: E402
from tests import quality as Q  # noqa: E402
DIM = 32
off, col, ww, ts, td = Q.sbm_problem()
V = len(off) - 1
iv = (np.random.default_rng(1).random((V, DIM)) - 0.5) / DIM
ic = (np.random.default_rng(3).random((V, DIM)) - 0.5) / DIM
gg = B.OracleGraph(B.SEM_GO, off, col, ww, max_line=len(col) // 2)
for W in [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "1,32,128").split(",")]:
    for seed in (13, 14, 15):
        a, c = iv.copy(), ic.copy()
        pairs = gg.train_deepwalk_go_streams(a, c, 6, 40, 5, 5, 0.025, seed, W)
        auc, rec = Q.evaluate_full(a, c, off, col, ts, td)
        print(json.dumps({"streams": W, "seed": seed, "pairs": pairs, "auc": round(auc, 4), "recall": round(rec, 4)}), flush=True)
