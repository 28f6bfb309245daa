#!/usr/bin/env python
"""Diagnostic: Go-semantics DeepWalk / node2vec quality on tests/quality.py's SBM problem as a function of the number of
Hogwild warps and of the table dtype (does the recall@10 gap to the one-stream CPU path come from concurrency?)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from smore_b200 import capi  # noqa: E402
from tests import quality as Q  # noqa: E402

DIM = 32
off, col, ww, ts, td = Q.sbm_problem()
V = len(off) - 1
Wv = (np.random.default_rng(1).random((V, DIM)) - 0.5) / DIM
Wc = (np.random.default_rng(3).random((V, DIM)) - 0.5) / DIM
g = capi.Graph.from_csr(off, col, ww, semantics=capi.SEM_GO, n_lines=len(col) // 2)
for model in ("deepwalk", "node2vec"):
    for dtype, dn in ((capi.F32, "f32"), (capi.F64, "f64")):
        for mw in (32, 128, 512, 0):
            if dtype == capi.F64 and mw not in (128, 0):
                continue
            res = []
            for seed in (13, 14, 15):
                m = capi.Model(g, DIM, 2, dtype)
                m.set_rows(0, Wv), m.set_rows(1, Wc)
                p = capi.default_params()
                p.semantics, p.mode, p.seed, p.alpha = capi.SEM_GO, capi.MODE_HOGWILD, seed, 0.025
                p.walk_times, p.walk_steps, p.window_min, p.window_max, p.negative_samples, p.max_warps = 6, 40, 1, 5, 5, mw
                if model == "node2vec":
                    p.n2v_p, p.n2v_q = 0.5, 2.0
                    st = m.train_node2vec(p)
                else:
                    st = m.train_deepwalk(p)
                res.append(Q.evaluate_full(m.get_rows(0), m.get_rows(1), off, col, ts, td))
            r = np.array(res)
            print(json.dumps({"model": model, "dtype": dn, "max_warps": mw, "pairs": int(st["pair_updates"]), "ms": st["kernel_ms"],
                              "auc": r[:, 0].round(4).tolist(), "recall": r[:, 1].round(4).tolist(),
                              "mean_recall": float(r[:, 1].mean())}), flush=True)
