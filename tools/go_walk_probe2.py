#!/usr/bin/env python
"""Diagnostic: Go DeepWalk on the SBM problem -- one deterministic stream vs Hogwild with 1, 2, 8 warps (same seed)."""
import json, os, sys
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
for mode, mw, dtype in ((capi.MODE_DETERMINISTIC, 1, capi.F64), (capi.MODE_HOGWILD, 1, capi.F32), (capi.MODE_HOGWILD, 2, capi.F32),
                        (capi.MODE_HOGWILD, 8, capi.F32)):
    for seed in (20261018, 11):
        m = capi.Model(g, DIM, 2, dtype)
        m.set_rows(0, Wv), m.set_rows(1, Wc)
        p = capi.default_params()
        p.semantics, p.mode, p.seed, p.alpha = capi.SEM_GO, mode, seed, 0.025
        p.walk_times, p.walk_steps, p.window_min, p.window_max, p.negative_samples, p.max_warps = 6, 40, 1, 5, 5, mw
        st = m.train_deepwalk(p)
        auc, rec = Q.evaluate_full(m.get_rows(0), m.get_rows(1), off, col, ts, td)
        print(json.dumps({"mode": mode, "max_warps": mw, "seed": seed, "pairs": int(st["pair_updates"]), "ms": st["kernel_ms"],
                          "auc": auc, "recall": rec}), flush=True)
