#!/usr/bin/env python
"""MF quality gate vs the number of concurrent warps (the linear model is the one most sensitive to stale concurrent
updates): held-out AUC / recall@10 on the SBM problem for max_warps in {64, 256, 512, policy}, 3 seeds each."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from smore_b200 import capi  # noqa: E402
from tests import quality as Q  # noqa: E402

ref = json.load(open(os.path.join(ROOT, "tests", "golden", "quality_baselines_v2.json")))["models"]["mf"]
off, col, ww, ts, td = Q.sbm_problem()
V, DIM = len(off) - 1, 32
Wv = (np.random.default_rng(1).random((V, DIM)) - 0.5) / DIM
g = capi.Graph.from_csr(off, col, ww, negative_method=capi.NEG_NO_DEGREES)
print(f"reference: AUC {ref['auc']:.4f} recall@10 {ref['recall_at_10']:.4f}")
for mw in (64, 256, 512, 1024, 0):
    res = []
    for seed in (13, 14, 15):
        m = capi.Model(g, DIM, 1, capi.F32)
        m.set_rows(0, Wv)
        p = capi.default_params()
        p.semantics, p.mode, p.seed, p.alpha, p.total = capi.SEM_CPP, capi.MODE_HOGWILD, seed, 0.025, ref["total"]
        p.negative_samples, p.lambda_, p.max_warps = ref["negative_samples"], ref["reg"], mw
        m.train_mf(p)
        W = m.get_rows(0)
        res.append(Q.evaluate_full(W, W, off, col, ts, td))
    r = np.array(res)
    print(f"max_warps {mw:5d}: AUC {r[:, 0].mean():.4f} +- {r[:, 0].std():.4f}  recall@10 {r[:, 1].mean():.4f} +- {r[:, 1].std():.4f}", flush=True)
