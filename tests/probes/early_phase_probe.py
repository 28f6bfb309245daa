#!/usr/bin/env python
"""Early-training AUC of LINE-2 on the 40 000-vertex SBM problem (dim 128) as a function of the worker count: the oracle on
one stream / all host cores vs the GPU with 16 / 512 / all warps, fp32 (atomic rows) and fp64 (stores). Question behind it:
the 1 M-vertex quality gate stops at 50 updates per vertex (AUC ~0.56) and the GPU came out 0.012 lower than the CPU
Hogwild run -- is that the number of concurrent workers, the arithmetic, or noise?"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import bindings as B  # noqa: E402
from smore_b200 import capi  # noqa: E402
from tests import quality as Q  # noqa: E402


def main():
    off, col, ww, ts, td = Q.sbm_problem(n_comm=500, comm_size=80, deg=24, seed=31)
    V, dim = len(off) - 1, 128
    Wv = (np.random.default_rng(1).random((V, dim)) - 0.5) / dim

    def auc_of(A, C):
        neg = np.random.default_rng(4).integers(0, V, len(ts))
        return Q.auc(np.einsum("ij,ij->i", A[ts], C[td]), np.einsum("ij,ij->i", A[ts], C[neg]))

    og = B.OracleGraph(B.SEM_CPP, off, col, ww)
    g = capi.Graph.from_csr(off, col, ww)
    for per_vertex in (50, 200):
        total = per_vertex * V
        for workers in (1, os.cpu_count() or 1):
            for seed in (11, 12):
                a, c = Wv.copy(), np.zeros((V, dim))
                og.time_line_cpp(a, c, 5, 0.025, total, seed, workers)
                print(f"{per_vertex:4d} updates/vertex  cpu oracle {workers:3d} threads seed {seed}: AUC {auc_of(a, c):.4f}", flush=True)
        for dtype, name in ((capi.F32, "f32"), (capi.F64, "f64")):
            for mw in (16, 512, 0):
                for seed in (13, 14):
                    m = capi.Model(g, dim, 2, dtype)
                    m.set_rows(0, Wv), m.set_rows(1, np.zeros((V, dim)))
                    p = capi.default_params()
                    p.semantics, p.mode, p.seed, p.alpha, p.total, p.negative_samples, p.max_warps = capi.SEM_CPP, capi.MODE_HOGWILD, seed, 0.025, total, 5, mw
                    st = m.train_line(p)
                    print(f"{per_vertex:4d} updates/vertex  gpu {name} max_warps {mw:4d} seed {seed}: AUC {auc_of(m.get_rows(0), m.get_rows(1)):.4f} "
                          f"({st['samples']} samples)", flush=True)


if __name__ == "__main__":
    main()
