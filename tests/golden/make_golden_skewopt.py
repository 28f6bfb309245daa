#!/usr/bin/env python
"""Generates tests/golden/golden_skewopt_v1.npz from the UNMODIFIED compiled reference (see make_golden.py): SPR::Train
(src/model/SkewOPT.cpp, UpdateSBPRPair src/proNet.cpp:1517-1566) on the bipartite graph of golden_v1 under the replayed
Philox stream. Groundwork for the Skew-OPT kernel: pins the oracle restatement today, the CUDA
path when it exists.

    make -C oracle ref && python tests/golden/make_golden_skewopt.py
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import bindings as B  # noqa: E402
from tests import graphs  # noqa: E402
from tests.golden.make_golden import SEED, ref_model  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "golden_skewopt_v1.npz")


def main():
    G = {}
    tmp = tempfile.mkdtemp()
    dim = 8
    bs, bd, bw = graphs.bipartite_graph(150, 90, 2500, seed=103)
    G["bip_src"], G["bip_dst"], G["bip_w"] = bs, bd, bw
    for tag, xi, omega, eta, scale in (("a", 10.0, 3.0, 3, 60.0), ("b", 0.5, 0.7, 2, 8.0)):
        r = ref_model(tmp, B.K_SKEWOPT, bs, bd, bw, 0, dim)
        W0, _ = graphs.init_tables(r.V, dim, seed=27)
        W0 = W0 * scale + 0.01
        G[f"{tag}_init"] = W0
        r.set_rows(0, W0)
        r.seed(SEED, 0)
        r.train_skewopt(1, xi, omega, eta, alpha=0.025, workers=1)
        G[f"{tag}_v"] = r.get_rows(0)
        G[f"{tag}_words"] = np.array(r.pos(), dtype=np.uint64)
        G[f"{tag}_args"] = np.array([xi, omega, eta])
    np.savez_compressed(OUT, **G)
    print("wrote", OUT, os.path.getsize(OUT), "bytes,", len(G), "arrays")


if __name__ == "__main__":
    main()
