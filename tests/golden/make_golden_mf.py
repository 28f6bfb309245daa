#!/usr/bin/env python
"""Generates tests/golden/golden_mf_v1.npz from the UNMODIFIED compiled reference (see make_golden.py): MF::Train
(src/model/MF.cpp:50-98) on the bipartite graph of golden_v1 and on a small general graph where negatives can be the
vertex itself (one shared table), under the replayed Philox stream.

    make -C oracle ref && python tests/golden/make_golden_mf.py
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

OUT = os.path.join(ROOT, "tests", "golden", "golden_mf_v1.npz")


def main():
    G = {}
    tmp = tempfile.mkdtemp()
    dim = 8
    bs, bd, bw = graphs.bipartite_graph(150, 90, 2500, seed=103)  # == golden_v1's bipartite graph
    gs, gd, gw = graphs.random_graph(40, 600, seed=107)           # every vertex has in-degree: negatives hit the vertex row
    for tag, (src, dst, w), K, reg in (("bip", (bs, bd, bw), 5, 0.01), ("small", (gs, gd, gw), 8, 0.05)):
        G[f"{tag}_src"], G[f"{tag}_dst"], G[f"{tag}_w"] = src, dst, w
        r = ref_model(tmp, B.K_MF, src, dst, w, 0, dim)
        W0, _ = graphs.init_tables(r.V, dim, seed=23)
        G[f"{tag}_init"] = W0
        r.set_rows(0, W0)
        r.seed(SEED, 0)
        r.train_mf(1, K, reg, alpha=0.025, workers=1)
        G[f"{tag}_v"] = r.get_rows(0)
        G[f"{tag}_words"] = np.array(r.pos(), dtype=np.uint64)
        G[f"{tag}_args"] = np.array([K, reg])
    np.savez_compressed(OUT, **G)
    print("wrote", OUT, os.path.getsize(OUT), "bytes,", len(G), "arrays")


if __name__ == "__main__":
    main()
