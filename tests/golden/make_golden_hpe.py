#!/usr/bin/env python
"""Generates tests/golden/golden_hpe_v1.npz from the UNMODIFIED compiled reference (see make_golden.py for the set-up):
HPE::Train (src/model/HPE.cpp:93-147) on the 300-vertex graph of golden_v1, under the replayed Philox stream.

    make -C oracle ref && python tests/golden/make_golden_hpe.py
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

OUT = os.path.join(ROOT, "tests", "golden", "golden_hpe_v1.npz")


def main():
    G = {}
    tmp = tempfile.mkdtemp()
    src, dst, w = graphs.random_graph(300, 3000, seed=101)  # == golden_v1's g300
    G["g300_src"], G["g300_dst"], G["g300_w"] = src, dst, w
    dim = 8
    for tag, steps, K, reg in (("a", 3, 5, 0.01), ("b", 5, 2, 0.05)):
        r = ref_model(tmp, B.K_HPE, src, dst, w, 1, dim)
        Wv, Wc = graphs.init_tables(r.V, dim, seed=21)
        G["init_v"], G["init_c"] = Wv, Wc
        r.set_rows(0, Wv), r.set_rows(1, Wc)
        r.seed(SEED, 0)
        r.train_hpe(1, steps, K, reg, alpha=0.025, workers=1)
        G[f"hpe_{tag}_v"], G[f"hpe_{tag}_c"] = r.get_rows(0), r.get_rows(1)
        G[f"hpe_{tag}_words"] = np.array(r.pos(), dtype=np.uint64)
        G[f"hpe_{tag}_args"] = np.array([steps, K, reg])
    np.savez_compressed(OUT, **G)
    print("wrote", OUT, os.path.getsize(OUT), "bytes,", len(G), "arrays")


if __name__ == "__main__":
    main()
