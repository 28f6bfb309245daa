#!/usr/bin/env python
"""Generates tests/golden/golden_v1.npz from the UNMODIFIED compiled reference (oracle/_ref/libsmore_ref.so =
/root/reference/src compiled -O2 with oracle/ref_shim.cpp replacing src/random.cpp). Run in the authoring container:

    make -C oracle ref && python tests/golden/make_golden.py

The reference ships no golden vectors of its own; these are outputs of the reference itself under the
replayed Philox stream, and are what pins oracle/smore_oracle.cpp (tests/test_oracle_golden.py) and the CUDA path
(tests/test_gpu_golden.py) where /root/reference is absent.
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import bindings as B  # noqa: E402
from tests import graphs  # noqa: E402

SEED = 20261018
OUT = os.path.join(ROOT, "tests", "golden", "golden_v1.npz")


def ref_model(tmp, kind, src, dst, w, undirected, dim, order=2, field=None):
    path = os.path.join(tmp, "g.txt")
    B.write_edge_list(path, src, dst, w)
    ffile = None
    if field is not None:
        ffile = os.path.join(tmp, "f.txt")
        with open(ffile, "w") as f:
            for name, fl in field:
                f.write(f"{name} {fl}\n")
    return B.Ref(kind, path, undirected, dim, order=order, field_file=ffile)


def main():
    G = {}
    tmp = tempfile.mkdtemp()
    # ---- Philox4x32-10 known answers (Random123 kat_vectors) ----
    G["philox_kat_ctr"] = np.array([[0, 0, 0, 0], [0xffffffff] * 4, [0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344]], dtype=np.uint32)
    G["philox_kat_key"] = np.array([[0, 0], [0xffffffff] * 2, [0xa4093822, 0x299f31d0]], dtype=np.uint32)
    G["philox_kat_out"] = np.array([[0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8],
                                    [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd],
                                    [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]], dtype=np.uint32)

    # ---- README graph (README.md:50-56): ingest, degrees, the three alias tables, sigmoid LUT, sampler replays ----
    src, dst, w = graphs.readme_graph()
    names = ["userA", "itemA", "itemC", "userB", "itemB", "userC"]
    path = os.path.join(tmp, "readme.txt")
    B.write_edge_list(path, src, dst, w, names=names)
    for und in (0, 1):
        ref = B.Ref(B.K_LINE, path, und, 4)
        off, col, ww = ref.csr()
        G[f"readme{und}_off"], G[f"readme{und}_col"], G[f"readme{und}_w"] = off, col, ww
        G[f"readme{und}_names"] = np.array(ref.names())
        for which, nm in ((0, "vertex"), (1, "negative"), (2, "context")):
            p, a = ref.alias(which)
            G[f"readme{und}_{nm}_prob"], G[f"readme{und}_{nm}_alias"] = p, a
        ref.seed(SEED, 1)
        G[f"readme{und}_source"] = ref.sample(0, 2000)
        ref.seed(SEED, 2)
        G[f"readme{und}_negative"] = ref.sample(1, 2000)
        ref.seed(SEED, 3)
        G[f"readme{und}_source_target"] = ref.sample(3, 2000)
        if und == 0:
            G["sigmoid_table"] = ref.sigmoid_table()

    # ---- a 300-vertex power-lawish graph: alias tables, samplers, walks ----
    src, dst, w = graphs.random_graph(300, 3000, seed=101)
    G["g300_src"], G["g300_dst"], G["g300_w"] = src, dst, w
    ref = ref_model(tmp, B.K_DEEPWALK, src, dst, w, 1, 8)
    for which, nm in ((0, "vertex"), (1, "negative"), (2, "context")):
        p, a = ref.alias(which)
        G[f"g300_{nm}_prob"], G[f"g300_{nm}_alias"] = p, a
    ref.seed(SEED, 4)
    G["g300_source_target"] = ref.sample(3, 20000)
    walks, pv, pc, lens, npairs = [], [], [], [], []
    for start in range(0, 300, 7):
        for mode, w0, w1 in ((0, 5, 0), (1, 2, 5)):
            ref.seed(SEED, 1000 + start)
            wk, a, b = ref.walk_pairs(start, 40, mode, w0, w1)
            walks.append(np.pad(wk, (0, 41 - len(wk)), constant_values=-1))
            lens.append(len(wk))
            npairs.append(len(a))
            pv.append(a)
            pc.append(b)
    G["g300_walks"], G["g300_walk_len"], G["g300_walk_npairs"] = np.array(walks), np.array(lens), np.array(npairs)
    G["g300_pair_v"], G["g300_pair_c"] = np.concatenate(pv), np.concatenate(pc)

    # ---- Train() of every model on the hot path: final tables after the CLI's minimum run ----
    dim = 8
    Wv, Wc = graphs.init_tables(len(ref.names()), dim, seed=11, context_zero=True)
    G["g300_init_v"], G["g300_init_c"] = Wv, Wc
    for order in (1, 2):
        r = ref_model(tmp, B.K_LINE, src, dst, w, 1, dim, order=order)
        r.set_rows(0, Wv)
        if order == 2:
            r.set_rows(1, Wc)
        r.seed(SEED, 0)
        r.train(1, 5, alpha=0.025, workers=1)
        G[f"line{order}_v"] = r.get_rows(0)
        G[f"line{order}_words"] = np.array(r.pos(), dtype=np.uint64)
        if order == 2:
            G["line2_c"] = r.get_rows(1)
    Wv2, Wc2 = graphs.init_tables(len(ref.names()), dim, seed=12)
    G["g300_init_v2"], G["g300_init_c2"] = Wv2, Wc2
    for kind, nm, args in ((B.K_DEEPWALK, "deepwalk", (3, 20, 5, 5, 0)), (B.K_WALKLETS, "walklets", (3, 20, 2, 4, 5))):
        r = ref_model(tmp, kind, src, dst, w, 1, dim)
        r.set_rows(0, Wv2)
        r.set_rows(1, Wc2)
        r.seed(SEED, 0)
        r.train(*args, alpha=0.025, workers=1)
        G[f"{nm}_v"], G[f"{nm}_c"] = r.get_rows(0), r.get_rows(1)
        G[f"{nm}_words"] = np.array(r.pos(), dtype=np.uint64)

    # bipartite graph for the ranking models
    nu, ni = 150, 90
    bs, bd, bw = graphs.bipartite_graph(nu, ni, 2500, seed=103)
    G["bip_src"], G["bip_dst"], G["bip_w"] = bs, bd, bw
    G["bip_nu"] = np.array(nu)
    for kind, nm in ((B.K_BPR, "bpr"), (B.K_WARP, "warp")):
        r = ref_model(tmp, kind, bs, bd, bw, 0, dim)
        W0, _ = graphs.init_tables(r.V, dim, seed=13)
        W0 = W0 * (30.0 if nm == "warp" else 1.0)  # WARP: margins around 1 so the scan goes beyond the first negative
        G[f"{nm}_init"] = W0
        r.set_rows(0, W0)
        r.seed(SEED, 0)
        r.train(1, 5, alpha=0.025, workers=1)
        G[f"{nm}_v"] = r.get_rows(0)
        G[f"{nm}_words"] = np.array(r.pos(), dtype=np.uint64)
    labels = sorted(set(bs.tolist()) | set(bd.tolist()))
    field = [(f"v{l}", "u" if l < nu else "i") for l in labels]
    r = ref_model(tmp, B.K_HOPREC, bs, bd, bw, 1, dim, field=field)
    W0, _ = graphs.init_tables(r.V, dim, seed=14)
    W0 = W0 * 8.0
    G["hoprec_init"] = W0
    G["hoprec_field"] = r.fields()
    r.set_rows(0, W0)
    r.seed(SEED, 0)
    r.train(1, 3, alpha=0.025, workers=1)
    G["hoprec_v"] = r.get_rows(0)
    G["hoprec_words"] = np.array(r.pos(), dtype=np.uint64)

    np.savez_compressed(OUT, **G)
    print("wrote", OUT, os.path.getsize(OUT), "bytes,", len(G), "arrays")


if __name__ == "__main__":
    main()
