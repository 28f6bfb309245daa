#!/usr/bin/env python
"""Per-model throughput of the hot path on one B200 (diagnostic companion of bench.py; same timing rules).

Prints one JSON line per model: updates/s (pair updates for the skip-gram models, samples for the ranking models),
algorithmic GB/s (SURVEY.md §8d byte model) and the fraction of the measured HBM peak.
Graph sizes are the BASELINE.json configs scaled by --scale (configs[3] is always scaled: 500M interactions do not
fit the host-side numpy generator's time budget).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from smore_b200 import capi, synth  # noqa: E402


def peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return float(json.load(open(p))["hbm_gbs"]) if os.path.exists(p) else 6650.0


def run(name, model, fn, p, bytes_per_unit, unit_key, steps, warmup):
    for i in range(warmup):
        p.stream_base = i * (1 << 20)
        fn(p)
    ms, units, tries = 0.0, 0, 0.0
    for i in range(steps):
        p.stream_base = (warmup + i) * (1 << 20)
        st = fn(p)
        ms += st["kernel_ms"]
        units += st[unit_key]
        tries = st["mean_tries"]
    b = bytes_per_unit(tries) if callable(bytes_per_unit) else bytes_per_unit
    gbs = units * b / (ms * 1e-3) / 1e9
    print(json.dumps({"model": name, "units_per_s": units / (ms * 1e-3), "unit": unit_key, "ms_per_step": ms / steps,
                      "algorithmic_bytes_per_unit": b, "algorithmic_GBps": gbs, "frac_of_measured_hbm": gbs / peak(),
                      "mean_tries": tries}), flush=True)


def device_bipartite_csr(nu, ni, ne, seed, zipf_s, undirected):
    """configs[3]-sized bipartite CSR built with torch ON THE GPU (the numpy generator + first-appearance relabelling of
    smore_b200/synth.py needs ~10 minutes of host sorting at 500 M interactions): users uniform, items Zipf(s) over a random
    permutation, integer weights 1..5; vids are the labels themselves (users [0, nu), items [nu, nu + ni)), entries of a row
    in generation order. -> (row_off int64, col int32, w float64, field int32) as numpy arrays."""
    import torch
    dev = torch.device("cuda:0")
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    cdf = torch.cumsum(torch.arange(1, ni + 1, device=dev, dtype=torch.float64) ** (-zipf_s), 0)
    cdf /= cdf[-1].clone()
    perm = torch.randperm(ni, device=dev, generator=gen)
    chunks_u, chunks_i, chunks_w = [], [], []
    step = 1 << 26
    for lo in range(0, ne, step):
        n = min(step, ne - lo)
        chunks_u.append(torch.randint(0, nu, (n,), device=dev, generator=gen, dtype=torch.int32))
        r = torch.rand(n, device=dev, generator=gen, dtype=torch.float64)
        chunks_i.append((nu + perm[torch.searchsorted(cdf, r, right=True).clamp_(max=ni - 1)]).to(torch.int32))
        chunks_w.append(torch.randint(1, 6, (n,), device=dev, generator=gen, dtype=torch.int8))
    u, it, w = torch.cat(chunks_u), torch.cat(chunks_i), torch.cat(chunks_w)
    del chunks_u, chunks_i, chunks_w
    if undirected:
        es, ed, w = torch.cat([u, it]), torch.cat([it, u]), torch.cat([w, w])
    else:
        es, ed = u, it
    del u, it
    V = nu + ni
    es_sorted, order = torch.sort(es, stable=True)
    del es
    col = ed[order].cpu().numpy()
    ww = w[order].to(torch.float64).cpu().numpy()
    del ed, w, order
    off = torch.zeros(V + 1, dtype=torch.int64, device=dev)
    off[1:] = torch.cumsum(torch.bincount(es_sorted.to(torch.int64), minlength=V), 0)
    off = off.cpu().numpy()
    del es_sorted
    torch.cuda.empty_cache()
    field = (np.arange(V) >= nu).astype(np.int32)
    return off, col, ww, field


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--only", default="")
    ap.add_argument("--zipf", type=float, default=1.0, help="item popularity exponent of the bipartite graphs")
    ap.add_argument("--c4-full", action="store_true",
                    help="BASELINE.json configs[3] at its named size: 10 M users x 2 M items, 500 M interactions (WARP, C++ BPR "
                         "directed: 500 M CSR entries; HOP-Rec undirected: 1 B), generated on the GPU")
    a = ap.parse_args()
    only = set(a.only.split(",")) if a.only else None

    def want(n):
        return only is None or n in only

    base = capi.default_params()
    base.mode, base.seed = capi.MODE_HOGWILD, 1

    # ---- configs[0]: BPR dim 64, 10k users x 10k items, 1M edges (and the same graph at dim 128) ----
    if want("bpr_go") or want("bpr_cpp"):
        src, dst, w = synth.bipartite_edges(10_000, 10_000, 1_000_000, 7, zipf_s=a.zipf)
        off, col, ww, _ = synth.csr_from_edges(src, dst, w, False)
        for dim in (64, 128):
            if want("bpr_go"):
                g = capi.Graph.from_csr(off, col, ww, semantics=capi.SEM_GO, n_lines=len(src))
                m = capi.Model(g, dim, 2, capi.F32)
                m.init(0, True, 1), m.init(1, True, 2)
                p = capi.default_params()
                p.semantics, p.mode, p.seed, p.total, p.lambda_ = capi.SEM_GO, capi.MODE_HOGWILD, 1, 1 << 25, 0.001
                run(f"bpr_go_d{dim}_c1", m, m.train_bpr, p, 6 * dim * 4 + 40, "samples", a.steps, a.warmup)
            if want("bpr_cpp"):
                g = capi.Graph.from_csr(off, col, ww, semantics=capi.SEM_CPP, negative_method=capi.NEG_NO_DEGREES)
                m = capi.Model(g, dim, 1, capi.F32)
                m.init(0, True, 1)
                p = capi.default_params()
                p.semantics, p.mode, p.seed, p.total = capi.SEM_CPP, capi.MODE_HOGWILD, 1, 1 << 24
                run(f"bpr_cpp_d{dim}_c1", m, m.train_bpr, p, 14 * dim * 4 + 60, "samples", a.steps, a.warmup)

    # ---- configs[3] at the named size ----
    if a.c4_full:
        nu, ni, ne, dim = 10_000_000, 2_000_000, 500_000_000, 128
        for und, names in ((False, ("bpr_cpp", "warp")), (True, ("hoprec",))):
            if not any(want(n) for n in names):
                continue
            t0 = time.time()
            off, col, ww, field = device_bipartite_csr(nu, ni, ne, 9, a.zipf, und)
            t1 = time.time()
            g = capi.Graph.from_csr(off, col, ww, semantics=capi.SEM_CPP, negative_method=capi.NEG_NO_DEGREES)
            print(json.dumps({"graph": "configs[3]", "undirected": und, "V": nu + ni, "csr_entries": int(len(col)),
                              "generate_s": round(t1 - t0, 1), "alias_build_and_upload_s": round(time.time() - t1, 1)}), flush=True)
            del off, col, ww
            if und:
                g.set_field(field)
            for nm in names:
                if not want(nm):
                    continue
                m = capi.Model(g, dim, 1, capi.F32)
                m.init(0, True, 1)
                p = capi.default_params()
                p.semantics, p.mode, p.seed = capi.SEM_CPP, capi.MODE_HOGWILD, 1
                if nm == "bpr_cpp":
                    p.total = 1 << 24
                    run("bpr_cpp_d128_c4", m, m.train_bpr, p, 14 * dim * 4 + 60, "samples", a.steps, a.warmup)
                elif nm == "warp":
                    p.total = 1 << 24
                    run("warp_d128_c4", m, m.train_warp, p, lambda tr: (2 + tr) * dim * 4 + 3 * dim * 4 + 40, "samples", a.steps, a.warmup)
                else:
                    p.total, p.walk_steps = 1 << 22, 5
                    run("hoprec_d128_c4", m, m.train_hoprec, p, 5 * 14 * dim * 4, "samples", a.steps, a.warmup)
                del m
            del g
        return

    # ---- a large bipartite graph (configs[3] scaled): BPR at dim 128 out of L2, WARP, HOP-Rec ----
    if want("bpr_go_big") or want("warp") or want("hoprec") or want("bpr_cpp_big") or want("mf") or want("skewopt"):
        nu, ni, ne = int(1_000_000 * a.scale), int(200_000 * a.scale), int(20_000_000 * a.scale)
        t0 = time.time()
        src, dst, w = synth.bipartite_edges(nu, ni, ne, 9, zipf_s=a.zipf)
        dim = 128
        if want("bpr_go_big"):
            off, col, ww, _ = synth.csr_from_edges(src, dst, w, False)
            g = capi.Graph.from_csr(off, col, ww, semantics=capi.SEM_GO, n_lines=len(src))
            m = capi.Model(g, dim, 2, capi.F32)
            m.init(0, True, 1), m.init(1, True, 2)
            p = capi.default_params()
            p.semantics, p.mode, p.seed, p.total, p.lambda_ = capi.SEM_GO, capi.MODE_HOGWILD, 1, 1 << 25, 0.001
            run(f"bpr_go_d{dim}_big", m, m.train_bpr, p, 6 * dim * 4 + 40, "samples", a.steps, a.warmup)
            del m, g
        if want("bpr_cpp_big") or want("warp") or want("mf") or want("skewopt"):
            off, col, ww, _ = synth.csr_from_edges(src, dst, w, False)
            g = capi.Graph.from_csr(off, col, ww, semantics=capi.SEM_CPP, negative_method=capi.NEG_NO_DEGREES)
            if want("mf"):
                m = capi.Model(g, dim, 1, capi.F32)
                m.init(0, True, 1)
                p = capi.default_params()
                p.semantics, p.mode, p.seed, p.total, p.lambda_ = capi.SEM_CPP, capi.MODE_HOGWILD, 1, 1 << 24, 0.01
                # MF::Train = LINE's sampling loop around UpdateFactorizedPair: 1 + (1 + K) rows R+W
                run(f"mf_d{dim}_big", m, m.train_mf, p, 2 * 7 * dim * 4 + 76, "samples", a.steps, a.warmup)
            if want("skewopt"):
                m = capi.Model(g, dim, 1, capi.F32)
                m.init(0, True, 1)
                p = capi.default_params()
                p.semantics, p.mode, p.seed, p.total = capi.SEM_CPP, capi.MODE_HOGWILD, 1, 1 << 23
                # UpdateSBPRPair: user + item + 16 negatives read; written: the accepted rounds' negatives + item + user
                # (mean_tries = accepted rounds per sample)
                run(f"skewopt_d{dim}_big", m, m.train_skewopt, p, lambda acc: (18 + acc + 2) * dim * 4 + 150, "samples",
                    a.steps, a.warmup)
            if want("bpr_cpp_big"):
                m = capi.Model(g, dim, 1, capi.F32)
                m.init(0, True, 1)
                p = capi.default_params()
                p.semantics, p.mode, p.seed, p.total = capi.SEM_CPP, capi.MODE_HOGWILD, 1, 1 << 24
                run(f"bpr_cpp_d{dim}_big", m, m.train_bpr, p, 14 * dim * 4 + 60, "samples", a.steps, a.warmup)
            if want("warp"):
                m = capi.Model(g, dim, 1, capi.F32)
                m.init(0, True, 1)
                p = capi.default_params()
                p.semantics, p.mode, p.seed, p.total = capi.SEM_CPP, capi.MODE_HOGWILD, 1, 1 << 24
                # (2 + tries) rows read, 3 written when violated; early in training every sample violates
                run(f"warp_d{dim}_big", m, m.train_warp, p, lambda tr: (2 + tr) * dim * 4 + 3 * dim * 4 + 40, "samples",
                    a.steps, a.warmup)
            del m, g
        if want("hoprec"):
            off, col, ww, labels = synth.csr_from_edges(src, dst, w, True)
            g = capi.Graph.from_csr(off, col, ww, semantics=capi.SEM_CPP, negative_method=capi.NEG_NO_DEGREES)
            g.set_field((labels >= nu).astype(np.int32))
            m = capi.Model(g, dim, 1, capi.F32)
            m.init(0, True, 1)
            p = capi.default_params()
            p.semantics, p.mode, p.seed, p.total, p.walk_steps = capi.SEM_CPP, capi.MODE_HOGWILD, 1, 1 << 22, 5
            # per hop: user + item + 5 negatives read, up to all 7 written: counted per FBPR round-group ("pair_updates"/5)
            run(f"hoprec_d{dim}_big", m, m.train_hoprec, p, 5 * 14 * dim * 4, "samples", a.steps, a.warmup)

    # ---- the Go tree's two-graph models on a planted-preference graph (200 k users x 100 k items, 3.6 M interactions,
    #      undirected as cmd/cpr / cmd/tpr load it) + their second graph; dim 128 ----
    if want("cpr") or want("tpr"):
        sys.path.insert(0, ROOT)
        from tests import quality as Q
        for kind in ("cpr", "tpr"):
            if not want(kind):
                continue
            off, col, ww, _, _, _, aoff, acol = Q.two_graph_problem(kind, n_comm=int(5000 * a.scale))
            V = len(off) - 1
            g = capi.Graph.from_csr(off, col, ww, semantics=capi.SEM_GO, n_lines=len(col) // 2)
            m = capi.Model(g, 128, 2, capi.F32)
            m.init(0, True, 1), m.init(1, True, 2)
            m.attach_aux(aoff, acol, rows=None, seed=3)
            p = capi.default_params()
            p.semantics, p.mode, p.seed, p.total = capi.SEM_GO, capi.MODE_HOGWILD, 1, 1 << 22
            d1 = np.diff(off).astype(np.float64)
            d2 = np.diff(aoff[:V + 1]).astype(np.float64)
            if kind == "cpr":
                # rows per sample: user + pos + neg read and written, + the user's neighbours in both graphs read;
                # users are drawn ~ out-degree, so E[neighbours] = sum d1 (d1 + d2) / sum d1
                p.alpha, p.lambda_, p.item_reg, p.margin = 0.1, 0.01, 0.01, 8.0
                rows = 6 + float((d1 * (d1 + d2)).sum() / d1.sum())
                run(f"cpr_d128", m, m.train_cpr, p, rows * 128 * 4 + 40, "samples", a.steps, a.warmup)
            else:
                # user + 2 items read and written; each item's words read for the enriched vector, read again and written
                # by the update: positives ~ in-degree given a degree-drawn user (~ d1), negatives ~ d1^0.75
                p.alpha, p.lambda_, p.text_weight = 0.025, 0.025, 0.5
                pos_w = float((d1 * d2).sum() / d1.sum())
                neg_w = float((d1 ** 0.75 * d2).sum() / (d1 ** 0.75).sum())
                rows = 6 + 3 * (pos_w + neg_w)
                run(f"tpr_d128", m, m.train_tpr, p, rows * 128 * 4 + 40, "samples", a.steps, a.warmup)
            del m, g

    # ---- configs[1] under Go semantics (CDF-scan neighbour sampling -> binary search over prefix sums, random contexts) ----
    if want("line_go") or want("line_cpp") or want("hpe"):
        nv = int(1_000_000 * a.scale)
        src, dst, w = synth.power_law_edges(nv, 10 * nv, 20261018)
        off, col, ww, _ = synth.csr_from_edges(src, dst, w, True)
        for nm, sem in (("line_cpp", capi.SEM_CPP), ("line_go", capi.SEM_GO)):
            if not want(nm):
                continue
            g = capi.Graph.from_csr(off, col, ww, semantics=sem, n_lines=len(src) if sem == capi.SEM_GO else 0)
            m = capi.Model(g, 128, 2, capi.F32)
            m.init(0, True, 1), m.init(1, sem == capi.SEM_GO, 2)
            p = capi.default_params()
            p.semantics, p.mode, p.seed, p.total = sem, capi.MODE_HOGWILD, 1, 1 << 24
            run(f"{nm}_d128_c2", m, m.train_line, p, 2 * 7 * 128 * 4 + 76, "samples", a.steps, a.warmup)
            del m, g
        if want("hpe"):
            # HPE (C++): walk_steps regularised skip-gram steps + one UpdatePair per sample = (steps + 1) x (K + 2) rows R+W
            g = capi.Graph.from_csr(off, col, ww, semantics=capi.SEM_CPP)
            m = capi.Model(g, 128, 2, capi.F32)
            m.init(0, True, 1), m.init(1, True, 2)
            p = capi.default_params()
            p.semantics, p.mode, p.seed, p.total, p.walk_steps, p.lambda_ = capi.SEM_CPP, capi.MODE_HOGWILD, 1, 1 << 22, 5, 0.01
            run("hpe_d128_c2", m, m.train_hpe, p, 2 * 7 * 128 * 4 + 28, "pair_updates", a.steps, a.warmup)
            del m, g

    # ---- configs[2]: DeepWalk dim 128, walk_steps 40, window 5 (V scaled; one epoch slice per step) ----
    if want("deepwalk") or want("walklets"):
        nv = int(1_000_000 * a.scale)
        src, dst, w = synth.power_law_edges(nv, 5 * nv, 11)
        off, col, ww, _ = synth.csr_from_edges(src, dst, w, True)
        g = capi.Graph.from_csr(off, col, ww, semantics=capi.SEM_CPP)
        dim = 128
        for nm in ("deepwalk", "walklets"):
            if not want(nm):
                continue
            m = capi.Model(g, dim, 2, capi.F32)
            m.init(0, True, 1), m.init(1, True, 2)
            p = capi.default_params()
            p.semantics, p.mode, p.seed = capi.SEM_CPP, capi.MODE_HOGWILD, 1
            p.walk_times, p.walk_steps, p.window_min, p.window_max, p.max_walks = 1, 40, (2 if nm == "walklets" else 1), 5, 200_000
            run(f"{nm}_d{dim}", m, m.train_deepwalk if nm == "deepwalk" else m.train_walklets, p, 2 * 7 * dim * 4 + 76,
                "pair_updates", a.steps, a.warmup)


if __name__ == "__main__":
    main()
