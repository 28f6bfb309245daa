"""Does the L1-allocating row gather (Row::load_ca, SMORE_L1_GATHER) cost model quality on a LARGE table?

Planted-preference bipartite graph: users and items live in clusters, a user buys from its cluster's items with
probability 0.8 (Zipf(s) popularity inside the cluster) and from the global Zipf(s) catalogue otherwise. 5 % of the
interactions are held out; AUC = P(score(u, held-out item) > score(u, random item)). Go-semantics BPR (users in the
vertex table, items in the context table) is trained twice on the same graph with the same draws: gathers through L1
(ld.global.ca) and around it (ld.global.cg). Prints one JSON line per run."""
import argparse
import json
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402


def planted(nu, ni, ne, n_clusters, zipf_s, seed):
    rng = np.random.default_rng(seed)
    per = ni // n_clusters
    p = np.arange(1, per + 1, dtype=np.float64) ** (-zipf_s)
    cdf = np.cumsum(p) / p.sum()
    pg = np.arange(1, ni + 1, dtype=np.float64) ** (-zipf_s)
    cdfg = np.cumsum(pg) / pg.sum()
    perm = rng.permutation(ni)
    users = rng.integers(0, nu, size=ne)
    cu = users % n_clusters
    inside = rng.random(ne) < 0.8
    rank_in = np.searchsorted(cdf, rng.random(ne), side="right").clip(max=per - 1)
    item_in = cu * per + rank_in
    item_gl = perm[np.searchsorted(cdfg, rng.random(ne), side="right").clip(max=ni - 1)]
    items = np.where(inside, item_in, item_gl)
    w = np.ones(ne)
    return users.astype(np.int64), (nu + items).astype(np.int64), w


def run(a):
    from smore_b200 import capi, synth

    capi.check(capi.lib().smore_init(0))
    src, dst, w = planted(a.users, a.items, a.edges, a.clusters, a.zipf, 11)
    (ts, td, tw), (hs, hd, _) = synth.split_edges(src, dst, w, 0.05, seed=12)
    off, col, ww, labels = synth.csr_from_edges(ts, td, tw, False)
    lab2id = np.full(a.users + a.items, -1, dtype=np.int64)
    lab2id[labels] = np.arange(len(labels))
    ok = (lab2id[hs] >= 0) & (lab2id[hd] >= 0)
    hu, hi = lab2id[hs[ok]], lab2id[hd[ok]]
    item_ids = lab2id[np.arange(a.users, a.users + a.items)]
    item_ids = item_ids[item_ids >= 0]
    g = capi.Graph.from_csr(off, col, ww, semantics=capi.SEM_GO, n_lines=len(ts))
    m = capi.Model(g, a.dim, 2, capi.F32)
    m.init(0, True, 1), m.init(1, True, 2)
    p = capi.default_params()
    p.semantics, p.mode, p.seed, p.total, p.lambda_, p.alpha = capi.SEM_GO, capi.MODE_HOGWILD, 1, a.total, 0.001, 0.025
    t0 = time.time()
    st = m.train_bpr(p)
    rng = np.random.default_rng(5)
    n_eval = min(len(hu), 400_000)
    sel = rng.choice(len(hu), n_eval, replace=False)
    Wv, Wc = m.get_rows(0, dtype=np.float32), m.get_rows(1, dtype=np.float32)
    u, ipos = hu[sel], hi[sel]
    ineg = item_ids[rng.integers(0, len(item_ids), n_eval)]
    sp = np.einsum("ij,ij->i", Wv[u], Wc[ipos])
    sn = np.einsum("ij,ij->i", Wv[u], Wc[ineg])
    auc = float((sp > sn).mean() + 0.5 * (sp == sn).mean())
    # the head of the catalogue is where the L1 copies live: AUC restricted to held-out positives among the 100 most popular items
    pop = np.bincount(col, minlength=g.V)
    head = np.argsort(-pop)[:100]
    hm = np.isin(ipos, head)
    auc_head = float((sp[hm] > sn[hm]).mean()) if hm.any() else None
    print(json.dumps({"l1_gather": os.environ.get("SMORE_L1_GATHER", "auto"), "users": a.users, "items": a.items,
                      "train_edges": int(len(ts)), "dim": a.dim, "zipf": a.zipf, "samples": st["samples"],
                      "M_samples_per_s": st["samples"] / st["kernel_ms"] / 1e3, "auc": auc, "auc_head100": auc_head,
                      "held_out_head100": int(hm.sum()), "wall_s": time.time() - t0}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--users", type=int, default=1_000_000)
    ap.add_argument("--items", type=int, default=200_000)
    ap.add_argument("--edges", type=int, default=20_000_000)
    ap.add_argument("--clusters", type=int, default=200)
    ap.add_argument("--zipf", type=float, default=1.0)
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--total", type=int, default=1 << 28)
    ap.add_argument("--child", action="store_true")
    a = ap.parse_args()
    if a.child:
        return run(a)
    for flag in ("1", "0"):
        env = dict(os.environ, SMORE_L1_GATHER=flag)
        subprocess.run([sys.executable, os.path.abspath(__file__), "--child"] + sys.argv[1:], env=env, check=True)


if __name__ == "__main__":
    main()
