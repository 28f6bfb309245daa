#!/usr/bin/env python
"""A/B matrix for the row-sharded LINE modes on ONE device (shards emulated in-process, as tests/test_gpu_sharded.py does).

Question: which ingredient of the sharded peer-access mode costs link-prediction quality --
  * the write-back of the vertex row (full-row store of a staged copy vs red.global.add of the delta), SMORE_SHARD_VRED
  * shard-local negatives vs negatives over all vertices, SMORE_SHARD_GLOBAL_NEG
  * the world size (2 / 4 / 8)
Every cell is averaged over --seeds Hogwild runs; AUC and recall@10 are evaluated on ALL held-out edges (the unit tests'
1500-source subsample has a standard error of ~5 % relative on recall@10, as large as the effect under study).

Output: one JSON line per cell on stdout (and a table on stderr).
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from smore_b200 import capi, synth  # noqa: E402
from smore_b200 import dist as sdist  # noqa: E402


def sbm_graph(n_comm, comm_size, deg, p_in, seed):
    rng = np.random.default_rng(seed)
    V = n_comm * comm_size
    n_edges = V * deg // 2
    src = rng.integers(0, V, n_edges)
    inside = rng.random(n_edges) < p_in
    dst_in = (src // comm_size) * comm_size + rng.integers(0, comm_size, n_edges)
    dst_out = rng.integers(0, V, n_edges)
    dst = np.where(inside, dst_in, dst_out)
    keep = src != dst
    w = rng.integers(1, 4, n_edges).astype(np.float64)
    return src[keep], dst[keep], w[keep]


def sbm_powerlaw_graph(n_comm, comm_size, deg, p_in, seed, expo=0.8):
    """Planted partition with heterogeneous degrees: both endpoints are drawn by a fitness ~ rank^-expo (random ranks), so
    most vertices have a handful of neighbours (fewer than the number of shards) and a few hubs have thousands."""
    rng = np.random.default_rng(seed)
    V = n_comm * comm_size
    n_edges = V * deg // 2
    fit = (np.arange(1, V + 1, dtype=np.float64) ** -expo)[rng.permutation(V)]
    cdf = np.cumsum(fit) / fit.sum()
    src = np.searchsorted(cdf, rng.random(n_edges)).clip(max=V - 1)
    inside = rng.random(n_edges) < p_in
    # inside a community: fitness-weighted choice among its members
    f2 = fit.reshape(n_comm, comm_size)
    ccdf = np.cumsum(f2, axis=1) / f2.sum(axis=1, keepdims=True)
    comm = src // comm_size
    u = rng.random(n_edges)
    dst_in = comm * comm_size + (ccdf[comm] < u[:, None]).sum(axis=1).clip(max=comm_size - 1)
    dst_out = np.searchsorted(cdf, rng.random(n_edges)).clip(max=V - 1)
    dst = np.where(inside, dst_in, dst_out)
    keep = src != dst
    w = rng.integers(1, 4, n_edges).astype(np.float64)
    return src[keep], dst[keep], w[keep]


def make_problem(n_comm=150, comm_size=80, deg=24, powerlaw=False):
    if powerlaw:
        src, dst, w = sbm_powerlaw_graph(n_comm=n_comm, comm_size=comm_size, deg=deg, p_in=0.85, seed=5)
    else:
        src, dst, w = sbm_graph(n_comm=n_comm, comm_size=comm_size, deg=deg, p_in=0.85, seed=5)
    (ts, td, tw), (hs, hd, _) = synth.split_edges(src, dst, w, 0.10, seed=6)
    off, col, ww, labels = synth.csr_from_edges(ts, td, tw, True)
    lab2id = {int(l): i for i, l in enumerate(labels)}
    ok = np.array([(int(a) in lab2id) and (int(b) in lab2id) for a, b in zip(hs, hd)])
    test_s = np.array([lab2id[int(a)] for a in hs[ok]])
    test_d = np.array([lab2id[int(b)] for b in hd[ok]])
    return off, col, ww, test_s, test_d


def auc(pos, neg):
    s = np.concatenate([pos, neg])
    ranks = s.argsort().argsort().astype(np.float64) + 1
    return (ranks[: len(pos)].sum() - len(pos) * (len(pos) + 1) / 2) / (len(pos) * len(neg))


def evaluate_full(Wv, Wc, off, col, test_s, test_d, seed=2):
    """AUC over all held-out edges vs as many random pairs; recall@10 over EVERY held-out source."""
    rng = np.random.default_rng(seed)
    Wv = np.asarray(Wv, dtype=np.float32)
    Wc = np.asarray(Wc, dtype=np.float32)
    pos = np.einsum("ij,ij->i", Wv[test_s], Wc[test_d])
    neg = np.einsum("ij,ij->i", Wv[test_s], Wc[rng.integers(0, len(Wc), len(test_s))])
    a = auc(pos, neg)
    srcs = np.unique(test_s)
    hits = tot = 0
    order = np.argsort(test_s, kind="stable")
    ts, td = test_s[order], test_d[order]
    starts = np.searchsorted(ts, srcs)
    ends = np.searchsorted(ts, srcs, side="right")
    for lo in range(0, len(srcs), 2048):
        blk = srcs[lo:lo + 2048]
        sc = Wv[blk] @ Wc.T
        for i, s in enumerate(blk):
            sc[i, col[off[s]:off[s + 1]]] = -np.inf
            sc[i, s] = -np.inf
        top = np.argpartition(-sc, 10, axis=1)[:, :10]
        for i in range(len(blk)):
            held = td[starts[lo + i]:ends[lo + i]]
            hits += len(np.intersect1d(held, top[i]))
            tot += min(len(np.unique(held)), 10)
    return float(a), hits / tot


def params(total, seed, max_warps):
    p = capi.default_params()
    p.semantics, p.mode, p.seed, p.alpha, p.total, p.negative_samples = capi.SEM_CPP, capi.MODE_HOGWILD, seed, 0.025, total, 5
    p.max_warps = max_warps
    return p


def run_unsharded(prob, dim, total, seed, max_warps):
    off, col, ww, test_s, test_d = prob
    V = len(off) - 1
    init = ((np.random.default_rng(1).random((V, dim)) - 0.5) / dim)
    g = capi.Graph.from_csr(off, col, ww)
    m = capi.Model(g, dim, 2, capi.F32)
    m.set_rows(0, init), m.set_rows(1, np.zeros((V, dim)))
    m.train_line(params(total, seed, max_warps))
    return evaluate_full(m.get_rows(0), m.get_rows(1), off, col, test_s, test_d)


def run_rotating(prob, dim, total, seed, world, cycles, max_warps, neg_mode):
    """Rotating shards, all shards of the ring in this process: `cycles` cycles of 2*world episodes, one LR schedule."""
    off, col, ww, test_s, test_d = prob
    V = len(off) - 1
    init = ((np.random.default_rng(1).random((V, dim)) - 0.5) / dim)
    ms = []
    for r in range(world):
        gr = capi.Graph.from_csr(off, col, ww)
        gr.set_shard_rotating(r, world)
        mr = capi.Model(gr, dim, 2, capi.F32)
        rows = sdist.owned_rows(V, r, world)
        mr.set_rows(0, init[rows]), mr.set_rows(1, np.zeros((len(rows), dim)))
        mr.enable_rotation()
        ms.append(mr)
    sdist.connect_rotation_local(ms)
    episodes = cycles * 2 * world
    p = params(total // episodes, seed, max_warps)
    p.sched_total, p.sched_offset, p.neg_mode = total, 0, neg_mode
    done, _ = sdist.train_line_rotating(ms, p, episodes)
    assert 0.85 * total <= sum(done) <= 1.01 * total, (sum(done), total)
    Wv, Wc = np.zeros((V, dim)), np.zeros((V, dim))
    for r, mr in enumerate(ms):
        assert mr.rot_position()["at_home"]
        rows = sdist.owned_rows(V, r, world)
        Wv[rows], Wc[rows] = mr.get_rows(0), mr.get_rows(1)
    return evaluate_full(Wv, Wc, off, col, test_s, test_d)


def run_peer(prob, dim, total, seed, world, rounds, max_warps, neg_mode=capi.PAIRING_COUPLED):
    """Peer-access mode, ranks launched one after the other in `rounds` rounds (one LR schedule over all of them)."""
    off, col, ww, test_s, test_d = prob
    V = len(off) - 1
    init = ((np.random.default_rng(1).random((V, dim)) - 0.5) / dim)
    ms = []
    for r in range(world):
        gr = capi.Graph.from_csr(off, col, ww)
        gr.set_shard(r, world)
        mr = capi.Model(gr, dim, 2, capi.F32)
        rows = sdist.owned_rows(V, r, world)
        mr.set_rows(0, init[rows]), mr.set_rows(1, np.zeros((len(rows), dim)))
        ms.append(mr)
    for t in range(2):
        ptrs = [mr.device_ptr(t) for mr in ms]
        for mr in ms:
            mr.set_peer_ptrs(t, ptrs)
    for k in range(rounds):
        for r, mr in enumerate(ms):
            p = params(total // rounds, seed * 1000 + k, max_warps)
            p.neg_mode = neg_mode
            p.stream_base = r * (1 << 20)
            p.sched_total, p.sched_offset = total, k * (total // rounds)
            mr.train_line(p)
    Wv, Wc = np.zeros((V, dim)), np.zeros((V, dim))
    for r, mr in enumerate(ms):
        rows = sdist.owned_rows(V, r, world)
        Wv[rows], Wc[rows] = mr.get_rows(0), mr.get_rows(1)
    return evaluate_full(Wv, Wc, off, col, test_s, test_d)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, default=3)
    ap.add_argument("--dim", type=int, default=32)
    ap.add_argument("--total", type=int, default=12_000_000)
    ap.add_argument("--rounds", type=int, default=20)
    ap.add_argument("--max-warps", type=int, default=512)
    ap.add_argument("--worlds", default="2,4,8")
    ap.add_argument("--cycles", default="5,20")
    ap.add_argument("--what", default="occupancy,peer,rotating", help="comma list of: occupancy, peer-r2a, peer, rotating")
    ap.add_argument("--powerlaw", action="store_true", help="heterogeneous degrees (most vertices have fewer neighbours than shards)")
    ap.add_argument("--deg", type=int, default=24)
    args = ap.parse_args()
    prob = make_problem(deg=args.deg, powerlaw=args.powerlaw)
    if args.powerlaw:
        d = np.diff(prob[0])
        print(f"power-law SBM: V={len(d)} entries={d.sum()} median degree {np.median(d):.0f}, "
              f"{(d <= 4).mean() * 100:.0f}% of the vertices have <= 4 neighbours, max {d.max()}", file=sys.stderr)
    cells = []

    def cell(name, fn):
        res = np.array([fn(s) for s in range(1, args.seeds + 1)])
        rec = {"cell": name, "auc_mean": res[:, 0].mean(), "auc_sd": res[:, 0].std(), "rec_mean": res[:, 1].mean(),
               "rec_sd": res[:, 1].std(), "runs": res.tolist()}
        cells.append(rec)
        print(json.dumps(rec), flush=True)
        print(f"{name:60s} AUC {rec['auc_mean']:.4f} +- {rec['auc_sd']:.4f}   recall@10 {rec['rec_mean']:.4f} +- {rec['rec_sd']:.4f}",
              file=sys.stderr, flush=True)

    worlds = [int(x) for x in args.worlds.split(",")]
    what = set(args.what.split(","))
    if "occupancy" in what:
        # how many concurrent warps does a 12 k-vertex table tolerate? (round 2b, plain stores: AUC 0.92 up to 1024 warps,
        # 0.81 at full occupancy; fp32 rows are updated with red.global.add since then)
        for mw in (512, 2048, 0):
            cell(f"unsharded max_warps={mw}", lambda s: run_unsharded(prob, args.dim, args.total, 10 + s, mw))
    else:
        cell(f"unsharded max_warps={args.max_warps}", lambda s: run_unsharded(prob, args.dim, args.total, 10 + s, args.max_warps))
    if "peer-r2a" in what:  # the round-2a matrix: write-back and negative-table ingredients of the coupled scheme
        for world in worlds:
            for vred in (0, 1):
                for gneg in (0, 1):
                    os.environ["SMORE_SHARD_VRED"] = str(vred)
                    os.environ["SMORE_SHARD_GLOBAL_NEG"] = str(gneg)
                    cell(f"peer world={world} vred={vred} global_neg={gneg} coupled rounds={args.rounds}",
                         lambda s: run_peer(prob, args.dim, args.total, 100 + s, world, args.rounds, args.max_warps))
        os.environ.pop("SMORE_SHARD_VRED", None)
        os.environ.pop("SMORE_SHARD_GLOBAL_NEG", None)
    if "peer" in what:
        for world in worlds:
            for nm, name in ((capi.PAIRING_COUPLED, "coupled"), (capi.PAIRING_SPLIT, "split")):
                cell(f"peer world={world} {name} rounds={args.rounds}",
                     lambda s: run_peer(prob, args.dim, args.total, 100 + s, world, args.rounds, args.max_warps, nm))
    if "rotating" in what:
        for world in worlds:
            for nm, name in ((capi.PAIRING_COUPLED, "coupled"), (capi.PAIRING_SPLIT, "split")):
                for cycles in [int(x) for x in args.cycles.split(",")]:
                    cell(f"rotating world={world} {name} cycles={cycles}",
                         lambda s: run_rotating(prob, args.dim, args.total, 200 + s, world, cycles, args.max_warps, nm))


if __name__ == "__main__":
    main()
