"""Shared by the link-prediction quality gates (TEST INFRASTRUCTURE): the planted-structure problems, the held-out
evaluation, and nothing else. Used by tests/test_zz_gpu_quality_gates.py, tests/golden/make_quality_baselines_v2.py (the
CPU reference side, precomputed) and tools/ab_sharded_quality.py."""
import numpy as np

from smore_b200 import synth


def sbm_graph(n_comm, comm_size, deg, p_in, seed):
    """Planted-partition graph: link structure an embedding can actually learn."""
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


def sbm_problem(n_comm=150, comm_size=80, deg=24, seed=5):
    """-> (row_off, col, w, test_src, test_dst): 90 % of the edges as an undirected CSR, 10 % held out (vertex ids)."""
    src, dst, w = sbm_graph(n_comm=n_comm, comm_size=comm_size, deg=deg, p_in=0.85, seed=seed)
    (ts, td, tw), (hs, hd, _) = synth.split_edges(src, dst, w, 0.10, seed=seed + 1)
    off, col, ww, labels = synth.csr_from_edges(ts, td, tw, True)
    lab2id = np.full(int(labels.max()) + 1, -1, dtype=np.int64)
    lab2id[labels] = np.arange(len(labels))
    hs, hd = hs[(hs <= labels.max()) & (hd <= labels.max())], hd[(hs <= labels.max()) & (hd <= labels.max())]
    ok = (lab2id[hs] >= 0) & (lab2id[hd] >= 0)
    return off, col, ww, lab2id[hs[ok]], lab2id[hd[ok]]


def bipartite_problem(n_comm=100, users_per=40, items_per=20, deg=20, p_in=0.8, seed=11, undirected=False, with_labels=False):
    """Planted preferences: the users of community c pick items of community c with probability p_in, any item otherwise.
    -> (row_off, col, w, test_user, test_item, is_item[V], field[V]) in vertex ids; 10 % of the interactions held out."""
    rng = np.random.default_rng(seed)
    nu, ni = n_comm * users_per, n_comm * items_per
    n_edges = nu * deg
    u = rng.integers(0, nu, n_edges)
    inside = rng.random(n_edges) < p_in
    it_in = (u // users_per) * items_per + rng.integers(0, items_per, n_edges)
    it_out = rng.integers(0, ni, n_edges)
    it = nu + np.where(inside, it_in, it_out)
    w = rng.integers(1, 6, n_edges).astype(np.float64)
    (ts, td, tw), (hs, hd, _) = synth.split_edges(u, it, w, 0.10, seed=seed + 1)
    off, col, ww, labels = synth.csr_from_edges(ts, td, tw, undirected)
    lab2id = np.full(nu + ni, -1, dtype=np.int64)
    lab2id[labels] = np.arange(len(labels))
    ok = (lab2id[hs] >= 0) & (lab2id[hd] >= 0)
    is_item = labels >= nu
    if with_labels:
        return off, col, ww, lab2id[hs[ok]], lab2id[hd[ok]], is_item, is_item.astype(np.int32), labels
    return off, col, ww, lab2id[hs[ok]], lab2id[hd[ok]], is_item, is_item.astype(np.int32)


def two_graph_problem(kind, n_comm=100, users_per=40, items_per=20, seed=11):
    """The planted-preference problem (undirected, as cmd/cpr and cmd/tpr load it) plus the SECOND graph of the Go tree's
    two-graph models, as bare adjacency over vids that extend the first graph's (shared entities keep their vids, which is
    what cpr.go:148-169 / tpr.go:108 rely on):
      "cpr": every user also has 6 source-domain items, 80 % from its community's 15 (aux vids V ..);
      "tpr": every item lists 4 words, 3 of its community's 5 topic words and 1 random word (aux vids V ..).
    -> (off, col, w, test_user, test_item, is_item, aux_off, aux_col)."""
    off, col, ww, tu, ti, is_item, _, labels = bipartite_problem(n_comm, users_per, items_per, seed=seed, undirected=True,
                                                                 with_labels=True)
    rng = np.random.default_rng(seed + 7)
    V, nu = len(labels), n_comm * users_per
    lists = [[] for _ in range(V)]
    if kind == "cpr":
        per, n_aux = 15, n_comm * 15
        for v in np.flatnonzero(~is_item):
            c = labels[v] // users_per
            inside = rng.random(6) < 0.8
            lists[v] = (V + np.where(inside, c * per + rng.integers(0, per, 6), rng.integers(0, n_aux, 6))).tolist()
    else:
        per, n_aux = 5, n_comm * 5
        for v in np.flatnonzero(is_item):
            c = (labels[v] - nu) // items_per
            lists[v] = (V + np.concatenate([c * per + rng.choice(per, 3, replace=False), rng.integers(0, n_aux, 1)])).tolist()
    aux_off = np.zeros(V + n_aux + 1, dtype=np.int64)
    aux_off[1:V + 1] = np.cumsum([len(x) for x in lists])
    aux_off[V + 1:] = aux_off[V]
    aux_col = np.array([x for lst in lists for x in lst], dtype=np.int32)
    return off, col, ww, tu, ti, is_item, aux_off, aux_col


def evaluate_two_graph(kind, U, I, A, off, col, aux_off, aux_col, test_u, test_i, is_item, text_weight=0.5, seed=2):
    """Held-out AUC under each model's own score: CPR ranks items by transformUser(u) . item (cpr.go:127-172, :225-229),
    TPR by user . text-enriched item (tpr.go:101-121, :186-190)."""
    U, I, A = (np.asarray(x, dtype=np.float64) for x in (U, I, A))
    V = len(off) - 1

    def row_sums(table, o, c, n_rows):
        out = np.zeros((n_rows, table.shape[1]))
        np.add.at(out, np.repeat(np.arange(n_rows), np.diff(o[:n_rows + 1])), table[c[:o[n_rows]]])
        return out

    if kind == "cpr":
        cnt = 1.0 + np.diff(off) + np.diff(aux_off[:V + 1])
        users = (U + row_sums(I, off, col, V) + row_sums(A, aux_off, aux_col, V)) / cnt[:, None]
        return evaluate_bipartite(users, I, test_u, test_i, is_item, seed)
    nw = np.diff(aux_off[:V + 1])
    enriched = np.where(nw[:, None] > 0, (1.0 - text_weight) * I + text_weight * row_sums(A, aux_off, aux_col, V) / np.maximum(nw, 1)[:, None], I)
    return evaluate_bipartite(U, enriched, test_u, test_i, is_item, seed)


def auc(pos, neg):
    s = np.concatenate([pos, neg])
    ranks = s.argsort().argsort().astype(np.float64) + 1
    return float((ranks[: len(pos)].sum() - len(pos) * (len(pos) + 1) / 2) / (len(pos) * len(neg)))


def evaluate_full(Wv, Wc, off, col, test_s, test_d, seed=2):
    """AUC over all held-out edges vs as many random pairs; recall@10 over EVERY held-out source (training neighbours and
    the source itself excluded from the ranking)."""
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
    return a, hits / tot


def evaluate_bipartite(Wu, Wi, test_u, test_i, is_item, seed=2):
    """AUC of held-out (user, item) pairs against (user, random item) pairs; Wu / Wi may be the same table."""
    rng = np.random.default_rng(seed)
    items = np.flatnonzero(is_item)
    Wu = np.asarray(Wu, dtype=np.float64)
    Wi = np.asarray(Wi, dtype=np.float64)
    pos = np.einsum("ij,ij->i", Wu[test_u], Wi[test_i])
    neg = np.einsum("ij,ij->i", Wu[test_u], Wi[items[rng.integers(0, len(items), len(test_u))]])
    return auc(pos, neg)


def evaluate_sampled(Wv, Wc, test_s, test_d, train_adj, rng):
    """The round-1 evaluation (AUC over all held-out edges, recall@10 over the first 1500 held-out sources): what
    tests/golden/quality_baselines_v1.json was computed with."""
    pos = np.einsum("ij,ij->i", Wv[test_s], Wc[test_d])
    neg = np.einsum("ij,ij->i", Wv[test_s], Wc[rng.integers(0, len(Wc), len(test_s))])
    a = auc(pos, neg)
    hits = tot = 0
    held = {}
    for s, d in zip(test_s.tolist(), test_d.tolist()):
        held.setdefault(s, set()).add(d)
    for s in list(held)[:1500]:
        sc = Wc @ Wv[s]
        sc[list(train_adj.get(s, ()))] = -np.inf
        sc[s] = -np.inf
        top = np.argpartition(-sc, 10)[:10]
        hits += len(held[s].intersection(top.tolist()))
        tot += min(len(held[s]), 10)
    return a, hits / tot
