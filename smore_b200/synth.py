"""Synthetic graphs of the BASELINE.json shapes (seeded, numpy only) and reference-order CSR construction.

Used by bench.py and the full-size property tests. All generators return *edge lists over labels*; `csr_from_edges`
then applies the reference's ingest rules (ids by first appearance, src before dst; adjacency in file order; reverse
entry right after the forward one when undirected -- src/proNet.cpp:178-215, pronet.go:145-155).
"""
from __future__ import annotations

import numpy as np


def power_law_edges(n_vertices: int, n_edges: int, seed: int, gamma: float = 2.5, max_w: int = 5):
    """Chung-Lu style: both endpoints drawn with P(i) ~ (i+1)^(-1/(gamma-1)); integer weights U{1..max_w}."""
    rng = np.random.default_rng(seed)
    p = np.arange(1, n_vertices + 1, dtype=np.float64) ** (-1.0 / (gamma - 1.0))
    cdf = np.cumsum(p)
    cdf /= cdf[-1]
    perm = rng.permutation(n_vertices)  # hot vertices are not the low labels
    src = perm[np.searchsorted(cdf, rng.random(n_edges), side="right").clip(max=n_vertices - 1)]
    dst = perm[np.searchsorted(cdf, rng.random(n_edges), side="right").clip(max=n_vertices - 1)]
    w = rng.integers(1, max_w + 1, size=n_edges).astype(np.float64)
    return src.astype(np.int64), dst.astype(np.int64), w


def bipartite_edges(n_users: int, n_items: int, n_edges: int, seed: int, zipf_s: float = 1.0, max_w: int = 5):
    """users uniform, items Zipf(s) over a random permutation; labels: users [0,n_users), items [n_users, ...)."""
    rng = np.random.default_rng(seed)
    p = np.arange(1, n_items + 1, dtype=np.float64) ** (-zipf_s)
    cdf = np.cumsum(p)
    cdf /= cdf[-1]
    perm = rng.permutation(n_items)
    src = rng.integers(0, n_users, size=n_edges)
    dst = n_users + perm[np.searchsorted(cdf, rng.random(n_edges), side="right").clip(max=n_items - 1)]
    w = rng.integers(1, max_w + 1, size=n_edges).astype(np.float64)
    return src.astype(np.int64), dst.astype(np.int64), w


def first_appearance_ids(src, dst):
    """Reference vertex ids: order of first appearance scanning src0, dst0, src1, dst1, ..."""
    inter = np.empty(2 * len(src), dtype=np.int64)
    inter[0::2] = src
    inter[1::2] = dst
    labels, first = np.unique(inter, return_index=True)
    order = np.argsort(first, kind="stable")
    new_id = np.empty(len(labels), dtype=np.int64)
    new_id[order] = np.arange(len(labels))
    pos = np.searchsorted(labels, inter)
    ids = new_id[pos]
    return ids[0::2], ids[1::2], labels[order]


def csr_from_edges(src, dst, w, undirected: bool):
    """(row_off int64[V+1], col int32[E], w float64[E], labels_by_id) in the reference's insertion order."""
    s, d, labels = first_appearance_ids(np.asarray(src), np.asarray(dst))
    w = np.asarray(w, dtype=np.float64)
    V = len(labels)
    if undirected:
        es = np.empty(2 * len(s), dtype=np.int64)
        ed = np.empty(2 * len(s), dtype=np.int64)
        ew = np.empty(2 * len(s), dtype=np.float64)
        es[0::2], es[1::2] = s, d
        ed[0::2], ed[1::2] = d, s
        ew[0::2], ew[1::2] = w, w
    else:
        es, ed, ew = s, d, w
    order = np.argsort(es, kind="stable")
    col = ed[order].astype(np.int32)
    ww = ew[order]
    row_off = np.zeros(V + 1, dtype=np.int64)
    row_off[1:] = np.cumsum(np.bincount(es, minlength=V))
    return row_off, col, ww, labels


def split_edges(src, dst, w, holdout: float, seed: int):
    """Random train / held-out split for the link-prediction quality gate."""
    rng = np.random.default_rng(seed)
    mask = rng.random(len(src)) < holdout
    return (src[~mask], dst[~mask], w[~mask]), (src[mask], dst[mask], w[mask])
