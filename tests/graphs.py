"""Small deterministic graphs shared by the CPU and GPU parity tests (TEST INFRASTRUCTURE)."""
import numpy as np


def random_graph(n_vertices, n_edges, seed, max_w=5, zipf=True):
    rng = np.random.RandomState(seed)
    if zipf:
        p = 1.0 / np.arange(1, n_vertices + 1)
        p /= p.sum()
        src = rng.choice(n_vertices, size=n_edges, p=p)
        dst = rng.choice(n_vertices, size=n_edges, p=p)
    else:
        src = rng.randint(0, n_vertices, size=n_edges)
        dst = rng.randint(0, n_vertices, size=n_edges)
    w = rng.randint(1, max_w + 1, size=n_edges).astype(np.float64)
    return src, dst, w


def bipartite_graph(n_users, n_items, n_edges, seed, max_w=5):
    """users are labels [0, n_users), items [n_users, n_users+n_items); directed user -> item."""
    rng = np.random.RandomState(seed)
    p = 1.0 / np.arange(1, n_items + 1)
    p /= p.sum()
    src = rng.randint(0, n_users, size=n_edges)
    dst = n_users + rng.choice(n_items, size=n_edges, p=p)
    w = rng.randint(1, max_w + 1, size=n_edges).astype(np.float64)
    return src, dst, w


def readme_graph():
    """The 5-edge example network of the reference README (README.md:50-56)."""
    names = {"userA": 0, "itemA": 1, "itemC": 2, "userB": 3, "itemB": 4, "userC": 5}
    edges = [("userA", "itemA", 3), ("userA", "itemC", 5), ("userB", "itemA", 1), ("userB", "itemB", 5),
             ("userC", "itemA", 4)]
    src = [names[a] for a, _, _ in edges]
    dst = [names[b] for _, b, _ in edges]
    w = [float(x) for _, _, x in edges]
    return np.array(src), np.array(dst), np.array(w)


def init_tables(V, dim, seed, context_zero=False):
    rng = np.random.RandomState(seed)
    Wv = (rng.random_sample((V, dim)) - 0.5) / dim
    Wc = np.zeros((V, dim)) if context_zero else (rng.random_sample((V, dim)) - 0.5) / dim
    return Wv, Wc
