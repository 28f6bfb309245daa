"""GPU: device-side construction (smore_b200/csrc/device_graph.cu) -- the parallel alias build and the synthetic
power-law graph generated straight into the rotating-shard tables, each against an independent numpy restatement."""
import numpy as np
import pytest

from smore_b200 import capi
from smore_b200 import dist as sdist
from tests import synth_ref

pytestmark = pytest.mark.gpu


def _implied_distribution(thr, alias):
    """P(i) the sampler realises: pick bucket j uniformly, keep j if k < thr[j] (k uniform 32-bit), else alias[j]."""
    n = len(thr)
    keep = thr.astype(np.float64) / 4294967296.0  # (a saturated bucket aliases itself: the remainder comes back to it)
    p = keep.copy()
    np.add.at(p, alias.astype(np.int64), 1.0 - keep)
    return p / n


@pytest.mark.parametrize("n,kind", [(1, "flat"), (2, "skew"), (7, "zeros"), (1000, "skew"), (4096, "flat"), (100_003, "zipf"),
                                    (3_000_000, "zipf"), (50_000, "onehot")])
def test_device_alias_table_encodes_the_distribution(n, kind):
    rng = np.random.default_rng(n)
    if kind == "flat":
        w = np.ones(n)
    elif kind == "skew":
        w = rng.random(n) ** 4 + 1e-3
    elif kind == "zeros":
        w = np.array([0.0, 3.0, 0.0, 1.0, 0.0, 0.0, 2.5])
    elif kind == "onehot":
        w = np.zeros(n)
        w[n // 3] = 5.0
    else:
        w = (np.arange(1, n + 1, dtype=np.float64) ** -1.0)[rng.permutation(n)]
    thr, alias = capi.alias_build_device(w)
    assert (alias < n).all()
    p = _implied_distribution(thr, alias)
    want = w / w.sum()
    # every bucket is quantised to 2^-32: an item that is the alias of c buckets is off by at most (c + 1) * 2^-32 / n
    cnt = np.bincount(alias.astype(np.int64), minlength=n)
    assert (np.abs(p - want) <= (cnt + 2) * 2.0 ** -32 / n + 1e-13).all(), np.abs(p - want).max()
    assert abs(p.sum() - 1.0) < 1e-9


def test_device_alias_all_zero_weights_fall_back_to_uniform():
    thr, alias = capi.alias_build_device(np.zeros(100))
    assert (thr == 0xFFFFFFFF).all() and np.array_equal(alias, np.arange(100, dtype=np.uint32))


def _blocks_numpy(V, E_lines, seed, rank, world):
    """Block sizes and masses of `rank` under C++ semantics, from the numpy restatement of the generator."""
    u, v, w = synth_ref.synth_edges(V, E_lines, seed)
    es, ed, ew = np.concatenate([u, v]), np.concatenate([v, u]), np.concatenate([w, w])
    out_deg = np.bincount(es, weights=ew, minlength=V)
    nrm = np.bincount(es, weights=ew ** 0.75, minlength=V)
    psrc = np.where(out_deg > 0, out_deg ** 0.75, 0.0)
    pe = psrc[es] / psrc.sum() / nrm[es] * ew ** 0.75
    shift = world.bit_length() - 1
    sub_cap = (-(-V // world) + 1) // 2
    loc = es >> shift
    q = 2 * (es & (world - 1)) + (loc >= sub_cap)
    mine = (ed & (world - 1)) == rank
    return (np.bincount(q[mine], minlength=2 * world), np.bincount(q[mine], weights=pe[mine], minlength=2 * world), (es, ed))


def test_synthetic_graph_matches_the_numpy_restatement():
    V, E, seed = 5003, 60_000, 20261018
    for world in (2, 4):
        tot = 0.0
        for r in range(world):
            g = capi.Graph.synthetic_rotating(V, E, seed, r, world)
            ri, info = g.rotation_info(), g.shard_info()
            cnt, mass, _ = _blocks_numpy(V, E, seed, r, world)
            assert np.array_equal(ri["block_edges"], cnt)  # the same edge list, entry for entry
            assert np.allclose(ri["block_mass"], mass, rtol=1e-9, atol=1e-15)
            assert info["n_local"] == len(sdist.owned_rows(V, r, world))
            tot += ri["block_mass"].sum()
        assert abs(tot - 1.0) < 1e-9
    with pytest.raises(capi.SmoreError):
        g.csr()  # no host CSR behind a device-generated graph
    with pytest.raises(capi.SmoreError):
        g.sample(capi.SAMPLE_SOURCE, 1, 0, 10)


def test_synthetic_blocks_train_rows_of_real_neighbours():
    """The K = 0 / one-source trick of test_gpu_rotation on the device-generated tables: a context fed by a single source
    must have been fed by one of its neighbours in the numpy restatement of the graph."""
    V, E, seed, dim = 6000, 40_000, 77, 32
    _, _, (es, ed) = _blocks_numpy(V, E, seed, 0, 2)
    adj = {}
    for a, b in zip(es.tolist(), ed.tolist()):
        adj.setdefault(b, set()).add(a)
    ids = np.arange(V) + 1.0
    init_v = np.full((V, dim), 1e-4)
    init_v[:, 0], init_v[:, 2] = ids * 1e-4, ids * ids * 1e-8
    for world in (2, 4):
        ms = []
        for r in range(world):
            g = capi.Graph.synthetic_rotating(V, E, seed, r, world)
            m = capi.Model(g, dim, 2, capi.F64)
            rows = sdist.owned_rows(V, r, world)
            m.set_rows(0, init_v[rows]), m.set_rows(1, np.zeros((len(rows), dim)))
            m.enable_rotation()
            ms.append(m)
        sdist.connect_rotation_local(ms)
        episodes = 2 * world
        p = capi.default_params()
        p.semantics, p.mode, p.seed, p.alpha, p.negative_samples, p.max_warps = capi.SEM_CPP, capi.MODE_HOGWILD, 9, 1e-6, 0, 64
        p.total = 60_000 // episodes
        done, _ = sdist.train_line_rotating(ms, p, episodes)
        assert sum(done) > 40_000
        Wc = np.zeros((V, dim))
        for r, m in enumerate(ms):
            Wc[sdist.owned_rows(V, r, world)] = m.get_rows(1)
        hit = np.flatnonzero(Wc[:, 1] > 0)
        mean = Wc[hit, 0] / Wc[hit, 1]
        var = Wc[hit, 2] / Wc[hit, 1] * 1e4 - mean * mean
        single = np.abs(var) < 1e-3  # (two sources one id apart with weights 1 : 99 already give 0.0099)
        x = np.round(mean[single] - 1).astype(int)
        good = np.array([xi in adj.get(c, ()) for c, xi in zip(hit[single].tolist(), x.tolist())])
        assert single.sum() > 300 and good.mean() > 0.995, (world, single.sum(), good.mean())
