"""GPU quality gates (BASELINE.json north_star: "Hogwild mode matches downstream link-prediction AUC / recall@10 within
0.5 %"), collected LAST: every test here compares a Hogwild fp32 run -- different draw streams, thousands of racing warps --
with the reference CPU path, so they are statistical; the deterministic parity tests must not hide behind them under -x.

Reference side: the oracle restatement (pinned bit-exactly against the compiled reference, tests/test_oracle_vs_ref.py) on
one stream, precomputed by tests/golden/make_quality_baselines_v2.py on the problems of tests/quality.py; the large-graph
gate runs the oracle's OpenMP Hogwild loop on the box's host cores instead.
Tolerance: 0.005 absolute (0.5 points) on AUC and on recall@10, as the north-star states it. Trainers run at the library's
own occupancy policy (max_warps = 0: every resident warp, capped at one warp per 8 table rows -- host_common.h) unless a
test says otherwise."""
import json
import os

import numpy as np
import pytest

from oracle import bindings as B
from smore_b200 import capi
from smore_b200 import dist as sdist
from tests import quality as Q

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(__file__)
Q2 = json.load(open(os.path.join(HERE, "golden", "quality_baselines_v2.json")))
DIM, TOL = 32, 0.005
_cache = {}


def sbm():
    if "sbm" not in _cache:
        _cache["sbm"] = Q.sbm_problem()
    return _cache["sbm"]


def bip(undirected=False):
    key = ("bip", undirected)
    if key not in _cache:
        _cache[key] = Q.bipartite_problem(undirected=undirected)
    return _cache[key]


def init_tables(V, context_zero=False):
    Wv = (np.random.default_rng(1).random((V, DIM)) - 0.5) / DIM
    Wc = np.zeros((V, DIM)) if context_zero else (np.random.default_rng(3).random((V, DIM)) - 0.5) / DIM
    return Wv, Wc


def hogwild(sem=capi.SEM_CPP, seed=13, **kw):
    p = capi.default_params()
    p.semantics, p.mode, p.seed, p.alpha = sem, capi.MODE_HOGWILD, seed, 0.025
    for k, v in kw.items():
        setattr(p, k, v)
    return p


SEEDS = (13, 14, 15)  # single-model gates: mean of 3 Hogwild runs (one run's recall@10 has a standard deviation of ~0.003)


def mean_of(run, seeds=SEEDS):
    """run(seed) -> (auc, recall) or auc; returns the mean over `seeds`."""
    return np.mean(np.array([run(sd) for sd in seeds], dtype=np.float64), axis=0)


def check(name, got, want, what="AUC"):
    print(f"{name}: {what} gpu {got:.4f} reference {want:.4f}")
    assert abs(got - want) < TOL, (name, what, got, want)


# ---- graph models on the planted-partition graph --------------------------------------------------------------------
@pytest.mark.parametrize("max_warps", [0, 512])
def test_line_cpp(max_warps):
    off, col, ww, ts, td = sbm()
    V = len(off) - 1
    ref = Q2["models"]["line_cpp"]
    Wv, Wc = init_tables(V, context_zero=True)
    g = capi.Graph.from_csr(off, col, ww)

    def run(seed):
        m = capi.Model(g, DIM, 2, capi.F32)
        m.set_rows(0, Wv), m.set_rows(1, Wc)
        m.train_line(hogwild(seed=seed, total=ref["total"], negative_samples=5, max_warps=max_warps))
        return Q.evaluate_full(m.get_rows(0), m.get_rows(1), off, col, ts, td)

    a, r = mean_of(run)
    check(f"LINE C++ (max_warps={max_warps})", a, ref["auc"])
    check(f"LINE C++ (max_warps={max_warps})", r, ref["recall_at_10"], "recall@10")


def test_line_go():
    off, col, ww, ts, td = sbm()
    V = len(off) - 1
    ref = Q2["models"]["line_go"]
    Wv, Wc = init_tables(V)
    g = capi.Graph.from_csr(off, col, ww, semantics=capi.SEM_GO, n_lines=len(col) // 2)

    def run(seed):
        m = capi.Model(g, DIM, 2, capi.F32)
        m.set_rows(0, Wv), m.set_rows(1, Wc)
        m.train_line(hogwild(capi.SEM_GO, seed=seed, total=ref["total"], negative_samples=5))
        return Q.evaluate_full(m.get_rows(0), m.get_rows(1), off, col, ts, td)

    a, r = mean_of(run)
    check("LINE Go", a, ref["auc"])
    check("LINE Go", r, ref["recall_at_10"], "recall@10")


def test_deepwalk_cpp():
    off, col, ww, ts, td = sbm()
    V = len(off) - 1
    ref = Q2["models"]["deepwalk_cpp"]
    Wv, Wc = init_tables(V)
    m = capi.Model(capi.Graph.from_csr(off, col, ww), DIM, 2, capi.F32)
    m.set_rows(0, Wv), m.set_rows(1, Wc)
    st = m.train_deepwalk(hogwild(walk_times=ref["walk_times"], walk_steps=ref["walk_steps"], window_min=1, window_max=ref["window"],
                                  negative_samples=5))
    assert 0.9 * ref["pairs"] < st["pair_updates"] < 1.1 * ref["pairs"]
    a, r = Q.evaluate_full(m.get_rows(0), m.get_rows(1), off, col, ts, td)
    check("DeepWalk", a, ref["auc"])
    check("DeepWalk", r, ref["recall_at_10"], "recall@10")


@pytest.mark.parametrize("model", ["deepwalk_go", "node2vec_go", "node2vec_unbiased"])
def test_go_walk_models(model):
    """DeepWalk under Go semantics (fixed full window, random contexts) and node2vec (Go tree only): biased walks with
    p = 0.5 / q = 2 against their own reference-path baseline, and p = q = 1 -- where the biased walk is distributed like
    the plain one -- against DeepWalk's."""
    off, col, ww, ts, td = sbm()
    V = len(off) - 1
    ref = Q2["models"]["node2vec_go" if model == "node2vec_go" else "deepwalk_go"]
    Wv, Wc = init_tables(V)
    g = capi.Graph.from_csr(off, col, ww, semantics=capi.SEM_GO, n_lines=len(col) // 2)

    def run(seed):
        m = capi.Model(g, DIM, 2, capi.F32)
        m.set_rows(0, Wv), m.set_rows(1, Wc)
        p = hogwild(capi.SEM_GO, seed=seed, walk_times=ref["walk_times"], walk_steps=ref["walk_steps"], window_min=1,
                    window_max=ref["window"], negative_samples=5)
        if model == "deepwalk_go":
            st = m.train_deepwalk(p)
        else:
            p.n2v_p, p.n2v_q = (ref["p"], ref["q"]) if model == "node2vec_go" else (1.0, 1.0)
            st = m.train_node2vec(p)
        assert 0.9 * ref["pairs"] < st["pair_updates"] < 1.1 * ref["pairs"]
        return Q.evaluate_full(m.get_rows(0), m.get_rows(1), off, col, ts, td)

    # Both sides are means here: the walk models draw few, long, correlated sample groups (72 000 walks), and one run's
    # recall@10 has a standard deviation of 0.002-0.003 on the CPU path (8 seeds: 0.0996-0.1074, quality_baselines_v2.json)
    # and up to 0.004 on the GPU (profiles/r2m_go_walk_probe.jsonl). tests/probes/go_walk_streams_probe.py shows that the GPU's
    # numbers are reproduced by the oracle when it splits the walks over the same number of draw streams sequentially.
    a, r = mean_of(run, seeds=range(13, 19))
    check(model, a, ref["auc"])
    check(model, r, ref["recall_at_10"], "recall@10")


def test_hpe():
    off, col, ww, ts, td = sbm()
    V, ref = len(off) - 1, Q2["models"]["hpe"]
    Wv, Wc = init_tables(V)
    g = capi.Graph.from_csr(off, col, ww)

    def run(seed):
        m = capi.Model(g, DIM, 2, capi.F32)
        m.set_rows(0, Wv), m.set_rows(1, Wc)
        m.train_hpe(hogwild(seed=seed, total=ref["total"], walk_steps=ref["walk_steps"], negative_samples=ref["negative_samples"],
                            lambda_=ref["reg"]))
        return Q.evaluate_full(m.get_rows(0), m.get_rows(1), off, col, ts, td)

    a, r = mean_of(run)
    check("HPE", a, ref["auc"])
    check("HPE", r, ref["recall_at_10"], "recall@10")


def test_mf():
    off, col, ww, ts, td = sbm()
    V, ref = len(off) - 1, Q2["models"]["mf"]
    Wv, _ = init_tables(V)
    g = capi.Graph.from_csr(off, col, ww, negative_method=capi.NEG_NO_DEGREES)

    def run(seed):
        m = capi.Model(g, DIM, 1, capi.F32)
        m.set_rows(0, Wv)
        m.train_mf(hogwild(seed=seed, total=ref["total"], negative_samples=ref["negative_samples"], lambda_=ref["reg"]))
        W = m.get_rows(0)
        return Q.evaluate_full(W, W, off, col, ts, td)

    a, r = mean_of(run)  # (profiles/r2j_mf_vs_concurrency.txt: no trend from 64 warps to the policy's grid)
    check("MF", a, ref["auc"])
    check("MF", r, ref["recall_at_10"], "recall@10")


def test_skewopt():
    off, col, ww, ts, td = sbm()
    V, ref = len(off) - 1, Q2["models"]["skewopt"]
    Wv, _ = init_tables(V)
    g = capi.Graph.from_csr(off, col, ww, negative_method=capi.NEG_NO_DEGREES)

    def run(seed):
        m = capi.Model(g, DIM, 1, capi.F32)
        m.set_rows(0, Wv + ref["init_offset"])
        m.train_skewopt(hogwild(seed=seed, total=ref["total"], xi=ref["xi"], omega=ref["omega"], eta=ref["eta"]))
        W = m.get_rows(0)
        return Q.evaluate_full(W, W, off, col, ts, td)

    a, r = mean_of(run)
    check("Skew-OPT", a, ref["auc"])
    check("Skew-OPT", r, ref["recall_at_10"], "recall@10")


# ---- ranking models on the planted-preference graph -------------------------------------------------------------------
def test_bpr_go():
    off, col, ww, tu, ti, is_item, _ = bip()
    V, ref = len(off) - 1, Q2["models"]["bpr_go"]
    Wv, Wc = init_tables(V)
    g = capi.Graph.from_csr(off, col, ww, semantics=capi.SEM_GO, n_lines=len(col))
    m = capi.Model(g, DIM, 2, capi.F32)
    m.set_rows(0, Wv), m.set_rows(1, Wc)
    m.train_bpr(hogwild(capi.SEM_GO, total=ref["total"], lambda_=ref["lam"]))
    check("BPR Go", Q.evaluate_bipartite(m.get_rows(0), m.get_rows(1), tu, ti, is_item), ref["auc"])


@pytest.mark.parametrize("kind", ["cpr", "tpr"])
def test_two_graph_models(kind):
    """CPR / TPR (Go tree only): Hogwild fp32 at the library's occupancy against the one-stream reference path, each under
    its own score (transformed user . item / user . text-enriched item), cmd/cpr and cmd/tpr default parameters."""
    off, col, ww, tu, ti, is_item, aoff, acol = Q.two_graph_problem(kind)
    V, Va, ref = len(off) - 1, len(aoff) - 1, Q2["models"][kind]
    U0 = (np.random.default_rng(1).random((V, DIM)) - 0.5) / DIM
    I0 = (np.random.default_rng(3).random((V, DIM)) - 0.5) / DIM
    A0 = (np.random.default_rng(5).random((Va, DIM)) - 0.5) / DIM
    g = capi.Graph.from_csr(off, col, ww, semantics=capi.SEM_GO, n_lines=len(col) // 2)

    def run(seed):
        m = capi.Model(g, DIM, 2, capi.F32)
        m.set_rows(0, U0), m.set_rows(1, I0)
        m.attach_aux(aoff, acol, rows=A0)
        if kind == "cpr":
            p = hogwild(capi.SEM_GO, seed=seed, total=ref["total"], lambda_=ref["user_reg"], item_reg=ref["item_reg"], margin=ref["margin"])
            p.alpha = ref["alpha"]
            st = m.train_cpr(p)
        else:
            p = hogwild(capi.SEM_GO, seed=seed, total=ref["total"], lambda_=ref["lam"], text_weight=ref["text_weight"])
            p.alpha = ref["alpha"]
            st = m.train_tpr(p)
        assert st["samples"] > 0.99 * ref["total"]
        return Q.evaluate_two_graph(kind, m.get_rows(0), m.get_rows(1), m.get_aux_rows(), off, col, aoff, acol, tu, ti, is_item)

    check(kind.upper(), float(mean_of(run)), ref["auc"])


@pytest.mark.parametrize("name", ["bpr_cpp", "warp"])
def test_cpp_ranking(name):
    off, col, ww, tu, ti, is_item, _ = bip()
    V, ref = len(off) - 1, Q2["models"][name]
    Wv, _ = init_tables(V)
    g = capi.Graph.from_csr(off, col, ww, negative_method=capi.NEG_NO_DEGREES)
    m = capi.Model(g, DIM, 1, capi.F32)
    m.set_rows(0, Wv)
    st = (m.train_bpr if name == "bpr_cpp" else m.train_warp)(hogwild(total=ref["total"]))
    if name == "warp":
        print(f"WARP mean tries: gpu {st['mean_tries']:.3f} reference {ref['mean_tries']:.3f}")
    W = m.get_rows(0)
    check(name, Q.evaluate_bipartite(W, W, tu, ti, is_item), ref["auc"])


def test_hoprec():
    off, col, ww, tu, ti, is_item, field = bip(undirected=True)
    V, ref = len(off) - 1, Q2["models"]["hoprec"]
    Wv, _ = init_tables(V)
    g = capi.Graph.from_csr(off, col, ww, negative_method=capi.NEG_NO_DEGREES)
    g.set_field(field)
    m = capi.Model(g, DIM, 1, capi.F32)
    m.set_rows(0, Wv)
    m.train_hoprec(hogwild(total=ref["total"], walk_steps=ref["walk_steps"]))
    W = m.get_rows(0)
    check("HOP-Rec", Q.evaluate_bipartite(W, W, tu, ti, is_item), ref["auc"])


# ---- the row-sharded modes: world 2 / 4 / 8, mean of 3 seeds against the same reference -------------------------------
def _sharded_run(world, mode, seed):
    off, col, ww, ts, td = sbm()
    V, total = len(off) - 1, Q2["models"]["line_cpp"]["total"]
    Wv, Wc = init_tables(V, context_zero=True)
    ms = []
    for r in range(world):
        g = capi.Graph.from_csr(off, col, ww)
        g.set_shard_rotating(r, world) if mode == "rotating" else g.set_shard(r, world)
        m = capi.Model(g, DIM, 2, capi.F32)
        rows = sdist.owned_rows(V, r, world)
        m.set_rows(0, Wv[rows]), m.set_rows(1, Wc[rows])
        if mode == "rotating":
            m.enable_rotation()
        ms.append(m)
    if mode == "rotating":
        sdist.connect_rotation_local(ms)
        cycles = 10
        episodes = cycles * 2 * world
        p = hogwild(seed=seed, total=total // episodes, negative_samples=5, sched_total=total)
        done, _ = sdist.train_line_rotating(ms, p, episodes)
        assert 0.9 * total <= sum(done) <= 1.01 * total
    else:
        for t in range(2):
            ptrs = [m.device_ptr(t) for m in ms]
            for m in ms:
                m.set_peer_ptrs(t, ptrs)
        rounds = 20  # the ranks take turns on the one device: 20 rounds under one LR schedule
        for k in range(rounds):
            for r, m in enumerate(ms):
                p = hogwild(seed=seed * 1000 + k, total=total // rounds, negative_samples=5, sched_total=total,
                            sched_offset=k * (total // rounds), stream_base=r << 20)
                m.train_line(p)
    Wv2, Wc2 = np.zeros((V, DIM)), np.zeros((V, DIM))
    for r, m in enumerate(ms):
        rows = sdist.owned_rows(V, r, world)
        Wv2[rows], Wc2[rows] = m.get_rows(0), m.get_rows(1)
    return Q.evaluate_full(Wv2, Wc2, off, col, ts, td)


@pytest.mark.parametrize("mode", ["peer", "rotating"])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_line(mode, world):
    ref = Q2["models"]["line_cpp"]
    res = np.array([_sharded_run(world, mode, seed) for seed in (101, 102, 103)])
    check(f"{mode} world={world}", res[:, 0].mean(), ref["auc"])
    check(f"{mode} world={world}", res[:, 1].mean(), ref["recall_at_10"], "recall@10")


@pytest.mark.parametrize("world", [2, 4])
def test_rotating_bpr_go(world):
    """The Go BPR on rotating shards (users travel, items stay): held-out AUC against the single-stream reference path."""
    off, col, ww, tu, ti, is_item, _ = bip()
    V, ref = len(off) - 1, Q2["models"]["bpr_go"]
    Wv, Wc = init_tables(V)
    ms = []
    for r in range(world):
        g = capi.Graph.from_csr(off, col, ww, semantics=capi.SEM_GO, n_lines=len(col))
        g.set_shard_rotating(r, world)
        m = capi.Model(g, DIM, 2, capi.F32)
        rows = sdist.owned_rows(V, r, world)
        m.set_rows(0, Wv[rows]), m.set_rows(1, Wc[rows])
        m.enable_rotation()
        ms.append(m)
    sdist.connect_rotation_local(ms)
    episodes = 20 * 2 * world
    p = hogwild(capi.SEM_GO, seed=31, total=ref["total"] // episodes, lambda_=ref["lam"], sched_total=ref["total"])
    done, _ = sdist.train_line_rotating(ms, p, episodes, trainer="bpr")
    assert 0.95 * ref["total"] <= sum(done) <= 1.01 * ref["total"]
    Wv2, Wc2 = np.zeros((V, DIM)), np.zeros((V, DIM))
    for r, m in enumerate(ms):
        rows = sdist.owned_rows(V, r, world)
        Wv2[rows], Wc2[rows] = m.get_rows(0), m.get_rows(1)
    check(f"rotating BPR Go world={world}", Q.evaluate_bipartite(Wv2, Wc2, tu, ti, is_item), ref["auc"])


@pytest.mark.parametrize("sb,hot", [(1 << 13, -1.0), (1 << 15, 12.0)])
def test_exchange_mode_line(sb, hot):
    """Bulk-exchange mode, 4 shards in one process: pure exchange with super-batches of 2^13 samples per shard, and the
    best-connected quarter of the vertices kept single-copy behind the peer pointers (hot threshold 12)."""
    off, col, ww, ts, td = sbm()
    V, ref, world = len(off) - 1, Q2["models"]["line_cpp"], 4
    Wv, Wc = init_tables(V, context_zero=True)
    ms = []
    for r in range(world):
        g = capi.Graph.from_csr(off, col, ww)
        g.set_shard(r, world)
        m = capi.Model(g, DIM, 2, capi.F32)
        rows = sdist.owned_rows(V, r, world)
        m.set_rows(0, Wv[rows]), m.set_rows(1, Wc[rows])
        m.enable_exchange(sb, hot)
        ms.append(m)
    if hot >= 0:
        for t in range(2):
            ptrs = [m.device_ptr(t) for m in ms]
            for m in ms:
                m.set_peer_ptrs(t, ptrs)
    stats = capi.train_line_group(ms, hogwild(seed=100, total=ref["total"], negative_samples=5, max_warps=512))
    assert 0.9 * ref["total"] <= sum(s["samples"] for s in stats) <= ref["total"]
    Wv2, Wc2 = np.zeros((V, DIM)), np.zeros((V, DIM))
    for r, m in enumerate(ms):
        rows = sdist.owned_rows(V, r, world)
        Wv2[rows], Wc2[rows] = m.get_rows(0), m.get_rows(1)
    a, rec = Q.evaluate_full(Wv2, Wc2, off, col, ts, td)
    check(f"exchange sb={sb} hot={hot}", a, ref["auc"])
    check(f"exchange sb={sb} hot={hot}", rec, ref["recall_at_10"], "recall@10")


# ---- the configuration bench.py times: dim 128 (RowCfg<float,4,1>), every resident warp -------------------------------
def test_full_occupancy_dim128_trained_to_convergence():
    """40 000 vertices (80 000 table rows: the occupancy policy lets all 3 552 warps run), dim 128, 1000 updates per vertex,
    against the precomputed single-stream reference path."""
    ref = Q2["models"]["line_cpp_d128_40k"]
    off, col, ww, ts, td = Q.sbm_problem(n_comm=500, comm_size=80, deg=24, seed=31)
    V, dim = len(off) - 1, ref["dim"]
    Wv = (np.random.default_rng(1).random((V, dim)) - 0.5) / dim
    m = capi.Model(capi.Graph.from_csr(off, col, ww), dim, 2, capi.F32)
    m.set_rows(0, Wv), m.set_rows(1, np.zeros((V, dim)))
    st = m.train_line(hogwild(total=ref["total"], negative_samples=5, max_warps=0))
    a, r = Q.evaluate_full(m.get_rows(0), m.get_rows(1), off, col, ts, td)
    print(f"{st['samples'] / st['kernel_ms'] / 1e3:.0f} M updates/s")
    check("LINE C++ dim 128, 3552 warps", a, ref["auc"])
    check("LINE C++ dim 128, 3552 warps", r, ref["recall_at_10"], "recall@10")


def test_full_occupancy_on_a_million_vertex_graph_matches_the_cpu_hogwild_path():
    """1 M vertices, dim 128, max_warps = 0 -> 3 552 warps, tables (1 GB) far larger than L2: bench.py's configuration.
    Against the oracle's OpenMP Hogwild loop (the reference's scheme, src/model/LINE.cpp:162-191) on all host cores: same
    graph, same update count, same LR schedule. Training such a graph to convergence on the CPU takes ~10 minutes, so both
    runs stop early (50 updates per vertex, AUC ~0.56) -- the phase in which AUC moves fastest with the number of
    effective updates: lost or over-applied updates would show as a different AUC."""
    off, col, ww, ts, td = Q.sbm_problem(n_comm=12_500, comm_size=80, deg=24, seed=21)
    V, dim, total = len(off) - 1, 128, 50_000_000
    rng = np.random.default_rng(1)
    Wv = ((rng.random((V, dim)) - 0.5) / dim)
    pick = np.random.default_rng(3).integers(0, len(ts), 200_000)
    ts, td = ts[pick], td[pick]

    def auc_of(A, C):
        neg = np.random.default_rng(4).integers(0, V, len(ts))
        return Q.auc(np.einsum("ij,ij->i", A[ts], C[td]), np.einsum("ij,ij->i", A[ts], C[neg]))

    og = B.OracleGraph(B.SEM_CPP, off, col, ww)
    a, c = Wv.copy(), np.zeros((V, dim))
    og.time_line_cpp(a, c, 5, 0.025, total, 11, os.cpu_count() or 1)
    cpu = auc_of(a, c)
    del a, c, og
    m = capi.Model(capi.Graph.from_csr(off, col, ww), dim, 2, capi.F32)
    m.set_rows(0, Wv), m.set_rows(1, np.zeros((V, dim)))
    st = m.train_line(hogwild(total=total, negative_samples=5, max_warps=0))
    gpu = auc_of(m.get_rows(0), m.get_rows(1))
    print(f"1M-vertex graph, {total} updates: AUC cpu-hogwild {cpu:.4f} gpu {gpu:.4f} ({st['samples'] / st['kernel_ms'] / 1e3:.0f} M updates/s)")
    assert cpu > 0.53, "both runs must have left chance level for the comparison to mean anything"
    assert abs(gpu - cpu) < TOL
