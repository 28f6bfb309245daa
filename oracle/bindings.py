"""TEST INFRASTRUCTURE ONLY -- ctypes bindings for the CPU oracle.

* ``Oracle``  : oracle/libsmore_oracle.so, our restatement (smore_oracle.cpp).
* ``Ref``     : oracle/_ref/libsmore_ref.so, the UNMODIFIED compiled C++ reference driven through the Philox shim
                (only present where `make -C oracle ref` ran, i.e. where /root/reference exists, or where the prebuilt
                file travelled with the snapshot).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this module.
"""
from __future__ import annotations

import contextlib
import ctypes as C
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libsmore_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libsmore_ref.so")
REF_FAST_SO = os.path.join(HERE, "_ref", "libsmore_ref_fast.so")

SEM_CPP, SEM_GO = 0, 1
NEG_DEGREES, NEG_IN_DEGREES, NEG_NO_DEGREES = 0, 1, 2
K_LINE, K_DEEPWALK, K_WALKLETS, K_BPR, K_WARP, K_HOPREC, K_HPE, K_MF, K_SKEWOPT = range(9)

u64 = C.c_uint64
i64 = C.c_int64
f64 = C.c_double
vp = C.c_void_p


def build(ref: bool | None = None) -> None:
    """Compile the restatement (always) and the compiled-reference libraries (when /root/reference is present)."""
    subprocess.run(["make", "-C", HERE, "oracle"], check=True, stdout=subprocess.DEVNULL)
    if ref is None:
        ref = os.path.isdir("/root/reference/src")
    if ref:
        subprocess.run(["make", "-C", HERE, "ref"], check=True, stdout=subprocess.DEVNULL)


def _ptr(a):
    return a.ctypes.data_as(vp) if a is not None else None


@contextlib.contextmanager
def quiet():
    """Silence the reference's progress printing (it writes to fd 1 from C++)."""
    sys.stdout.flush()
    saved = os.dup(1)
    devnull = os.open(os.devnull, os.O_WRONLY)
    try:
        os.dup2(devnull, 1)
        yield
    finally:
        os.dup2(saved, 1)
        os.close(devnull)
        os.close(saved)


class Oracle:
    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            if not os.path.exists(ORACLE_SO):
                build(ref=False)
            L = C.CDLL(ORACLE_SO)
            L.orc_graph_create.restype = vp
            L.orc_graph_create.argtypes = [C.c_int, i64, i64, vp, vp, vp, i64, C.c_int]
            L.orc_graph_free.argtypes = [vp]
            L.orc_graph_set_field.argtypes = [vp, vp]
            L.orc_graph_degrees.argtypes = [vp, vp, vp]
            L.orc_graph_alias.argtypes = [vp, C.c_int, vp, vp]
            L.orc_sigmoid_table.argtypes = [vp, vp]
            L.orc_fast_sigmoid.argtypes = [vp, f64]
            L.orc_fast_sigmoid.restype = f64
            L.orc_alias_build.argtypes = [C.c_int, vp, i64, f64, vp, vp]
            L.orc_philox_block.argtypes = [vp, vp, vp]
            L.orc_stream_words.argtypes = [u64, u64, u64, i64, vp]
            L.orc_sample.argtypes = [vp, C.c_int, u64, u64, i64, vp, vp]
            L.orc_sample.restype = u64
            L.orc_walk_pairs.argtypes = [vp, u64, u64, i64, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, i64]
            L.orc_walk_pairs.restype = i64
            L.orc_train_line_cpp.argtypes = [vp, vp, vp, C.c_int, C.c_int, f64, u64, u64, u64]
            L.orc_train_line_cpp.restype = u64
            L.orc_train_bpr_cpp.argtypes = [vp, vp, C.c_int, f64, u64, u64, u64]
            L.orc_train_bpr_cpp.restype = u64
            L.orc_train_warp_cpp.argtypes = [vp, vp, C.c_int, f64, u64, u64, u64, vp]
            L.orc_train_warp_cpp.restype = u64
            L.orc_train_hoprec_cpp.argtypes = [vp, vp, C.c_int, C.c_int, f64, u64, u64, u64]
            L.orc_train_hoprec_cpp.restype = u64
            L.orc_train_skewopt_cpp.argtypes = [vp, vp, C.c_int, f64, f64, C.c_int, f64, u64, u64, u64]
            L.orc_train_skewopt_cpp.restype = u64
            L.orc_update_sbpr_pair_cpp.argtypes = [vp, vp, i64, i64, C.c_int, f64, f64, C.c_int, f64, u64, u64]
            L.orc_update_sbpr_pair_cpp.restype = u64
            L.orc_train_mf_cpp.argtypes = [vp, vp, C.c_int, C.c_int, f64, f64, u64, u64, u64]
            L.orc_train_mf_cpp.restype = u64
            L.orc_update_factorized_pair_cpp.argtypes = [vp, vp, i64, i64, C.c_int, f64, C.c_int, f64, u64, u64]
            L.orc_update_factorized_pair_cpp.restype = u64
            L.orc_train_hpe_cpp.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_int, f64, f64, u64, u64, u64]
            L.orc_train_hpe_cpp.restype = u64
            L.orc_update_community_cpp.argtypes = [vp, vp, vp, i64, i64, C.c_int, f64, C.c_int, C.c_int, f64, u64, u64]
            L.orc_update_community_cpp.restype = u64
            L.orc_train_walk_cpp.argtypes = [vp, C.c_int, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                             f64, u64, u64, i64, vp]
            L.orc_train_walk_cpp.restype = u64
            L.orc_train_line_go.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_int, f64, u64, u64, u64]
            L.orc_train_line_go.restype = u64
            L.orc_train_bpr_go.argtypes = [vp, vp, vp, C.c_int, f64, f64, u64, u64, u64]
            L.orc_train_bpr_go.restype = u64
            L.orc_train_deepwalk_go.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, f64, u64, u64,
                                                i64, vp]
            L.orc_train_deepwalk_go.restype = u64
            L.orc_train_cpr_go.argtypes = [vp, vp, vp, vp, vp, C.c_int, f64, f64, f64, f64, u64, u64, u64, u64]
            L.orc_train_cpr_go.restype = u64
            L.orc_train_tpr_go.argtypes = [vp, vp, vp, vp, vp, C.c_int, f64, f64, f64, u64, u64, u64, u64]
            L.orc_train_tpr_go.restype = u64
            L.orc_train_deepwalk_go_streams.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, f64, u64, C.c_int, vp]
            L.orc_train_deepwalk_go_streams.restype = u64
            L.orc_train_node2vec_go.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, f64, f64, f64, u64, u64,
                                                i64, vp]
            L.orc_train_node2vec_go.restype = u64
            L.orc_biased_walk_go.argtypes = [vp, i64, C.c_int, f64, f64, u64, u64, vp, vp]
            L.orc_biased_walk_go.restype = i64
            L.orc_update_pair_cpp.argtypes = [vp, vp, vp, i64, i64, C.c_int, C.c_int, f64, u64, u64]
            L.orc_update_pair_cpp.restype = u64
            L.orc_update_bpr_pair_cpp.argtypes = [vp, vp, i64, i64, i64, C.c_int, f64, u64, u64]
            L.orc_update_bpr_pair_cpp.restype = u64
            L.orc_update_warp_pair_cpp.argtypes = [vp, vp, i64, i64, i64, C.c_int, f64, u64, u64]
            L.orc_update_warp_pair_cpp.restype = u64
            L.orc_update_fbpr_pair_cpp.argtypes = [vp, vp, i64, i64, i64, C.c_int, f64, f64, u64, u64]
            L.orc_update_fbpr_pair_cpp.restype = u64
            L.orc_time_line_cpp.argtypes = [vp, vp, vp, C.c_int, C.c_int, f64, u64, u64, C.c_int]
            L.orc_time_line_cpp.restype = f64
            L.orc_write_edge_list.argtypes = [C.c_char_p, vp, vp, vp, i64]
            cls._lib = L
        return cls._lib


def philox_block(ctr, key):
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    Oracle.lib().orc_philox_block(_ptr(c), _ptr(k), _ptr(out))
    return out


def stream_words(seed, stream, first, n):
    out = np.zeros(n, dtype=np.uint32)
    Oracle.lib().orc_stream_words(seed, stream, first, n, _ptr(out))
    return out


def alias_build(sem, dist, power=1.0):
    dist = np.ascontiguousarray(dist, dtype=np.float64)
    prob = np.zeros(len(dist))
    alias = np.zeros(len(dist), dtype=np.int64)
    Oracle.lib().orc_alias_build(sem, _ptr(dist), len(dist), power, _ptr(prob), _ptr(alias))
    return prob, alias


class OracleGraph:
    """CSR graph (reference insertion order) + the reference's alias tables, restated."""

    def __init__(self, sem, row_off, col, w, max_line=None, neg_method=NEG_DEGREES):
        self.L = Oracle.lib()
        self.sem = sem
        self.row_off = np.ascontiguousarray(row_off, dtype=np.int64)
        self.col = np.ascontiguousarray(col, dtype=np.int32)
        self.w = np.ascontiguousarray(w, dtype=np.float64)
        self.V = len(self.row_off) - 1
        self.E = len(self.col)
        if max_line is None:
            max_line = self.E
        self.h = vp(self.L.orc_graph_create(sem, self.V, self.E, _ptr(self.row_off), _ptr(self.col), _ptr(self.w),
                                            max_line, neg_method))

    def __del__(self):
        try:
            self.L.orc_graph_free(self.h)
        except Exception:
            pass

    def set_field(self, field):
        f = np.ascontiguousarray(field, dtype=np.int32)
        assert len(f) == self.V
        self.L.orc_graph_set_field(self.h, _ptr(f))

    def degrees(self):
        o, i = np.zeros(self.V), np.zeros(self.V)
        self.L.orc_graph_degrees(self.h, _ptr(o), _ptr(i))
        return o, i

    def alias(self, which):
        n = self.E if which == 2 else self.V
        prob = np.zeros(n)
        alias = np.zeros(n, dtype=np.int64)
        self.L.orc_graph_alias(self.h, which, _ptr(prob), _ptr(alias))
        return prob, alias

    def sigmoid_table(self):
        t = np.zeros(1001)
        self.L.orc_sigmoid_table(self.h, _ptr(t))
        return t

    def fast_sigmoid(self, x):
        return self.L.orc_fast_sigmoid(self.h, float(x))

    def sample(self, which, seed, stream, n, arg=None):
        out = np.zeros(2 * n if which == 3 else n, dtype=np.int64)
        a = np.ascontiguousarray(arg, dtype=np.int64) if arg is not None else None
        pos = self.L.orc_sample(self.h, which, seed, stream, n, _ptr(a), _ptr(out))
        return out, pos

    def walk_pairs(self, seed, stream, start, steps, mode, w0, w1=0, cap=1 << 16):
        walk = np.zeros(steps + 1, dtype=np.int64)
        wl = i64(0)
        pv = np.zeros(cap, dtype=np.int64)
        pc = np.zeros(cap, dtype=np.int64)
        n = self.L.orc_walk_pairs(self.h, seed, stream, start, steps, mode, w0, w1, _ptr(walk), C.byref(wl), _ptr(pv),
                                  _ptr(pc), cap)
        return walk[: wl.value], pv[:n], pc[:n]

    # train loops: tables are float64 [V, dim] C-contiguous arrays, updated in place; returns words consumed
    def train_line_cpp(self, Wv, Wc, K, alpha, total, seed, stream=0):
        return self.L.orc_train_line_cpp(self.h, _ptr(Wv), _ptr(Wc), Wv.shape[1], K, alpha, total, seed, stream)

    def train_bpr_cpp(self, W, alpha, total, seed, stream=0):
        return self.L.orc_train_bpr_cpp(self.h, _ptr(W), W.shape[1], alpha, total, seed, stream)

    def train_warp_cpp(self, W, alpha, total, seed, stream=0):
        tries = u64(0)
        pos = self.L.orc_train_warp_cpp(self.h, _ptr(W), W.shape[1], alpha, total, seed, stream, C.byref(tries))
        return pos, tries.value

    def train_hoprec_cpp(self, W, walk_steps, alpha, total, seed, stream=0):
        return self.L.orc_train_hoprec_cpp(self.h, _ptr(W), W.shape[1], walk_steps, alpha, total, seed, stream)

    def train_skewopt_cpp(self, W, xi, omega, eta, alpha, total, seed, stream=0):
        return self.L.orc_train_skewopt_cpp(self.h, _ptr(W), W.shape[1], xi, omega, eta, alpha, total, seed, stream)

    def update_sbpr_pair_cpp(self, W, v, ci, xi, omega, eta, alpha, seed, stream=0):
        return self.L.orc_update_sbpr_pair_cpp(self.h, _ptr(W), v, ci, W.shape[1], xi, omega, eta, alpha, seed, stream)

    def train_mf_cpp(self, W, K, reg, alpha, total, seed, stream=0):
        return self.L.orc_train_mf_cpp(self.h, _ptr(W), W.shape[1], K, reg, alpha, total, seed, stream)

    def update_factorized_pair_cpp(self, W, v, c, reg, K, alpha, seed, stream=0):
        return self.L.orc_update_factorized_pair_cpp(self.h, _ptr(W), v, c, W.shape[1], reg, K, alpha, seed, stream)

    def train_hpe_cpp(self, Wv, Wc, walk_steps, K, reg, alpha, total, seed, stream=0):
        return self.L.orc_train_hpe_cpp(self.h, _ptr(Wv), _ptr(Wc), Wv.shape[1], walk_steps, K, reg, alpha, total, seed, stream)

    def update_community_cpp(self, Wv, Wc, v, c, reg, walk_steps, K, alpha, seed, stream=0):
        return self.L.orc_update_community_cpp(self.h, _ptr(Wv), _ptr(Wc), v, c, Wv.shape[1], reg, walk_steps, K, alpha,
                                               seed, stream)

    def train_walk_cpp(self, walklets, Wv, Wc, walk_times, walk_steps, w0, w1, K, alpha, seed, stream=0, max_walks=-1):
        pairs = u64(0)
        pos = self.L.orc_train_walk_cpp(self.h, int(walklets), _ptr(Wv), _ptr(Wc), Wv.shape[1], walk_times, walk_steps,
                                        w0, w1, K, alpha, seed, stream, max_walks, C.byref(pairs))
        return pos, pairs.value

    def train_line_go(self, Wv, Wc, order, K, alpha, iterations, seed, stream=0):
        return self.L.orc_train_line_go(self.h, _ptr(Wv), _ptr(Wc), Wv.shape[1], order, K, alpha, iterations, seed,
                                        stream)

    def train_bpr_go(self, Wv, Wc, alpha, lam, iterations, seed, stream=0):
        return self.L.orc_train_bpr_go(self.h, _ptr(Wv), _ptr(Wc), Wv.shape[1], alpha, lam, iterations, seed, stream)

    def train_deepwalk_go(self, Wv, Wc, walk_times, walk_steps, window, K, alpha, seed, stream=0, max_walks=-1):
        pairs = u64(0)
        pos = self.L.orc_train_deepwalk_go(self.h, _ptr(Wv), _ptr(Wc), Wv.shape[1], walk_times, walk_steps, window, K,
                                           alpha, seed, stream, max_walks, C.byref(pairs))
        return pos, pairs.value

    def train_cpr_go(self, source, U, T, S, alpha, user_reg, item_reg, margin, iterations, total, seed, stream=0):
        """self = the target-domain graph, source = the source-domain OracleGraph; S is read-only."""
        return self.L.orc_train_cpr_go(self.h, source.h, _ptr(U), _ptr(T), _ptr(S), U.shape[1], alpha, user_reg, item_reg, margin,
                                       iterations, total, seed, stream)

    def train_tpr_go(self, words, U, I, Wd, alpha, lam, text_weight, iterations, total, seed, stream=0):
        """self = the user-item graph, words = the item-word OracleGraph."""
        return self.L.orc_train_tpr_go(self.h, words.h, _ptr(U), _ptr(I), _ptr(Wd), U.shape[1], alpha, lam, text_weight,
                                       iterations, total, seed, stream)

    def train_deepwalk_go_streams(self, Wv, Wc, walk_times, walk_steps, window, K, alpha, seed, n_streams):
        """Diagnostic: the device's W-stream work split of the Go DeepWalk loop, run sequentially (not a reference path)."""
        pairs = u64(0)
        self.L.orc_train_deepwalk_go_streams(self.h, _ptr(Wv), _ptr(Wc), Wv.shape[1], walk_times, walk_steps, window, K,
                                             alpha, seed, n_streams, C.byref(pairs))
        return pairs.value

    def train_node2vec_go(self, Wv, Wc, walk_times, walk_steps, window, K, alpha, p, q, seed, stream=0, max_walks=-1):
        pairs = u64(0)
        pos = self.L.orc_train_node2vec_go(self.h, _ptr(Wv), _ptr(Wc), Wv.shape[1], walk_times, walk_steps, window, K,
                                           alpha, p, q, seed, stream, max_walks, C.byref(pairs))
        return pos, pairs.value

    def biased_walk_go(self, start, steps, p, q, seed, stream=0):
        out = np.zeros(steps + 1, dtype=np.int64)
        words = u64(0)
        n = self.L.orc_biased_walk_go(self.h, start, steps, p, q, seed, stream, _ptr(out), C.byref(words))
        return out[:n], words.value

    def update_pair_cpp(self, Wv, Wc, v, c, K, alpha, seed, stream=0):
        return self.L.orc_update_pair_cpp(self.h, _ptr(Wv), _ptr(Wc), v, c, Wv.shape[1], K, alpha, seed, stream)

    def update_bpr_pair_cpp(self, W, v, ci, cj, alpha, seed, stream=0):
        return self.L.orc_update_bpr_pair_cpp(self.h, _ptr(W), v, ci, cj, W.shape[1], alpha, seed, stream)

    def update_warp_pair_cpp(self, W, v, ci, cj, alpha, seed, stream=0):
        return self.L.orc_update_warp_pair_cpp(self.h, _ptr(W), v, ci, cj, W.shape[1], alpha, seed, stream)

    def update_fbpr_pair_cpp(self, W, v, ci, cj, alpha, margin, seed, stream=0):
        return self.L.orc_update_fbpr_pair_cpp(self.h, _ptr(W), v, ci, cj, W.shape[1], alpha, margin, seed, stream)

    def time_line_cpp(self, Wv, Wc, K, alpha, total, seed, workers):
        return self.L.orc_time_line_cpp(self.h, _ptr(Wv), _ptr(Wc), Wv.shape[1], K, alpha, total, seed, workers)


def ref_available(fast=False) -> bool:
    return os.path.exists(REF_FAST_SO if fast else REF_SO)


class Ref:
    """One model object of the compiled reference (LINE / DeepWalk / Walklets / BPR / WARP / HBPR)."""

    _libs = {}

    @classmethod
    def lib(cls, fast=False):
        path = REF_FAST_SO if fast else REF_SO
        if path not in cls._libs:
            L = C.CDLL(path)
            L.ref_new.restype = vp
            L.ref_new.argtypes = [C.c_int]
            L.ref_free.argtypes = [vp]
            L.ref_seed.argtypes = [u64, u64]
            L.ref_stream_pos.restype = u64
            L.ref_load_edge_list.argtypes = [vp, C.c_char_p, C.c_int]
            L.ref_load_field.argtypes = [vp, C.c_char_p]
            L.ref_init.argtypes = [vp, C.c_int, C.c_int]
            L.ref_num_vertices.argtypes = [vp]
            L.ref_num_vertices.restype = i64
            L.ref_num_lines.argtypes = [vp]
            L.ref_num_lines.restype = i64
            L.ref_get_csr.argtypes = [vp, vp, vp, vp]
            L.ref_get_degrees.argtypes = [vp, vp, vp]
            L.ref_vertex_name.argtypes = [vp, i64]
            L.ref_vertex_name.restype = C.c_char_p
            L.ref_field_of.argtypes = [vp, i64]
            L.ref_field_of.restype = C.c_int32
            L.ref_alias_size.argtypes = [vp, C.c_int]
            L.ref_alias_size.restype = i64
            L.ref_get_alias.argtypes = [vp, C.c_int, vp, vp]
            L.ref_alias_method.argtypes = [vp, i64, f64, vp, vp]
            L.ref_fast_sigmoid.argtypes = [vp, f64]
            L.ref_fast_sigmoid.restype = f64
            L.ref_sigmoid_table.argtypes = [vp, vp]
            L.ref_sample.argtypes = [vp, C.c_int, i64, vp, vp]
            L.ref_walk_pairs.argtypes = [vp, i64, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, i64]
            L.ref_walk_pairs.restype = i64
            L.ref_set_rows.argtypes = [vp, C.c_int, vp]
            L.ref_get_rows.argtypes = [vp, C.c_int, vp]
            L.ref_update_pair.argtypes = [vp, i64, i64, C.c_int, f64]
            L.ref_update_bpr_pair.argtypes = [vp, i64, i64, i64, f64]
            L.ref_update_warp_pair.argtypes = [vp, i64, i64, i64, f64]
            L.ref_update_fbpr_pair.argtypes = [vp, i64, i64, i64, f64, f64]
            L.ref_train.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, f64, C.c_int]
            L.ref_save_weights.argtypes = [vp, C.c_char_p]
            L.ref_train_skewopt.argtypes = [vp, C.c_int, C.c_int, f64, f64, f64, f64, C.c_int, C.c_int]
            L.ref_update_sbpr_pair.argtypes = [vp, i64, i64, f64, f64, C.c_int, f64]
            L.ref_train_mf.argtypes = [vp, C.c_int, C.c_int, f64, f64, C.c_int]
            L.ref_update_factorized_pair.argtypes = [vp, i64, i64, f64, C.c_int, f64]
            L.ref_train_hpe.argtypes = [vp, C.c_int, C.c_int, C.c_int, f64, f64, C.c_int]
            L.ref_update_community.argtypes = [vp, i64, i64, f64, C.c_int, C.c_int, f64]
            cls._libs[path] = L
        return cls._libs[path]

    def __init__(self, kind, edge_file, undirected, dim, order=2, field_file=None, fast=False):
        self.L = self.lib(fast)
        self.kind = kind
        self.dim = dim
        self.order = order
        with quiet():
            self.h = vp(self.L.ref_new(kind))
            self.L.ref_load_edge_list(self.h, edge_file.encode(), int(undirected))
            if field_file is not None:
                self.L.ref_load_field(self.h, field_file.encode())
            self.L.ref_init(self.h, dim, order)
        self.V = self.L.ref_num_vertices(self.h)
        self.max_line = self.L.ref_num_lines(self.h)

    def __del__(self):
        try:
            self.L.ref_free(self.h)
        except Exception:
            pass

    def seed(self, seed, stream_base=0):
        self.L.ref_seed(seed, stream_base)

    def pos(self):
        return self.L.ref_stream_pos()

    def csr(self):
        E = self.L.ref_alias_size(self.h, 2)
        off = np.zeros(self.V + 1, dtype=np.int64)
        col = np.zeros(E, dtype=np.int32)
        w = np.zeros(E)
        self.L.ref_get_csr(self.h, _ptr(off), _ptr(col), _ptr(w))
        return off, col, w

    def degrees(self):
        o, i = np.zeros(self.V), np.zeros(self.V)
        self.L.ref_get_degrees(self.h, _ptr(o), _ptr(i))
        return o, i

    def names(self):
        return [self.L.ref_vertex_name(self.h, v).decode() for v in range(self.V)]

    def fields(self):
        return np.array([self.L.ref_field_of(self.h, v) for v in range(self.V)], dtype=np.int32)

    def alias(self, which):
        n = self.L.ref_alias_size(self.h, which)
        prob = np.zeros(n)
        alias = np.zeros(n, dtype=np.int64)
        self.L.ref_get_alias(self.h, which, _ptr(prob), _ptr(alias))
        return prob, alias

    @classmethod
    def alias_method(cls, dist, power=1.0):
        L = cls.lib()
        dist = np.ascontiguousarray(dist, dtype=np.float64)
        prob = np.zeros(len(dist))
        alias = np.zeros(len(dist), dtype=np.int64)
        with quiet():
            L.ref_alias_method(_ptr(dist), len(dist), power, _ptr(prob), _ptr(alias))
        return prob, alias

    def sigmoid_table(self):
        t = np.zeros(1001)
        self.L.ref_sigmoid_table(self.h, _ptr(t))
        return t

    def fast_sigmoid(self, x):
        return self.L.ref_fast_sigmoid(self.h, float(x))

    def sample(self, which, n, arg=None):
        out = np.zeros(2 * n if which == 3 else n, dtype=np.int64)
        a = np.ascontiguousarray(arg, dtype=np.int64) if arg is not None else None
        self.L.ref_sample(self.h, which, n, _ptr(a), _ptr(out))
        return out

    def walk_pairs(self, start, steps, mode, w0, w1=0, cap=1 << 16):
        walk = np.zeros(steps + 1, dtype=np.int64)
        wl = i64(0)
        pv = np.zeros(cap, dtype=np.int64)
        pc = np.zeros(cap, dtype=np.int64)
        n = self.L.ref_walk_pairs(self.h, start, steps, mode, w0, w1, _ptr(walk), C.byref(wl), _ptr(pv), _ptr(pc), cap)
        return walk[: wl.value], pv[:n], pc[:n]

    def set_rows(self, table, W):
        W = np.ascontiguousarray(W, dtype=np.float64)
        assert W.shape == (self.V, self.dim)
        self.L.ref_set_rows(self.h, table, _ptr(W))

    def get_rows(self, table):
        W = np.zeros((self.V, self.dim))
        self.L.ref_get_rows(self.h, table, _ptr(W))
        return W

    def update_pair(self, v, c, K, alpha):
        self.L.ref_update_pair(self.h, v, c, K, alpha)

    def update_bpr_pair(self, v, ci, cj, alpha):
        self.L.ref_update_bpr_pair(self.h, v, ci, cj, alpha)

    def update_warp_pair(self, v, ci, cj, alpha):
        self.L.ref_update_warp_pair(self.h, v, ci, cj, alpha)

    def update_fbpr_pair(self, v, ci, cj, alpha, margin):
        self.L.ref_update_fbpr_pair(self.h, v, ci, cj, alpha, margin)

    def train(self, a, b=0, c=0, d=0, e=0, alpha=0.025, workers=1):
        with quiet():
            self.L.ref_train(self.h, a, b, c, d, e, alpha, workers)

    def train_skewopt(self, sample_times, xi, omega, eta, alpha=0.025, workers=1):
        with quiet():
            self.L.ref_train_skewopt(self.h, sample_times, 5, alpha, 0.01, xi, omega, eta, workers)

    def update_sbpr_pair(self, v, ci, xi, omega, eta, alpha):
        self.L.ref_update_sbpr_pair(self.h, v, ci, xi, omega, eta, alpha)

    def train_mf(self, sample_times, K, reg, alpha=0.025, workers=1):
        with quiet():
            self.L.ref_train_mf(self.h, sample_times, K, alpha, reg, workers)

    def update_factorized_pair(self, v, c, reg, K, alpha):
        self.L.ref_update_factorized_pair(self.h, v, c, reg, K, alpha)

    def train_hpe(self, sample_times, walk_steps, K, reg, alpha=0.025, workers=1):
        with quiet():
            self.L.ref_train_hpe(self.h, sample_times, walk_steps, K, reg, alpha, workers)

    def update_community(self, v, c, reg, walk_steps, K, alpha):
        self.L.ref_update_community(self.h, v, c, reg, walk_steps, K, alpha)

    def save_weights(self, path):
        with quiet():
            self.L.ref_save_weights(self.h, path.encode())


def write_edge_list_fast(path, src, dst, w):
    """Same format as write_edge_list (names "v<label>"), written from C for multi-million-line files."""
    src = np.ascontiguousarray(src, dtype=np.int64)
    dst = np.ascontiguousarray(dst, dtype=np.int64)
    w = np.ascontiguousarray(w, dtype=np.float64)
    if Oracle.lib().orc_write_edge_list(os.fsencode(path), _ptr(src), _ptr(dst), _ptr(w), len(src)) != 0:
        raise OSError(f"cannot write {path}")


def write_edge_list(path, src, dst, w, names=None):
    """Text edge list the reference ingests: `name name weight` per line."""
    with open(path, "w") as f:
        if names is None:
            for a, b, x in zip(src, dst, w):
                f.write(f"v{a} v{b} {x:g}\n")
        else:
            for a, b, x in zip(src, dst, w):
                f.write(f"{names[a]} {names[b]} {x:g}\n")


def edges_to_csr(src, dst, w, undirected):
    """Reference ingest order (src/proNet.cpp:178-215, pronet.go:145-155): ids by first appearance (src before dst),
    per-vertex adjacency in file order, reverse entry appended right after the forward one when undirected.
    Returns (row_off, col, w, id_of_label) where id_of_label maps the input labels to reference vertex ids."""
    src = np.asarray(src)
    dst = np.asarray(dst)
    w = np.asarray(w, dtype=np.float64)
    ids = {}
    s2 = np.empty(len(src), dtype=np.int64)
    d2 = np.empty(len(src), dtype=np.int64)
    for i, (a, b) in enumerate(zip(src.tolist(), dst.tolist())):
        if a not in ids:
            ids[a] = len(ids)
        if b not in ids:
            ids[b] = len(ids)
        s2[i] = ids[a]
        d2[i] = ids[b]
    V = len(ids)
    if undirected:
        es = np.empty(2 * len(src), dtype=np.int64)
        ed = np.empty(2 * len(src), dtype=np.int64)
        ew = np.empty(2 * len(src))
        es[0::2], es[1::2] = s2, d2
        ed[0::2], ed[1::2] = d2, s2
        ew[0::2], ew[1::2] = w, w
    else:
        es, ed, ew = s2, d2, w
    order = np.argsort(es, kind="stable")
    col = ed[order].astype(np.int32)
    ww = ew[order]
    row_off = np.zeros(V + 1, dtype=np.int64)
    np.add.at(row_off, es + 1, 1)
    row_off = np.cumsum(row_off)
    return row_off, col, ww, ids
