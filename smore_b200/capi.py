"""ctypes binding of include/smore_b200.h -- the same entry points a cgo host would bind.

The shared library is the product; this module only loads it and marshals numpy buffers. There is no fallback:
if the library is missing, or a call fails (no sm_100 device, bad argument ...), a SmoreError is raised.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SMORE_B200_LIB") or os.path.join(HERE, "lib", "libsmore_b200.so")  # (override: experiment builds)

SEM_CPP, SEM_GO = 0, 1
NEG_DEGREES, NEG_IN_DEGREES, NEG_NO_DEGREES = 0, 1, 2
MODE_DETERMINISTIC, MODE_HOGWILD = 0, 1
F32, F64 = 0, 1
AT_VERTEX, AT_NEGATIVE, AT_CONTEXT = 0, 1, 2
PAIRING_AUTO, PAIRING_COUPLED, PAIRING_SPLIT = 0, 1, 2
SAMPLE_SOURCE, SAMPLE_NEGATIVE, SAMPLE_TARGET, SAMPLE_SOURCE_TARGET = 0, 1, 2, 3

u64, i64, f64, vp = C.c_uint64, C.c_int64, C.c_double, C.c_void_p

# Every symbol include/smore_b200.h declares (tests check the library exports each one).
EXPORTS = [
    "smore_init", "smore_last_error", "smore_version", "smore_kernel_launches",
    "smore_graph_create", "smore_graph_load_edge_list", "smore_edge_list_to_csr", "smore_graph_load_field", "smore_graph_set_field",
    "smore_graph_info", "smore_graph_get_csr", "smore_graph_vertex_name", "smore_graph_get_alias",
    "smore_graph_get_field", "smore_graph_destroy", "smore_sample_debug", "smore_walk_debug",
    "smore_model_create", "smore_model_init", "smore_model_set_rows", "smore_model_get_rows",
    "smore_model_set_rows_f32", "smore_model_get_rows_f32", "smore_model_device_ptr", "smore_model_destroy",
    "smore_graph_set_shard", "smore_graph_shard_info", "smore_model_ipc_handle", "smore_model_open_peers",
    "smore_model_set_peer_ptrs", "smore_model_enable_replica", "smore_model_refresh_replica",
    "smore_model_enable_exchange", "smore_dist_nccl_unique_id", "smore_dist_nccl_init", "smore_dist_nccl_shutdown",
    "smore_graph_set_shard_rotating", "smore_graph_rotation_info", "smore_model_enable_rotation", "smore_model_rot_ipc_handles",
    "smore_model_rot_open_next", "smore_model_rot_slot_ptrs", "smore_model_rot_set_next_ptrs", "smore_rot_send_begin",
    "smore_train_line_episode", "smore_train_bpr_episode", "smore_rot_send_end", "smore_rot_position",
    "smore_alias_build_device", "smore_graph_create_synthetic_rotating",
    "smore_model_set_rows_f32_async", "smore_model_get_rows_f32_async", "smore_model_wait_copies",
    "smore_train_line_group", "smore_exchange_stats", "smore_debug_sm_partition", "smore_model_save_weights", "smore_format_rows",
    "smore_model_load_pretrain", "smore_model_save_checkpoint", "smore_model_load_checkpoint", "smore_model_progress", "smore_progress", "smore_train_params_default", "smore_train_line", "smore_train_bpr",
    "smore_train_warp", "smore_train_hoprec", "smore_train_hpe", "smore_train_mf", "smore_train_skewopt", "smore_train_deepwalk", "smore_train_walklets", "smore_train_node2vec", "smore_train_stats", "smore_model_attach_aux", "smore_model_get_aux_rows", "smore_train_cpr", "smore_train_tpr",
]


class SmoreError(RuntimeError):
    pass


class TrainParams(C.Structure):
    _fields_ = [
        ("semantics", C.c_int), ("mode", C.c_int), ("seed", u64), ("stream_base", u64), ("alpha", f64),
        ("total", u64), ("negative_samples", C.c_int), ("order", C.c_int), ("lambda_", f64),
        ("walk_times", C.c_int), ("walk_steps", C.c_int), ("window_min", C.c_int), ("window_max", C.c_int),
        ("max_warps", C.c_int), ("max_walks", i64), ("sched_total", u64), ("sched_offset", u64),
        ("xi", f64), ("omega", f64), ("eta", C.c_int), ("neg_mode", C.c_int),
        ("n2v_p", f64), ("n2v_q", f64), ("item_reg", f64), ("margin", f64), ("text_weight", f64),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SmoreError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                             "(make -C smore_b200/csrc). There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        L.smore_last_error.restype = C.c_char_p
        L.smore_version.restype = C.c_char_p
        L.smore_kernel_launches.restype = u64
        L.smore_init.argtypes = [C.c_int]
        L.smore_graph_create.argtypes = [i64, i64, vp, vp, vp, i64, C.c_int, C.c_int, C.POINTER(vp)]
        L.smore_graph_load_edge_list.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
        L.smore_edge_list_to_csr.argtypes = [C.c_char_p, C.c_int, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64), vp, vp, vp, vp,
                                             i64, C.POINTER(i64)]
        L.smore_graph_load_field.argtypes = [vp, C.c_char_p]
        L.smore_graph_set_field.argtypes = [vp, vp]
        L.smore_graph_info.argtypes = [vp, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64)]
        L.smore_graph_get_csr.argtypes = [vp, vp, vp, vp]
        L.smore_graph_vertex_name.argtypes = [vp, i64]
        L.smore_graph_vertex_name.restype = C.c_char_p
        L.smore_graph_get_alias.argtypes = [vp, C.c_int, vp, vp]
        L.smore_graph_get_field.argtypes = [vp, vp]
        L.smore_graph_destroy.argtypes = [vp]
        L.smore_graph_destroy.restype = None
        L.smore_sample_debug.argtypes = [vp, C.c_int, u64, u64, i64, vp, vp, C.POINTER(u64)]
        L.smore_walk_debug.argtypes = [vp, u64, u64, i64, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.POINTER(i64), vp, vp,
                                       i64, C.POINTER(i64)]
        L.smore_model_create.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
        L.smore_model_init.argtypes = [vp, C.c_int, C.c_int, u64]
        for name in ("smore_model_set_rows", "smore_model_get_rows", "smore_model_set_rows_f32",
                     "smore_model_get_rows_f32"):
            getattr(L, name).argtypes = [vp, C.c_int, i64, i64, vp]
        L.smore_model_set_rows_f32_async.argtypes = [vp, C.c_int, i64, i64, vp]
        L.smore_model_get_rows_f32_async.argtypes = [vp, C.c_int, i64, i64, vp]
        L.smore_model_wait_copies.argtypes = [vp, C.c_int, C.c_int]
        L.smore_model_device_ptr.argtypes = [vp, C.c_int, C.POINTER(vp)]
        L.smore_model_destroy.argtypes = [vp]
        L.smore_model_destroy.restype = None
        L.smore_model_save_weights.argtypes = [vp, C.c_int, C.c_char_p, C.c_int]
        L.smore_graph_set_shard.argtypes = [vp, C.c_int, C.c_int]
        L.smore_graph_shard_info.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(i64), C.POINTER(f64)]
        L.smore_model_ipc_handle.argtypes = [vp, C.c_int, vp]
        L.smore_model_open_peers.argtypes = [vp, C.c_int, vp]
        L.smore_model_set_peer_ptrs.argtypes = [vp, C.c_int, vp]
        L.smore_model_enable_replica.argtypes = [vp, C.c_int]
        L.smore_model_refresh_replica.argtypes = [vp, C.c_int]
        L.smore_model_enable_exchange.argtypes = [vp, i64, f64]
        L.smore_graph_set_shard_rotating.argtypes = [vp, C.c_int, C.c_int]
        L.smore_graph_rotation_info.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(i64), vp, vp]
        L.smore_model_enable_rotation.argtypes = [vp]
        L.smore_model_rot_ipc_handles.argtypes = [vp, vp]
        L.smore_model_rot_open_next.argtypes = [vp, vp]
        L.smore_model_rot_slot_ptrs.argtypes = [vp, vp]
        L.smore_model_rot_set_next_ptrs.argtypes = [vp, vp]
        L.smore_rot_send_begin.argtypes = [vp, i64]
        L.smore_train_line_episode.argtypes = [vp, C.POINTER(TrainParams), i64]
        L.smore_train_bpr_episode.argtypes = [vp, C.POINTER(TrainParams), i64]
        L.smore_rot_send_end.argtypes = [vp, i64]
        L.smore_rot_position.argtypes = [vp, C.POINTER(i64), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.smore_alias_build_device.argtypes = [vp, i64, vp, vp]
        L.smore_graph_create_synthetic_rotating.argtypes = [i64, i64, u64, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
        L.smore_dist_nccl_unique_id.argtypes = [vp]
        L.smore_dist_nccl_init.argtypes = [vp, C.c_int, C.c_int]
        L.smore_dist_nccl_shutdown.argtypes = []
        L.smore_train_line_group.argtypes = [vp, C.c_int, C.POINTER(TrainParams)]
        L.smore_exchange_stats.argtypes = [vp, C.POINTER(u64), C.POINTER(u64), C.POINTER(i64)]
        L.smore_debug_sm_partition.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.smore_model_load_pretrain.argtypes = [vp, C.c_int, C.c_char_p, C.POINTER(i64)]
        L.smore_model_save_checkpoint.argtypes = [vp, C.c_char_p]
        L.smore_model_load_checkpoint.argtypes = [vp, C.c_char_p]
        L.smore_model_progress.argtypes = [vp, C.POINTER(u64), C.POINTER(u64), C.POINTER(u64), C.POINTER(u64)]
        L.smore_model_attach_aux.argtypes = [vp, i64, vp, vp, vp, u64]
        L.smore_model_get_aux_rows.argtypes = [vp, i64, i64, vp]
        L.smore_progress.argtypes = [vp, C.POINTER(u64), C.POINTER(u64), C.POINTER(C.c_double), C.POINTER(C.c_int)]
        L.smore_format_rows.argtypes = [vp, i64, C.c_int, i64, C.c_int, vp, i64]
        L.smore_format_rows.restype = i64
        L.smore_train_params_default.argtypes = [C.POINTER(TrainParams)]
        L.smore_train_params_default.restype = None
        for name in ("smore_train_line", "smore_train_bpr", "smore_train_warp", "smore_train_hoprec", "smore_train_hpe", "smore_train_mf",
                     "smore_train_skewopt", "smore_train_deepwalk", "smore_train_walklets", "smore_train_node2vec", "smore_train_cpr",
                     "smore_train_tpr"):
            getattr(L, name).argtypes = [vp, C.POINTER(TrainParams)]
        L.smore_train_stats.argtypes = [vp, C.POINTER(u64), C.POINTER(u64), C.POINTER(u64), C.POINTER(f64),
                                        C.POINTER(f64)]
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise SmoreError(f"smore_b200 error {rc}: {lib().smore_last_error().decode()}")


def _ptr(a):
    return a.ctypes.data_as(vp) if a is not None else None


def default_params() -> TrainParams:
    p = TrainParams()
    lib().smore_train_params_default(C.byref(p))
    return p


def train_line_group(models, p):
    """All shards of one row-sharded model in this process (exchange mode, device copies instead of NCCL)."""
    arr = (vp * len(models))(*[m.h for m in models])
    check(lib().smore_train_line_group(C.cast(arr, vp), len(models), C.byref(p)))
    return [m.stats() for m in models]


def nccl_unique_id() -> bytes:
    buf = (C.c_ubyte * 128)()
    check(lib().smore_dist_nccl_unique_id(C.cast(buf, vp)))
    return bytes(buf)


def nccl_init(uid: bytes, rank: int, world: int):
    buf = (C.c_ubyte * 128).from_buffer_copy(uid)
    check(lib().smore_dist_nccl_init(C.cast(buf, vp), rank, world))


def nccl_shutdown():
    check(lib().smore_dist_nccl_shutdown())


def edge_list_to_csr(path, undirected):
    """Host-only ingest: (row_off, col, w, names, n_lines) exactly as Graph.from_edge_list would build them."""
    V, E, n, nb = i64(), i64(), i64(), i64()
    check(lib().smore_edge_list_to_csr(os.fsencode(path), int(undirected), C.byref(V), C.byref(E), C.byref(n), None, None, None,
                                       None, 0, C.byref(nb)))
    off = np.zeros(V.value + 1, dtype=np.int64)
    col = np.zeros(E.value, dtype=np.int32)
    w = np.zeros(E.value)
    names = C.create_string_buffer(max(1, nb.value))
    check(lib().smore_edge_list_to_csr(os.fsencode(path), int(undirected), C.byref(V), C.byref(E), C.byref(n), _ptr(off), _ptr(col),
                                       _ptr(w), C.cast(names, vp), nb.value, C.byref(nb)))
    return off, col, w, names.raw[: nb.value].decode().split("\n")[:-1], n.value


def format_rows(rows, first_id=0, fmt=0) -> bytes:
    """The embedding writer's text for host rows (host-only: usable without a GPU)."""
    rows = np.ascontiguousarray(rows, dtype=np.float64)
    n, dim = rows.shape
    need = lib().smore_format_rows(_ptr(rows), n, dim, first_id, fmt, None, 0)
    if need < 0:
        check(int(need))
    buf = C.create_string_buffer(int(need))
    got = lib().smore_format_rows(_ptr(rows), n, dim, first_id, fmt, C.cast(buf, vp), need)
    assert got == need
    return buf.raw[:need]


def alias_build_device(weights):
    """(thr uint32[n], alias uint32[n]) of the GPU-built alias table for `weights` (include/smore_b200.h)."""
    w = np.ascontiguousarray(weights, dtype=np.float64)
    thr = np.zeros(len(w), dtype=np.uint32)
    alias = np.zeros(len(w), dtype=np.uint32)
    check(lib().smore_alias_build_device(_ptr(w), len(w), _ptr(thr), _ptr(alias)))
    return thr, alias


def kernel_launches() -> int:
    return int(lib().smore_kernel_launches())


class Graph:
    """CSR + alias tables resident in HBM (smore_graph_t)."""

    def __init__(self, handle):
        self.h = handle
        V, E, n = i64(), i64(), i64()
        check(lib().smore_graph_info(self.h, C.byref(V), C.byref(E), C.byref(n)))
        self.V, self.E, self.n_lines = V.value, E.value, n.value

    @classmethod
    def from_csr(cls, row_off, col, w, semantics=SEM_CPP, negative_method=NEG_DEGREES, n_lines=0):
        row_off = np.ascontiguousarray(row_off, dtype=np.int64)
        col = np.ascontiguousarray(col, dtype=np.int32)
        w = np.ascontiguousarray(w, dtype=np.float64)
        h = vp()
        check(lib().smore_graph_create(len(row_off) - 1, len(col), _ptr(row_off), _ptr(col), _ptr(w), n_lines, semantics,
                                       negative_method, C.byref(h)))
        return cls(h)

    @classmethod
    def from_edge_list(cls, path, undirected, semantics=SEM_CPP, negative_method=NEG_DEGREES):
        h = vp()
        check(lib().smore_graph_load_edge_list(os.fsencode(path), int(undirected), semantics, negative_method,
                                               C.byref(h)))
        return cls(h)

    @classmethod
    def synthetic_rotating(cls, V, E_lines, seed, rank, world, semantics=SEM_CPP):
        """Power-law graph generated on the device into this rank's rotating-shard tables (no host CSR)."""
        h = vp()
        check(lib().smore_graph_create_synthetic_rotating(V, E_lines, seed, semantics, rank, world, C.byref(h)))
        return cls(h)

    def close(self):
        if self.h:
            lib().smore_graph_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_shard(self, rank, world):
        """Row-shard: this rank owns vertices v with v % world == rank (world in {1,2,4,8})."""
        check(lib().smore_graph_set_shard(self.h, rank, world))
        return self.shard_info()

    def set_shard_rotating(self, rank, world):
        """Rotating shards: as set_shard, plus one edge table per (vertex sub-part, this rank's contexts) block."""
        check(lib().smore_graph_set_shard_rotating(self.h, rank, world))
        return self.shard_info()

    def rotation_info(self):
        n, cap = C.c_int(), i64()
        check(lib().smore_graph_rotation_info(self.h, C.byref(n), C.byref(cap), None, None))
        mass = np.zeros(n.value)
        edges = np.zeros(n.value, dtype=np.int64)
        check(lib().smore_graph_rotation_info(self.h, C.byref(n), C.byref(cap), _ptr(mass), _ptr(edges)))
        return {"n_sub": n.value, "sub_cap": cap.value, "block_mass": mass, "block_edges": edges}

    def shard_info(self):
        r, w, n, f = C.c_int(), C.c_int(), i64(), f64()
        check(lib().smore_graph_shard_info(self.h, C.byref(r), C.byref(w), C.byref(n), C.byref(f)))
        return {"rank": r.value, "world": w.value, "n_local": n.value, "source_mass_fraction": f.value}

    def load_field(self, path):
        check(lib().smore_graph_load_field(self.h, os.fsencode(path)))

    def set_field(self, field):
        f = np.ascontiguousarray(field, dtype=np.int32)
        assert len(f) == self.V
        check(lib().smore_graph_set_field(self.h, _ptr(f)))

    def field(self):
        f = np.zeros(self.V, dtype=np.int32)
        check(lib().smore_graph_get_field(self.h, _ptr(f)))
        return f

    def csr(self):
        off = np.zeros(self.V + 1, dtype=np.int64)
        col = np.zeros(self.E, dtype=np.int32)
        w = np.zeros(self.E)
        check(lib().smore_graph_get_csr(self.h, _ptr(off), _ptr(col), _ptr(w)))
        return off, col, w

    def names(self):
        out = []
        for v in range(self.V):
            s = lib().smore_graph_vertex_name(self.h, v)
            out.append(s.decode() if s is not None else None)
        return out

    def alias(self, which):
        n = self.E if which == AT_CONTEXT else self.V
        prob = np.zeros(n)
        alias = np.zeros(n, dtype=np.int64)
        check(lib().smore_graph_get_alias(self.h, which, _ptr(prob), _ptr(alias)))
        return prob, alias

    def sample(self, which, seed, stream, n, arg=None):
        out = np.zeros(2 * n if which == SAMPLE_SOURCE_TARGET else n, dtype=np.int64)
        a = np.ascontiguousarray(arg, dtype=np.int64) if arg is not None else None
        used = u64()
        check(lib().smore_sample_debug(self.h, which, seed, stream, n, _ptr(a), _ptr(out), C.byref(used)))
        return out, used.value

    def walk_pairs(self, seed, stream, start, steps, mode, w0, w1=0, cap=1 << 16):
        walk = np.zeros(steps + 1, dtype=np.int64)
        wl, npairs = i64(), i64()
        pv = np.zeros(cap, dtype=np.int64)
        pc = np.zeros(cap, dtype=np.int64)
        check(lib().smore_walk_debug(self.h, seed, stream, start, steps, mode, w0, w1, _ptr(walk), C.byref(wl), _ptr(pv),
                                     _ptr(pc), cap, C.byref(npairs)))
        n = min(npairs.value, cap)
        return walk[: wl.value], pv[:n], pc[:n]


class Model:
    """Embedding tables in HBM (smore_model_t) + the train entry points."""

    def __init__(self, graph: Graph, dim, n_tables=2, dtype=F32):
        self.graph = graph
        self.dim, self.n_tables, self.dtype = dim, n_tables, dtype
        self.h = vp()
        check(lib().smore_model_create(graph.h, dim, n_tables, dtype, C.byref(self.h)))

    def close(self):
        if self.h:
            lib().smore_model_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def init(self, table, random=True, seed=1):
        check(lib().smore_model_init(self.h, table, int(random), seed))

    def set_rows(self, table, W, first=0):
        if W.dtype == np.float32:
            W = np.ascontiguousarray(W)
            check(lib().smore_model_set_rows_f32(self.h, table, first, W.shape[0], _ptr(W)))
        else:
            W = np.ascontiguousarray(W, dtype=np.float64)
            check(lib().smore_model_set_rows(self.h, table, first, W.shape[0], _ptr(W)))

    @property
    def rows(self):
        return self.graph.shard_info()["n_local"]

    def get_rows(self, table, first=0, n=None, dtype=np.float64, out=None):
        n = self.rows - first if n is None else n
        W = out if out is not None else np.zeros((n, self.dim), dtype=dtype)
        fn = lib().smore_model_get_rows_f32 if W.dtype == np.float32 else lib().smore_model_get_rows
        check(fn(self.h, table, first, n, _ptr(W)))
        return W

    def set_rows_async(self, table, W, first=0):
        """W: float32, C-contiguous, ideally pinned. Queued on the upload stream (wait_copies before training)."""
        assert W.dtype == np.float32 and W.flags["C_CONTIGUOUS"]
        check(lib().smore_model_set_rows_f32_async(self.h, table, first, W.shape[0], _ptr(W)))

    def get_rows_async(self, table, out, first=0):
        assert out.dtype == np.float32 and out.flags["C_CONTIGUOUS"]
        check(lib().smore_model_get_rows_f32_async(self.h, table, first, out.shape[0], _ptr(out)))

    def wait_copies(self, uploads=True, readbacks=True):
        check(lib().smore_model_wait_copies(self.h, int(uploads), int(readbacks)))

    def ipc_handle(self, table) -> bytes:
        buf = (C.c_ubyte * 64)()
        check(lib().smore_model_ipc_handle(self.h, table, C.cast(buf, vp)))
        return bytes(buf)

    def open_peers(self, table, handles: bytes):
        """handles = world x 64 bytes (cudaIpcMemHandle_t per rank, in rank order)."""
        buf = (C.c_ubyte * len(handles)).from_buffer_copy(handles)
        check(lib().smore_model_open_peers(self.h, table, C.cast(buf, vp)))

    def set_peer_ptrs(self, table, ptrs):
        arr = (vp * len(ptrs))(*[vp(p) for p in ptrs])
        check(lib().smore_model_set_peer_ptrs(self.h, table, C.cast(arr, vp)))

    def enable_replica(self, table=0):
        check(lib().smore_model_enable_replica(self.h, table))

    def refresh_replica(self, table=0):
        check(lib().smore_model_refresh_replica(self.h, table))

    def enable_exchange(self, superbatch=0, hot_threshold=64.0):
        """Bulk-exchange mode (row-sharded LINE): remote vertex rows move in per-super-batch all-to-alls; vertices
        expected as a source >= hot_threshold times per super-batch stay single-copy behind the peer mappings
        (hot_threshold < 0: none)."""
        check(lib().smore_model_enable_exchange(self.h, superbatch, hot_threshold))

    def exchange_stats(self):
        sb, rows, hot = u64(), u64(), i64()
        check(lib().smore_exchange_stats(self.h, C.byref(sb), C.byref(rows), C.byref(hot)))
        return {"superbatches": sb.value, "rows_requested": rows.value, "hot_vertices": hot.value}

    # ---- rotating shards (include/smore_b200.h) ----
    def enable_rotation(self):
        check(lib().smore_model_enable_rotation(self.h))

    def rot_ipc_handles(self) -> bytes:
        buf = (C.c_ubyte * 192)()
        check(lib().smore_model_rot_ipc_handles(self.h, C.cast(buf, vp)))
        return bytes(buf)

    def rot_open_next(self, handles: bytes):
        buf = (C.c_ubyte * 192).from_buffer_copy(handles)
        check(lib().smore_model_rot_open_next(self.h, C.cast(buf, vp)))

    def rot_slot_ptrs(self):
        arr = (vp * 3)()
        check(lib().smore_model_rot_slot_ptrs(self.h, C.cast(arr, vp)))
        return [arr[k] for k in range(3)]

    def rot_set_next_ptrs(self, ptrs):
        arr = (vp * 3)(*[vp(p) for p in ptrs])
        check(lib().smore_model_rot_set_next_ptrs(self.h, C.cast(arr, vp)))

    def rot_send_begin(self, episode):
        check(lib().smore_rot_send_begin(self.h, episode))

    def train_line_episode(self, p, episode):
        check(lib().smore_train_line_episode(self.h, C.byref(p), episode))
        return self.stats()

    def train_bpr_episode(self, p, episode):
        check(lib().smore_train_bpr_episode(self.h, C.byref(p), episode))
        return self.stats()

    def rot_send_end(self, episode):
        check(lib().smore_rot_send_end(self.h, episode))

    def rot_position(self):
        e, h, q = i64(), C.c_int(), C.c_int()
        check(lib().smore_rot_position(self.h, C.byref(e), C.byref(h), C.byref(q)))
        return {"episode": e.value, "at_home": bool(h.value), "training_subpart": q.value}

    def device_ptr(self, table):
        p = vp()
        check(lib().smore_model_device_ptr(self.h, table, C.byref(p)))
        return p.value

    def save_weights(self, path, table=0, fmt=0):
        check(lib().smore_model_save_weights(self.h, table, os.fsencode(path), fmt))

    def load_pretrain(self, path, table=0) -> int:
        n = i64()
        check(lib().smore_model_load_pretrain(self.h, table, os.fsencode(path), C.byref(n)))
        return n.value

    def save_checkpoint(self, path):
        check(lib().smore_model_save_checkpoint(self.h, os.fsencode(path)))

    def load_checkpoint(self, path):
        check(lib().smore_model_load_checkpoint(self.h, os.fsencode(path)))

    def progress(self):
        a, b, c, d = u64(), u64(), u64(), u64()
        check(lib().smore_model_progress(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        return {"seed": a.value, "next_stream": b.value, "sched_total": c.value, "sched_done": d.value}

    def live_progress(self):
        """smore_progress: callable from another thread while a train call on this model blocks (ctypes releases the GIL)."""
        a, b, al, r = u64(), u64(), C.c_double(), C.c_int()
        check(lib().smore_progress(self.h, C.byref(a), C.byref(b), C.byref(al), C.byref(r)))
        return {"done": a.value, "total": b.value, "alpha": al.value, "running": bool(r.value)}

    def _train(self, fn, p):
        check(fn(self.h, C.byref(p)))
        return self.stats()

    def train_line(self, p): return self._train(lib().smore_train_line, p)
    def train_bpr(self, p): return self._train(lib().smore_train_bpr, p)
    def train_warp(self, p): return self._train(lib().smore_train_warp, p)
    def train_hoprec(self, p): return self._train(lib().smore_train_hoprec, p)
    def train_hpe(self, p): return self._train(lib().smore_train_hpe, p)
    def train_mf(self, p): return self._train(lib().smore_train_mf, p)
    def train_skewopt(self, p): return self._train(lib().smore_train_skewopt, p)
    def train_deepwalk(self, p): return self._train(lib().smore_train_deepwalk, p)
    def train_walklets(self, p): return self._train(lib().smore_train_walklets, p)
    def train_node2vec(self, p): return self._train(lib().smore_train_node2vec, p)
    def train_cpr(self, p): return self._train(lib().smore_train_cpr, p)
    def train_tpr(self, p): return self._train(lib().smore_train_tpr, p)

    def attach_aux(self, row_off, col, rows=None, seed=0):
        """CPR / TPR: adjacency of the second graph (CSR over its own vids) + the third table (None: random init)."""
        off = np.ascontiguousarray(row_off, dtype=np.int64)
        cl = np.ascontiguousarray(col, dtype=np.int32)
        self.aux_V = len(off) - 1
        r = None
        if rows is not None:
            r = np.ascontiguousarray(rows, dtype=np.float64)
            assert r.shape == (self.aux_V, self.dim)
        check(lib().smore_model_attach_aux(self.h, self.aux_V, off.ctypes.data_as(vp), cl.ctypes.data_as(vp),
                                           r.ctypes.data_as(vp) if r is not None else None, seed))

    def get_aux_rows(self, first=0, n=None):
        n = self.aux_V - first if n is None else n
        out = np.empty((n, self.dim), dtype=np.float64)
        check(lib().smore_model_get_aux_rows(self.h, first, n, out.ctypes.data_as(vp)))
        return out

    def stats(self):
        s, pr, w0, ms, tr = u64(), u64(), u64(), f64(), f64()
        check(lib().smore_train_stats(self.h, C.byref(s), C.byref(pr), C.byref(w0), C.byref(tr), C.byref(ms)))
        return {"samples": s.value, "pair_updates": pr.value, "words_stream0": w0.value, "mean_tries": tr.value,
                "kernel_ms": ms.value}
