"""CPU-only checks of the drop-in boundary: the library loads, exports every symbol the header declares, and refuses
to compute without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

from smore_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "smore_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(smore_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = capi.lib()
    declared = _header_symbols()
    assert sorted(capi.EXPORTS) == declared
    for name in declared:
        assert hasattr(L, name), name
    assert b"sm_100a" in L.smore_version()


def test_train_params_layout_and_defaults():
    p = capi.default_params()
    # reference CLI defaults: cli/line.cpp:53-54, cmd/line/main.go:15-21
    assert (p.alpha, p.negative_samples, p.order, p.total) == (0.025, 5, 2, 10_000_000)
    assert (p.walk_times, p.walk_steps, p.window_max) == (10, 40, 5)
    assert p.max_walks == -1 and p.lambda_ == 0.001


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    off = np.array([0, 1, 2], dtype=np.int64)
    col = np.array([1, 0], dtype=np.int32)
    w = np.ones(2)
    with pytest.raises(capi.SmoreError, match="no CUDA device|CUDA"):
        capi.Graph.from_csr(off, col, w)


def test_product_never_touches_oracle():
    """The product tree (and the tools/ scripts) must not import, link or open anything under oracle/: only tests/,
    __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm may."""
    for top in ("smore_b200", "tools"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if not f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".sh")):
                    continue
                for line in open(os.path.join(dirpath, f), errors="ignore").read().splitlines():
                    if re.search(r"^\s*(#\s*include|import|from)\b.*oracle", line) or "libsmore_oracle" in line or "libsmore_ref" in line:
                        raise AssertionError(f"{f}: {line}")


def test_argument_validation_happens_before_any_device_work():
    """Shape and argument checks of the C ABI run before the library touches a device (or the caller's arrays), so they are
    the same with and without a GPU: every call below must come back with SMORE_E_INVALID and a message, never crash
    (round-1 advisor finding: smore_graph_create copied the caller's arrays before validating V / E)."""
    import ctypes as C

    L = capi.lib()
    vp = C.c_void_p
    off = np.array([0, 1, 2], dtype=np.int64)
    col = np.array([1, 0], dtype=np.int32)
    w = np.ones(2)
    P = lambda a: a.ctypes.data_as(vp)  # noqa: E731
    h = vp()

    def invalid(rc, pattern):
        assert rc == -1, rc  # SMORE_E_INVALID (include/smore_b200.h)
        assert re.search(pattern, L.smore_last_error().decode()), L.smore_last_error()

    invalid(L.smore_graph_create(2, 2, None, P(col), P(w), 2, capi.SEM_CPP, capi.NEG_DEGREES, C.byref(h)), "null")
    invalid(L.smore_graph_create(2, 2, P(off), None, P(w), 2, capi.SEM_CPP, capi.NEG_DEGREES, C.byref(h)), "null")
    invalid(L.smore_graph_create(2, 2, P(off), P(col), P(w), 2, capi.SEM_CPP, capi.NEG_DEGREES, None), "null")
    invalid(L.smore_graph_create(2, 2, P(off), P(col), P(w), 2, 7, capi.NEG_DEGREES, C.byref(h)), "semantics")
    invalid(L.smore_graph_create(2, 2, P(off), P(col), P(w), 2, capi.SEM_CPP, 9, C.byref(h)), "negative_method")
    invalid(L.smore_graph_create(0, 2, P(off), P(col), P(w), 2, capi.SEM_CPP, capi.NEG_DEGREES, C.byref(h)), "V=0")
    invalid(L.smore_graph_create(-3, 2, P(off), P(col), P(w), 2, capi.SEM_CPP, capi.NEG_DEGREES, C.byref(h)), "V=-3")
    invalid(L.smore_graph_create(1 << 31, 2, P(off), P(col), P(w), 2, capi.SEM_CPP, capi.NEG_DEGREES, C.byref(h)), "out of range")
    invalid(L.smore_graph_create(2, -1, P(off), P(col), P(w), 2, capi.SEM_CPP, capi.NEG_DEGREES, C.byref(h)), "E=-1")
    invalid(L.smore_graph_create(2, 1 << 32, P(off), P(col), P(w), 2, capi.SEM_CPP, capi.NEG_DEGREES, C.byref(h)), "out of range")
    invalid(L.smore_graph_create(2, 3, P(off), P(col), P(w), 3, capi.SEM_CPP, capi.NEG_DEGREES, C.byref(h)), "row_off")
    assert not h.value
    invalid(L.smore_graph_load_edge_list(None, 1, capi.SEM_CPP, capi.NEG_DEGREES, C.byref(h)), "null")
    invalid(L.smore_graph_load_edge_list(b"/tmp/x", 1, 5, capi.NEG_DEGREES, C.byref(h)), "semantics")
    # null handles
    invalid(L.smore_model_wait_copies(None, 1, 1), "null")
    invalid(L.smore_rot_send_begin(None, 0), "rotation")
    invalid(L.smore_model_set_rows_f32_async(None, 0, 0, 1, P(w)), "bad model")
