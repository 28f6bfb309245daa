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
