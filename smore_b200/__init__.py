"""smore_b200 -- B200-native sampled-SGD backend for SMORe (LINE / DeepWalk / Walklets / BPR / WARP / HOP-Rec).

The product is smore_b200/lib/libsmore_b200.so (hand-written sm_100a CUDA behind the C ABI in include/smore_b200.h);
`capi` is its ctypes binding and `models` mirrors the reference's model classes on top of it.
"""
from . import capi  # noqa: F401

__all__ = ["capi"]
