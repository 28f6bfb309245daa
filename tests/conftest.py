import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "ref: needs oracle/_ref/libsmore_ref.so (the compiled reference)")


def pytest_collection_modifyitems(config, items):
    from oracle import bindings

    have_ref = bindings.ref_available()
    skip_ref = pytest.mark.skip(reason="oracle/_ref/libsmore_ref.so not built (no /root/reference here)")
    for item in items:
        if "ref" in item.keywords and not have_ref:
            item.add_marker(skip_ref)
