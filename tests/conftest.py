import os
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: takes more than ~30 s on CPU")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def pytest_sessionstart(session):
    """WESEP_TC_FLAGS=<int>: debug / A-B switches of the tcgen05 GEMMs for a whole test session (e.g. 1024 = 2-CTA kernels in
    3xTF32 instead of the mixed tf32 + bf16 split product)."""
    import os
    fl = os.environ.get("WESEP_TC_FLAGS")
    if fl:
        from wesep_b200 import _lib
        _lib.lib().wesep_b200_set_tc_flags(int(fl))
