import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture
def emu_backend():
    """Inject the pure-PyTorch ABI emulation (tests/emu.py) for CPU tests of the host-side logic."""
    from audio8_b200 import ops
    import emu
    prev = ops._BACKEND
    ops.set_backend(emu.EmuOps())
    yield ops._BACKEND
    ops.set_backend(prev)
