import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: long-running CPU test")


def pytest_sessionstart(session):
    """The libraries are build artefacts (git-ignored): on a checkout without them, build before collecting -- the tests
    bind libsfx.so (the product; never a fallback) and the oracle (the checker) through ctypes."""
    need = [os.path.join(ROOT, "symforce_b200", "lib", "libsfx.so"), os.path.join(ROOT, "oracle", "_build", "liboracle.so")]
    if any(not os.path.exists(p) for p in need):
        import __graft_entry__ as g

        g.build()
