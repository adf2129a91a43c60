import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: takes more than a few seconds on CPU")


def pytest_sessionstart(session):
    """Build the native artefacts once if they are missing (nvcc cross-compiles without a GPU; ~1-2 min the first time)."""
    lib = os.path.join(ROOT, "lqg_b200", "csrc", "liblqgk.so")
    emul = os.path.join(ROOT, "tests", "emul", "liblqgk_emul.so")
    if not (os.path.exists(lib) and os.path.exists(emul)):
        import __graft_entry__
        __graft_entry__.build()


@pytest.fixture(scope="session")
def repo_root():
    return ROOT
