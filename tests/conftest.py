import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "slow: long-running statistical pin (needs /root/reference)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def built_lib():
    """libmho.so built in-tree (nvcc cross-compiles without a GPU)."""
    from multihop_offload_b200 import build
    return build.build()


def pytest_collection_modifyitems(config, items):
    """`slow` tests (minutes of CPU) run only when asked for: `-m slow` or MHO_SLOW=1."""
    if "slow" in (config.getoption("-m") or "") or os.environ.get("MHO_SLOW"):
        return
    skip = pytest.mark.skip(reason="slow: run with -m slow or MHO_SLOW=1")
    for it in items:
        if "slow" in it.keywords:
            it.add_marker(skip)
