import os
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.dirname(__file__)):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


# The round-end GPU run uses `-x`: run the parity tests proper first, the multi-process / drop-in integration tests (which
# also depend on the prebuilt reference binaries and host threads) after them.
_ORDER = {"test_gpu_parity.py": 0, "test_gpu_sweeps.py": 1, "test_gpu_checkpoint.py": 2, "test_gpu_grad_stats.py": 3,
          "test_gpu_full_size.py": 4, "test_gpu_multirank.py": 5, "test_gpu_dropin.py": 6}


def pytest_collection_modifyitems(session, config, items):
    keyed = [(_ORDER.get(os.path.basename(str(it.fspath)), -1), i, it) for i, it in enumerate(items)]
    keyed.sort(key=lambda k: (k[0], k[1]))
    items[:] = [it for _, _, it in keyed]


@pytest.fixture(scope="session")
def built_library():
    """The C-ABI shared library, built in-tree on demand (nvcc cross-compiles without a GPU)."""
    from smarties_b200 import build as b
    return b.build()
