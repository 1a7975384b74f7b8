import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure both shared libraries exist (the driver calls __graft_entry__.build() first; this is a safety net)."""
    import oracle_lib
    import barbell_b200
    if not os.path.exists(barbell_b200.lib_path()) or not os.path.exists(oracle_lib._SO):
        sys.path.insert(0, ROOT)
        import __graft_entry__ as g
        g.build()
