import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "production-stack_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built_libraries():
    """The C-ABI library and the C oracle are built in-tree (no JIT cache): build if missing."""
    import __graft_entry__ as g
    g.build(quiet=True)
    yield


@pytest.fixture
def shm_name():
    name = f"/b200kv-test-{os.getpid()}-{os.urandom(4).hex()}"
    yield name
    from b200kv import KVPool
    KVPool.unlink(name)
