import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def orc():
    from oracle import orc as _orc
    _orc.lib()
    return _orc


@pytest.fixture(scope="session")
def gen():
    from tools import gen as _gen
    _gen.lib()
    return _gen


@pytest.fixture(scope="session")
def mm2():
    import minimap2_rs_b200 as m
    m.lib()
    return m


@pytest.fixture(scope="session")
def ctx(mm2):
    c = mm2.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="session")
def dense_ctx(mm2):
    """a context whose chaining sends EVERY read with >= 1 anchors to the CTA-per-read kernel (chain_dense_kernel)"""
    os.environ["MM2_CHAIN_DENSE_MIN"] = "1"
    try:
        c = mm2.Context(0)
    finally:
        del os.environ["MM2_CHAIN_DENSE_MIN"]
    yield c
    c.close()
