"""pytest configuration: the ``gpu`` marker and shared fixtures.

CPU suite  (-m "not gpu"): oracle vs golden vectors / live reference, host logic, C-ABI load + symbol check.
GPU suite  (-m gpu)      : parity tests proper — every call goes through the C ABI of libare_b200.so.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle_binding import Oracle, build_oracle
    build_oracle()
    return Oracle()


@pytest.fixture(scope="session")
def reference():
    from oracle_binding import REF_SO, Reference
    if not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref/libare_ref.so not present (reference tree absent and no prebuilt copy shipped)")
    return Reference()


@pytest.fixture(scope="session")
def lib():
    import __graft_entry__ as g
    g.build(oracle=False)
    from aurora_rendering_engine_b200 import capi
    return capi.load_library()


@pytest.fixture()
def ctx(lib):
    from aurora_rendering_engine_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()
