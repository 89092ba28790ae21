import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    import oracle_lib
    oracle_lib.lib()
    return oracle_lib


@pytest.fixture(scope="session")
def apa():
    import astar_pairwise_aligner_b200 as A
    if not os.path.exists(A.lib_path()):
        import __graft_entry__
        __graft_entry__.build()
    A.load_library()
    return A


@pytest.fixture(scope="session")
def engine(apa):
    return apa.Engine(0)
