import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_addoption(parser):
    parser.addoption("--aai-lib", default="", help="developer A/B: run the suite against another build of libaai_b200.so")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    lib = config.getoption("--aai-lib")
    if lib:
        import area_average_interpolation_b200 as aai

        aai.LIB_PATH = os.path.abspath(lib)


@pytest.fixture(scope="session")
def built():
    """Make sure the in-tree CUDA library and the CPU checkers exist (cross-compiles without a GPU)."""
    import __graft_entry__ as g

    g.build()
    return True
