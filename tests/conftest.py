import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with `-m gpu` under gpurun)")


@pytest.fixture(scope="session")
def engine():
    """The initialised engine; fails loudly (no CPU fallback) without a GPU."""
    from firedrake_b200 import _lib
    return _lib.init(0)


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as orc
    orc.build()
    return orc


@pytest.fixture(params=[0, 1], ids=["sumfact", "dmma"])
def matrix_kernel(request, engine):
    """Both element-matrix kernels: the sum-factorised column kernel (action_hex.cu, MATRIX mode)
    and the dense B^T D B kernel on the fp64 tensor pipe (bdb_matrix.cu, degrees 2..4)."""
    from firedrake_b200 import _lib
    _lib.check(engine.fdb_set_option(b"matrix_kernel", request.param))
    yield request.param
    _lib.check(engine.fdb_set_option(b"matrix_kernel", -1))
