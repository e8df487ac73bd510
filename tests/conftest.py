import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with `-m gpu` under gpurun)")


@pytest.fixture(scope="session")
def oracle_service():
    from oracle_service import OracleService
    return OracleService()


@pytest.fixture(scope="session")
def cuda_service():
    """The product service on cuda:0.  Never skips: a GPU test on a box without a working
    B200 / librindow_b200.so must FAIL (there is no CPU fallback to fall back to)."""
    import rindow_math_matrix_b200 as rb
    svc = rb.MatlibCuda()
    assert svc.serviceLevel() >= rb.Service.LV_ADVANCED, "no B200 visible / librindow_b200.so missing"
    return svc
