import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from tests import _oracle
    return _oracle.load_oracle()


@pytest.fixture(scope="session")
def ref_host():
    from tests import _oracle
    lib = _oracle.load_ref_host()
    if lib is None:
        pytest.skip("oracle/_ref/libref_tsdf_host.so not built (reference tree absent)")
    return lib


@pytest.fixture(scope="session")
def ref_coloration():
    from tests import _oracle
    lib = _oracle.load_ref_coloration()
    if lib is None:
        pytest.skip("oracle/_ref/libref_coloration.so not built (reference tree absent)")
    return lib


@pytest.fixture(scope="session")
def gpu_ctx():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from cudadepthmapintegration_b200 import Context
    ctx = Context(0)
    yield ctx
    ctx.close()
