import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure): built on demand from oracle/einsums_oracle.c."""
    from oracle import binding
    binding.lib()
    return binding


@pytest.fixture(scope="session")
def eb():
    """The product library through its Python door; builds it in-tree when missing."""
    import einsums_b200
    from einsums_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    _lib.lib()
    return einsums_b200


@pytest.fixture(scope="session")
def gpu(eb):
    """Initialised device 0; fails loudly (never skips to a CPU path) when there is no sm_100 device."""
    import torch
    from einsums_b200 import ops
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    ops.init(0)
    torch.cuda.set_device(0)
    return ops
