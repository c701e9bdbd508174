import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture
def host_ops_on_cpu():
    """TEST-ONLY: let the host-side mirror run on CPU tensors through the library statements
    (tests/ops_lib.py) so its wiring can be checked against the reference goldens without a GPU.  The product
    has no such switch: ops.py has one (sm_100a) implementation per op and require_cuda raises on CPU tensors."""
    from gedepth_b200 import ops
    from tests import ops_lib
    restore = ops_lib.install(ops)
    yield
    restore()


@pytest.fixture(autouse=True)
def _deterministic_inputs():
    """Every test draws its random tensors (including `randn_like` upstream gradients on the GPU) from the same
    seeds on every run, so a pass on one box is a pass on the next."""
    import torch
    torch.manual_seed(20240917)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(20240917)
    yield
