import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "slow: full-size CPU oracle run")


@pytest.fixture(scope="session")
def repo_root():
    return REPO


@pytest.fixture(autouse=True)
def _reset_step_rng_salt(request):
    """The dropout / sampler kernels add a device "salt" to their seeds (ops.StepRng; training steps advance it).  Kernel
    tests compare against CPU replicas that assume salt 0, so every GPU test starts from a zero salt whatever ran before."""
    if request.node.get_closest_marker("gpu") is not None:
        import torch
        if torch.cuda.is_available():
            from spmm_b200 import ops
            ops.step_rng("cuda").reset(0)
    yield
