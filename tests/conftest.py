import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(autouse=True)
def _fresh_globals():
    """Isolate the process-wide singletons between tests."""
    from paddlefleetx_b200.distributed.apis import env
    from paddlefleetx_b200.parallel.rng import get_rng_state_tracker

    env._hcg = None
    get_rng_state_tracker().reset()
    yield
    env._hcg = None
