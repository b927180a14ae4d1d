import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(autouse=True)
def _reset_dropout_salt(request):
    """GPU tests compare against the host dropout mirror at salt 0; captured training steps advance the device salt."""
    yield
    if request.node.get_closest_marker("gpu") is not None:
        import torch
        if torch.cuda.is_available():
            from get_b200 import ops
            ops.dropout_salt_set(0)
            ops.GRAD_SINK.clear()
            ops.GRAD_READY_HOOK = None
            ops.set_precision("fp32")
