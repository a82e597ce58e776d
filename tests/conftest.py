import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """A plain `pytest` on a machine without CUDA reports the GPU tests as skipped, not as errors.  On the GPU box
    (`-m gpu`, CUDA present) nothing is skipped here; LQPB_REQUIRE_GPU=1 turns the skip into a hard failure so that a
    box that lost its device cannot pass silently."""
    import torch
    if torch.cuda.is_available():
        return
    if os.environ.get("LQPB_REQUIRE_GPU") == "1":
        raise pytest.UsageError("LQPB_REQUIRE_GPU=1 but torch.cuda.is_available() is False")
    skip = pytest.mark.skip(reason="needs a CUDA device (B200); there is no CPU fallback")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
