import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "multigpu: needs at least two CUDA devices (NCCL edge of the batch sharding)")


def _cuda_devices():
    try:
        import torch
        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    """`gpu` tests are skipped (not failed) on a box without a CUDA device, `multigpu` tests with fewer than two.
    On a GPU box nothing is skipped: the CUDA path is the only path (no CPU fallback exists)."""
    n = _cuda_devices()
    no_gpu = pytest.mark.skip(reason="no CUDA device visible (the product path has no CPU fallback)")
    no_multi = pytest.mark.skip(reason="needs >= 2 CUDA devices")
    for item in items:
        if "gpu" in item.keywords and n == 0:
            item.add_marker(no_gpu)
        if "multigpu" in item.keywords and n < 2:
            item.add_marker(no_multi)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
