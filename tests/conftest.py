import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run on the GPU box via gpurun)")


def _no_b200_reason():
    """None when an sm_100 device is visible; otherwise why the gpu-marked tests cannot run here.  A missing
    library on a box that HAS a B200 is not a reason to skip: those tests must then fail loudly."""
    try:
        import torch
        if not torch.cuda.is_available():
            return "no CUDA device"
        major, _ = torch.cuda.get_device_capability(0)
        if major != 10:
            return f"device is sm_{major}x, the kernels are sm_100a"
    except Exception as e:   # pragma: no cover
        return f"torch.cuda unusable: {e}"
    return None


def pytest_collection_modifyitems(config, items):
    reason = _no_b200_reason()
    if reason is None:
        return
    skip = pytest.mark.skip(reason=f"needs a B200: {reason}")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
