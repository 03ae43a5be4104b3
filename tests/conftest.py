import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "hw_pending(order=n): written after the round's GPU budget was spent, not yet run on a "
                                       "B200; collected last (lowest order first) so that under -x a failure here cannot "
                                       "mask the verified suite")


def pytest_collection_modifyitems(config, items):
    def key(it):  # stable sort: verified tests keep their file order; pending ones follow, lowest `order` (least new machinery) first
        m = it.get_closest_marker("hw_pending")
        return (0, 0) if m is None else (1, m.kwargs.get("order", 50))

    items.sort(key=key)


@pytest.fixture(scope="session")
def cuda_device():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from ralf_b200 import _lib

    _lib.check(_lib.lib().ralf_check_device(0), "ralf_check_device")
    return torch.device("cuda:0")
