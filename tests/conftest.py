import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "hw_pending(order=n, optin=False): written after the round's GPU budget was spent, "
                                       "not yet run on a B200; collected last (lowest order first) so that under -x a "
                                       "failure here cannot mask the verified suite; optin=True additionally needs "
                                       "RALF_TEST_OPTIN=1")


def pytest_collection_modifyitems(config, items):
    def key(it):  # stable sort: verified tests keep their file order; pending ones follow, lowest `order` (least new machinery) first
        m = it.get_closest_marker("hw_pending")
        return (0, 0) if m is None else (1, m.kwargs.get("order", 50))

    items.sort(key=key)
    if not os.environ.get("RALF_TEST_OPTIN"):
        # hw_pending(optin=True): first run belongs in a deliberate, bounded GPU call (profiles/run_r2_first_call.sh sets
        # RALF_TEST_OPTIN=1), not in the unattended suite: opt-in throughput variants, new tensor-core kernels (a wrong
        # barrier there is a hang, not a failure), and features whose host side is already pinned on the CPU.
        skip = pytest.mark.skip(reason="first hardware run is opt-in: RALF_TEST_OPTIN=1 (profiles/run_r2_first_call.sh)")
        for it in items:
            m = it.get_closest_marker("hw_pending")
            if m is not None and m.kwargs.get("optin"):
                it.add_marker(skip)


@pytest.fixture(scope="session")
def cuda_device():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from ralf_b200 import _lib

    _lib.check(_lib.lib().ralf_check_device(0), "ralf_check_device")
    return torch.device("cuda:0")
