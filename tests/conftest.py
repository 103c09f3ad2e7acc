import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _cuda_up(request):
    """before the first GPU test: ride out a transient CUDA initialisation failure of a fresh box"""
    if any(item.get_closest_marker("gpu") for item in request.session.items):
        import shutil
        if shutil.which("nvidia-smi"):
            import __graft_entry__ as graft
            graft.wait_for_cuda()
    yield
