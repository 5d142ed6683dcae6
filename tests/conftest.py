import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _ensure_library():
    """The C-ABI library is a build artefact (git-ignored).  Build it once if it is missing so that the CPU suite
    can check the ABI; on the GPU box the prebuilt .so travels with the snapshot."""
    so = os.path.join(ROOT, "microaligner_b200", "libmicroaligner_b200.so")
    if not os.path.exists(so):
        import build
        build.build()


def pytest_configure(config):
    _ensure_library()
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs the read-only reference checkout at /root/reference")


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
