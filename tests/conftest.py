import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# no SD2.1 checkpoint exists offline: the suites run on deterministic random-init weights (an explicit opt-in, see
# faceposegenerator_b200.pipeline.random_weights_allowed)
os.environ.setdefault("IDB_ALLOW_RANDOM_WEIGHTS", "1")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def cuda_dev():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("gpu-marked test started without a CUDA device")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    from faceposegenerator_b200 import _lib
    _lib.check(_lib.load().idb_device_check(), "idb_device_check")
    return torch.device("cuda:0")
