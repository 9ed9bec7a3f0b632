import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "simple-diffusion-model_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")

# The captured graphs time the 128-channel conv variants and switch the library's process-wide choice (b200/autotune.py); tests
# that compare an eager run with a later graph replay bit for bit need ONE choice per process, so the suite pins the default and
# tests/test_autotune_gpu.py exercises the tuner explicitly.
os.environ.setdefault("SDM_B200_AUTOTUNE", "0")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def load_golden(name):
    import torch
    return torch.load(os.path.join(GOLDEN, name), map_location="cpu", weights_only=False)


def rel_l2(a, b):
    import torch
    a = a.detach().double().flatten()
    b = b.detach().double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))
