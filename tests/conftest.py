import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def ckpts():
    """Seeded synthetic checkpoints in the reference's state-dict format (talkshow_b200/synth.py)."""
    import torch

    from talkshow_b200 import synth

    torch.set_grad_enabled(False)
    return {
        "pixel": synth.body_pixel_checkpoint(0),
        "vq": synth.body_vq_checkpoint(0),
        "face": synth.face_checkpoint(0),
    }


def draw_noise(steps, B, seed, K=2048):
    """RNG contract: one exponential_ of shape [B,K] per sampled position, reference order
    (gated_pixelcnn_v2.py:167-176)."""
    import torch

    torch.manual_seed(seed)
    out = torch.empty(steps, B, K)
    for s in range(steps):
        out[s] = torch.empty(B, K).exponential_(1)
    return out


def noise_fp(noise):
    import numpy as np

    return np.array([float(noise.double().sum()), float(noise[0, 0, :8].double().sum()),
                     float(noise[-1, -1, -8:].double().sum())])
