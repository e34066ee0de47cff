"""GPU (-m gpu), run last: non-default executors of the PixelCNN sampler.

The cluster plan (ts_set_pixelcnn_mode(3)) was validated on a B200 with scratch/test_cluster.py (same 9 600 codes as
the default kernel at B=64 x T=75).  Kept in a separate file so that the default-path parity tests all run first.
"""
import pytest
import torch

import talkshow_oracle as O
from conftest import draw_noise
from talkshow_b200 import synth

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(scope="module")
def eng(ckpts):
    from talkshow_b200.engine import Engine

    e = Engine(0)
    e.load_pixelcnn(ckpts["pixel"]["generator"])
    yield e
    torch.cuda.synchronize()
    e.close()


def test_pixelcnn_cluster_plan(eng, ckpts):
    """EXPERIMENTAL executor (ts_set_pixelcnn_mode(3) before the load): 4-CTA clusters split K and reduce through
    distributed shared memory.  Same sampled sequences as the oracle and the default kernel."""
    from talkshow_b200.engine import Engine

    e = Engine(0)
    e.set_pixelcnn_mode(3)
    try:
        try:
            e.load_pixelcnn(ckpts["pixel"]["generator"])
        except RuntimeError as ex:          # fewer resident clusters than the plan needs on this part
            pytest.skip("cluster plan not available: %s" % ex)
        B, T = 5, 20
        label = torch.tensor([0, 1, 2, 3, 1])
        aud = O.audio_encoder(ckpts["pixel"]["audioencoder"], synth.synth_mfcc(B, 4 * T, seed=41))
        noise = draw_noise(2 * T, B, 17)
        ref = O.pixelcnn_generate(ckpts["pixel"]["generator"], label, T, B, aud.unsqueeze(-1).repeat(1, 1, 1, 2),
                                  noise=noise, window=18)
        got, lc = e.pixelcnn_generate(aud, label, noise, want_logits=True)
        base, lb = eng.pixelcnn_generate(aud, label, noise, want_logits=True)
        assert torch.equal(got.cpu(), ref)
        assert torch.equal(base.cpu(), ref)
        assert (lc - lb).abs().max().item() <= TOL
    finally:
        torch.cuda.synchronize()
        e.close()
