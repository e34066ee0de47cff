"""Shared pieces of the three wrappers: device resolution, engine sharing, checkpoint key clean-up,
sampler-noise draw (the RNG contract of the C ABI)."""
from __future__ import annotations

from collections import OrderedDict

import torch

from ..engine import Engine

_ENGINES = {}


def resolve_device(gpu):
    """The reference does ``torch.device(args.gpu)`` with ``--gpu`` an int (trainer/options.py:5).
    Only CUDA devices are served: there is no CPU path."""
    dev = torch.device(gpu) if not isinstance(gpu, torch.device) else gpu
    if dev.type != "cuda":
        raise RuntimeError("talkshow_b200 runs on CUDA devices only (args.gpu=%r); there is no CPU fallback" % (gpu,))
    return torch.device("cuda", dev.index or 0)


def shared_engine(device):
    """One ts_engine per device, shared by the wrappers of a process (weights of body and face live
    side by side like in scripts/demo.py)."""
    idx = device.index or 0
    if idx not in _ENGINES:
        _ENGINES[idx] = Engine(idx)
    return _ENGINES[idx]


def strip_module(sd):
    """'module.' prefixes come from the reference's nn.DataParallel wrapping
    (nets/smplx_body_pixel.py:117-126)."""
    return OrderedDict((k.replace("module.", ""), v) for k, v in sd.items())


def draw_sampler_noise(T, B, device, generator=None, per_step=True):
    """Exp(1) noise ``q`` for the 2T categorical draws, [2T,B,2048].

    The reference draws ``probs.multinomial(1)`` once per sampled position, which consumes one
    ``exponential_`` of shape [B,2048] from the default generator of the tensor's device
    (gated_pixelcnn_v2.py:173-176).  ``per_step=True`` issues exactly those calls, so under the same
    ``torch.manual_seed`` the stream equals the reference's on that device type; ``False`` draws the
    whole block in one call (same distribution, different stream)."""
    noise = torch.empty(2 * T, B, 2048, device=device)
    if per_step:
        for s in range(2 * T):
            noise[s] = torch.empty(B, 2048, device=device).exponential_(1, generator=generator)
    else:
        noise.exponential_(1, generator=generator)
    return noise
