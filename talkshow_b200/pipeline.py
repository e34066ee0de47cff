"""Whole-body batch generation (what scripts/demo.py:158-229 does per clip, batched) and its
multi-GPU sharding.

A batch of clips / diversity samples is a set of independent autoregressive chains:
rank r of G takes a contiguous slice of the batch, runs face + body + pose assembly on its own
GPU, and ONE NCCL all-gather of the final [b,F,265] pose tensor rebuilds the batch on every rank.
No other collective is on the path.
"""
from __future__ import annotations

import torch

from .engine import Engine
from .nets.base import draw_sampler_noise


def shard_range(B, rank, world):
    """Contiguous slice of the batch owned by ``rank`` (first B % world ranks get one extra)."""
    base, rem = divmod(B, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class WholeBody:
    """face (jaw+expression) + body/hands + part2full assembly -> SMPL-X parameters [B,F,265].

    ``overlap_batch`` > 0: batches up to that size run the body path and the face path SIDE BY SIDE on two streams.  The
    sampler is a serial chain of 3 900 stages that is latency-bound at every batch size (on 96 CTAs it takes 27 ms for 64
    samples, 26 ms on all 148), so a second engine holds a sampler plan for ``overlap_ctas`` persistent CTAs launched as CTA
    pairs (whole TPCs, high-priority stream) and the face kernels fill the remaining TPCs, then the whole GPU.  Measured
    (one B200, 10 s clips): 8 clips 27.3 -> 24.7 ms, 32: 40.9 -> 35.2, 64: 59.2 -> 51.0.  Results are bit-identical to the
    sequential order (same kernels, same arithmetic; tests/test_gpu_baseline_shapes.py)."""

    def __init__(self, engine: Engine, overlap_batch=64, overlap_ctas=96):
        self.e = engine
        self.device = engine.device
        self.overlap_batch = overlap_batch if not getattr(engine, "host_only", True) else 0
        self.overlap_ctas = overlap_ctas
        self.e2 = None            # body path engine with the partial-GPU sampler plan (same device)
        self._side = None
        self._body_engine = engine   # the engine whose sampler ran in the last generate()

    def load(self, pixel_ckpt, vq_ckpt, face_ckpt):
        """Checkpoint dicts in the reference's formats (talkshow_b200/synth.py docstring)."""
        self.e.load_pixelcnn(pixel_ckpt["generator"])
        self.e.load_audioenc(pixel_ckpt["audioencoder"])
        self.e.load_vq(0, vq_ckpt["g_body"])
        self.e.load_vq(1, vq_ckpt["g_hand"])
        self.e.load_face(face_ckpt["generator"])
        dim = pixel_ckpt["generator"]["embedding.weight"].shape[1]
        if self.overlap_batch > 0 and dim == 256 and self.overlap_ctas < self.e.sm_count:
            idx = self.device.index or 0
            self.e2 = Engine(idx)
            self.e2.set_pixelcnn_ctas(self.overlap_ctas)
            self.e2.load_pixelcnn(pixel_ckpt["generator"])
            self.e2.load_audioenc(pixel_ckpt["audioencoder"])
            self.e2.load_vq(0, vq_ckpt["g_body"])
            self.e2.load_vq(1, vq_ckpt["g_hand"])
            self._side = torch.cuda.Stream(device=self.device, priority=-1)   # the sampler CTAs are placed before queued face CTAs

    @property
    def launches(self):
        """Kernels launched by the library so far (both engines)."""
        return self.e.launches + (self.e2.launches if self.e2 is not None else 0)

    def pixelcnn_timing(self, enable=True):
        self.e.pixelcnn_timing(enable)
        if self.e2 is not None:
            self.e2.pixelcnn_timing(enable)

    def pixelcnn_last_ms(self):
        """Device time of the sampler launch of the last ``generate`` (events on its launch stream, inside the library)."""
        return self._body_engine.pixelcnn_last_ms()

    def close(self):
        if self.e2 is not None:
            torch.cuda.synchronize(self.device)
            self.e2.close()
            self.e2 = None

    def generate(self, mfcc, wave, label, noise=None, stand=False, per_step_noise=True):
        """mfcc [B,64,M], wave [B,N] (16 kHz), label [B] on the device -> poses [B,F,265] (device).
        F = N*30//16000 face frames; the body (4*T frames) is padded/truncated to F like demo.py:207-211."""
        B, _, M = mfcc.shape
        frame = wave.shape[1] * 30 // 16000
        T = self.e.latent_rows(M)
        if noise is None:
            noise = draw_sampler_noise(T, B, self.device, per_step=per_step_noise)
        idz = torch.zeros(B, 4, device=self.device)                    # demo.py:173-181: id not passed
        if self.e2 is not None and B <= self.overlap_batch:
            # small batch: sampler (on overlap_ctas SMs, launched first) and face regressor (on the free TPCs) side by side
            cur = torch.cuda.current_stream(self.device)
            self._side.wait_stream(cur)
            self._body_engine = self.e2
            with torch.cuda.stream(self._side):
                _, body = self.e2.body_generate(mfcc, label, noise, want_codes=False)
            face = self.e.face_forward(wave, idz, frame)
            cur.wait_stream(self._side)
            body.record_stream(cur)                                      # allocated on the side stream, consumed here
            for t in (mfcc, label, noise):
                if torch.is_tensor(t) and t.is_cuda:
                    t.record_stream(self._side)                          # allocated here, consumed on the side stream
            return self.e.assemble_pose(face, body, stand)
        self._body_engine = self.e
        face = self.e.face_forward(wave, idz, frame)
        _, body = self.e.body_generate(mfcc, label, noise, want_codes=False)
        return self.e.assemble_pose(face, body, stand)

    def generate_sharded(self, mfcc, wave, label, rank, world, noise_full=None, seed=None, stand=False, gather=True,
                         group=None):
        """Multi-GPU form of ``generate`` whose RESULT DOES NOT DEPEND ON ``world``: the arguments are
        the FULL batch (host or device tensors, identical on every rank), the sampler noise is drawn for the full batch in
        the reference's order — ``noise_full`` [2T,B,2048], or drawn here from ``seed`` on this rank's device generator
        (same seed on every rank gives every rank the same stream) — and rank r keeps the slice of its contiguous shard
        ``shard_range(B, r, world)``.  Each rank runs its shard; ONE all-gather rebuilds [B,F,265] everywhere."""
        B, _, M = mfcc.shape
        lo, hi = shard_range(B, rank, world)
        T = self.e.latent_rows(M)
        if noise_full is None:
            g = None
            if seed is not None:
                g = torch.Generator(device=self.device)
                g.manual_seed(int(seed))
            noise_full = draw_sampler_noise(T, B, self.device, generator=g)
        if hi == lo:                                                     # more ranks than samples: nothing to run here
            frame = wave.shape[1] * 30 // 16000
            local = torch.empty(0, frame, 265, device=self.device)
        else:
            dev = lambda t: t[lo:hi].to(self.device, non_blocking=True)
            local = self.generate(dev(mfcc), dev(wave), dev(label), noise=noise_full[:, lo:hi].to(self.device).contiguous(),
                                  stand=stand)
        return allgather_poses(local, B, world, group=group, engine=self.e) if gather else local

    def generate_host(self, mfcc_host, wave_host, label_host, out_host=None, **kw):
        """Public end-to-end call with HOST buffers (pinned for async copies): H2D inputs, generate,
        D2H result.  Returns the host tensor [B,F,265]."""
        mfcc = mfcc_host.to(self.device, non_blocking=True)
        wave = wave_host.to(self.device, non_blocking=True)
        label = label_host.to(self.device, non_blocking=True)
        poses = self.generate(mfcc, wave, label, **kw)
        if out_host is None:
            out_host = torch.empty(poses.shape, dtype=poses.dtype, pin_memory=True)
        out_host.copy_(poses, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return out_host


def allgather_poses(local, B_total, world, group=None, engine=None):
    """ONE all-gather of the pose tensor over NCCL (NVLink/NVSwitch).  local [b,F,265] -> [B_total,F,265].
    Uneven shards are padded to the largest shard for the collective and trimmed afterwards.  With an ``engine`` whose
    communicator was created (``Engine.nccl_init``) the collective is the library's own ``ts_allgather``; otherwise
    torch.distributed's ``all_gather_into_tensor`` (NCCL on GPUs, gloo in the CPU tests)."""
    import torch.distributed as dist

    if world == 1:
        return local
    bmax = -(-B_total // world)
    if local.shape[0] < bmax:
        pad = torch.zeros((bmax - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        local = torch.cat([local, pad], 0)
    if engine is not None and getattr(engine, "nccl_world", 0) == world:
        out = engine.allgather(local.contiguous())
    else:
        out = torch.empty((world * bmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    if B_total == world * bmax:
        return out
    parts = []
    for r in range(world):
        lo, hi = shard_range(B_total, r, world)
        parts.append(out[r * bmax: r * bmax + (hi - lo)])
    return torch.cat(parts, 0)
