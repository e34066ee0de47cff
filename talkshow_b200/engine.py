"""Thin object wrapper over the C ABI: PyTorch CUDA tensors in, PyTorch CUDA tensors out.

Module-level mirror of the reference network modules (nets/spg/*.py) — used by the wrapper classes
in talkshow_b200/nets/ and directly by the parity tests.  Every method enqueues on the current
CUDA stream of the engine's device and returns without synchronising.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib

PIX_TILE = 64  # samples per PixelCNN launch (csrc/pixelcnn.h PIX_MB)


class Engine:
    def __init__(self, device=0):
        self.L = _lib.lib()
        h = C.c_void_p()
        dev = device if isinstance(device, int) else torch.device(device).index or 0
        rc = self.L.ts_engine_create(C.byref(h), dev)
        if rc:
            raise RuntimeError("ts_engine_create(%d): %s" % (dev, self.L.ts_last_error(None).decode()))
        self.h = h
        self.host_only = dev < 0
        self.device = torch.device("cuda", dev) if dev >= 0 else torch.device("cpu")
        self.loaded = set()

    def close(self):
        if getattr(self, "h", None):
            self.L.ts_engine_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- helpers ---------------------------------------------------------------------------------
    def _check(self, rc, what):
        if rc:
            raise RuntimeError("%s failed (status %d): %s" % (what, rc, self.L.ts_last_error(self.h).decode()))

    def _s(self):
        return _lib.stream_ptr(self.device)

    def _dev(self, t, dtype):
        return t.to(device=self.device, dtype=dtype).contiguous()

    @property
    def sm_count(self):
        return self.L.ts_engine_sm_count(self.h)

    @property
    def launches(self):
        return int(self.L.ts_launch_count(self.h))

    @property
    def pixelcnn_row_bytes(self):
        return int(self.L.ts_pixelcnn_row_bytes(self.h))

    @property
    def pixelcnn_staged_row_bytes(self):
        return int(self.L.ts_pixelcnn_staged_row_bytes(self.h))

    def pixelcnn_timing(self, enable=True):
        self._check(self.L.ts_pixelcnn_timing(self.h, int(enable)), "ts_pixelcnn_timing")

    def pixelcnn_last_ms(self):
        return float(self.L.ts_pixelcnn_last_ms(self.h))

    def set_tensor_cores(self, enable=True):
        self._check(self.L.ts_set_tensor_cores(self.h, int(enable)), "ts_set_tensor_cores")

    def pixelcnn_trace(self, row):
        """Arm (row >= 0) / disarm (row < 0) the per-stage, per-CTA time stamps of one latent row (debug)."""
        self._check(self.L.ts_pixelcnn_trace(self.h, int(row)), "ts_pixelcnn_trace")

    def pixelcnn_trace_read(self):
        """-> int64 tensor [stages, ctas, 4] (ns): weights ready, left the grid barrier, task done, arrived."""
        import ctypes as C
        n = C.c_int64(0)
        self._check(self.L.ts_pixelcnn_trace_read(self.h, None, C.byref(n)), "ts_pixelcnn_trace_read")
        buf = torch.zeros(n.value, dtype=torch.int64)
        self._check(self.L.ts_pixelcnn_trace_read(self.h, buf.data_ptr(), C.byref(n)), "ts_pixelcnn_trace_read")
        ns, nc = C.c_int(0), C.c_int(0)
        self._check(self.L.ts_pixelcnn_plan_shape(self.h, C.byref(ns), C.byref(nc)), "ts_pixelcnn_plan_shape")
        return buf.view(ns.value, nc.value, 4)

    def set_pixelcnn_fusion(self, on):
        """Plan built by the next load_pixelcnn: 1/True fused 52-stage (default), 0/False plain 84-stage,
        2 fused with vert_to_horiz scheduled in the horizontal pass (experimental)."""
        self._check(self.L.ts_set_pixelcnn_fusion(self.h, int(on)), "ts_set_pixelcnn_fusion")

    def set_pixelcnn_ctas(self, n):
        """Persistent CTAs of the grid-wide sampler plan built by the next load_pixelcnn (0 = one per SM)."""
        self._check(self.L.ts_set_pixelcnn_ctas(self.h, int(n)), "ts_set_pixelcnn_ctas")

    def set_vq_parallel(self, max_batch):
        """body_generate: run the two VQ decoders of batches <= max_batch side by side on two streams (0: one after the other)."""
        self._check(self.L.ts_set_vq_parallel(self.h, int(max_batch)), "ts_set_vq_parallel")

    def set_pixelcnn_mode(self, mode):
        self._check(self.L.ts_set_pixelcnn_mode(self.h, mode), "ts_set_pixelcnn_mode")

    # -- weights ---------------------------------------------------------------------------------
    def load_pixelcnn(self, sd):
        arr, keep = _lib.pack_tensors(sd)
        self._check(self.L.ts_load_pixelcnn(self.h, arr, len(sd)), "ts_load_pixelcnn")
        self.loaded.add("pixelcnn")

    def load_audioenc(self, sd):
        arr, keep = _lib.pack_tensors(sd)
        self._check(self.L.ts_load_audioenc(self.h, arr, len(sd)), "ts_load_audioenc")
        self.loaded.add("audioenc")

    def load_vq(self, which, sd):
        arr, keep = _lib.pack_tensors(sd)
        self._check(self.L.ts_load_vq(self.h, which, arr, len(sd)), "ts_load_vq")
        self.loaded.add("vq%d" % which)

    def load_face(self, sd):
        sd = dict(sd)
        p = "audio_encoder.encoder.pos_conv_embed.conv."
        if p + "weight" not in sd:      # resolve the weight_norm parametrisation (dim=2), old or new names
            if p + "parametrizations.weight.original0" in sd:
                g, v = sd.pop(p + "parametrizations.weight.original0"), sd.pop(p + "parametrizations.weight.original1")
            elif p + "weight_g" in sd and p + "weight_v" in sd:
                g, v = sd.pop(p + "weight_g"), sd.pop(p + "weight_v")
            else:
                raise RuntimeError("ts_load_face: checkpoint tensor '%sweight' (or its weight_norm pair weight_g / weight_v, "
                                   "parametrizations.weight.original0 / original1) missing" % p)
            sd[p + "weight"] = torch._weight_norm(v.float().cpu(), g.float().cpu(), 2)
        arr, keep = _lib.pack_tensors(sd)
        self._check(self.L.ts_load_face(self.h, arr, len(sd)), "ts_load_face")
        self.loaded.add("face")

    # -- modules ---------------------------------------------------------------------------------
    def latent_rows(self, M):
        return self.L.ts_latent_rows(M)

    def mfcc(self, wave, sr):
        """get_mfcc_ta's transform chain on the device: wave [B,N] mono at ``sr`` Hz -> [B,64,M] (30 fps)."""
        wave = self._dev(wave, torch.float32)
        B, N = wave.shape
        M = self.L.ts_mfcc_frames(N, int(sr))
        out = torch.empty(B, 64, M, device=self.device)
        self._check(self.L.ts_mfcc(self.h, _lib.ptr(wave), _lib.ptr(out), B, N, int(sr), self._s()), "ts_mfcc")
        return out

    def audio_encode(self, mfcc):
        """AudioEncoder.forward: [B,64,M] -> [B,256,T]."""
        mfcc = self._dev(mfcc, torch.float32)
        B, _, M = mfcc.shape
        out = torch.empty(B, 256, self.latent_rows(M), device=self.device)
        self._check(self.L.ts_audio_encode(self.h, _lib.ptr(mfcc), _lib.ptr(out), B, M, self._s()), "ts_audio_encode")
        return out

    def pixelcnn_generate(self, aud, label, noise, T=None, pre_latents=None, want_logits=False):
        """GatedPixelCNN.generate: aud [B,256,T0+T], label [B], noise [2T,B,2048] -> codes [B,T,2]
        (+ logits [2T,B,2048]).  Batches above the 64-sample tile are chunked here."""
        aud = self._dev(aud, torch.float32)
        label = self._dev(label, torch.int64)
        noise = self._dev(noise, torch.float32)
        B = aud.shape[0]
        T0 = 0 if pre_latents is None else pre_latents.shape[1]
        T = aud.shape[2] - T0 if T is None else T
        if label.numel() == 1 and B > 1:
            label = label.expand(B).contiguous()
        pre = None if pre_latents is None else self._dev(pre_latents, torch.int64)
        codes = torch.empty(B, T, 2, dtype=torch.int64, device=self.device)
        logits = torch.empty(2 * T, B, 2048, device=self.device) if want_logits else None
        aud_c = aud.shape[1]
        assert aud_c == 256, "AudioEncoder output has 256 channels"
        for b0 in range(0, B, PIX_TILE):
            b1 = min(B, b0 + PIX_TILE)
            nb = b1 - b0
            full = nb == B
            nz = noise if full else noise[:, b0:b1].contiguous()
            cz = codes if full else torch.empty(nb, T, 2, dtype=torch.int64, device=self.device)
            lz = logits if (full or logits is None) else torch.empty(2 * T, nb, 2048, device=self.device)
            self._check(self.L.ts_pixelcnn_generate(
                self.h, _lib.ptr(aud[b0:b1].contiguous()), _lib.ptr(label[b0:b1].contiguous()), _lib.ptr(nz),
                _lib.ptr(cz), _lib.ptr(lz), nb, T, _lib.ptr(None if pre is None else pre[b0:b1].contiguous()), T0,
                self._s()), "ts_pixelcnn_generate")
            if not full:
                codes[b0:b1] = cz
                if logits is not None:
                    logits[:, b0:b1] = lz
        return (codes, logits) if want_logits else codes

    def pixelcnn_logits(self, aud, label, codes):
        """GatedPixelCNN.forward (teacher forced): -> logits [B,2048,T,2]."""
        aud = self._dev(aud, torch.float32)
        label = self._dev(label, torch.int64)
        codes = self._dev(codes, torch.int64)
        B, _, T = aud.shape
        assert B <= PIX_TILE
        out = torch.empty(B, 2048, T, 2, device=self.device)
        self._check(self.L.ts_pixelcnn_logits(self.h, _lib.ptr(aud), _lib.ptr(label), _lib.ptr(codes), _lib.ptr(out),
                                              B, T, self._s()), "ts_pixelcnn_logits")
        return out

    def vq_dim(self, which):
        """pose channels of the loaded VQ-VAE: 39 / 90 (axis-angle), 78 / 180 (convert_to_6d)."""
        c = int(self.L.ts_vq_dim(self.h, which))
        if c <= 0:
            raise RuntimeError("vq weights %d not loaded" % which)
        return c

    def vq_decode(self, which, idx):
        """VQVAE.decode(latents=idx): [B,T] -> [B,C,4T]."""
        idx = self._dev(idx, torch.int64)
        B, T = idx.shape
        C_ = self.vq_dim(which)
        out = torch.empty(B, C_, 4 * T, device=self.device)
        self._check(self.L.ts_vq_decode(self.h, which, _lib.ptr(idx), _lib.ptr(out), B, T, self._s()), "ts_vq_decode")
        return out

    def vq_encode(self, which, poses, want_e=False):
        """VQVAE.encode: poses [B,F,C] -> idx [B,T] (and e [B,64,T])."""
        poses = self._dev(poses, torch.float32)
        B, F, _ = poses.shape
        T = self.latent_rows(F)
        idx = torch.empty(B, T, dtype=torch.int64, device=self.device)
        e = torch.empty(B, 64, T, device=self.device) if want_e else None
        self._check(self.L.ts_vq_encode(self.h, which, _lib.ptr(poses), _lib.ptr(idx), _lib.ptr(e), B, F, self._s()),
                    "ts_vq_encode")
        return (idx, e) if want_e else idx

    def face_forward(self, wave, id_onehot, frame):
        """s2g_face.Generator.forward: wave [B,N], id [B,4] -> [B,frame,103]."""
        wave = self._dev(wave, torch.float32)
        B, N = wave.shape
        idv = self._dev(id_onehot, torch.float32)
        if idv.shape[0] == 1 and B > 1:
            idv = idv.expand(B, -1).contiguous()
        out = torch.empty(B, frame, 103, device=self.device)
        self._check(self.L.ts_face_forward(self.h, _lib.ptr(wave), _lib.ptr(idv), _lib.ptr(out), B, N, frame,
                                           self._s()), "ts_face_forward")
        return out

    def body_generate(self, mfcc, label, noise, want_codes=True):
        """fused s2g_body_pixel core: mfcc [B,64,M] -> (codes [B,T,2], poses [B,4T,C]), C = 129 (258 for 6-D)."""
        mfcc = self._dev(mfcc, torch.float32)
        label = self._dev(label, torch.int64)
        noise = self._dev(noise, torch.float32)
        B, _, M = mfcc.shape
        if label.numel() == 1 and B > 1:
            label = label.expand(B).contiguous()
        T = self.latent_rows(M)
        codes = torch.empty(B, T, 2, dtype=torch.int64, device=self.device) if want_codes else None
        poses = torch.empty(B, 4 * T, self.vq_dim(0) + self.vq_dim(1), device=self.device)
        for b0 in range(0, B, PIX_TILE):
            b1 = min(B, b0 + PIX_TILE)
            full = (b1 - b0) == B
            nz = noise if full else noise[:, b0:b1].contiguous()
            self._check(self.L.ts_body_generate(
                self.h, _lib.ptr(mfcc[b0:b1]), _lib.ptr(label[b0:b1]), _lib.ptr(nz),
                _lib.ptr(None if codes is None else codes[b0:b1]), _lib.ptr(poses[b0:b1]), b1 - b0, M, self._s()),
                "ts_body_generate")
        return codes, poses

    # -- the single collective of the path, inside the library (ts_allgather over a dlopen'ed NCCL) ----------------
    @staticmethod
    def _libnccl_path():
        import glob
        import os

        base = os.path.dirname(os.path.dirname(torch.__file__))
        hits = glob.glob(os.path.join(base, "nvidia", "nccl", "lib", "libnccl.so*"))
        return (hits[0] if hits else "libnccl.so.2").encode()

    def nccl_init(self, rank, world, group=None):
        """Create this engine's NCCL communicator.  The 128-byte unique id is made on rank 0 and broadcast with
        torch.distributed (any backend: it is 128 bytes of host data), every rank then joins."""
        import torch.distributed as dist

        path = self._libnccl_path()
        buf = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            self._check(self.L.ts_nccl_unique_id(self.h, path, buf.data_ptr()), "ts_nccl_unique_id")
        if world > 1:
            obj = [buf.numpy().tobytes()]
            dist.broadcast_object_list(obj, src=0, group=group)
            buf = torch.frombuffer(bytearray(obj[0]), dtype=torch.uint8).clone()
        self._check(self.L.ts_nccl_init(self.h, path, buf.data_ptr(), rank, world), "ts_nccl_init")
        self.nccl_world = world
        return self

    def allgather(self, local):
        """local [n, ...] fp32 (same shape on every rank) -> [world * n, ...] through ts_allgather (NCCL, current stream)."""
        local = self._dev(local, torch.float32)
        world = getattr(self, "nccl_world", 0)
        if not world:
            raise RuntimeError("Engine.nccl_init was not called")
        out = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), device=self.device)
        self._check(self.L.ts_allgather(self.h, _lib.ptr(local), _lib.ptr(out), local.numel(), self._s()), "ts_allgather")
        return out

    def rot6d_to_axis_angle(self, d6):
        """matrix_to_axis_angle(rotation_6d_to_matrix(d6)) (rotation_conversion.py:512-533,433-447): [...,6] -> [...,3]."""
        d6 = self._dev(d6, torch.float32)
        if d6.shape[-1] != 6:
            raise ValueError("rot6d_to_axis_angle: last dim must be 6, got %s" % (tuple(d6.shape),))
        out = torch.empty(d6.shape[:-1] + (3,), device=self.device)
        self._check(self.L.ts_rot6d_to_axis_angle(self.h, _lib.ptr(d6), _lib.ptr(out), d6.numel() // 6, self._s()),
                    "ts_rot6d_to_axis_angle")
        return out

    def assemble_pose(self, face, body, stand=False):
        """demo.py:182-229 + part2full: face [B,Ff,103], body [B,Fb,129] -> [B,Ff,265]."""
        face = self._dev(face, torch.float32)
        body = self._dev(body, torch.float32)
        B, Ff, _ = face.shape
        out = torch.empty(B, Ff, 265, device=self.device)
        self._check(self.L.ts_assemble_pose(self.h, _lib.ptr(face), _lib.ptr(body), _lib.ptr(out), B, Ff,
                                            body.shape[1], int(stand), self._s()), "ts_assemble_pose")
        return out
