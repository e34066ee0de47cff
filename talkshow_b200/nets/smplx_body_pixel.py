"""s2g_body_pixel — host mirror of the reference wrapper nets/smplx_body_pixel.py (inference
methods).  Audio features -> AudioEncoder -> gated-PixelCNN sampler -> two VQ-VAE decoders, all
inside the CUDA engine (C ABI: ts_body_generate / ts_pixelcnn_generate / ts_vq_decode)."""
from __future__ import annotations

import os

import numpy as np
import torch

from ..data_utils.lower_body import c_index_3d, c_index_6d
from ..data_utils.utils import get_mfcc_sepa, get_mfcc_ta, load_wav
from .base import draw_sampler_noise, resolve_device, shared_engine, strip_module


class TrainWrapper:
    """Constructor/attribute surface of nets/smplx_body_pixel.py:30-75 (args.gpu, args.infer,
    config.Data.pose.*, config.Model.{composition,bh_model,code_num,vq_path})."""

    def __init__(self, args, config, engine=None):
        self.args = args
        self.config = config
        self.device = resolve_device(self.args.gpu)
        self.global_step = 0
        self.convert_to_6d = self.config.Data.pose.convert_to_6d
        self.expression = self.config.Data.pose.expression
        self.epoch = 0
        self.init_params()
        self.num_classes = 4
        self.audio = True
        self.composition = self.config.Model.composition
        self.bh_model = self.config.Model.bh_model
        if not self.bh_model or not self.composition:
            raise NotImplementedError("talkshow_b200 builds the bh_model=true, composition=true prior "
                                      "(config/body_pixel.json); convert_to_6d selects the dim 512 x 10-layer geometry")
        self.c_index = c_index_6d if self.convert_to_6d else c_index_3d      # :72-75
        self.engine = engine or shared_engine(self.device)
        self.noise_device = self.device      # 'cpu' reproduces the CPU reference's RNG stream
        self.noise_per_step = True
        self.device_mfcc = True              # MFCC features on the device (ts_mfcc); False = host torchaudio chain
        # the reference loads the VQ-VAE checkpoint at construction (:59-62); honour it when present
        vq_path = getattr(self.config.Model, "vq_path", None)
        if vq_path and os.path.exists(vq_path):
            ck = torch.load(vq_path, map_location="cpu")["generator"]
            self.load_vq_state_dict(ck)

    def init_params(self):
        """nets/smplx_body_pixel.py:144-174: body 39 + hands 90 axis-angle dims, doubled for the 6-D layout
        (convert_to_6d: the prior is pixelcnn(2048, 512, 10, ...) instead of (2048, 256, 15, ...), :49-52)."""
        scale = 2 if self.convert_to_6d else 1
        body, hand = 39 * scale, 90 * scale
        self.each_dim = [0, body, hand, 100 if self.expression else 0]
        self.dim_list = [0, 0, 0, body, body + hand]
        self.full_dim = body + hand
        self.pose = int(self.full_dim / round(3 * scale))

    # -- checkpoints -----------------------------------------------------------------------------
    def load_vq_state_dict(self, sd):
        """{'g_body': ..., 'g_hand': ...} as saved by s2g_body_vq.state_dict()."""
        self.engine.load_vq(0, strip_module(sd["g_body"]))
        self.engine.load_vq(1, strip_module(sd["g_hand"]))

    def load_state_dict(self, state_dict):
        """ckpt['generator'] of a body-pixel checkpoint (nets/smplx_body_pixel.py:115-142)."""
        sd = {k: (strip_module(v) if v is not None else None) for k, v in state_dict.items() if isinstance(v, dict) or v is None}
        gen = sd["generator"] if "generator" in sd else strip_module(state_dict)
        self.engine.load_pixelcnn(gen)
        self._loaded = {"generator": gen, "audioencoder": None}
        if sd.get("audioencoder") is not None:
            self.engine.load_audioenc(sd["audioencoder"])
            self._loaded["audioencoder"] = sd["audioencoder"]

    def state_dict(self):
        """The nested layout the reference saves (nets/smplx_body_pixel.py:104-113): the weights that were loaded (the engine
        keeps repacked copies only), optimizer / discriminator slots empty — this is an inference engine."""
        loaded = getattr(self, "_loaded", {"generator": None, "audioencoder": None})
        return {"generator": loaded["generator"], "generator_optim": None, "audioencoder": loaded["audioencoder"],
                "audioencoder_optim": None, "discriminator": None, "discriminator_optim": None}

    # -- inference -------------------------------------------------------------------------------
    def _noise(self, T, B):
        return draw_sampler_noise(T, B, self.noise_device, per_step=self.noise_per_step)

    def infer_on_audio(self, aud_fn, initial_pose=None, norm_stats=None, exp=None, var=None, w_pre=False, rand=None,
                       continuity=False, id=None, fps=15, sr=22000, B=1, am=None, am_sr=None, frame=0, **kwargs):
        """(aud_fn) -> generated motion, numpy (B, F, 129).  Reference: :232-289."""
        assert self.args.infer, "train mode"
        if continuity:
            return self._infer_continuity(aud_fn, id, fps, sr, B)
        if torch.is_tensor(aud_fn) or isinstance(aud_fn, np.ndarray):
            aud_feat = np.asarray(aud_fn, dtype=np.float32)          # already (M, 64) features
            mfcc = torch.from_numpy(np.ascontiguousarray(aud_feat.T))[None].repeat(B, 1, 1)   # [B,64,M]
        elif fps == 30 and sr == 22000 and self.device_mfcc:
            audio, sr_0 = load_wav(aud_fn)                           # decode on the host, features on the device
            mfcc = self.engine.mfcc(audio.mean(0, keepdim=True), sr_0).repeat(B, 1, 1)
        else:
            aud_feat = get_mfcc_ta(aud_fn, sr=sr, fps=fps, smlpx=True, type="mfcc", am=am)
            mfcc = torch.from_numpy(np.ascontiguousarray(aud_feat.T))[None].repeat(B, 1, 1)
        label = torch.tensor([0]) if id is None else id.reshape(-1).repeat(B)[:B] if id.numel() == 1 else id
        return self.generate(mfcc, label, noise_fn=kwargs.get("noise_fn")).cpu().numpy()

    def generate(self, aud, id, frame_num=0, noise_fn=None):
        """tensor API (:306-326): aud [B,64,M] features, id [B] -> torch (B, F, 129) on the device.
        (The reference's version passes decode()'s tuple to torch.cat and raises; this returns the
        tensor its infer_on_audio builds.)"""
        mfcc = aud.to(torch.float32)
        B, _, M = mfcc.shape
        T = self.engine.latent_rows(M)
        noise = self._noise(T, B) if noise_fn is None else noise_fn(T, B)      # [2T,B,2048] Exp(1) draws (RNG contract)
        codes, poses = self.engine.body_generate(mfcc, id.to(torch.int64), noise)
        self.last_codes = codes
        return poses

    def infer(self, aud_feat, frame, id, B, pre_latents=None, pre_audio=None, pre_pose=None):
        """The reference's inner call (:291-304): aud_feat [B,M,64] (time-major features of ONE chunk) ->
        (latents [B,T,2], audio [B,256,T,2], body [B,39,4T], hand [B,90,4T]).  ``pre_latents`` / ``pre_audio`` (the first
        two results of the previous chunk's call) condition the sampler on that chunk (GatedPixelCNN.generate :158-165);
        ``pre_pose`` is accepted and unused like in the reference, whose Decoder.forward ignores ``pre_state``
        (nets/spg/vqvae_1d.py:139-149) — each chunk is decoded on its own."""
        a = self.engine.audio_encode(aud_feat.transpose(1, 2).contiguous())            # [B,256,T]
        T = a.shape[2]
        label = id.reshape(-1)
        if pre_latents is None:
            latents = self.engine.pixelcnn_generate(a, label, self._noise(T, B))
        else:
            pa = pre_audio[..., 0] if pre_audio.dim() == 4 else pre_audio              # [B,256,T0]
            latents = self.engine.pixelcnn_generate(torch.cat([pa.to(a.device), a], 2), label, self._noise(T, B), T=T,
                                                    pre_latents=pre_latents)
        body = self.engine.vq_decode(0, latents[..., 0].contiguous())
        hand = self.engine.vq_decode(1, latents[..., 1].contiguous())
        return latents, a.unsqueeze(-1).repeat(1, 1, 1, 2), body, hand

    def _infer_continuity(self, aud_fn, id, fps, sr, B):
        """continuity=True (:244-269): a 2 s prefix, then the rest conditioned on the prefix's latents and audio.  The two
        chunks go through the audio encoder separately and are DECODED separately, so each chunk sees zero padding at the
        2 s seam."""
        aud_feat, gap = get_mfcc_sepa(aud_fn, sr=sr, fps=fps)                 # (M0+M1, 64), M0
        feat = torch.from_numpy(np.ascontiguousarray(np.asarray(aud_feat, dtype=np.float32)))[None].repeat(B, 1, 1)   # [B,M,64]
        label = torch.tensor([0]) if id is None else id.reshape(-1).repeat(B)
        pre_pose = {"b": None, "h": None}
        lat0, audio0, body_0, hand_0 = self.infer(feat[:, :gap], 0, label, B, pre_pose=pre_pose)
        pre_pose["b"], pre_pose["h"] = body_0[:, :, -4:].transpose(1, 2), hand_0[:, :, -4:].transpose(1, 2)
        lat1, _, body_1, hand_1 = self.infer(feat[:, gap:], 0, label, B, lat0, audio0, pre_pose)
        self.last_codes = torch.cat([lat0, lat1], 1)
        body, hand = torch.cat([body_0, body_1], 2), torch.cat([hand_0, hand_1], 2)
        return torch.cat([body, hand], 1).transpose(1, 2).cpu().numpy()
