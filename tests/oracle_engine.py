"""TEST INFRASTRUCTURE: a CPU stand-in for ``talkshow_b200.engine.Engine`` whose methods are answered by the oracle
(oracle/talkshow_oracle.py).  It exists so that the HOST logic around the C ABI — the wrapper classes of
``talkshow_b200.nets``, ``scripts/demo.py``, ``pipeline.WholeBody`` — can be driven on a machine without a GPU and compared
with the reference-generated goldens: what the wrappers do with checkpoints, features, ids, noise, chunks and layouts is the
same code whichever object answers ``audio_encode`` / ``pixelcnn_generate`` / ``vq_decode`` / ``face_forward``.  Never imported
by the package; the product path has no CPU fallback (tests/test_cabi_and_host.py::test_wrappers_refuse_cpu_device)."""
import torch

import talkshow_oracle as O


class OracleEngine:
    host_only = True                      # pipeline.WholeBody: no second engine / stream
    sm_count = 148

    def __init__(self, window=18):
        self.device = torch.device("cpu")
        self.window = window              # >= the 17-row receptive field: identical to the reference's full-grid loop
        self.sd = {}
        self.launches = 0

    # -- weights (same names as Engine) --------------------------------------------------------
    def load_pixelcnn(self, sd):
        self.sd["pixelcnn"] = dict(sd)

    def load_audioenc(self, sd):
        self.sd["audioenc"] = dict(sd)

    def load_vq(self, which, sd):
        self.sd["vq%d" % which] = dict(sd)

    def load_face(self, sd):
        self.sd["face"] = dict(sd)

    def close(self):
        pass

    # -- the module-level boundary -------------------------------------------------------------
    @staticmethod
    def latent_rows(M):
        return O.latent_rows(M)

    def vq_dim(self, which):
        return int(self.sd["vq%d" % which]["decoder.project.weight"].shape[0])

    def mfcc(self, wave, sr):
        """[1,N] at sr -> [1,64,M]: the host torchaudio chain (the device MFCC is compared with it in the GPU tests)."""
        from talkshow_b200.data_utils.utils import mfcc_from_wave

        return torch.from_numpy(mfcc_from_wave(wave, sr, sr=22000, fps=30).T.copy())[None]

    def audio_encode(self, mfcc):
        if mfcc.shape[2] < 4:
            raise RuntimeError("ts_audio_encode: M=%d" % mfcc.shape[2])
        return O.audio_encoder(self.sd["audioenc"], mfcc.float())

    def pixelcnn_generate(self, aud, label, noise, T=None, pre_latents=None, want_logits=False):
        assert not want_logits
        B = aud.shape[0]
        a2 = aud.float().unsqueeze(-1).repeat(1, 1, 1, 2)
        label = label.reshape(-1).to(torch.int64)
        if label.numel() == 1 and B > 1:
            label = label.expand(B)
        if pre_latents is None:
            return O.pixelcnn_generate(self.sd["pixelcnn"], label, a2.shape[2], B, a2, noise=noise, window=self.window)
        T0 = pre_latents.shape[1]
        T = a2.shape[2] - T0 if T is None else T
        return O.pixelcnn_generate(self.sd["pixelcnn"], label, T, B, a2[:, :, T0:], noise=noise, pre_latents=pre_latents,
                                   pre_audio=a2[:, :, :T0], window=self.window)

    def vq_decode(self, which, idx):
        return O.vq_decode(self.sd["vq%d" % which], idx.to(torch.int64))

    def vq_encode(self, which, poses, want_e=False):
        q, idx = O.vq_encode(self.sd["vq%d" % which], poses.float())
        return (idx, q) if want_e else idx

    def face_forward(self, wave, id_onehot, frame):
        wave = wave.float().reshape(wave.shape[0], -1)
        idv = id_onehot.float()
        if idv.shape[0] == 1 and wave.shape[0] > 1:
            idv = idv.expand(wave.shape[0], -1)
        return O.face_forward(self.sd["face"], wave, idv, frame)

    def body_generate(self, mfcc, label, noise, want_codes=True):
        B = mfcc.shape[0]
        label = label.reshape(-1).to(torch.int64)
        if label.numel() == 1 and B > 1:
            label = label.expand(B)
        codes, poses = O.body_generate({"generator": self.sd["pixelcnn"], "audioencoder": self.sd["audioenc"]},
                                       {"g_body": self.sd["vq0"], "g_hand": self.sd["vq1"]}, mfcc.float(), label, noise=noise,
                                       window=self.window)
        return (codes if want_codes else None), poses

    def assemble_pose(self, face, body, stand=False):
        return torch.stack([O.assemble_pose(face[b], body[b], stand) for b in range(face.shape[0])])

    def rot6d_to_axis_angle(self, d6):
        return O.rot6d_to_axis_angle(d6)
