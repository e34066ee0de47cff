"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on seeded
synthetic checkpoints and inputs.  Run in the build container only:

    python tests/golden/make_golden.py

The fixtures pin the oracle (tests/test_oracle_golden.py) and, on the GPU box, the CUDA engine.
Checkpoints/inputs are regenerated from seeds by talkshow_b200/synth.py (uniform-only arithmetic,
host independent); the fixtures store the small inputs, the outputs and fingerprints.
"""
import argparse
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import torch  # noqa: E402

from ref_import import import_reference  # noqa: E402
from talkshow_b200 import synth  # noqa: E402

SAMPLER_SEED = 2024


def draw_noise(steps, B, K=2048):
    """The reference consumes one ``exponential_`` of shape [B,K] per sampled position
    (gated_pixelcnn_v2.py:175 -> ATen multinomial)."""
    out = torch.empty(steps, B, K)
    for s in range(steps):
        out[s] = torch.empty(B, K).exponential_(1)
    return out


def noise_fp(noise):
    return np.array([float(noise.double().sum()), float(noise[0, 0, :8].double().sum()),
                     float(noise[-1, -1, -8:].double().sum())])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    cwd = os.getcwd()
    nets = import_reference()
    REF_DIR = os.getcwd()
    import nets.smplx_body_pixel as ref_bp
    import nets.smplx_body_vq as ref_vq
    import nets.smplx_face as ref_face
    from trainer.config import load_JsonConfig
    torch.set_grad_enabled(False)

    tmp = tempfile.mkdtemp()
    vq_ckpt = synth.body_vq_checkpoint(0)
    vq_path = os.path.join(tmp, "vq.pth")
    torch.save({"generator": vq_ckpt}, vq_path)
    bp_ckpt = synth.body_pixel_checkpoint(0)
    face_ckpt = synth.face_checkpoint(0)

    cfg_pixel = load_JsonConfig("config/body_pixel.json")
    cfg_pixel.Model.vq_path = vq_path
    cfg_vq = load_JsonConfig("config/body_vq.json")
    cfg_face = load_JsonConfig("config/face.json")
    a = types.SimpleNamespace(gpu="cpu", infer=True)

    def want(n):
        return not args.only or n in args.only.split(",")

    out = {}

    # ---- body pixel: wrapper end to end (config 3) ------------------------------------------
    if want("pixel_b1_t30") or want("pixel_b3_t75") or want("pixel_cont") or want("wrapper_cont") or want("wrapper_b3"):
        g = ref_bp.TrainWrapper(a, cfg_pixel)
        g.load_state_dict(bp_ckpt)            # demo.py:54-62 passes ckpt['generator'] = this dict
    if want("pixel_b1_t30"):
        mfcc = synth.synth_mfcc(1, 120)       # [1,64,120]
        ref_bp.get_mfcc_ta = lambda *aa, **kk: mfcc[0].transpose(0, 1).numpy()     # (M,64) like utils.py:177
        torch.manual_seed(SAMPLER_SEED)
        pred = g.infer_on_audio("synthetic.wav", id=torch.tensor([0]), fps=30, B=1)  # (1,120,129)
        torch.manual_seed(SAMPLER_SEED)
        noise = draw_noise(60, 1)
        # codes: recompute through the same modules with the same seed
        torch.manual_seed(SAMPLER_SEED)
        audio = g.audioencoder(mfcc).unsqueeze(-1).repeat(1, 1, 1, 2)
        lat = g.generator.generate(torch.tensor([0]), shape=[30, 2], batch_size=1, aud_feat=audio)
        logits = g.generator(lat, torch.tensor([0]), audio)                       # teacher-forced [1,2048,30,2]
        rows = [0, 1, 2, 3, 17, 18, 29]
        np.savez_compressed(os.path.join(HERE, "pixel_b1_t30.npz"), mfcc=mfcc.numpy(), label=np.array([0]),
                            audio=audio[..., 0].numpy(), codes=lat.numpy(), pred=pred,
                            logit_rows=np.array(rows), logits=logits[0][:, rows, :].numpy(),
                            noise_fp=noise_fp(noise), sampler_seed=SAMPLER_SEED,
                            fp=np.array(list(synth.fingerprint(bp_ckpt).values())),
                            fp_vq=np.array(list(synth.fingerprint(vq_ckpt).values())))
        print("pixel_b1_t30", lat[0, :6].tolist(), pred.shape)

    if want("wrapper_b3"):
        # diversity through the wrapper's own batch argument (test_body.py:139-146 passes B): one clip, B = 3 samples, speaker 2,
        # 15 fps default features injected as (M, 64)
        mfcc = synth.synth_mfcc(1, 40, seed=515)
        ref_bp.get_mfcc_ta = lambda *aa, **kk: mfcc[0].transpose(0, 1).numpy()
        torch.manual_seed(SAMPLER_SEED + 5)
        pred = g.infer_on_audio("synthetic.wav", id=torch.tensor([2]), B=3)        # (3, 40, 129)
        torch.manual_seed(SAMPLER_SEED + 5)
        noise = draw_noise(20, 3)
        np.savez_compressed(os.path.join(HERE, "wrapper_b3.npz"), pred=pred, noise_fp=noise_fp(noise), sampler_seed=SAMPLER_SEED + 5, mfcc_seed=515)
        print("wrapper_b3", pred.shape)

    if want("pixel_b3_t75"):
        mfcc = synth.synth_mfcc(3, 300, seed=77)
        label = torch.tensor([0, 1, 3])
        torch.manual_seed(SAMPLER_SEED + 1)
        audio = g.audioencoder(mfcc).unsqueeze(-1).repeat(1, 1, 1, 2)
        lat = g.generator.generate(label, shape=[75, 2], batch_size=3, aud_feat=audio)
        body, _ = g.g_body.decode(b=3, w=75, latents=lat[..., 0])
        hand, _ = g.g_hand.decode(b=3, w=75, latents=lat[..., 1])
        pred = torch.cat([body, hand], 1).transpose(1, 2).numpy()
        torch.manual_seed(SAMPLER_SEED + 1)
        noise = draw_noise(150, 3)
        np.savez_compressed(os.path.join(HERE, "pixel_b3_t75.npz"), label=label.numpy(), codes=lat.numpy(),
                            pred=pred[:, ::7].copy(), pred_stride=7, noise_fp=noise_fp(noise),
                            sampler_seed=SAMPLER_SEED + 1, mfcc_seed=77)
        print("pixel_b3_t75", lat[:, :3].tolist())

    if want("pixel_cont"):
        # continuity path: generate(pre_latents, pre_audio), gated_pixelcnn_v2.py:158-165
        mfcc = synth.synth_mfcc(2, 160, seed=99)
        label = torch.tensor([2, 1])
        audio = g.audioencoder(mfcc).unsqueeze(-1).repeat(1, 1, 1, 2)             # [2,256,40,2]
        torch.manual_seed(SAMPLER_SEED + 2)
        lat0 = g.generator.generate(label, shape=[15, 2], batch_size=2, aud_feat=audio[:, :, :15])
        lat1 = g.generator.generate(label, shape=[25, 2], batch_size=2, aud_feat=audio[:, :, 15:],
                                    pre_latents=lat0, pre_audio=audio[:, :, :15])
        torch.manual_seed(SAMPLER_SEED + 2)
        noise = draw_noise(80, 2)
        np.savez_compressed(os.path.join(HERE, "pixel_cont.npz"), label=label.numpy(), codes0=lat0.numpy(),
                            codes1=lat1.numpy(), noise_fp=noise_fp(noise), sampler_seed=SAMPLER_SEED + 2,
                            mfcc_seed=99)
        print("pixel_cont", lat1[:, :3].tolist())


    if want("wrapper_cont"):
        # wrapper-level continuity=True (smplx_body_pixel.py:244-269): 2 s prefix + remainder, features of the two
        # chunks computed separately (get_mfcc_sepa, injected here like get_mfcc_ta above), each chunk DECODED SEPARATELY
        # (Decoder.forward ignores pre_state, vqvae_1d.py:139-149) and concatenated.
        f0 = synth.synth_mfcc(1, 60, seed=311)[0].transpose(0, 1).numpy()       # (60,64): 2 s at 30 fps
        f1 = synth.synth_mfcc(1, 100, seed=312)[0].transpose(0, 1).numpy()      # (100,64)
        ref_bp.get_mfcc_sepa = lambda *aa, **kk: (np.concatenate((f0, f1), 0), f0.shape[0])
        torch.manual_seed(SAMPLER_SEED + 3)
        pred = g.infer_on_audio("synthetic.wav", continuity=True, id=torch.tensor([2]), fps=30, B=2)   # (2,160,129)
        torch.manual_seed(SAMPLER_SEED + 3)
        n0 = draw_noise(30, 2)
        n1 = draw_noise(50, 2)
        # VQ wrapper continuity=True (smplx_body_vq.py:256-271): five 60-frame chunks round-tripped separately
        gv = ref_vq.TrainWrapper(a, cfg_vq)
        gv.load_state_dict(vq_ckpt)
        poses = synth.synth_poses(2, 300, seed=313)
        outv = gv.infer_on_audio(torch.zeros(2, 64, 300), initial_pose=poses, continuity=True, fps=30)   # (300, 258)
        outs = gv.infer_on_audio(torch.zeros(1, 64, 300), initial_pose=synth.synth_poses(1, 300, seed=5), fps=30, smooth=True)    # (300, 129): :283-291
        np.savez_compressed(os.path.join(HERE, "wrapper_cont.npz"), vq_smooth=outs[140:170].copy(), pred=pred[:, ::3].copy(), pred_stride=3,
                            pred_seam=pred[:, 52:68].copy(), label=np.array([2]),
                            noise_fp0=noise_fp(n0), noise_fp1=noise_fp(n1), sampler_seed=SAMPLER_SEED + 3,
                            vq_out=outv[::3].copy(), vq_seam=outv[56:64].copy())
        print("wrapper_cont", pred.shape, outv.shape)

    # ---- 6-D geometry: pixelcnn(2048, 512, 10, ...) + VQ-VAEs over 78 / 180 channels (smplx_body_pixel.py:49-57) ----
    if want("pixel_6d"):
        vq6 = synth.body_vq_checkpoint_6d(0)
        vq6_path = os.path.join(tmp, "vq6.pth")
        torch.save({"generator": vq6}, vq6_path)
        bp6 = synth.body_pixel_checkpoint_6d(0)
        cfg6 = load_JsonConfig("config/body_pixel.json")
        cfg6.Data.pose.convert_to_6d = True
        cfg6.Model.vq_path = vq6_path
        g6 = ref_bp.TrainWrapper(a, cfg6)
        g6.load_state_dict(bp6)
        assert g6.generator.dim == 512 and len(g6.generator.layers) == 10
        mfcc = synth.synth_mfcc(1, 32, seed=611)                                   # T = 8 latent rows
        ref_bp.get_mfcc_ta = lambda *aa, **kk: mfcc[0].transpose(0, 1).numpy()
        torch.manual_seed(SAMPLER_SEED + 4)
        pred = g6.infer_on_audio("synthetic.wav", id=torch.tensor([1]), fps=30, B=2)   # (2, 32, 258)
        torch.manual_seed(SAMPLER_SEED + 4)
        noise = draw_noise(16, 2)
        torch.manual_seed(SAMPLER_SEED + 4)
        audio = g6.audioencoder(mfcc.repeat(2, 1, 1)).unsqueeze(-1).repeat(1, 1, 1, 2)
        lab = torch.tensor([1, 1])
        lat = g6.generator.generate(lab, shape=[8, 2], batch_size=2, aud_feat=audio)
        logits = g6.generator(lat, lab, audio)                                     # teacher-forced [2,2048,8,2]
        np.savez_compressed(os.path.join(HERE, "pixel_6d.npz"), codes=lat.numpy(), pred=pred, logits=logits[:, :, [0, 3, 7], :].numpy(),
                            logit_rows=np.array([0, 3, 7]), noise_fp=noise_fp(noise), sampler_seed=SAMPLER_SEED + 4,
                            fp=np.array(list(synth.fingerprint(bp6).values())), fp_vq=np.array(list(synth.fingerprint(vq6).values())))
        print("pixel_6d", lat[0, :3].tolist(), pred.shape)

    # ---- VQ roundtrip (config 2) -------------------------------------------------------------
    if want("vq_roundtrip"):
        gv = ref_vq.TrainWrapper(a, cfg_vq)
        gv.load_state_dict(vq_ckpt)
        poses = synth.synth_poses(2, 88)
        outv = gv.infer_on_audio(torch.zeros(2, 64, 88), initial_pose=poses, fps=30)     # (88, 129)
        gt = poses[:, gv.c_index].permute(0, 2, 1)
        eb, ib = gv.g_body.encode(gt_poses=gt[..., :39])
        eh, ih = gv.g_hand.encode(gt_poses=gt[..., 39:])
        np.savez_compressed(os.path.join(HERE, "vq_roundtrip.npz"), poses=poses.numpy(), out=outv,
                            idx_body=ib.numpy(), idx_hand=ih.numpy(), e_body=eb.numpy(),
                            fp_vq=np.array(list(synth.fingerprint(vq_ckpt).values())))
        print("vq_roundtrip", ib.tolist(), outv.shape)

    # ---- face (config 1 stand-in + batch) ----------------------------------------------------
    if want("face"):
        gf = ref_face.TrainWrapper(a, cfg_face)
        gf.load_state_dict(face_ckpt)
        wave = synth.synth_wave(1, 64000)
        o1 = gf.infer_on_audio(wave[:, None, :])                                       # (1,120,103), id=None
        wave2 = synth.synth_wave(2, 24000, seed=5)
        ids = torch.tensor([1, 3])
        o2 = gf.generator(wave2[:, None, :], None, torch.nn.functional.one_hot(ids, 4), time_steps=45)[0].numpy()
        np.savez_compressed(os.path.join(HERE, "face.npz"), out_4s=o1, out_b2=o2, ids_b2=ids.numpy(),
                            wave_head=wave[0, :64].numpy(),
                            fp=np.array(list(synth.fingerprint(face_ckpt).values())))
        print("face", o1.shape, float(np.abs(o1).max()), o2.shape)
    # ---- face at the benchmarked clip length: 10 s, two clips with speaker ids (config 5's face half, per clip) ---------------
    if want("face_10s"):
        gf = ref_face.TrainWrapper(a, cfg_face)
        gf.load_state_dict(face_ckpt)
        gf.generator.eval()
        wave = synth.synth_wave(2, 160000, seed=8)
        ids = torch.tensor([2, 0])
        o = gf.generator(wave[:, None, :], None, torch.nn.functional.one_hot(ids, 4), time_steps=300)[0].numpy()     # (2,300,103)
        np.savez_compressed(os.path.join(HERE, "face_10s.npz"), out=o[:, ::4].copy(), out_stride=4, out_head=o[:, :8].copy(), ids=ids.numpy(),
                            wave_seed=8, fp=np.array(list(synth.fingerprint(face_ckpt).values())))
        print("face_10s", o.shape, float(np.abs(o).max()))
    # ---- 6-D rotation -> axis-angle (demo.py:185-188,216-219 for convert_to_6d configs) ----------
    if want("rot6d"):
        from data_utils.rotation_conversion import matrix_to_axis_angle, rotation_6d_to_matrix, axis_angle_to_matrix, matrix_to_rotation_6d
        g = torch.Generator().manual_seed(23)
        d6 = torch.randn(4096, 6, generator=g)
        # edge cases: exact identity, tiny angles (small-angle branch), near-pi turns, unnormalised / nearly parallel inputs
        aa = torch.randn(64, 3, generator=g)
        aa = aa / aa.norm(dim=-1, keepdim=True)
        ang = torch.cat([torch.zeros(8), torch.logspace(-9, -3, 24), 3.14159 - torch.logspace(-6, -1, 16), torch.full((16,), 1.0)])
        special = matrix_to_rotation_6d(axis_angle_to_matrix(aa * ang[:, None]))
        special[-16:] *= torch.logspace(-3, 3, 16)[:, None]                  # scale invariance of the Gram-Schmidt step
        d6 = torch.cat([d6, special], 0)
        out = matrix_to_axis_angle(rotation_6d_to_matrix(d6))
        np.savez_compressed(os.path.join(HERE, "rot6d.npz"), d6=d6.numpy(), aa=out.numpy())
        print("rot6d", d6.shape, float(out.abs().max()), bool(torch.isfinite(out).all()))
    # ---- the whole demo flow: scripts/demo.py:158-246 (infer) of the reference, file in -> .npy out ---------------------
    if want("demo_flow"):
        import importlib.machinery
        from scipy.io import wavfile

        class _RenderTool:                              # visualise/rendering.py needs pyrender / ffmpeg: out of scope
            def __init__(self, *aa, **kk):
                self.calls = []

            def _render_sequences(self, wav, vertices_list, **kk):
                self.calls.append((wav, len(vertices_list), kk))

        for name, attrs in (("visualise", {}), ("visualise.rendering", {"RenderTool": _RenderTool})):
            m = types.ModuleType(name)
            m.__spec__ = importlib.machinery.ModuleSpec(name, None)
            m.__path__ = []
            for k, v in attrs.items():
                setattr(m, k, v)
            sys.modules[name] = m
        import data_utils.utils as ref_du
        import scripts.demo as ref_demo
        import transformers
        from talkshow_b200.nets.smplx_face import hf_layerdrop_uses_torch_rng   # reads transformers' source only
        from talkshow_b200.data_utils.utils import load_wav

        # file decoding: torchaudio.load needs torchcodec and librosa is a stub here; both readers return the
        # samples of the 16 kHz int16 file unchanged, which is what the real ones do for such a file
        ref_du.ta.load = lambda fn: load_wav(fn)
        ref_du.librosa.load = lambda fn, sr=16000: (load_wav(fn)[0].mean(0).numpy(), 16000)
        ref_demo.Wav2Vec2Processor.from_pretrained = staticmethod(lambda name: object())      # hub download; only `am is None` is tested
        # the module-level stubs above replaced the patched readers of the earlier sections: restore the real feature code
        ref_bp.get_mfcc_ta = ref_du.get_mfcc_ta
        sec, nsamp, spk, seed = 3, 3, 2, 321
        x = (synth.synth_wave(1, 16000 * sec, seed=41)[0].numpy() * 20000).astype(np.int16)
        wav = os.path.join(tmp, "clip one.wav")
        wavfile.write(wav, 16000, x)
        torch.save({"generator": bp_ckpt}, os.path.join(tmp, "body.pth"))
        torch.save({"generator": face_ckpt}, os.path.join(tmp, "face.pth"))
        from trainer.options import parse_args as ref_parse_args
        dargs = ref_parse_args().parse_args(["--config_file", "config/body_pixel.json", "--infer", "--audio_file", wav, "--id", str(spk),
                                             "--num_sample", str(nsamp), "--body_model_path", os.path.join(tmp, "body.pth"),
                                             "--face_model_path", os.path.join(tmp, "face.pth")])
        dargs.gpu = "cpu"
        g_body = ref_demo.init_model(dargs.body_model_name, dargs.body_model_path, dargs, cfg_pixel)
        g_face = ref_demo.init_model(dargs.face_model_name, dargs.face_model_path, dargs, cfg_face)

        class _Smplx:                                   # smplx is not installed: get_vertices only needs .vertices / .body_pose
            def __call__(self, **kw):
                return types.SimpleNamespace(vertices=torch.zeros(1, 4, 3), body_pose=kw["body_pose"])

        rt = _RenderTool()
        os.makedirs(os.path.join(tmp, "visualise", "video", cfg_pixel.Log.name), exist_ok=True)
        os.chdir(tmp)                                   # infer() saves relative to the working directory
        torch.manual_seed(seed)
        ref_demo.infer(g_body, g_face, _Smplx(), rt, cfg_pixel, dargs)
        saved = np.load(os.path.join(tmp, "visualise", "video", cfg_pixel.Log.name, "clip one.npy"))
        frame = 16000 * sec * 30 // 16000
        assert saved.shape == (nsamp * frame, 265) and len(rt.calls) == 1 and rt.calls[0][1] == nsamp
        # the --stand and --only_face branches (scripts/demo.py:165-169,224-227; part2full(pred, stand))
        variants = {}
        for key, flags in (("stand", ["--stand"]), ("only_face", ["--only_face"])):
            vargs = ref_parse_args().parse_args(["--config_file", "config/body_pixel.json", "--infer", "--audio_file", wav, "--id", str(spk),
                                                 "--num_sample", "1", "--body_model_path", os.path.join(tmp, "body.pth"),
                                                 "--face_model_path", os.path.join(tmp, "face.pth")] + flags)
            torch.manual_seed(seed)
            ref_demo.infer(g_body, g_face, _Smplx(), rt, cfg_pixel, vargs)
            variants[key] = np.load(os.path.join(tmp, "visualise", "video", cfg_pixel.Log.name, "clip one.npy"))
            assert variants[key].shape == (frame, 265)
        torch.manual_seed(seed)
        n_first = draw_noise(2 * 22, 1)                 # sample 0's draws: M = 90 feature frames -> T = 22 latent rows
        np.savez_compressed(os.path.join(HERE, "demo_flow.npz"), saved=saved[::2].copy(), saved_stride=2, seconds=sec, num_sample=nsamp,
                            speaker=spk, seed=seed, wave_seed=41, log_name=str(cfg_pixel.Log.name), frame=frame, noise_fp=noise_fp(n_first),
                            saved_stand=variants["stand"][::3].copy(), saved_only_face=variants["only_face"][::3].copy(),
                            hf_torch_layerdrop=int(hf_layerdrop_uses_torch_rng()), transformers_version=transformers.__version__)
        os.chdir(REF_DIR)
        print("demo_flow", saved.shape, float(np.abs(saved).max()))
    # ---- the outputs the reference SHIPS (demo/**/*.npy, written by scripts/demo.py:239-245 with the authors' checkpoints): layout,
    #      the constant lower-body columns part2full inserts (data_utils/lower_body.py:68-87), the --only_face static block ----
    if want("demo_npy_layout"):
        import glob
        import json
        lower = [i for i in range(265) if not (i < 3 or 18 <= i < 21 or 27 <= i < 30 or 36 <= i < 39 or i >= 45)]      # 33 inserted columns
        files = {}
        for f in sorted(glob.glob("demo/**/*.npy", recursive=True)):
            arr = np.load(f)
            assert arr.ndim == 2 and arr.shape[1] == 265 and arr.dtype == np.float32
            lc = arr[:, lower]
            assert (lc == lc[0]).all()                                                   # constant over the frames of a file
            static = arr[:, 3:165]
            files[f] = {"rows": int(arr.shape[0]), "lower": [float(v) for v in lc[0]],
                        "only_face_static": [float(v) for v in static[0]] if (static == static[0]).all() else None}
        json.dump({"lower_columns": lower, "files": files}, open(os.path.join(HERE, "demo_npy_layout.json"), "w"), indent=1, sort_keys=True)
        print("demo_npy_layout", len(files))
    # ---- host audio front-end: data_utils/utils.py:148-231 (get_mfcc_ta) and :234-263 (get_mfcc_sepa) on a stereo 44.1 kHz file ----
    if want("frontend"):
        from scipy.io import wavfile
        import data_utils.utils as ref_du
        from talkshow_b200.data_utils.utils import load_wav

        ref_du.ta.load = lambda fn: load_wav(fn)          # torchaudio.load needs torchcodec here; same samples
        x = (synth.synth_wave(2, 44100 * 5, seed=3).numpy().T * 20000).astype(np.int16)
        wav = os.path.join(tmp, "stereo44k.wav")
        wavfile.write(wav, 44100, x)
        fe = {}
        for fps in (30, 15):
            fe["ta_%d" % fps] = ref_du.get_mfcc_ta(wav, sr=22000, fps=fps, smlpx=True, type="mfcc", am=None)
            fe["sepa_%d" % fps], fe["gap_%d" % fps] = ref_du.get_mfcc_sepa(wav, sr=22000, fps=fps)
        np.savez_compressed(os.path.join(HERE, "frontend.npz"), wave_seed=3, seconds=5, sr=44100,
                            **{k: (v[::5].copy() if isinstance(v, np.ndarray) else v) for k, v in fe.items()},
                            shapes=np.array([fe["ta_30"].shape[0], fe["ta_15"].shape[0], fe["sepa_30"].shape[0], fe["sepa_15"].shape[0]]))
        print("frontend", fe["ta_30"].shape, fe["gap_30"], fe["gap_15"])
    # ---- SMPL-X call semantics: the reference's get_vertices (scripts/demo.py:122-152) and get_joints (data_utils/get_j.py:20-51)
    #      run on a stand-in smplx_model that answers each per-frame keyword call with the float64 restatement of smplx 0.1.28
    #      (oracle/smplx_oracle.py; the package and the licensed model file are absent).  Pins WHICH columns of the 265-vector the
    #      reference hands to which argument, the per-frame loop and the output layouts — not the body model's arithmetic. ----
    if want("smplx_calls"):
        import importlib.machinery
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import smplx_oracle as SO
        from talkshow_b200 import smplx_lbs
        if "scripts.demo" not in sys.modules:
            for name, attrs in (("visualise", {}), ("visualise.rendering", {"RenderTool": object})):
                m = types.ModuleType(name)
                m.__spec__ = importlib.machinery.ModuleSpec(name, None)
                m.__path__ = []
                for k, v in attrs.items():
                    setattr(m, k, v)
                sys.modules[name] = m
        import scripts.demo as ref_demo2
        import data_utils.get_j as ref_getj
        model = smplx_lbs.synthetic_model(V=300, seed=11, nfaces=500)

        class _Model:
            batch_size = 1
            calls = 0

            def __call__(self, betas=None, expression=None, jaw_pose=None, leye_pose=None, reye_pose=None, global_orient=None,
                         body_pose=None, left_hand_pose=None, right_hand_pose=None, return_verts=True):
                _Model.calls += 1
                p = torch.cat([jaw_pose, leye_pose, reye_pose, global_orient, body_pose, left_hand_pose, right_hand_pose, expression], 1)
                v, j = SO.smplx_forward(model, p, betas[:1])
                out = {"vertices": v, "joints": j, "body_pose": body_pose}
                return type("Out", (dict,), {"__getattr__": dict.__getitem__})(out)

        gen = torch.Generator().manual_seed(12)
        res = [((torch.rand(5, 265, generator=gen, dtype=torch.float64) * 2 - 1) * 0.4) for _ in range(2)]
        betas = (torch.rand(1, 300, generator=gen, dtype=torch.float64) - 0.5) * 0.2
        verts, poses = ref_demo2.get_vertices(_Model(), betas, [r.clone() for r in res], True, require_pose=True)
        assert _Model.calls == 10                                   # one model call per frame
        joints3 = ref_getj.get_joints(_Model(), betas, torch.stack(res).clone())        # [2,5,127,3]
        joints2 = ref_getj.get_joints(_Model(), betas, res[0].clone())                  # [5,127,3]
        np.savez_compressed(os.path.join(HERE, "smplx_calls.npz"), verts=np.stack(verts).astype(np.float32),
                            poses=torch.stack(poses).numpy().astype(np.float32), joints3=joints3.numpy().astype(np.float32),
                            joints2=joints2.numpy().astype(np.float32), model_seed=11, V=300, nfaces=500, pose_seed=12)
        print("smplx_calls", np.stack(verts).shape, tuple(joints3.shape))
    # ---- pose layout helpers of data_utils/lower_body.py (imported by scripts/demo.py:24, diversity.py, test_body.py) ----------
    if want("lower_body"):
        import data_utils.lower_body as ref_lb
        gen = torch.Generator().manual_seed(77)
        pred = torch.rand(6, 232, generator=gen)
        full = torch.rand(6, 265, generator=gen)
        gt = torch.rand(4, 265, generator=gen)
        np.savez_compressed(os.path.join(HERE, "lower_body.npz"), seed=77, c_index_3d=np.asarray(ref_lb.c_index_3d), c_index_6d=np.asarray(ref_lb.c_index_6d),
                            part2full=ref_lb.part2full(pred).numpy(), part2full_stand=ref_lb.part2full(pred, True).numpy(),
                            pred2poses=ref_lb.pred2poses(pred, gt).numpy(), poses2poses=ref_lb.poses2poses(full, gt).numpy(),
                            poses2pred=ref_lb.poses2pred(full).numpy(), poses2pred_stand=ref_lb.poses2pred(full, True).numpy())
        print("lower_body", len(ref_lb.c_index_3d), len(ref_lb.c_index_6d))
    # ---- CLI surface: trainer/options.py:3-37 (demo.py:251-252 does parse_args().parse_args()) ------------------
    if want("options"):
        import json
        from trainer.options import parse_args as ref_parse_args
        demo_cmd = ["--config_file", "./config/body_pixel.json", "--infer", "--audio_file", "./demo_audio/1st-page.wav",
                    "--id", "2", "--whole_body", "--num_sample", "12"]
        json.dump({"defaults": vars(ref_parse_args().parse_args([])), "demo_cmd": demo_cmd,
                   "demo": vars(ref_parse_args().parse_args(demo_cmd))},
                  open(os.path.join(HERE, "options.json"), "w"), indent=1, sort_keys=True)
        print("options", len(vars(ref_parse_args().parse_args([]))))
    os.chdir(cwd)


if __name__ == "__main__":
    main()
