"""CPU (-m "not gpu"): the HOST logic of the drop-in layer — wrapper classes (talkshow_b200.nets), scripts/demo.py flow,
pipeline.WholeBody — driven with an oracle-backed stand-in for the engine (tests/oracle_engine.py) and compared with the
reference-generated goldens.  What the wrappers do with checkpoints ('module.' prefixes, nested dicts, vq_path), features,
speaker ids, the noise contract, continuity chunks and output layouts is the same Python whichever object answers the
module-level calls; the GPU tests (tests/test_gpu_wrappers.py) run the same flows on the CUDA engine."""
import os
import types

import numpy as np
import pytest
import torch

import talkshow_oracle as O
from conftest import GOLDEN, ROOT, draw_noise, noise_fp
from oracle_engine import OracleEngine
from talkshow_b200 import synth


def _cfg(name):
    from talkshow_b200.trainer.config import load_JsonConfig

    return load_JsonConfig(os.path.join(ROOT, "config", name + ".json"))


def _args():
    return types.SimpleNamespace(gpu=0, infer=True)      # resolve_device(0) -> cuda:0 (no CUDA call); the engine is injected


def _load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def _reference_face_pass_rng(g_face):
    """What the reference's face pass does to torch's default generator before the body samples are drawn
    (scripts/demo.py:173-204): with the installed transformers the wav2vec2 encoder draws one torch.rand([]) per layer
    (LayerDrop probability, eval mode included); the release the reference pins drew from numpy's generator instead."""
    from talkshow_b200.nets.smplx_face import hf_layerdrop_uses_torch_rng

    assert g_face.layerdrop_rng_draws is None and g_face.encoder_layers == 12
    if hf_layerdrop_uses_torch_rng():
        for _ in range(12):
            torch.rand([])


@pytest.fixture()
def shared(monkeypatch):
    """one oracle-backed engine behind nets.base.shared_engine (what scripts/demo.py's wrappers share per device)."""
    import talkshow_b200.nets.base as base

    eng = OracleEngine()
    monkeypatch.setattr(base, "_ENGINES", {0: eng})
    return eng


def test_body_pixel_wrapper_golden(ckpts, tmp_path, shared):
    """BASELINE config 3 through the wrapper's reference signature: vq checkpoint from config.Model.vq_path, DataParallel
    'module.' prefixes, features as an array, CPU-generator noise -> the reference's codes and poses."""
    from talkshow_b200.nets import init_model

    gold = _load("pixel_b1_t30")
    cfg = _cfg("body_pixel")
    vq_path = str(tmp_path / "vq.pth")
    torch.save({"generator": ckpts["vq"]}, vq_path)
    cfg.Model.vq_path = vq_path
    g = init_model("s2g_body_pixel", _args(), cfg)
    assert g.engine is shared and set(shared.sd) == {"vq0", "vq1"}
    g.load_state_dict({"generator": {"module." + k: v for k, v in ckpts["pixel"]["generator"].items()},
                       "audioencoder": ckpts["pixel"]["audioencoder"], "generator_optim": None})
    g.noise_device = "cpu"
    mfcc = synth.synth_mfcc(1, 120)
    torch.manual_seed(int(gold["sampler_seed"]))
    pred = g.infer_on_audio(mfcc[0].t().numpy(), id=torch.tensor([0]), fps=30, B=1)
    assert isinstance(pred, np.ndarray) and pred.shape == (1, 120, 129)
    assert np.allclose(gold["noise_fp"], noise_fp(draw_noise(60, 1, int(gold["sampler_seed"]))), rtol=0, atol=1e-9)
    assert np.array_equal(g.last_codes.numpy(), gold["codes"])
    assert np.abs(pred - gold["pred"]).max() <= 1e-5
    # state_dict(): the nested layout the reference saves (nets/smplx_body_pixel.py:104-113), optimizer slots empty
    sd = g.state_dict()
    assert set(sd) == {"generator", "generator_optim", "audioencoder", "audioencoder_optim", "discriminator", "discriminator_optim"}
    assert sd["generator_optim"] is None and sd["discriminator"] is None
    assert all(not k.startswith("module.") for k in sd["generator"])
    assert torch.equal(sd["generator"]["embedding.weight"], ckpts["pixel"]["generator"]["embedding.weight"])
    # infer(): the reference's inner call (:291-304): features [B,M,64] -> (latents, audio [B,256,T,2], body, hand)
    torch.manual_seed(int(gold["sampler_seed"]))
    lat, audio, body, hand = g.infer(mfcc.transpose(1, 2), 0, torch.tensor([0]), 1, pre_pose={"b": None, "h": None})
    assert np.array_equal(lat.numpy(), gold["codes"]) and audio.shape == (1, 256, 30, 2)
    assert body.shape == (1, 39, 120) and hand.shape == (1, 90, 120)
    assert np.abs(torch.cat([body, hand], 1).transpose(1, 2).numpy() - gold["pred"]).max() <= 1e-5
    # diversity through the wrapper's batch argument: one clip, B = 3, speaker 2 (id [1] repeated like :253-256), reference golden
    gb = _load("wrapper_b3")
    feats = synth.synth_mfcc(1, 40, seed=int(gb["mfcc_seed"]))[0].t().numpy()
    assert np.allclose(gb["noise_fp"], noise_fp(draw_noise(20, 3, int(gb["sampler_seed"]))), rtol=0, atol=1e-9)
    torch.manual_seed(int(gb["sampler_seed"]))
    pred3 = g.infer_on_audio(feats, id=torch.tensor([2]), B=3)
    assert pred3.shape == (3, 40, 129) and np.abs(pred3 - gb["pred"]).max() <= 1e-5
    assert not np.array_equal(pred3[0], pred3[1])
    with pytest.raises(NotImplementedError):
        init_model("s2g_LS3DCG", _args(), cfg)
    g.args.infer = False
    with pytest.raises(AssertionError):
        g.infer_on_audio(mfcc[0].t().numpy(), id=torch.tensor([0]), fps=30, B=1)


def test_body_pixel_wrapper_continuity_golden(ckpts, tmp_path, monkeypatch, shared):
    """continuity=True: the 2 s prefix and the remainder are sampled with carried latents / audio and decoded separately
    (nets/smplx_body_pixel.py:244-269); seam frames and a strided subset of all frames against the reference."""
    import talkshow_b200.nets.smplx_body_pixel as bp

    gold = _load("wrapper_cont")
    cfg = _cfg("body_pixel")
    cfg.Model.vq_path = str(tmp_path / "missing.pth")
    g = bp.TrainWrapper(_args(), cfg)
    g.load_vq_state_dict(ckpts["vq"])
    g.load_state_dict(ckpts["pixel"])
    g.noise_device = "cpu"
    f0 = synth.synth_mfcc(1, 60, seed=311)[0].t().numpy()
    f1 = synth.synth_mfcc(1, 100, seed=312)[0].t().numpy()
    monkeypatch.setattr(bp, "get_mfcc_sepa", lambda *a, **k: (np.concatenate((f0, f1), 0), f0.shape[0]))
    seed = int(gold["sampler_seed"])
    torch.manual_seed(seed)
    n0 = torch.stack([torch.empty(2, 2048).exponential_(1) for _ in range(30)])
    n1 = torch.stack([torch.empty(2, 2048).exponential_(1) for _ in range(50)])
    assert np.allclose(gold["noise_fp0"], noise_fp(n0), rtol=0, atol=1e-9) and np.allclose(gold["noise_fp1"], noise_fp(n1), rtol=0, atol=1e-9)
    torch.manual_seed(seed)
    pred = g.infer_on_audio("synthetic.wav", continuity=True, id=torch.tensor(gold["label"]), fps=30, B=2)
    assert pred.shape == (2, 160, 129)
    assert np.abs(pred[:, ::int(gold["pred_stride"])] - gold["pred"]).max() <= 1e-5
    assert np.abs(pred[:, 52:68] - gold["pred_seam"]).max() <= 1e-5


def test_body_vq_wrapper_golden(ckpts, shared):
    from talkshow_b200.nets import s2g_body_vq

    gold = _load("vq_roundtrip")
    g = s2g_body_vq(_args(), _cfg("body_vq"))
    g.load_state_dict({"g_body": {"module." + k: v for k, v in ckpts["vq"]["g_body"].items()}, "g_hand": ckpts["vq"]["g_hand"]})
    poses = synth.synth_poses(2, 88)
    out = g.infer_on_audio(torch.zeros(2, 64, 88), initial_pose=poses, fps=30)
    assert out.shape == gold["out"].shape == (88, 258)
    assert np.abs(out - gold["out"]).max() <= 1e-5
    ib, ih = g.encode(poses)
    assert np.array_equal(ib.numpy(), gold["idx_body"]) and np.array_equal(ih.numpy(), gold["idx_hand"])
    gc = _load("wrapper_cont")
    outc = g.infer_on_audio(torch.zeros(2, 64, 300), initial_pose=synth.synth_poses(2, 300, seed=313), continuity=True, fps=30)
    assert outc.shape == (300, 258)
    assert np.abs(outc[::3] - gc["vq_out"]).max() <= 1e-5
    assert np.abs(outc[56:64] - gc["vq_seam"]).max() <= 1e-5
    sd = g.state_dict()                                        # nets/smplx_body_vq.py:77-94
    assert {"g_body", "g_hand"} <= set(sd) and sd["g_body_optim"] is None
    assert all(not k.startswith("module.") for k in sd["g_body"])
    # smooth=True blends 10 frames from frame 149 on (:283-291) and changes nothing else
    base = g.infer_on_audio(None, initial_pose=synth.synth_poses(1, 300, seed=5), fps=30)
    sm = g.infer_on_audio(None, initial_pose=synth.synth_poses(1, 300, seed=5), fps=30, smooth=True)
    assert np.array_equal(sm[:149], base[:149]) and np.array_equal(sm[159:], base[159:]) and not np.array_equal(sm[149:159], base[149:159])
    assert np.abs(sm[140:170] - gc["vq_smooth"]).max() <= 1e-5                  # the reference's smooth=True output around frame 149


def test_face_wrapper_golden(ckpts, shared):
    from talkshow_b200.nets import s2g_face

    gold = _load("face")
    g = s2g_face(_args(), _cfg("face"))
    g.load_state_dict({"generator": {"module." + k: v for k, v in ckpts["face"]["generator"].items()}})
    wave = synth.synth_wave(1, 64000)
    out = g.infer_on_audio(wave[:, None, :])                  # tensor input like smplx_face.py:195-197, id=None
    assert out.shape == (1, 120, 103)
    assert np.abs(out - gold["out_4s"]).max() <= 1e-5
    out2 = g.generate(synth.synth_wave(2, 24000, seed=5)[:, None, :], 45)
    assert out2.shape == (2, 45, 103)
    assert set(g.state_dict()) >= {"generator", "generator_optim"}


def test_demo_flow_on_the_oracle_engine(ckpts, tmp_path, monkeypatch, shared):
    """scripts/demo.py:158-246 through talkshow_b200.scripts.demo with checkpoint FILES, a wav file and the reference's
    command line: num_sample diversity samples as one batched call draw the noise sample by sample like the reference's
    loop, jaw | body | expression -> part2full -> (num_sample*F, 265) .npy."""
    from scipy.io import wavfile

    from talkshow_b200.data_utils.utils import load_wav, mfcc_from_wave
    from talkshow_b200.scripts import demo
    from talkshow_b200.trainer.options import parse_args

    gold = _load("demo_flow")                                   # the reference's own scripts/demo.py:infer on the same files
    sec, nsamp, spk, seed = int(gold["seconds"]), int(gold["num_sample"]), int(gold["speaker"]), int(gold["seed"])
    x = (synth.synth_wave(1, 16000 * sec, seed=int(gold["wave_seed"]))[0].numpy() * 20000).astype(np.int16)
    wav = str(tmp_path / "clip one.wav")
    wavfile.write(wav, 16000, x)
    torch.save({"generator": ckpts["pixel"]}, str(tmp_path / "body.pth"))
    torch.save({"generator": ckpts["face"]}, str(tmp_path / "face.pth"))
    torch.save({"generator": ckpts["vq"]}, str(tmp_path / "vq.pth"))
    args = parse_args().parse_args(["--config_file", os.path.join(ROOT, "config", "body_pixel.json"), "--infer", "--audio_file", wav,
                                    "--id", str(spk), "--num_sample", str(nsamp), "--body_model_path", str(tmp_path / "body.pth"),
                                    "--face_model_path", str(tmp_path / "face.pth")])
    config = _cfg("body_pixel")
    config.Model.vq_path = str(tmp_path / "vq.pth")
    g_body = demo.init_model(args.body_model_name, args.body_model_path, args, config)
    g_face = demo.init_model(args.face_model_name, args.face_model_path, args, _cfg("face"))
    g_body.noise_device = "cpu"
    g_body.device = g_face.device = torch.device("cpu")         # demo.infer moves its results to the wrapper's device
    monkeypatch.chdir(tmp_path)
    torch.manual_seed(seed)
    result_list, verts = demo.infer(g_body, g_face, None, None, config, args)
    assert verts is None and len(result_list) == nsamp
    audio, sr = load_wav(wav)
    frame = audio.shape[1] * 30 // 16000
    face = O.face_forward(ckpts["face"]["generator"], audio, torch.zeros(1, 4), frame)[0]
    mfcc = torch.from_numpy(mfcc_from_wave(audio, sr, sr=22000, fps=30).T.copy())[None]
    T = O.latent_rows(mfcc.shape[2])
    torch.manual_seed(seed)
    _reference_face_pass_rng(g_face)
    for i in range(nsamp):                                      # the reference's order: one sample after the other
        noise = torch.stack([torch.empty(1, 2048).exponential_(1) for _ in range(2 * T)])
        _, body = O.body_generate(ckpts["pixel"], ckpts["vq"], mfcc, torch.tensor([spk]), noise=noise, window=18)
        ref = O.assemble_pose(face, body[0])
        assert result_list[i].shape == ref.shape == (frame, 265)
        assert (result_list[i] - ref).abs().max().item() <= 1e-5
    saved = np.load(str(tmp_path / "visualise" / "video" / config.Log.name / "clip one.npy"))
    assert saved.shape == (nsamp * frame, 265) and config.Log.name == str(gold["log_name"])
    assert np.array_equal(saved, np.concatenate([r.numpy() for r in result_list], 0))
    torch.manual_seed(seed)
    assert np.allclose(gold["noise_fp"], noise_fp(draw_noise(2 * T, 1, seed)), rtol=0, atol=1e-9)
    from talkshow_b200.nets.smplx_face import hf_layerdrop_uses_torch_rng
    if hf_layerdrop_uses_torch_rng() == bool(gold["hf_torch_layerdrop"]):       # same transformers behaviour as when the fixture was made
        assert np.abs(saved[::int(gold["saved_stride"])] - gold["saved"]).max() <= 1e-5      # == the file the reference's demo.py writes
        # --stand / --only_face (scripts/demo.py:165-169,224-227), one sample each, same seed
        for key, flag in (("saved_stand", "--stand"), ("saved_only_face", "--only_face")):
            vargs = parse_args().parse_args(["--config_file", os.path.join(ROOT, "config", "body_pixel.json"), "--infer", "--audio_file", wav,
                                             "--id", str(spk), "--num_sample", "1", "--body_model_path", str(tmp_path / "body.pth"),
                                             "--face_model_path", str(tmp_path / "face.pth"), flag])
            torch.manual_seed(seed)
            res, _ = demo.infer(g_body, g_face, None, None, config, vargs, save=False)
            assert np.abs(res[0].numpy()[::3] - gold[key]).max() <= 1e-5, key
    # the command-line entry point (scripts/demo.py:250-300 main): config file -> both wrappers from checkpoint files -> infer
    import json
    cfg_json = json.load(open(os.path.join(ROOT, "config", "body_pixel.json")))
    cfg_json["Model"]["vq_path"] = str(tmp_path / "vq.pth")
    json.dump(cfg_json, open(str(tmp_path / "cfg.json"), "w"))
    import talkshow_b200.nets.base as base
    monkeypatch.setattr(base, "resolve_device", lambda gpu: torch.device("cpu"))
    for mod in ("smplx_body_pixel", "smplx_face"):
        monkeypatch.setattr(__import__("talkshow_b200.nets." + mod, fromlist=["x"]), "resolve_device", base.resolve_device)
    torch.manual_seed(seed)
    orig_init = demo.init_model

    def init_cpu_noise(name, path, a, c):
        g = orig_init(name, path, a, c)
        g.noise_device = "cpu"
        return g

    monkeypatch.setattr(demo, "init_model", init_cpu_noise)
    main_list, main_verts = demo.main(["--config_file", str(tmp_path / "cfg.json"), "--infer", "--audio_file", wav, "--id", str(spk),
                                       "--num_sample", str(nsamp), "--body_model_path", str(tmp_path / "body.pth"),
                                       "--face_model_path", str(tmp_path / "face.pth")])
    assert main_verts is None and all(torch.equal(a, b) for a, b in zip(main_list, result_list))
    # the pinned transformers (no torch draws in the face pass): forced by the attribute, differs from the auto mode exactly then
    g_face.layerdrop_rng_draws = 0
    torch.manual_seed(seed)
    again, _ = demo.infer(g_body, g_face, None, None, config, args, save=False)
    assert torch.equal(again[0], result_list[0]) != hf_layerdrop_uses_torch_rng()
    assert torch.equal(again[0][:, :3], result_list[0][:, :3]) and torch.equal(again[0][:, 165:], result_list[0][:, 165:])   # face columns: deterministic


def test_whole_body_pipeline_host_logic(ckpts):
    """pipeline.WholeBody on the stand-in engine: face + body + assembly per sample, body padded / truncated to the face
    length, sharded generation equal to the unsharded one."""
    from talkshow_b200.pipeline import WholeBody

    wb = WholeBody(OracleEngine())
    wb.load(ckpts["pixel"], ckpts["vq"], ckpts["face"])
    assert wb.e2 is None
    B, sec = 3, 1
    wave = synth.synth_wave(B, 16000 * sec + 800, seed=12)              # 31 face frames, 28 body frames: last body frame repeated
    mfcc = synth.synth_mfcc(B, 30, seed=13)
    label = torch.tensor([1, 3, 0])
    noise = draw_noise(2 * O.latent_rows(30), B, 5)
    got = wb.generate(mfcc, wave, label, noise=noise)
    face = O.face_forward(ckpts["face"]["generator"], wave, torch.zeros(B, 4), 31)
    _, body = O.body_generate(ckpts["pixel"], ckpts["vq"], mfcc, label, noise=noise, window=18)
    assert body.shape[1] == 28
    ref = torch.stack([O.assemble_pose(face[b], body[b]) for b in range(B)])
    assert got.shape == (B, 31, 265) and torch.equal(got, ref)
    assert torch.equal(got[:, 30, 45:165], got[:, 27, 45:165])          # body columns of the padded frames repeat frame 27
    parts = [wb.generate_sharded(mfcc, wave, label, r, 2, noise_full=noise, gather=False) for r in range(2)]
    assert [p.shape[0] for p in parts] == [2, 1]
    assert (torch.cat(parts, 0) - got).abs().max().item() <= 1e-5      # ATen's CPU convs round differently per batch size
