"""GPU (-m gpu): the drop-in wrapper classes (talkshow_b200.nets) through their reference
signatures — constructor(args, config), load_state_dict(reference-format dict), infer_on_audio /
generate — against reference-generated goldens and the oracle."""
import os
import types

import numpy as np
import pytest
import torch

import talkshow_oracle as O
from conftest import GOLDEN, ROOT, draw_noise, noise_fp
from talkshow_b200 import synth

pytestmark = pytest.mark.gpu


def _cfg(name):
    from talkshow_b200.trainer.config import load_JsonConfig

    return load_JsonConfig(os.path.join(ROOT, "config", name + ".json"))


def _args():
    return types.SimpleNamespace(gpu=0, infer=True)


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


def test_body_pixel_wrapper_infer_on_audio(ckpts, tmp_path):
    """BASELINE config 3 through the wrapper: vq checkpoint picked up from config.Model.vq_path like the
    reference ctor (smplx_body_pixel.py:59-62), features passed as an array, CPU-generator noise."""
    from talkshow_b200.nets import init_model

    gold = _load("pixel_b1_t30")
    cfg = _cfg("body_pixel")
    vq_path = str(tmp_path / "vq.pth")
    torch.save({"generator": ckpts["vq"]}, vq_path)
    cfg.Model.vq_path = vq_path
    g = init_model("s2g_body_pixel", _args(), cfg)
    g.load_state_dict({"generator": {"module." + k: v for k, v in ckpts["pixel"]["generator"].items()},
                       "audioencoder": ckpts["pixel"]["audioencoder"], "generator_optim": None})
    g.noise_device = "cpu"                      # reproduce the CPU reference's sampling stream
    mfcc = synth.synth_mfcc(1, 120)
    torch.manual_seed(int(gold["sampler_seed"]))
    pred = g.infer_on_audio(mfcc[0].t().numpy(), id=torch.tensor([0]), fps=30, B=1)
    assert isinstance(pred, np.ndarray) and pred.shape == (1, 120, 129)
    probe = draw_noise(60, 1, int(gold["sampler_seed"]))
    if np.allclose(gold["noise_fp"], noise_fp(probe), rtol=0, atol=1e-9):
        assert np.array_equal(g.last_codes.cpu().numpy(), gold["codes"])
        assert np.abs(pred - gold["pred"]).max() <= 1e-4
    # tensor API returns the tensor the reference's infer_on_audio builds
    torch.manual_seed(3)
    out = g.generate(mfcc.repeat(2, 1, 1), torch.tensor([0, 1]))
    assert out.shape == (2, 120, 129) and out.is_cuda
    with pytest.raises(NotImplementedError):
        init_model("s2g_LS3DCG", _args(), cfg)


def test_body_pixel_wrapper_wav_and_continuity(ckpts, tmp_path):
    from scipy.io import wavfile

    from talkshow_b200.nets import s2g_body_pixel

    cfg = _cfg("body_pixel")
    cfg.Model.vq_path = str(tmp_path / "missing.pth")       # absent: weights injected instead
    g = s2g_body_pixel(_args(), cfg)
    g.load_vq_state_dict(ckpts["vq"])
    g.load_state_dict(ckpts["pixel"])
    x = (synth.synth_wave(1, 16000 * 5)[0].numpy() * 20000).astype(np.int16)
    p = str(tmp_path / "clip.wav")
    wavfile.write(p, 16000, x)
    torch.manual_seed(1)
    pred = g.infer_on_audio(p, id=torch.tensor([2]), fps=30, B=3)          # 3 diversity samples of one clip
    assert pred.shape == (3, 148, 129) and np.isfinite(pred).all()           # M=150 frames -> T=37 rows -> 148
    assert not np.array_equal(pred[0], pred[1])
    torch.manual_seed(1)
    cont = g.infer_on_audio(p, id=torch.tensor([2]), fps=30, B=1, continuity=True)
    assert cont.shape[0] == 1 and cont.shape[2] == 129 and np.isfinite(cont).all()


def test_body_pixel_wrapper_continuity_golden(ckpts, tmp_path, monkeypatch):
    """continuity=True against the reference-generated golden (tests/golden/make_golden.py --only wrapper_cont):
    the 2 s prefix and the remainder are decoded separately (nets/smplx_body_pixel.py:262-269), so the frames on
    both sides of the seam (rows 52..67 stored in full) must match, not just the shapes."""
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
    torch.manual_seed(seed)
    pred = g.infer_on_audio("synthetic.wav", continuity=True, id=torch.tensor(gold["label"]), fps=30, B=2)
    assert pred.shape == (2, 160, 129)
    # always: the oracle's per-chunk restatement with the same noise
    m0 = torch.from_numpy(f0.T.copy())[None].repeat(2, 1, 1)
    m1 = torch.from_numpy(f1.T.copy())[None].repeat(2, 1, 1)
    l0, l1, ref = O.body_infer_continuity(ckpts["pixel"], ckpts["vq"], m0, m1, torch.tensor(gold["label"]).repeat(2), n0, n1, window=18)
    assert torch.equal(g.last_codes.cpu(), torch.cat([l0, l1], 1))
    assert np.abs(pred - ref.numpy()).max() <= 1e-4
    if np.allclose(gold["noise_fp0"], noise_fp(n0), rtol=0, atol=1e-9) and np.allclose(gold["noise_fp1"], noise_fp(n1), rtol=0, atol=1e-9):
        assert np.abs(pred[:, ::int(gold["pred_stride"])] - gold["pred"]).max() <= 1e-4
        assert np.abs(pred[:, 52:68] - gold["pred_seam"]).max() <= 1e-4


def test_body_vq_wrapper(ckpts):
    from talkshow_b200.nets import s2g_body_vq

    gold = _load("vq_roundtrip")
    g = s2g_body_vq(_args(), _cfg("body_vq"))
    g.load_state_dict(ckpts["vq"])
    poses = synth.synth_poses(2, 88)
    out = g.infer_on_audio(torch.zeros(2, 64, 88), initial_pose=poses, fps=30)
    assert out.shape == gold["out"].shape == (88, 258)
    assert np.abs(out - gold["out"]).max() <= 1e-4
    ib, ih = g.encode(poses)
    assert np.array_equal(ib.cpu().numpy(), gold["idx_body"]) and np.array_equal(ih.cpu().numpy(), gold["idx_hand"])
    # continuity=True: five 60-frame chunks round-tripped separately (nets/smplx_body_vq.py:256-271), reference golden
    gc = _load("wrapper_cont")
    outc = g.infer_on_audio(torch.zeros(2, 64, 300), initial_pose=synth.synth_poses(2, 300, seed=313), continuity=True, fps=30)
    assert outc.shape == (300, 258)
    assert np.abs(outc[::3] - gc["vq_out"]).max() <= 1e-4
    assert np.abs(outc[56:64] - gc["vq_seam"]).max() <= 1e-4


def test_face_wrapper(ckpts):
    from talkshow_b200.nets import s2g_face

    gold = _load("face")
    g = s2g_face(_args(), _cfg("face"))
    g.load_state_dict(ckpts["face"])
    wave = synth.synth_wave(1, 64000)
    out = g.infer_on_audio(wave[:, None, :])                 # tensor input like smplx_face.py:195-197, id=None
    assert out.shape == (1, 120, 103)
    assert np.abs(out - gold["out_4s"]).max() <= 1e-4
    out2 = g.generate(synth.synth_wave(2, 24000, seed=5)[:, None, :], 45)
    assert out2.shape == (2, 45, 103) and out2.is_cuda


def test_whole_body_pipeline_matches_oracle(ckpts):
    """face + body + part2full on the device == demo.py's per-sample assembly on the oracle outputs."""
    from talkshow_b200.engine import Engine
    from talkshow_b200.pipeline import WholeBody

    e = Engine(0)
    wb = WholeBody(e)
    wb.load(ckpts["pixel"], ckpts["vq"], ckpts["face"])
    B, sec = 2, 2
    wave = synth.synth_wave(B, 16000 * sec, seed=12)
    mfcc = synth.synth_mfcc(B, 60, seed=13)
    label = torch.tensor([1, 3])
    noise = draw_noise(2 * O.latent_rows(60), B, 5)
    got = wb.generate(mfcc.cuda(), wave.cuda(), label.cuda(), noise=noise.cuda()).cpu()
    face = O.face_forward(ckpts["face"]["generator"], wave, torch.zeros(B, 4), sec * 30)
    _, body = O.body_generate(ckpts["pixel"], ckpts["vq"], mfcc, label, noise=noise, window=18)
    ref = torch.stack([O.assemble_pose(face[b], body[b]) for b in range(B)])
    assert got.shape == ref.shape == (B, 60, 265)
    assert (got - ref).abs().max().item() <= 1e-4
    host = wb.generate_host(mfcc.pin_memory(), wave.pin_memory(), label.pin_memory(), noise=noise.cuda())
    assert torch.equal(host, got)
    torch.cuda.synchronize()
    e.close()


def test_edge_shapes(ckpts):
    """shortest clips the path accepts, ragged lengths, batch > one PixelCNN tile."""
    from talkshow_b200.engine import Engine

    e = Engine(0)
    e.load_pixelcnn(ckpts["pixel"]["generator"])
    e.load_audioenc(ckpts["pixel"]["audioencoder"])
    e.load_vq(0, ckpts["vq"]["g_body"])
    e.load_vq(1, ckpts["vq"]["g_hand"])
    for M in (4, 7, 13):                                      # T = 1, 1, 3 latent rows
        mfcc = synth.synth_mfcc(2, M, seed=M)
        T = O.latent_rows(M)
        noise = draw_noise(2 * T, 2, M)
        ref_c, ref_p = O.body_generate(ckpts["pixel"], ckpts["vq"], mfcc, torch.tensor([0, 1]), noise=noise)
        c, p = e.body_generate(mfcc, torch.tensor([0, 1]), noise)
        assert torch.equal(c.cpu(), ref_c) and (p.cpu() - ref_p).abs().max().item() <= 1e-4
    with pytest.raises(RuntimeError):
        e.audio_encode(torch.zeros(1, 64, 3))                 # below the two stride-2 convs' minimum
    # 70 samples = two batch tiles; the second tile must see its own noise slice
    B, M = 70, 8
    mfcc = synth.synth_mfcc(1, M, seed=3).repeat(B, 1, 1)
    noise = torch.empty(2 * O.latent_rows(M), B, 2048).exponential_(1, generator=torch.Generator().manual_seed(9))
    c, _ = e.body_generate(mfcc, torch.zeros(B, dtype=torch.int64), noise)
    ref_c, _ = O.body_generate(ckpts["pixel"], ckpts["vq"], mfcc[64:], torch.zeros(6, dtype=torch.int64), noise=noise[:, 64:])
    assert torch.equal(c[64:].cpu(), ref_c)
    torch.cuda.synchronize()
    e.close()


def test_demo_flow_matches_reference_flow(ckpts, tmp_path, monkeypatch):
    """The body of the reference's scripts/demo.py (:158-246: face pass, num_sample body samples, jaw | body | expression ->
    part2full -> (num_sample*F, 265) .npy) through talkshow_b200.scripts.demo with checkpoint FILES, a wav file and the
    reference's command line, against the oracle run sample by sample with the reference's RNG order."""
    from scipy.io import wavfile

    from talkshow_b200.data_utils.utils import load_wav, mfcc_from_wave
    from talkshow_b200.scripts import demo
    from talkshow_b200.trainer.options import parse_args

    sec, nsamp, spk = 3, 3, 2
    x = (synth.synth_wave(1, 16000 * sec, seed=41)[0].numpy() * 20000).astype(np.int16)
    wav = str(tmp_path / "clip one.wav")
    wavfile.write(wav, 16000, x)
    torch.save({"generator": ckpts["pixel"]}, str(tmp_path / "body.pth"))
    torch.save({"generator": ckpts["face"]}, str(tmp_path / "face.pth"))
    torch.save({"generator": ckpts["vq"]}, str(tmp_path / "vq.pth"))
    args = parse_args().parse_args(["--config_file", os.path.join(ROOT, "config", "body_pixel.json"), "--infer", "--audio_file", wav,
                                    "--id", str(spk), "--num_sample", str(nsamp), "--body_model_path", str(tmp_path / "body.pth"),
                                    "--face_model_path", str(tmp_path / "face.pth")])
    assert args.body_model_name == "s2g_body_pixel" and args.face_model_name == "s2g_face"      # reference defaults
    config = _cfg("body_pixel")
    config.Model.vq_path = str(tmp_path / "vq.pth")
    g_body = demo.init_model(args.body_model_name, args.body_model_path, args, config)
    g_face = demo.init_model(args.face_model_name, args.face_model_path, args, _cfg("face"))
    g_body.noise_device = "cpu"
    g_body.device_mfcc = False            # the oracle sees the host torchaudio features
    monkeypatch.chdir(tmp_path)
    seed = 321
    torch.manual_seed(seed)
    result_list, verts = demo.infer(g_body, g_face, None, None, config, args)
    assert verts is None and len(result_list) == nsamp
    # oracle, the reference's order: face once, then one body sample after the other from the same generator
    audio, sr = load_wav(wav)
    frame = audio.shape[1] * 30 // 16000
    face = O.face_forward(ckpts["face"]["generator"], audio, torch.zeros(1, 4), frame)[0]
    mfcc = torch.from_numpy(mfcc_from_wave(audio, sr, sr=22000, fps=30).T.copy())[None]
    T = O.latent_rows(mfcc.shape[2])
    torch.manual_seed(seed)
    _reference_face_pass_rng(g_face)
    for i in range(nsamp):
        noise = torch.stack([torch.empty(1, 2048).exponential_(1) for _ in range(2 * T)])
        _, body = O.body_generate(ckpts["pixel"], ckpts["vq"], mfcc, torch.tensor([spk]), noise=noise, window=18)
        ref = O.assemble_pose(face, body[0])
        assert result_list[i].shape == ref.shape == (frame, 265)
        assert (result_list[i].cpu() - ref).abs().max().item() <= 1e-4
    saved = np.load(str(tmp_path / "visualise" / "video" / config.Log.name / "clip one.npy"))
    assert saved.shape == (nsamp * frame, 265)                       # scripts/demo.py:239-245
    assert np.array_equal(saved, np.concatenate([r.cpu().numpy() for r in result_list], 0))
    assert not np.array_equal(saved[:frame], saved[frame:2 * frame])      # diversity samples differ
    # the file the REFERENCE's own scripts/demo.py:infer writes for the same wav, checkpoints, command line and seed
    # (tests/golden/make_golden.py --only demo_flow), when this host's CPU generator gives the stream it was made with
    gold = _load("demo_flow")
    assert (sec, nsamp, spk, seed) == (int(gold["seconds"]), int(gold["num_sample"]), int(gold["speaker"]), int(gold["seed"]))
    from talkshow_b200.nets.smplx_face import hf_layerdrop_uses_torch_rng
    if np.allclose(gold["noise_fp"], noise_fp(draw_noise(2 * T, 1, seed)), rtol=0, atol=1e-9) and hf_layerdrop_uses_torch_rng() == bool(gold["hf_torch_layerdrop"]):
        assert np.abs(saved[::int(gold["saved_stride"])] - gold["saved"]).max() <= 1e-4
