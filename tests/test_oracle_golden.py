"""CPU: the oracle restatement (oracle/talkshow_oracle.py) reproduces the golden vectors that
tests/golden/make_golden.py produced with the UNMODIFIED reference modules."""
import os

import numpy as np
import pytest
import torch

import talkshow_oracle as O
from conftest import GOLDEN, draw_noise, noise_fp
from talkshow_b200 import synth


def _load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def _same_noise(g, noise):
    return np.allclose(g["noise_fp"], noise_fp(noise), rtol=0, atol=1e-9)


def test_fingerprints(ckpts):
    g = _load("pixel_b1_t30")
    assert np.allclose(list(synth.fingerprint(ckpts["pixel"]).values()), g["fp"], rtol=1e-12)
    assert np.allclose(list(synth.fingerprint(ckpts["vq"]).values()), g["fp_vq"], rtol=1e-12)
    assert np.allclose(list(synth.fingerprint(ckpts["face"]).values()), _load("face")["fp"], rtol=1e-12)


def test_pixel_b1_t30(ckpts):
    g = _load("pixel_b1_t30")
    mfcc = synth.synth_mfcc(1, 120)
    assert np.array_equal(mfcc.numpy(), g["mfcc"])
    noise = draw_noise(60, 1, int(g["sampler_seed"]))
    if not _same_noise(g, noise):
        pytest.skip("host RNG stream differs from the fixture machine")
    audio = O.audio_encoder(ckpts["pixel"]["audioencoder"], mfcc)
    assert np.abs(audio.numpy() - g["audio"]).max() <= 1e-6
    # literal O(T^2) generate with pre-drawn noise == reference generate with its own multinomial
    lat, pred = O.body_generate(ckpts["pixel"], ckpts["vq"], mfcc, torch.tensor([0]), noise=noise)
    assert np.array_equal(lat.numpy(), g["codes"])
    assert np.abs(pred.numpy() - g["pred"]).max() <= 1e-6
    # windowed forward is bit-identical
    lat_w, _ = O.body_generate(ckpts["pixel"], ckpts["vq"], mfcc, torch.tensor([0]), noise=noise, window=18)
    assert np.array_equal(lat_w.numpy(), g["codes"])
    # teacher-forced logits
    a2 = audio.unsqueeze(-1).repeat(1, 1, 1, 2)
    logits = O.pixelcnn_forward(ckpts["pixel"]["generator"], lat, torch.tensor([0]), a2)
    assert np.abs(logits[0][:, g["logit_rows"].tolist(), :].numpy() - g["logits"]).max() <= 1e-5


def test_pixel_b3_t75(ckpts):
    g = _load("pixel_b3_t75")
    mfcc = synth.synth_mfcc(3, 300, seed=int(g["mfcc_seed"]))
    noise = draw_noise(150, 3, int(g["sampler_seed"]))
    if not _same_noise(g, noise):
        pytest.skip("host RNG stream differs from the fixture machine")
    lat, pred = O.body_generate(ckpts["pixel"], ckpts["vq"], mfcc, torch.tensor(g["label"]), noise=noise, window=18)
    assert np.array_equal(lat.numpy(), g["codes"])
    assert np.abs(pred.numpy()[:, ::int(g["pred_stride"])] - g["pred"]).max() <= 1e-6


def test_pixel_continuity(ckpts):
    g = _load("pixel_cont")
    mfcc = synth.synth_mfcc(2, 160, seed=int(g["mfcc_seed"]))
    noise = draw_noise(80, 2, int(g["sampler_seed"]))
    if not _same_noise(g, noise):
        pytest.skip("host RNG stream differs from the fixture machine")
    sd = ckpts["pixel"]["generator"]
    label = torch.tensor(g["label"])
    audio = O.audio_encoder(ckpts["pixel"]["audioencoder"], mfcc).unsqueeze(-1).repeat(1, 1, 1, 2)
    lat0 = O.pixelcnn_generate(sd, label, 15, 2, audio[:, :, :15], noise=noise[:30])
    lat1 = O.pixelcnn_generate(sd, label, 25, 2, audio[:, :, 15:], noise=noise[30:], pre_latents=lat0,
                               pre_audio=audio[:, :, :15], window=18)
    assert np.array_equal(lat0.numpy(), g["codes0"])
    assert np.array_equal(lat1.numpy(), g["codes1"])


def test_wrapper_continuity(ckpts):
    """wrapper-level continuity=True goldens (body-pixel: per-chunk decode; VQ: five 60-frame chunks)."""
    g = _load("wrapper_cont")
    seed = int(g["sampler_seed"])
    torch.manual_seed(seed)
    n0 = torch.stack([torch.empty(2, 2048).exponential_(1) for _ in range(30)])
    n1 = torch.stack([torch.empty(2, 2048).exponential_(1) for _ in range(50)])
    if not (np.allclose(g["noise_fp0"], noise_fp(n0), rtol=0, atol=1e-9) and np.allclose(g["noise_fp1"], noise_fp(n1), rtol=0, atol=1e-9)):
        pytest.skip("host RNG stream differs from the fixture machine")
    m0 = synth.synth_mfcc(1, 60, seed=311).repeat(2, 1, 1)
    m1 = synth.synth_mfcc(1, 100, seed=312).repeat(2, 1, 1)
    label = torch.tensor(g["label"]).repeat(2)
    _, _, pred = O.body_infer_continuity(ckpts["pixel"], ckpts["vq"], m0, m1, label, n0, n1, window=18)
    assert pred.shape == (2, 160, 129)
    assert np.abs(pred.numpy()[:, ::int(g["pred_stride"])] - g["pred"]).max() <= 1e-6
    assert np.abs(pred.numpy()[:, 52:68] - g["pred_seam"]).max() <= 1e-6
    out = O.body_vq_continuity(ckpts["vq"], synth.synth_poses(2, 300, seed=313)).numpy()
    assert out.shape == (300, 258)
    assert np.abs(out[::3] - g["vq_out"]).max() <= 1e-6
    assert np.abs(out[56:64] - g["vq_seam"]).max() <= 1e-6


def test_pixel_6d_geometry():
    """f2: the convert_to_6d geometry pixelcnn(2048, 512, 10, ...) + VQ-VAEs over 78 / 180 channels
    (nets/smplx_body_pixel.py:49-57), reference wrapper run by make_golden.py --only pixel_6d."""
    g = _load("pixel_6d")
    bp6, vq6 = synth.body_pixel_checkpoint_6d(0), synth.body_vq_checkpoint_6d(0)
    assert np.allclose(list(synth.fingerprint(bp6).values()), g["fp"], rtol=1e-12)
    assert np.allclose(list(synth.fingerprint(vq6).values()), g["fp_vq"], rtol=1e-12)
    noise = draw_noise(16, 2, int(g["sampler_seed"]))
    if not _same_noise(g, noise):
        pytest.skip("host RNG stream differs from the fixture machine")
    mfcc = synth.synth_mfcc(1, 32, seed=611).repeat(2, 1, 1)
    lab = torch.tensor([1, 1])
    lat, pred = O.body_generate(bp6, vq6, mfcc, lab, noise=noise)
    assert np.array_equal(lat.numpy(), g["codes"])
    assert pred.shape == (2, 32, 258)
    assert np.abs(pred.numpy() - g["pred"]).max() <= 1e-6
    audio = O.audio_encoder(bp6["audioencoder"], mfcc).unsqueeze(-1).repeat(1, 1, 1, 2)
    logits = O.pixelcnn_forward(bp6["generator"], lat, lab, audio)
    assert np.abs(logits[:, :, g["logit_rows"].tolist(), :].numpy() - g["logits"]).max() <= 1e-5


def test_vq_roundtrip(ckpts):
    g = _load("vq_roundtrip")
    poses = synth.synth_poses(2, 88)
    assert np.array_equal(poses.numpy(), g["poses"])
    ib, ih, out = O.body_vq_roundtrip(ckpts["vq"], poses)
    assert np.array_equal(ib.numpy(), g["idx_body"])
    assert np.array_equal(ih.numpy(), g["idx_hand"])
    assert out.shape == g["out"].shape
    assert np.abs(out.numpy() - g["out"]).max() <= 1e-6


def test_face(ckpts):
    g = _load("face")
    wave = synth.synth_wave(1, 64000)
    assert np.array_equal(wave[0, :64].numpy(), g["wave_head"])
    out = O.face_forward(ckpts["face"]["generator"], wave, torch.zeros(1, 4), 64000 * 30 // 16000)
    assert out.shape == (1, 120, 103)
    assert np.abs(out.numpy() - g["out_4s"]).max() <= 2e-5
    wave2 = synth.synth_wave(2, 24000, seed=5)
    ids = torch.nn.functional.one_hot(torch.tensor(g["ids_b2"]), 4).float()
    out2 = O.face_forward(ckpts["face"]["generator"], wave2, ids, 45)
    assert np.abs(out2.numpy() - g["out_b2"]).max() <= 2e-5


def test_face_10s(ckpts):
    """the benchmarked clip length (10 s -> 300 frames), two clips with speaker ids, reference-generated golden."""
    g = _load("face_10s")
    wave = synth.synth_wave(2, 160000, seed=int(g["wave_seed"]))
    ids = torch.nn.functional.one_hot(torch.tensor(g["ids"]), 4).float()
    out = O.face_forward(ckpts["face"]["generator"], wave, ids, 300).numpy()
    assert out.shape == (2, 300, 103)
    err = max(np.abs(out[:, ::int(g["out_stride"])] - g["out"]).max(), np.abs(out[:, :8] - g["out_head"]).max())
    assert err <= 2e-5, err


def test_pose_assembly_layout():
    face = torch.arange(5 * 103, dtype=torch.float32).view(5, 103)
    body = torch.arange(4 * 129, dtype=torch.float32).view(4, 129) + 1000
    full = O.assemble_pose(face, body)
    assert full.shape == (5, 265)
    assert torch.equal(full[:, :3], face[:, :3])
    assert torch.allclose(full[:, 9:12], torch.tensor([3.0747, -0.0158, -0.0152]).expand(5, 3))
    assert torch.equal(full[4, -100:], face[4, 3:])
    assert torch.equal(full[4, 3 + 15:3 + 18], body[3, :3])      # last body frame repeated


def test_rot6d_to_axis_angle_golden():
    """f2 piece: oracle restatement == the reference's matrix_to_axis_angle(rotation_6d_to_matrix(x)) run here
    (tests/golden/make_golden.py --only rot6d): random inputs, identity, tiny angles, near-pi turns, scaled inputs."""
    gold = np.load(os.path.join(GOLDEN, "rot6d.npz"))
    got = O.rot6d_to_axis_angle(torch.tensor(gold["d6"])).numpy()
    assert np.isfinite(got).all()
    assert np.abs(got - gold["aa"]).max() <= 1e-6
