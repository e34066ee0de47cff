"""GPU (-m gpu): oracle parity AT the shapes bench.py measures (BASELINE.json configs 3, 4, 5).

The sampler is compared bit-exact with the windowed oracle (window=18 >= the 17-row receptive
field, identical to the reference's full-grid loop: tests/test_oracle_golden.py), the decoded
poses and the face regressor within 1e-4 max-abs, the whole-body result through the public
host-buffer call.  The CPU oracle needs about a minute for the 64 x 75 sampler.
"""
import pytest
import torch

import talkshow_oracle as O
from conftest import draw_noise
from talkshow_b200 import synth

pytestmark = pytest.mark.gpu
TOL = 1e-4
FACE_CLIPS = (0, 31, 63)


@pytest.fixture(scope="module")
def wb(ckpts):
    from talkshow_b200.engine import Engine
    from talkshow_b200.pipeline import WholeBody

    e = Engine(0)
    w = WholeBody(e)
    w.load(ckpts["pixel"], ckpts["vq"], ckpts["face"])
    yield w
    torch.cuda.synchronize()
    e.close()


def _explain_mismatch(got, ref, ref_fn):
    """first differing draw, with the oracle's top-2 gap of p/q there (a near tie is the only legitimate cause)."""
    bad = (got != ref).nonzero()
    b, t, c = bad[0].tolist()
    return "first mismatch at sample %d row %d col %d: engine %d oracle %d (%d of %d draws differ)%s" % (
        b, t, c, int(got[b, t, c]), int(ref[b, t, c]), bad.shape[0], ref.numel(), ref_fn(b, t, c) if ref_fn else "")


@pytest.fixture(scope="module")
def cfg5(ckpts):
    """BASELINE config 5 on one GPU: 64 clips x 10 s, speaker ids arange(64) % 4."""
    B, M = 64, 300
    mfcc = synth.synth_mfcc(B, M, seed=1234)
    wave = synth.synth_wave(B, 160000, seed=1234)
    label = (torch.arange(B) % 4).to(torch.int64)
    T = O.latent_rows(M)
    noise = draw_noise(2 * T, B, 2024)
    torch.set_num_threads(min(32, torch.get_num_threads()))
    ref_codes, ref_body = O.body_generate(ckpts["pixel"], ckpts["vq"], mfcc, label, noise=noise, window=18)
    return dict(B=B, M=M, T=T, mfcc=mfcc, wave=wave, label=label, noise=noise, ref_codes=ref_codes, ref_body=ref_body)


def test_config5_sampler_b64_t75(wb, cfg5):
    """the benchmarked sampler shape: 9 600 draws bit-exact, decoded poses <= 1e-4."""
    c = cfg5
    codes, poses = wb.e.body_generate(c["mfcc"], c["label"], c["noise"])
    codes = codes.cpu()
    assert codes.shape == (64, 75, 2)
    assert torch.equal(codes, c["ref_codes"]), _explain_mismatch(codes, c["ref_codes"], None)
    err = (poses.cpu() - c["ref_body"]).abs().max().item()
    print("config 5 body poses max-abs err vs oracle: %.3e" % err)
    assert err <= TOL


def test_config4_diversity_b12_t75(wb, ckpts):
    """BASELINE config 4: 12 diversity samples of one 10 s clip, id 0."""
    B, M = 12, 300
    mfcc = synth.synth_mfcc(1, M, seed=77).repeat(B, 1, 1)
    label = torch.zeros(B, dtype=torch.int64)
    T = O.latent_rows(M)
    noise = draw_noise(2 * T, B, 2024)
    ref_codes, ref_body = O.body_generate(ckpts["pixel"], ckpts["vq"], mfcc, label, noise=noise, window=18)
    codes, poses = wb.e.body_generate(mfcc, label, noise)
    codes = codes.cpu()
    assert torch.equal(codes, ref_codes), _explain_mismatch(codes, ref_codes, None)
    assert (poses.cpu() - ref_body).abs().max().item() <= TOL
    assert len({tuple(x.flatten().tolist()) for x in codes}) == B          # 12 different sequences


def test_config3_b1_t30_4s(wb, ckpts):
    """BASELINE config 3: one 4 s clip, one sample, id 0 (fused body call)."""
    mfcc = synth.synth_mfcc(1, 120, seed=5)
    noise = draw_noise(60, 1, 2024)
    ref_codes, ref_body = O.body_generate(ckpts["pixel"], ckpts["vq"], mfcc, torch.tensor([0]), noise=noise, window=18)
    codes, poses = wb.e.body_generate(mfcc, torch.tensor([0]), noise)
    assert torch.equal(codes.cpu(), ref_codes)
    assert (poses.cpu() - ref_body).abs().max().item() <= TOL


@pytest.fixture(scope="module")
def face5(ckpts, cfg5):
    sel = list(FACE_CLIPS)
    return O.face_forward(ckpts["face"]["generator"], cfg5["wave"][sel], torch.zeros(len(sel), 4), 300)


def test_config5_face_b64(wb, cfg5, face5):
    """the B=64 x 10 s face forward (75 M-tiles, multi-GB workspace): clips 0, 31, 63 against the oracle run on
    those clips alone (per-clip independence of the regressor is part of what this checks)."""
    got = wb.e.face_forward(cfg5["wave"], torch.zeros(64, 4), 300).cpu()
    assert got.shape == (64, 300, 103) and torch.isfinite(got).all()
    err = (got[list(FACE_CLIPS)] - face5).abs().max().item()
    print("config 5 face max-abs err vs oracle (clips %s): %.3e" % (FACE_CLIPS, err))
    assert err <= TOL


def test_config5_whole_body_generate_host(wb, cfg5, face5):
    """config 5 through the public host-buffer call (what bench.py's e2e leg times)."""
    c = cfg5
    out = wb.generate_host(c["mfcc"].pin_memory(), c["wave"].pin_memory(), c["label"].pin_memory(), noise=c["noise"].cuda())
    assert out.shape == (64, 300, 265)
    for i, b in enumerate(FACE_CLIPS):
        ref = O.assemble_pose(face5[i], c["ref_body"][b])
        assert (out[b] - ref).abs().max().item() <= TOL
    # body columns of every clip (jaw 0:3 | pose | expression 165:265 come from the face regressor)
    ref_full = torch.stack([O.assemble_pose(torch.zeros(300, 103), c["ref_body"][b]) for b in range(64)])
    body_cols = [i for i in range(3, 165)]
    assert (out[:, :, body_cols] - ref_full[:, :, body_cols]).abs().max().item() <= TOL


@pytest.mark.parametrize("B,world", [(12, 8), (16, 2)])
def test_sharded_equals_unsharded(wb, B, world):
    """SURVEY.md §8e G-independence on the device: the full-batch noise is drawn once and sliced per rank; the shards
    of every rank (config 4's uneven 12 -> 2,2,2,2,1,1,1,1 split; an even 2-way split), run one after the other here,
    concatenate to exactly the one-GPU result — sampled codes and poses bit-identical, whatever position a sample has
    inside the sampler's batch tile."""
    from talkshow_b200.pipeline import shard_range

    M, N = 80, 16000 * 80 // 30
    mfcc = synth.synth_mfcc(B, M, seed=501).cuda()
    wave = synth.synth_wave(B, N, seed=502).cuda()
    label = (torch.arange(B) % 4).cuda()
    noise = draw_noise(2 * O.latent_rows(M), B, 503).cuda()
    full = wb.generate(mfcc, wave, label, noise=noise)
    parts = [wb.generate_sharded(mfcc, wave, label, r, world, noise_full=noise, gather=False) for r in range(world)]
    assert [p.shape[0] for p in parts] == [shard_range(B, r, world)[1] - shard_range(B, r, world)[0] for r in range(world)]
    got = torch.cat(parts, 0)
    body_cols = list(range(3, 165))
    assert torch.equal(got[:, :, body_cols], full[:, :, body_cols])          # sampler + VQ decoders: bit-identical
    assert (got - full).abs().max().item() <= 1e-5                           # face GEMM tiles depend on the batch size


def test_overlapped_small_batch_equals_sequential(ckpts):
    """8 clips per GPU (config 5 on 8 GPUs): the body path on a 100-CTA sampler plan and the face path run side by side on
    two streams; the result is bit-identical to the sequential order."""
    from talkshow_b200.engine import Engine
    from talkshow_b200.pipeline import WholeBody

    B, M = 8, 300
    mfcc = synth.synth_mfcc(B, M, seed=901).cuda()
    wave = synth.synth_wave(B, 160000, seed=902).cuda()
    label = (torch.arange(B) % 4).cuda()
    noise = draw_noise(2 * O.latent_rows(M), B, 903).cuda()
    outs = []
    for ob in (0, 8):
        e = Engine(0)
        w = WholeBody(e, overlap_batch=ob)
        w.load(ckpts["pixel"], ckpts["vq"], ckpts["face"])
        assert (w.e2 is not None) == (ob > 0)
        for _ in range(2):
            out = w.generate(mfcc, wave, label, noise=noise)
        torch.cuda.synchronize()
        ts = []
        for _ in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); out = w.generate(mfcc, wave, label, noise=noise); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        print("whole body, 8 clips x 10 s, overlap_batch=%d: %.2f ms per step" % (ob, min(ts)))
        outs.append(out.clone())
        w.close()
        e.close()
    assert torch.equal(outs[0], outs[1])
