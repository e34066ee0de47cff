"""CPU: the PixelCNN execution plan (stage table + packed weights built by the C++ packer in
host-only mode) reproduces the oracle's logits and sampled codes when interpreted in numpy."""
import numpy as np
import pytest
import torch

import plan_emulator as PE
import talkshow_oracle as O
from conftest import draw_noise
from talkshow_b200 import _lib, synth
from talkshow_b200.engine import Engine


# name -> (fusion level, persistent CTAs or 0 = one per SM).  fused96: the partial-GPU plan of the two-stream step (output_conv.2
# over two stages); sched2_*: schedule 2 on partial-GPU plans (88: + output_conv.2 split; 80: + layer-0 vertical stack one column
# per stage, 55 stages)
PLANS = {"fused": (1, 0), "plain": (0, 0), "sched2": (2, 0), "fused96": (1, 96), "sched2_96": (2, 96), "sched2_88": (2, 88),
         "sched2_80": (2, 80)}


@pytest.fixture(scope="module", params=list(PLANS))
def plan(ckpts, request):
    e = Engine(-148)            # host-only planning engine sized for 148 SMs
    level, ctas = PLANS[request.param]
    e.set_pixelcnn_fusion(level)
    if ctas:
        e.set_pixelcnn_ctas(ctas)
    e.load_pixelcnn(ckpts["pixel"]["generator"])
    table, blob = _lib.plan_to_numpy(e.h)
    rb = e.pixelcnn_row_bytes
    e.close()
    return PE.Plan(table, blob), rb


def _audio_terms(sd, aud):
    """aud [B,256,T] -> AUDV, AUDH [B,T,256] (what the engine precomputes with three GEMMs)."""
    a = torch.einsum("oc,bct->bto", sd["embedding_aud.weight"][:, :, 0, 0], aud) + sd["embedding_aud.bias"]
    av = torch.einsum("oc,btc->bto", sd["fusion_v.weight"][:, 256:, 0, 0], a) + sd["fusion_v.bias"]
    ah = torch.einsum("oc,btc->bto", sd["fusion_h.weight"][:, 256:, 0, 0], a) + sd["fusion_h.bias"]
    return av.numpy(), ah.numpy()


def test_plan_shape(plan):
    p, row_bytes = plan
    assert p.ncta in (80, 88, 96, 148) and p.L == 15 and p.nstages in (52, 54, 55, 84)
    # 2048 / (16 rows x CTAs) -> output_conv.2 takes two stages per column below 128 CTAs; below 86 CTAs the layer-0 vertical
    # stack (K = 1536: 12 rows per CTA) takes one stage per column
    assert p.nstages == {148: p.nstages, 96: 54, 88: 54, 80: 55}[p.ncta]
    fused = p.nstages in (52, 54, 55)
    assert (p.table[:, :, 0] >= 11).any() == fused      # EPI_HRESF / EPI_HGATE2 / EPI_OUT1F only in the fused plan
    assert row_bytes == 89774080        # SURVEY.md §8d algorithmic bytes per latent row
    t = p.table
    assert (t[:, :, 4][t[:, :, 6] > 0] <= (64 if p.cl == 4 else 16)).all()     # weight rows per CTA / per cluster
    t = t[:, ::p.cl]                                        # one entry per work unit
    # every output row of every stage is owned by exactly one CTA
    for s in range(p.nstages):
        for epi, layer, col in {tuple(x) for x in t[s][:, :3].tolist() if x[0] not in (0, 10)}:
            sel = t[s][(t[s][:, 0] == epi) & (t[s][:, 1] == layer) & (t[s][:, 2] == col)]
            rows = sorted((int(r0), int(n)) for r0, n in sel[:, 3:5])
            pos = rows[0][0]
            assert pos in (0, 1024)                          # 1024: second half of a split output_conv.2
            for r0, n in rows:
                assert r0 == pos
                pos += n
            assert pos - rows[0][0] in (256, 512, 1024, 2048)
    if p.hvslots > 2:                                       # schedule 2: no vert_to_horiz in the vertical stages 2..15
        vert = np.array([(t[s, :, 0] == 2).any() for s in range(p.nstages)])     # stages holding EPI_VERT tasks (layers >= 1)
        assert vert.sum() == p.L - 1
        assert not ((t[vert][:, :, 0] == PE.EPI_V2H) | (t[vert][:, :, 0] == PE.EPI_V2H1)).any()
        last = int(np.nonzero(vert)[0].max())
        assert (t[last + 1:, :, 0] == PE.EPI_V2H1).any()


def test_plan_teacher_forced_logits(plan, ckpts):
    p, _ = plan
    sd = ckpts["pixel"]["generator"]
    B, T = 3, 6
    g = torch.Generator().manual_seed(3)
    codes = torch.randint(0, 2048, (B, T, 2), generator=g)
    label = torch.tensor([0, 3, 1])
    aud = O.audio_encoder(ckpts["pixel"]["audioencoder"], synth.synth_mfcc(B, 4 * T, seed=5))
    ref = O.pixelcnn_forward(sd, codes, label, aud.unsqueeze(-1).repeat(1, 1, 1, 2))     # [B,2048,T,2]
    av, ah = _audio_terms(sd, aud)
    cls_w = [sd["layers.%d.class_cond_embedding.weight" % l].numpy() for l in range(p.L)]
    got_codes, logits = PE.run(p, sd["embedding.weight"].numpy(), cls_w, av, ah, label.numpy(), codes.numpy(), T)
    assert np.array_equal(got_codes, codes.numpy())
    got = np.transpose(logits, (2, 3, 0, 1))                                             # [B,2048,T,2]
    err = np.abs(got - ref.numpy()).max()
    assert err <= 2e-4, err


def test_plan_sampling_with_prefix(plan, ckpts):
    """free-running sampling after a forced prefix == oracle generate(pre_latents=...)."""
    p, _ = plan
    sd = ckpts["pixel"]["generator"]
    B, T0, T = 2, 2, 3
    label = torch.tensor([2, 0])
    aud = O.audio_encoder(ckpts["pixel"]["audioencoder"], synth.synth_mfcc(B, 4 * (T0 + T), seed=6))
    a2 = aud.unsqueeze(-1).repeat(1, 1, 1, 2)
    g = torch.Generator().manual_seed(4)
    pre = torch.randint(0, 2048, (B, T0, 2), generator=g)
    noise = draw_noise(2 * T, B, 11)
    ref = O.pixelcnn_generate(sd, label, T, B, a2[:, :, T0:], noise=noise, pre_latents=pre, pre_audio=a2[:, :, :T0])
    av, ah = _audio_terms(sd, aud)
    cls_w = [sd["layers.%d.class_cond_embedding.weight" % l].numpy() for l in range(p.L)]
    got, _ = PE.run(p, sd["embedding.weight"].numpy(), cls_w, av, ah, label.numpy(), pre.numpy(), T0 + T,
                    noise=noise.numpy(), T0=T0)
    assert np.array_equal(got[:, T0:], ref.numpy())


WBUF, MAXROWS = 18560, 16        # csrc/pixelcnn.h: floats per weight staging buffer, accumulator rows per CTA


@pytest.mark.parametrize("level,ctas", [(1, 96), (1, 98), (1, 110), (1, 126), (1, 128), (1, 146), (2, 80), (2, 82), (2, 86),
                                        (2, 94), (2, 112), (2, 128), (2, 130), (2, 148)])
def test_plan_invariants_over_cta_counts(ckpts, level, ctas):
    """ts_set_pixelcnn_ctas: every even count the stage table allows gives a plan the kernel can run — each task's weight slab
    fits one staging buffer, at most 16 rows per CTA, every output row of every job owned exactly once, the same set of
    (job, layer, column) per row whatever the count."""
    e = Engine(-148)
    e.set_pixelcnn_fusion(level)
    e.set_pixelcnn_ctas(ctas)
    e.load_pixelcnn(ckpts["pixel"]["generator"])
    table, blob = _lib.plan_to_numpy(e.h)
    e.close()
    p = PE.Plan(table, blob)
    assert p.ncta == ctas
    t = p.table
    mm = t[:, :, 6] > 0
    assert ((t[:, :, 6] + 1) * t[:, :, 7])[mm].max() <= WBUF            # (K + 1) x padded rows
    assert t[:, :, 4][mm].max() <= MAXROWS and (t[:, :, 7] % 4 == 0).all()
    assert (t[:, :, 5][mm] + ((t[:, :, 6] + 1) * t[:, :, 7])[mm]).max() <= len(blob)
    jobs = {}
    for s in range(p.nstages):
        for epi, layer, col, r0, n in t[s][:, :5].tolist():
            if epi not in (0, 10):
                jobs.setdefault((epi, layer, col), []).append((r0, n))
    want = {1: 512, 2: 512, 3: 512, 4: 256, 5: 512, 6: 256, 11: 256, 12: 512, 13: 512, 9: 2048, 14: 512}   # output rows per job kind
    for (epi, layer, col), rows in jobs.items():
        rows.sort()
        pos = 0
        for r0, n in rows:
            assert r0 == pos and n > 0, (epi, layer, col, rows)
            pos += n
        assert pos == want[epi], (epi, layer, col, pos)
    n_v2h = sum(1 for k in jobs if k[0] in (3, 14))
    assert n_v2h == (15 if level == 1 else 1 + 2 * 14)                 # per layer (both columns in two passes) / per layer and column


@pytest.mark.parametrize("level,ctas", [(1, 94), (1, 64), (2, 78), (0, 96), (1, 32), (1, 150)])
def test_plan_rejects_too_few_ctas(ckpts, level, ctas):
    e = Engine(-148)
    e.set_pixelcnn_fusion(level)
    e.set_pixelcnn_ctas(ctas)
    with pytest.raises(RuntimeError, match="pixelcnn plan"):
        e.load_pixelcnn(ckpts["pixel"]["generator"])
    e.close()


def test_default_plan_tables_are_pinned(ckpts):
    """The stage tables of the plans the product runs by default (one CTA per SM, and the 96-CTA plan of the overlapped step) and of
    the two cross-check plans: any change to the builder that moves a task shows up here before it reaches a GPU (the tables of
    these four plans were the same before and after the round-2 builder changes; the kernels ran on exactly these)."""
    import hashlib

    want = {("fused", 0): "d7da3b16adf5", ("plain", 0): "b3d2b9c3c967", ("sched2", 0): "46eeee405bdc", ("fused", 96): "87f66d560666"}
    for (name, ctas), digest in want.items():
        e = Engine(-148)
        e.set_pixelcnn_fusion({"plain": 0, "sched2": 2}.get(name, 1))
        if ctas:
            e.set_pixelcnn_ctas(ctas)
        e.load_pixelcnn(ckpts["pixel"]["generator"])
        table, _ = _lib.plan_to_numpy(e.h)
        e.close()
        assert hashlib.sha1(table.tobytes()).hexdigest()[:12] == digest, (name, ctas)
