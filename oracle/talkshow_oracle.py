"""CPU oracle for the TalkSHOW speech-to-motion hot path.

TEST INFRASTRUCTURE ONLY.  This file is a CPU (PyTorch fp32, ATen ops) restatement of the
reference's algorithm, used by ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` as the *checker* and the timed CPU arm.  Nothing under
``talkshow_b200/`` imports it; the product path is CUDA only and fails loudly without its
extension.

Each function cites the reference file:line it follows (paths relative to yhw-yhw/TalkSHOW).
Pinning: the reference holds no golden vectors.  The oracle is pinned against
outputs of the reference's own modules run in the build container on seeded synthetic
checkpoints: ``tests/golden/make_golden.py`` (imports /root/reference) wrote
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks this file reproduces them
(bit-exact indices, <=1e-6 floats).  Third-party arithmetic restated here: HF transformers
``Wav2Vec2Model`` (reference pins transformers~=4.22.1, requirements.txt:2; restated from the
published wav2vec2-base architecture: 7-layer conv feature extractor with GroupNorm on layer 0,
post-LN transformer), torch ``multinomial`` (= argmax(p / Exp(1) noise), ATen
native/Distributions.cpp / Sampling).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

# ---------------------------------------------------------------------------------------------
# Gated PixelCNN prior  (nets/spg/gated_pixelcnn_v2.py)
# ---------------------------------------------------------------------------------------------


def _gate(x):
    """GatedActivation, gated_pixelcnn_v2.py:16-22."""
    a, b = x.chunk(2, dim=1)
    return torch.tanh(a) * torch.sigmoid(b)


def _n_layers(sd):
    n = 0
    while "layers.%d.vert_stack.weight" % n in sd:
        n += 1
    return n


def pixelcnn_forward(sd, x, label, aud):
    """GatedPixelCNN.forward, gated_pixelcnn_v2.py:130-150 with GatedMaskedConv2d.forward :61-87
    (bh_model=True, audio=True, eval: the Dropout on the ones vector :139-141 is identity).

    x [B,H,2] int64 codes, label [B] int64, aud [B,256,H,2] -> logits [B,2048,H,2]."""
    dim = sd["embedding.weight"].shape[1]
    shp = x.size() + (-1,)
    e = F.embedding(x.reshape(-1), sd["embedding.weight"]).view(shp).permute(0, 3, 1, 2)
    x_v, x_h = e, e
    for l in range(_n_layers(sd)):
        p = "layers.%d." % l
        k = 7 if l == 0 else 3
        if l == 1:
            a = F.conv2d(aud, sd["embedding_aud.weight"], sd["embedding_aud.bias"])
            x_v = F.conv2d(torch.cat([x_v, a], 1), sd["fusion_v.weight"], sd["fusion_v.bias"])
            x_h = F.conv2d(torch.cat([x_h, a], 1), sd["fusion_h.weight"], sd["fusion_h.bias"])
        wv, wh = sd[p + "vert_stack.weight"], sd[p + "horiz_stack.weight"]
        if l == 0:  # make_causal(), :57-59 (mask 'A')
            wv = wv.clone()
            wv[:, :, -1].zero_()
            wh = wh.clone()
            wh[:, :, :, -1].zero_()
        h = F.embedding(label, sd[p + "class_cond_embedding.weight"])
        h_vert = F.conv2d(x_v, wv, sd[p + "vert_stack.bias"], 1, (k // 2, 1))[:, :, : x_v.size(-2), :]
        out_v = _gate(h_vert + h[:, :, None, None])
        h_horiz = F.conv2d(x_h, wh, sd[p + "horiz_stack.bias"], 1, (0, 1))[:, :, :, : x_h.size(-1)]
        v2h = F.conv2d(h_vert, sd[p + "vert_to_horiz.weight"], sd[p + "vert_to_horiz.bias"])
        out = _gate(v2h + h_horiz + h[:, :, None, None])
        out_h = F.conv2d(out, sd[p + "horiz_resid.weight"], sd[p + "horiz_resid.bias"])
        if l > 0:
            out_h = out_h + x_h
        x_v, x_h = out_v, out_h
    y = F.relu(F.conv2d(x_h, sd["output_conv.0.weight"], sd["output_conv.0.bias"]))
    return F.conv2d(y, sd["output_conv.2.weight"], sd["output_conv.2.bias"])


def draw(probs, noise=None):
    """probs.multinomial(1) (gated_pixelcnn_v2.py:175).  ATen's single-sample path is
    ``argmax(probs / q)``, ``q = empty_like(probs).exponential_(1)`` from the default generator of
    the tensor's device; when ``noise`` is given it is that q (drawn by the caller in the same
    order and shape), which is the RNG contract of the C-ABI (include/talkshow_b200.h)."""
    if noise is None:
        return probs.multinomial(1).squeeze(-1)
    return torch.argmax(probs / noise, dim=-1)


def pixelcnn_generate(sd, label, T, B, aud_feat, noise=None, pre_latents=None, pre_audio=None,
                      window=None, return_logits=False):
    """GatedPixelCNN.generate, gated_pixelcnn_v2.py:152-177: 2*T full forwards, one sampled
    position each (the reference's O(T^2) loop).

    ``window=W`` evaluates each forward on rows [i-W+1, i] only.  The network's receptive field is
    17 rows back, so W>=18 is bit-identical on the same host and is what the tests
    use for the larger cases; ``window=None`` is the literal reference cost (used for timing).
    ``noise`` [2T,B,2048] replaces the generator (see ``draw``)."""
    x = torch.zeros((B, T, 2), dtype=torch.int64)
    h0 = 0
    if pre_latents is not None:
        x = torch.cat([pre_latents, x], 1)
        aud_feat = torch.cat([pre_audio, aud_feat], 2)
        h0 = pre_latents.shape[1]
    h = h0 + T
    logs = []
    step = 0
    for i in range(h0, h):
        lo = 0 if window is None else max(0, i - window + 1)
        for j in range(2):
            logits = pixelcnn_forward(sd, x[:, lo:i + 1] if window else x, label,
                                      aud_feat[:, :, lo:i + 1] if window else aud_feat)
            lg = logits[:, :, (i - lo) if window else i, j]
            if return_logits:
                logs.append(lg.clone())
            probs = F.softmax(lg, -1)
            x[:, i, j] = draw(probs, None if noise is None else noise[step])
            step += 1
    out = x[:, h0:h]
    return (out, torch.stack(logs, 0)) if return_logits else out


# ---------------------------------------------------------------------------------------------
# VQ-VAE 1-D  (nets/spg/vqvae_1d.py, nets/spg/vqvae_modules.py)
# ---------------------------------------------------------------------------------------------


def _bn(sd, p, x):
    return F.batch_norm(x, sd[p + "running_mean"], sd[p + "running_var"], sd[p + "weight"], sd[p + "bias"],
                        False, 0.0, 1e-5)


def conv_norm_relu(sd, p, x, sample="none", residual=False, slope=0.2):
    """vqvae_modules.ConvNormRelu.forward :167-172 (leaky=True everywhere on this path)."""
    if sample == "up":
        out = F.conv_transpose1d(x, sd[p + "conv.weight"], sd[p + "conv.bias"], 2, 1)
    elif sample == "down":
        out = F.conv1d(x, sd[p + "conv.weight"], sd[p + "conv.bias"], 2, 1)
    else:
        out = F.conv1d(x, sd[p + "conv.weight"], sd[p + "conv.bias"], 1, 1)
    out = _bn(sd, p + "norm.", out)
    if residual:
        if sample == "up":
            out = out + F.conv_transpose1d(x, sd[p + "residual_layer.weight"], sd[p + "residual_layer.bias"], 2, 1)
        else:
            out = out + F.conv1d(x, sd[p + "residual_layer.weight"], sd[p + "residual_layer.bias"], 2, 1)
    return F.leaky_relu(out, slope)


def res_stack(sd, p, x):
    """Res_CNR_Stack.forward, vqvae_modules.py:204-212."""
    h = x
    i = 0
    while p + "_layers.%d.conv.weight" % i in sd:
        h = conv_norm_relu(sd, p + "_layers.%d." % i, h)
        i += 1
    h = _bn(sd, p + "norm.", F.conv1d(h, sd[p + "conv.weight"], sd[p + "conv.bias"], 1, 1))
    return F.relu(h + x)


def encoder_trunk(sd, p, x):
    """AudioEncoder.forward vqvae_1d.py:27-34 == Encoder.forward :84-91 without pre_vq_conv."""
    h = conv_norm_relu(sd, p + "project.", x)
    h = res_stack(sd, p + "_enc_1.", h)
    h = conv_norm_relu(sd, p + "_down_1.", h, "down", True)
    h = res_stack(sd, p + "_enc_2.", h)
    h = conv_norm_relu(sd, p + "_down_2.", h, "down", True)
    return res_stack(sd, p + "_enc_3.", h)


def audio_encoder(sd, mfcc):
    """mfcc [B,64,M] -> [B,256,M//4]."""
    return encoder_trunk(sd, "", mfcc)


def vq_code_indices(sd, flat_x):
    """VectorQuantizerEMA.get_code_indices, vqvae_modules.py:311-319."""
    emb = sd["vq_layer.embeddings"]
    d = torch.sum(flat_x ** 2, 1, keepdim=True) + torch.sum(emb ** 2, 1) - 2.0 * torch.matmul(flat_x, emb.t())
    return torch.argmin(d, 1)


def vq_encode(sd, gt_poses):
    """VQVAE.encode, vqvae_1d.py:196-199 (eval): gt_poses [B,F,C] -> (e [B,64,T], idx [B,T])."""
    z = encoder_trunk(sd, "encoder.", gt_poses.transpose(1, 2))
    z = F.conv1d(z, sd["encoder.pre_vq_conv.weight"], sd["encoder.pre_vq_conv.bias"])
    x = z.permute(0, 2, 1).contiguous()
    idx = vq_code_indices(sd, x.reshape(-1, x.shape[-1]))
    q = F.embedding(idx, sd["vq_layer.embeddings"]).view_as(x).permute(0, 2, 1).contiguous()
    return q, idx.view(q.shape[0], q.shape[2])


def vq_decoder(sd, e):
    """Decoder.forward, vqvae_1d.py:139-149: e [B,64,T] -> [B,C,4T]."""
    h = F.conv1d(e, sd["decoder.aft_vq_conv.weight"], sd["decoder.aft_vq_conv.bias"])
    h = res_stack(sd, "decoder._dec_1.", h)
    h = conv_norm_relu(sd, "decoder._up_2.", h, "up", True)
    h = res_stack(sd, "decoder._dec_2.", h)
    h = conv_norm_relu(sd, "decoder._up_3.", h, "up", True)
    h = res_stack(sd, "decoder._dec_3.", h)
    return F.conv1d(h, sd["decoder.project.weight"], sd["decoder.project.bias"])


def vq_decode(sd, latents):
    """VQVAE.decode(latents=...), vqvae_1d.py:201-208 + quantize vqvae_modules.py:321-323."""
    b, w = latents.shape
    e = F.embedding(latents, sd["vq_layer.embeddings"]).view(b, w, -1).permute(0, 2, 1).contiguous()
    return vq_decoder(sd, e)


def vq_roundtrip(sd, gt_poses):
    """VQVAE.forward in eval, vqvae_1d.py:184-189 -> (idx [B,T], recon [B,C,F])."""
    q, idx = vq_encode(sd, gt_poses)
    return idx, vq_decoder(sd, q)


# ---------------------------------------------------------------------------------------------
# body wrappers  (nets/smplx_body_pixel.py, nets/smplx_body_vq.py)
# ---------------------------------------------------------------------------------------------


def body_generate(ckpt, vq_ckpt, mfcc, label, noise=None, window=None):
    """s2g_body_pixel.infer_on_audio :270-285 after feature extraction:
    mfcc [B,64,M], label [B] -> (codes [B,T,2], poses [B,F,129])."""
    B = mfcc.shape[0]
    audio = audio_encoder(ckpt["audioencoder"], mfcc).unsqueeze(-1).repeat(1, 1, 1, 2)
    lat = pixelcnn_generate(ckpt["generator"], label, audio.shape[2], B, audio, noise=noise, window=window)
    body = vq_decode(vq_ckpt["g_body"], lat[..., 0])
    hand = vq_decode(vq_ckpt["g_hand"], lat[..., 1])
    return lat, torch.cat([body, hand], 1).transpose(1, 2)


def body_infer_continuity(ckpt, vq_ckpt, mfcc0, mfcc1, label, noise0=None, noise1=None, window=None):
    """s2g_body_pixel.infer_on_audio(continuity=True), nets/smplx_body_pixel.py:244-269 + infer :291-304, after
    get_mfcc_sepa (data_utils/utils.py:234-263): the 2 s prefix and the remainder are encoded SEPARATELY by the
    audio encoder, the remainder is sampled with pre_latents / pre_audio of the prefix, and each chunk is decoded
    SEPARATELY (Decoder.forward ignores pre_state, vqvae_1d.py:139-149) before the time-axis concat.
    mfcc0 [B,64,M0], mfcc1 [B,64,M1] -> (lat0, lat1, poses [B,F0+F1,129])."""
    B = mfcc0.shape[0]
    a0 = audio_encoder(ckpt["audioencoder"], mfcc0).unsqueeze(-1).repeat(1, 1, 1, 2)
    a1 = audio_encoder(ckpt["audioencoder"], mfcc1).unsqueeze(-1).repeat(1, 1, 1, 2)
    lat0 = pixelcnn_generate(ckpt["generator"], label, a0.shape[2], B, a0, noise=noise0, window=window)
    lat1 = pixelcnn_generate(ckpt["generator"], label, a1.shape[2], B, a1, noise=noise1, pre_latents=lat0, pre_audio=a0,
                             window=window)
    body = torch.cat([vq_decode(vq_ckpt["g_body"], lat0[..., 0]), vq_decode(vq_ckpt["g_body"], lat1[..., 0])], 2)
    hand = torch.cat([vq_decode(vq_ckpt["g_hand"], lat0[..., 1]), vq_decode(vq_ckpt["g_hand"], lat1[..., 1])], 2)
    return lat0, lat1, torch.cat([body, hand], 1).transpose(1, 2)


_FIX_3D = [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 21, 22, 23, 24, 25, 26,
           30, 31, 32, 33, 34, 35, 45, 46, 47, 48, 49, 50]
C_INDEX_3D = [i for i in range(165) if i not in _FIX_3D]          # data_utils/lower_body.py:44-56


def body_vq_roundtrip(vq_ckpt, initial_pose):
    """s2g_body_vq.infer_on_audio :208-295 (composition, no continuity/smooth):
    initial_pose [B,165,F] -> (idx_body [B,T], idx_hand [B,T], out numpy-layout [F, B*129])."""
    gt = initial_pose[:, C_INDEX_3D].permute(0, 2, 1)
    ib, rb = vq_roundtrip(vq_ckpt["g_body"], gt[..., :39])
    ih, rh = vq_roundtrip(vq_ckpt["g_hand"], gt[..., 39:])
    pred = torch.cat([rb, rh], 1).transpose(1, 2)                  # [B,F,129]
    return ib, ih, torch.cat(list(pred), 1)                        # np.concatenate(output, axis=1), :293


def body_vq_continuity(vq_ckpt, initial_pose, chunk=60, nchunks=5):
    """s2g_body_vq.infer_on_audio(continuity=True) :256-271: five 60-frame chunks, each round-tripped on its own
    (pre_state is ignored by Decoder.forward), concatenated in time -> numpy-layout [5*60, B*129]."""
    gt = initial_pose[:, C_INDEX_3D].permute(0, 2, 1)
    bs, hs = [], []
    for i in range(nchunks):
        seg = gt[:, i * chunk:(i + 1) * chunk]
        bs.append(vq_roundtrip(vq_ckpt["g_body"], seg[..., :39])[1])
        hs.append(vq_roundtrip(vq_ckpt["g_hand"], seg[..., 39:])[1])
    pred = torch.cat([torch.cat(bs, 2), torch.cat(hs, 2)], 1).transpose(1, 2)
    return torch.cat(list(pred), 1)


# ---------------------------------------------------------------------------------------------
# face regressor  (nets/spg/s2g_face.py, nets/spg/wav2vec.py, nets/layers.py, HF Wav2Vec2Model)
# ---------------------------------------------------------------------------------------------

W2V_KERNEL = (10, 3, 3, 3, 3, 2, 2)
W2V_STRIDE = (5, 2, 2, 2, 2, 2, 2)


def pos_conv_weight(sd, p):
    """Effective weight of the weight_norm'd positional conv (dim=2): w = g * v / ||v|| with the
    norm over every dim but 2 (torch._weight_norm).  Accepts new (parametrizations.weight.
    original0/1) and old (weight_g / weight_v) names."""
    if p + "parametrizations.weight.original0" in sd:
        g, v = sd[p + "parametrizations.weight.original0"], sd[p + "parametrizations.weight.original1"]
    else:
        g, v = sd[p + "weight_g"], sd[p + "weight_v"]
    return torch._weight_norm(v, g, 2)


def wav2vec2(sd, p, wave, frame_num):
    """reference Wav2Vec2Model.forward, nets/spg/wav2vec.py:76-143 (eval, no attention mask):
    feature extractor -> linear_interpolation(50->30 fps, output_len=frame_num) :64-70,92-95 ->
    feature_projection -> encoder (pos-conv, LN, 12 post-LN layers).  wave [B,N] -> [B,frame,768]."""
    h = wave[:, None, :]
    for i, (k, s) in enumerate(zip(W2V_KERNEL, W2V_STRIDE)):
        q = p + "feature_extractor.conv_layers.%d." % i
        h = F.conv1d(h, sd[q + "conv.weight"], None, s)
        if i == 0:
            h = F.group_norm(h, 512, sd[q + "layer_norm.weight"], sd[q + "layer_norm.bias"], 1e-5)
        h = F.gelu(h)
    h = F.interpolate(h, size=frame_num, align_corners=False, mode="linear")       # [B,512,frame]
    h = h.transpose(1, 2)
    q = p + "feature_projection."
    h = F.layer_norm(h, (512,), sd[q + "layer_norm.weight"], sd[q + "layer_norm.bias"], 1e-5)
    h = F.linear(h, sd[q + "projection.weight"], sd[q + "projection.bias"])
    e = p + "encoder."
    pc = F.conv1d(h.transpose(1, 2), pos_conv_weight(sd, e + "pos_conv_embed.conv."),
                  sd[e + "pos_conv_embed.conv.bias"], 1, 64, 1, 16)[:, :, :-1]
    h = h + F.gelu(pc).transpose(1, 2)
    h = F.layer_norm(h, (768,), sd[e + "layer_norm.weight"], sd[e + "layer_norm.bias"], 1e-5)
    B, T, _ = h.shape
    l = 0
    while e + "layers.%d.attention.q_proj.weight" % l in sd:
        q = e + "layers.%d." % l
        a = q + "attention."
        qs = F.linear(h, sd[a + "q_proj.weight"], sd[a + "q_proj.bias"]).view(B, T, 12, 64).transpose(1, 2)
        ks = F.linear(h, sd[a + "k_proj.weight"], sd[a + "k_proj.bias"]).view(B, T, 12, 64).transpose(1, 2)
        vs = F.linear(h, sd[a + "v_proj.weight"], sd[a + "v_proj.bias"]).view(B, T, 12, 64).transpose(1, 2)
        w = F.softmax(torch.matmul(qs, ks.transpose(2, 3)) * (64 ** -0.5), dim=-1)
        o = torch.matmul(w, vs).transpose(1, 2).reshape(B, T, 768)
        h = h + F.linear(o, sd[a + "out_proj.weight"], sd[a + "out_proj.bias"])
        h = F.layer_norm(h, (768,), sd[q + "layer_norm.weight"], sd[q + "layer_norm.bias"], 1e-5)
        f = F.gelu(F.linear(h, sd[q + "feed_forward.intermediate_dense.weight"],
                            sd[q + "feed_forward.intermediate_dense.bias"]))
        h = h + F.linear(f, sd[q + "feed_forward.output_dense.weight"], sd[q + "feed_forward.output_dense.bias"])
        h = F.layer_norm(h, (768,), sd[q + "final_layer_norm.weight"], sd[q + "final_layer_norm.bias"], 1e-5)
        l += 1
    return h


def _cnr_ln(sd, p, x, residual=None):
    """layers.ConvNormRelu.forward with norm='ln', nets/layers.py:142-151:
    relu(LN_C(conv(x)) + residual(x))."""
    out = F.conv1d(x, sd[p + "conv.weight"], sd[p + "conv.bias"], 1, 1)
    c = out.shape[1]
    out = F.layer_norm(out.transpose(1, 2), (c,), sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-5).transpose(1, 2)
    if residual == "conv":
        out = out + F.conv1d(x, sd[p + "residual_layer.0.weight"], sd[p + "residual_layer.0.bias"], 1, 1)
    elif residual == "id":
        out = out + x
    return F.relu(out)


def face_forward(sd, wave, id_onehot, frame):
    """s2g_face.Generator.forward, nets/spg/s2g_face.py:196-224 (eval, encoder_choice
    'faceformer', identity=True).  wave [B,N], id_onehot [B,4] float -> [B,frame,103]."""
    h = wav2vec2(sd, "audio_encoder.", wave, frame)
    feat = F.linear(h, sd["audio_feature_map.weight"], sd["audio_feature_map.bias"]).transpose(1, 2)
    idv = id_onehot.reshape(id_onehot.shape[0], -1, 1).repeat(1, 1, feat.shape[2]).to(feat.dtype)
    idv = F.conv1d(idv, sd["audio_middle.id_mlp.weight"], sd["audio_middle.id_mlp.bias"])
    x = torch.cat([feat, idv.expand(feat.shape[0], -1, -1)], 1)
    f = "audio_middle.first_net.conv_layers."
    x = _cnr_ln(sd, f + "0.", x, "conv")
    x = _cnr_ln(sd, f + "1.", x, "id")
    x = _cnr_ln(sd, f + "2.", x, "id")
    outs = []
    for b in range(2):
        m = x
        for i in range(3):
            m = _cnr_ln(sd, "decoder.%d.%d." % (b, i), m)
        outs.append(F.conv1d(m, sd["final_out.%d.weight" % b], sd["final_out.%d.bias" % b]))
    return torch.cat(outs, 1).transpose(1, 2)


# ---------------------------------------------------------------------------------------------
# pose assembly  (scripts/demo.py:182-229, data_utils/lower_body.py:68-87)
# ---------------------------------------------------------------------------------------------

LOWER_POSE = [0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 3.0747, -0.0158, -0.0152, -1.1826512813568115, 0.23866955935955048,
              0.15146760642528534, -1.2604516744613647, -0.3160211145877838, -0.1603458970785141,
              1.1654603481292725, 0.0, 0.0, 1.2521806955337524, 0.041598282754421234, -0.06312154978513718,
              0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0]


def part2full(x, stand=False):
    """data_utils/lower_body.py:68-87: [F,232] -> [F,265]."""
    lp = torch.tensor(LOWER_POSE, dtype=x.dtype)
    if stand:
        lp = torch.zeros_like(lp)
        lp[6:9] = torch.tensor([3.0747, -0.0158, -0.0152])
    lp = lp[None].repeat(x.shape[0], 1)
    return torch.cat([x[:, :3], lp[:, :15], x[:, 3:6], lp[:, 15:21], x[:, 6:9], lp[:, 21:27], x[:, 9:12],
                      lp[:, 27:], x[:, 12:]], 1)


def assemble_pose(face, body, stand=False):
    """scripts/demo.py:182-229 for one sample: face [Ff,103], body [Fb,129] -> [Ff,265]."""
    jaw, exp = face[:, :3], face[:, 3:]
    if body.shape[0] < face.shape[0]:
        body = torch.cat([body, body[-1:].repeat(face.shape[0] - body.shape[0], 1)], 0)
    else:
        body = body[: face.shape[0]]
    return part2full(torch.cat([jaw, body, exp], -1), stand)


def rot6d_to_axis_angle(d6):
    """matrix_to_axis_angle(rotation_6d_to_matrix(d6)) restated: data_utils/rotation_conversion.py:512-533
    (Gram-Schmidt), :98-118 (matrix_to_quaternion with _sqrt_positive_part / _copysign), :481-507
    (quaternion_to_axis_angle, small-angle branch at 1e-6).  d6 [...,6] -> [...,3]."""
    a1, a2 = d6[..., :3], d6[..., 3:]
    b1 = F.normalize(a1, dim=-1)
    b2 = a2 - (b1 * a2).sum(-1, keepdim=True) * b1
    b2 = F.normalize(b2, dim=-1)
    b3 = torch.cross(b1, b2, dim=-1)
    m = torch.stack((b1, b2, b3), dim=-2)
    m00, m11, m22 = m[..., 0, 0], m[..., 1, 1], m[..., 2, 2]

    def sqp(x):
        return torch.where(x > 0, torch.sqrt(x.clamp_min(0)), torch.zeros_like(x))

    def cps(a, b):
        return torch.where((a < 0) != (b < 0), -a, a)

    q0 = 0.5 * sqp(1 + m00 + m11 + m22)
    q1 = cps(0.5 * sqp(1 + m00 - m11 - m22), m[..., 2, 1] - m[..., 1, 2])
    q2 = cps(0.5 * sqp(1 - m00 + m11 - m22), m[..., 0, 2] - m[..., 2, 0])
    q3 = cps(0.5 * sqp(1 - m00 - m11 + m22), m[..., 1, 0] - m[..., 0, 1])
    v = torch.stack((q1, q2, q3), -1)
    norms = torch.norm(v, p=2, dim=-1, keepdim=True)
    half = torch.atan2(norms, q0.unsqueeze(-1))
    ang = 2 * half
    small = ang.abs() < 1e-6
    s = torch.where(small, 0.5 - ang * ang / 48, torch.sin(half) / torch.where(small, torch.ones_like(ang), ang))
    return v / s


def latent_rows(n_mfcc_frames):
    """T for M MFCC frames: two k4/s2/p1 downsamples (vqvae_modules.py:105-106)."""
    m = (n_mfcc_frames + 2 - 4) // 2 + 1
    return (m + 2 - 4) // 2 + 1
