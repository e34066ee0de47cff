"""tiny end-to-end run for compute-sanitizer (memcheck / racecheck): both PixelCNN executors, VQ decode, face, LBS."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from talkshow_b200 import synth, smplx_lbs
from talkshow_b200.engine import Engine
torch.set_grad_enabled(False)
what = sys.argv[1] if len(sys.argv) > 1 else "all"
e = Engine(0)
ck, vq = synth.body_pixel_checkpoint(0), synth.body_vq_checkpoint(0)
e.load_pixelcnn(ck["generator"]); e.load_audioenc(ck["audioencoder"]); e.load_vq(0, vq["g_body"]); e.load_vq(1, vq["g_hand"])
B, M = 3, 12
mfcc = synth.synth_mfcc(B, M, seed=3); label = torch.tensor([0, 2, 1])
T = e.latent_rows(M)
noise = torch.empty(2 * T, B, 2048).exponential_(1, generator=torch.Generator().manual_seed(1))
for mode in ((0, 2) if what in ("all", "pix") else ()):
    e.set_pixelcnn_mode(mode)
    codes, poses = e.body_generate(mfcc, label, noise)
    torch.cuda.synchronize()
    print("mode", mode, codes.flatten()[:6].tolist(), float(poses.abs().max()))
e.set_pixelcnn_mode(0)
if what in ("all", "face"):
    e.load_face(synth.face_checkpoint(0)["generator"])
    out = e.face_forward(synth.synth_wave(1, 16000, seed=2), torch.zeros(1, 4), 30)
    torch.cuda.synchronize()
    print("face", float(out.abs().max()))
if what in ("all", "lbs"):
    sm = smplx_lbs.SmplxModel(smplx_lbs.synthetic_model(V=1200, seed=1, nfaces=2000), engine=e)
    v, j = sm.forward_pose265(torch.rand(5, 265) * 0.3)
    torch.cuda.synchronize()
    print("lbs", float(v.abs().max()), tuple(j.shape))
e.close()
print("done")

if what in ("all", "wb"):
    from talkshow_b200.pipeline import WholeBody
    e3 = Engine(0)
    w = WholeBody(e3)                       # overlapped: sampler on a 96-CTA plan (second engine, side stream) next to the face path
    w.load(ck, vq, synth.face_checkpoint(0))
    out = w.generate(synth.synth_mfcc(2, 12, seed=5).cuda(), synth.synth_wave(2, 16000, seed=6).cuda(), torch.tensor([1, 3]).cuda())
    torch.cuda.synchronize()
    print("whole body (overlapped)", tuple(out.shape), float(out.abs().max()))
    w.close(); e3.close()
