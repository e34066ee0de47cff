"""BUILD CONTAINER ONLY (imports the unmodified reference from /root/reference): CPU time of the reference's own modules vs the oracle
port that bench.py --impl reference times on the GPU box, same synthetic checkpoints and inputs, same thread count."""
import sys, os, time, types, numpy as np, torch
sys.path.insert(0,'/root/repo/tests/golden'); sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/oracle')
import ref_import
nets = ref_import.import_reference()
import nets.smplx_body_pixel as ref_bp, nets.smplx_face as ref_face
from trainer.config import load_JsonConfig
from talkshow_b200 import synth
import talkshow_oracle as O
torch.set_grad_enabled(False); torch.set_flush_denormal(True)
nthr = int(sys.argv[1]) if len(sys.argv) > 1 else 8
torch.set_num_threads(nthr)
import tempfile
tmp = tempfile.mkdtemp()
vq = synth.body_vq_checkpoint(0); torch.save({"generator": vq}, tmp + "/vq.pth")
bp = synth.body_pixel_checkpoint(0); fc = synth.face_checkpoint(0)
cfg = load_JsonConfig("config/body_pixel.json"); cfg.Model.vq_path = tmp + "/vq.pth"
a = types.SimpleNamespace(gpu="cpu", infer=True)
g = ref_bp.TrainWrapper(a, cfg); g.load_state_dict(bp)
gf = ref_face.TrainWrapper(a, load_JsonConfig("config/face.json")); gf.load_state_dict(fc)
B, sec = 2, 10
wave = synth.synth_wave(B, 16000 * sec, seed=4321); mfcc = synth.synth_mfcc(B, 300, seed=9); label = torch.arange(B) % 4
for m in (g.generator, g.audioencoder, g.g_body, g.g_hand, gf.generator): m.eval()
def ref_step():
    face = gf.generator(wave[:, None, :], None, torch.zeros(B, 4), time_steps=300)[0]
    audio = g.audioencoder(mfcc).unsqueeze(-1).repeat(1, 1, 1, 2)
    lat = g.generator.generate(label, shape=[75, 2], batch_size=B, aud_feat=audio)
    body, _ = g.g_body.decode(b=B, w=75, latents=lat[..., 0]); hand, _ = g.g_hand.decode(b=B, w=75, latents=lat[..., 1])
    return face, body, hand
def port_step():
    face = O.face_forward(fc["generator"], wave, torch.zeros(B, 4), 300)
    _, body = O.body_generate(bp, vq, mfcc, label, noise=None, window=None)
    return face, body
for name, fn in (("reference modules", ref_step), ("oracle port", port_step), ("reference modules", ref_step), ("oracle port", port_step)):
    torch.manual_seed(1); t0 = time.perf_counter(); fn(); dt = time.perf_counter() - t0
    print("%s: %.2f s for %d clips x %d s (%d threads) = %.1f frames/s" % (name, dt, B, sec, nthr, B * 300 / dt), flush=True)
