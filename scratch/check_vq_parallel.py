"""ts_set_vq_parallel: the two VQ decoders of the fused body path side by side on two streams — same result, time per batch size."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
LOG = open(os.path.join(ROOT, "gpurun_out", "check_vq_parallel.log"), "a")
def say(m):
    print(m, flush=True); LOG.write(m + "\n"); LOG.flush()
import torch
from talkshow_b200 import synth
from talkshow_b200.engine import Engine
from talkshow_b200.pipeline import WholeBody
torch.set_grad_enabled(False)
ck = dict(pixel=synth.body_pixel_checkpoint(0), vq=synth.body_vq_checkpoint(0), face=synth.face_checkpoint(0))
def ev():
    x = torch.cuda.Event(enable_timing=True); x.record(); return x
e = Engine(0)
e.load_pixelcnn(ck["pixel"]["generator"]); e.load_audioenc(ck["pixel"]["audioencoder"]); e.load_vq(0, ck["vq"]["g_body"]); e.load_vq(1, ck["vq"]["g_hand"])
def body(B, M, par):
    e.set_vq_parallel(64 if par else 0)
    mfcc = synth.synth_mfcc(B, M, seed=1).cuda(); label = (torch.arange(B) % 4).cuda(); T = e.latent_rows(M)
    noise = torch.empty(2 * T, B, 2048, device="cuda").exponential_(1, generator=torch.Generator(device="cuda").manual_seed(3))
    best = 1e9
    for it in range(5):
        a = ev(); c, p = e.body_generate(mfcc, label, noise); b = ev(); torch.cuda.synchronize()
        if it: best = min(best, a.elapsed_time(b))
    return c.clone(), p.clone(), best
for B, M in ((1, 120), (8, 300), (12, 300), (16, 300), (32, 300), (64, 300)):
    c0, p0, t0 = body(B, M, False)
    c1, p1, t1 = body(B, M, True)
    c2, p2, t2 = body(B, M, False)
    say("B=%d M=%d body_generate: sequential decoders %.3f / %.3f ms, side by side %.3f ms; codes equal %s, poses bit-identical %s"
        % (B, M, t0, t2, t1, torch.equal(c0, c1), torch.equal(p0, p1) and torch.equal(p0, p2)))
e.close()
# whole-body step of the 8-clip shard (two engines, two streams): decoders side by side on the body engine
eng = Engine(0); wb = WholeBody(eng); wb.load(ck["pixel"], ck["vq"], ck["face"])
for B in (8, 12):
    mfcc = synth.synth_mfcc(B, 300, seed=1).cuda(); wave = synth.synth_wave(B, 160000, seed=2).cuda(); label = (torch.arange(B) % 4).cuda()
    noise = torch.empty(150, B, 2048, device="cuda").exponential_(1, generator=torch.Generator(device="cuda").manual_seed(7))
    first = None
    for par in (0, 16, 0, 16):
        wb.e2.set_vq_parallel(par); eng.set_vq_parallel(par)
        for _ in range(2): o = wb.generate(mfcc, wave, label, noise=noise)
        torch.cuda.synchronize(); ts = []
        for _ in range(5):
            a = ev(); o = wb.generate(mfcc, wave, label, noise=noise); b = ev(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
        first = o.clone() if first is None else first
        say("B=%d whole-body step, vq_parallel=%d: %.3f ms (median %.3f); equal to the first result: %s" % (B, par, min(ts), sorted(ts)[2], torch.equal(o, first)))
wb.close(); eng.close()
