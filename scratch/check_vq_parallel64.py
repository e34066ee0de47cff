"""whole-body step at 64 / 32 clips with the two VQ decoders of the body engine side by side (ts_set_vq_parallel) or not."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
LOG = open(os.path.join(ROOT, "gpurun_out", "check_vq_parallel64.log"), "a")
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
eng = Engine(0); wb = WholeBody(eng); wb.load(ck["pixel"], ck["vq"], ck["face"])
for B in (64, 32):
    mfcc = synth.synth_mfcc(B, 300, seed=1).cuda(); wave = synth.synth_wave(B, 160000, seed=2).cuda(); label = (torch.arange(B) % 4).cuda()
    noise = torch.empty(150, B, 2048, device="cuda").exponential_(1, generator=torch.Generator(device="cuda").manual_seed(7))
    first = None
    for par in (0, 64, 0, 64):
        wb.e2.set_vq_parallel(par); eng.set_vq_parallel(par)
        o = wb.generate(mfcc, wave, label, noise=noise)
        torch.cuda.synchronize(); ts = []
        for _ in range(4):
            a = ev(); o = wb.generate(mfcc, wave, label, noise=noise); b = ev(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
        first = o.clone() if first is None else first
        say("B=%d whole-body step, vq_parallel=%d: %.3f ms (median %.3f); equal to the first result: %s" % (B, par, min(ts), sorted(ts)[2], torch.equal(o, first)))
wb.close(); eng.close()
