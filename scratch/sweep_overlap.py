"""whole-body step time vs (clips per GPU, sampler CTAs of the overlapped small-batch path)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from talkshow_b200 import synth
from talkshow_b200.engine import Engine
from talkshow_b200.pipeline import WholeBody
from talkshow_b200.nets.base import draw_sampler_noise
torch.set_grad_enabled(False)
ck = dict(pixel=synth.body_pixel_checkpoint(0), vq=synth.body_vq_checkpoint(0), face=synth.face_checkpoint(0))
for B in (8, 12, 16, 32):
    mfcc = synth.synth_mfcc(B, 300, seed=1).cuda(); wave = synth.synth_wave(B, 160000, seed=2).cuda(); label = (torch.arange(B) % 4).cuda()
    for ctas in (0, 48, 56, 96, 100, 120):
        e = Engine(0)
        w = WholeBody(e, overlap_batch=(64 if ctas else 0), overlap_ctas=ctas or 100)
        try:
            w.load(ck["pixel"], ck["vq"], ck["face"])
        except Exception as ex:
            print("B=%d ctas %d: load failed" % (B, ctas)); e.close(); continue
        for _ in range(2): w.generate(mfcc, wave, label)
        torch.cuda.synchronize()
        ts = []
        for _ in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); w.generate(mfcc, wave, label); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
        print("B=%d sampler CTAs %s: %.2f ms" % (B, ctas or "148 (sequential)", min(ts)), flush=True)
        w.close(); e.close()
