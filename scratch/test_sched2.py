"""EXPERIMENTAL schedule 2 (ts_set_pixelcnn_fusion(2): vert_to_horiz moved into the horizontal pass) vs the default persistent kernel: codes, logits, time."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from talkshow_b200 import synth
from talkshow_b200.engine import Engine
torch.set_grad_enabled(False)
ck = synth.body_pixel_checkpoint(0)
def ev():
    x = torch.cuda.Event(enable_timing=True); x.record(); return x
e0 = Engine(0); e0.load_pixelcnn(ck["generator"]); e0.load_audioenc(ck["audioencoder"])
e3 = Engine(0); e3.set_pixelcnn_fusion(2); e3.load_pixelcnn(ck["generator"])
print("schedule-2 plan loaded"); sys.stdout.flush()
for B, T in ((3, 6), (64, 75)):
    mfcc = synth.synth_mfcc(B, 4 * T).cuda(); label = (torch.arange(B) % 4).cuda()
    g = torch.Generator(device='cuda').manual_seed(5)
    noise = torch.empty(2 * T, B, 2048, device='cuda').exponential_(1, generator=g)
    a = e0.audio_encode(mfcc)
    c0, l0 = e0.pixelcnn_generate(a, label, noise, want_logits=True)
    torch.cuda.synchronize()
    c3, l3 = e3.pixelcnn_generate(a, label, noise, want_logits=True)
    torch.cuda.synchronize()
    print("B=%d T=%d codes equal: %s (mismatches %d), logits max-abs diff %.3e" % (B, T, torch.equal(c0, c3), (c0 != c3).sum().item(), (l0 - l3).abs().max().item()))
    sys.stdout.flush()
    for name, e in (("default", e0), ("sched2", e3)):
        best = 1e9
        for it in range(3):
            t1 = ev(); e.pixelcnn_generate(a, label, noise); t2 = ev(); torch.cuda.synchronize()
            best = min(best, t1.elapsed_time(t2))
        print("  %s: %.3f ms (%.1f us/row)" % (name, best, best * 1000 / T)); sys.stdout.flush()
