"""cluster-resident executor (ts_set_pixelcnn_mode(2)) vs the grid-wide persistent kernel (mode 0): logits, codes, time."""
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
e2 = Engine(0); e2.load_pixelcnn(ck["generator"]); e2.set_pixelcnn_mode(2)
print("plans loaded"); sys.stdout.flush()
# teacher-forced logits
B, T = 3, 6
g = torch.Generator().manual_seed(3)
codes = torch.randint(0, 2048, (B, T, 2), generator=g)
label = torch.tensor([0, 3, 1])
a = e0.audio_encode(synth.synth_mfcc(B, 4 * T, seed=5))
l0 = e0.pixelcnn_logits(a, label, codes); torch.cuda.synchronize()
l2 = e2.pixelcnn_logits(a, label, codes); torch.cuda.synchronize()
print("teacher-forced logits max-abs diff v3 vs v1: %.3e (|l| max %.2f)" % ((l0 - l2).abs().max().item(), l0.abs().max().item())); sys.stdout.flush()
shapes = [(3, 6), (5, 20), (64, 75), (8, 75), (12, 75), (1, 30), (33, 10)]
if len(sys.argv) > 1:
    shapes = [tuple(int(v) for v in s.split("x")) for s in sys.argv[1:]]
for B, T in shapes:
    mfcc = synth.synth_mfcc(B, 4 * T).cuda(); label = (torch.arange(B) % 4).cuda()
    gg = torch.Generator(device='cuda').manual_seed(5)
    noise = torch.empty(2 * T, B, 2048, device='cuda').exponential_(1, generator=gg)
    a = e0.audio_encode(mfcc)
    c0, l0 = e0.pixelcnn_generate(a, label, noise, want_logits=True)
    torch.cuda.synchronize()
    c2, l2 = e2.pixelcnn_generate(a, label, noise, want_logits=True)
    torch.cuda.synchronize()
    print("B=%d T=%d codes equal: %s (mismatches %d), logits max-abs diff %.3e" % (B, T, torch.equal(c0, c2), (c0 != c2).sum().item(), (l0 - l2).abs().max().item()))
    sys.stdout.flush()
    for name, e in (("v1 grid-wide", e0), ("v3 cluster", e2)):
        best = 1e9
        for it in range(3):
            t1 = ev(); e.pixelcnn_generate(a, label, noise); t2 = ev(); torch.cuda.synchronize()
            best = min(best, t1.elapsed_time(t2))
        print("  %s: %.3f ms (%.1f us/row)" % (name, best, best * 1000 / T)); sys.stdout.flush()
