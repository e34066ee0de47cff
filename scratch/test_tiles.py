"""grid-wide executor: batch tiles 8/16/32 (auto) vs the 64-sample tile (TS_PIX_TILE=64): logits, codes, time; mode 1 cross-check."""
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
shapes = [(3, 6), (8, 75), (12, 75), (16, 75), (20, 20), (32, 75), (1, 30)]
for B, T in shapes:
    mfcc = synth.synth_mfcc(B, 4 * T).cuda(); label = (torch.arange(B) % 4).cuda()
    gg = torch.Generator(device='cuda').manual_seed(5)
    noise = torch.empty(2 * T, B, 2048, device='cuda').exponential_(1, generator=gg)
    a = e0.audio_encode(mfcc)
    os.environ["TS_PIX_TILE"] = "64"
    c0, l0 = e0.pixelcnn_generate(a, label, noise, want_logits=True); torch.cuda.synchronize()
    best64 = 1e9
    for it in range(3):
        t1 = ev(); e0.pixelcnn_generate(a, label, noise); t2 = ev(); torch.cuda.synchronize(); best64 = min(best64, t1.elapsed_time(t2))
    del os.environ["TS_PIX_TILE"]
    c2, l2 = e0.pixelcnn_generate(a, label, noise, want_logits=True); torch.cuda.synchronize()
    best = 1e9
    for it in range(3):
        t1 = ev(); e0.pixelcnn_generate(a, label, noise); t2 = ev(); torch.cuda.synchronize(); best = min(best, t1.elapsed_time(t2))
    msg = ""
    if T <= 20:
        e0.set_pixelcnn_mode(1)
        c1, l1 = e0.pixelcnn_generate(a, label, noise, want_logits=True); torch.cuda.synchronize()
        e0.set_pixelcnn_mode(0)
        msg = " | per-stage debug mode equal: %s" % torch.equal(c1, c2)
    print("B=%d T=%d codes equal: %s, logits max-abs diff %.3e | tile 64: %.3f ms (%.1f us/row), auto tile: %.3f ms (%.1f us/row)%s"
          % (B, T, torch.equal(c0, c2), (l0 - l2).abs().max().item(), best64, best64 * 1e3 / T, best, best * 1e3 / T, msg))
    sys.stdout.flush()
