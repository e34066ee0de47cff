import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from talkshow_b200 import synth
from talkshow_b200.engine import Engine
torch.set_grad_enabled(False)
ck = synth.body_pixel_checkpoint(0)
def ev():
    x = torch.cuda.Event(enable_timing=True); x.record(); return x
res = {}
for fused in [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else '1,0').split(',')]:
    e = Engine(0); e.set_pixelcnn_fusion(fused)
    e.load_pixelcnn(ck["generator"]); e.load_audioenc(ck["audioencoder"])
    for B in (64, 8):
        T = 75
        mfcc = synth.synth_mfcc(B, 300).cuda(); label = (torch.arange(B) % 4).cuda()
        g = torch.Generator(device='cuda').manual_seed(5)
        noise = torch.empty(2*T, B, 2048, device='cuda').exponential_(1, generator=g)
        a = e.audio_encode(mfcc)
        best = 1e9
        for it in range(4):
            t1 = ev(); codes = e.pixelcnn_generate(a, label, noise); t2 = ev(); torch.cuda.synchronize()
            best = min(best, t1.elapsed_time(t2))
        res[(fused, B)] = codes.cpu()
        print("pipe=%s fused=%d B=%d pixelcnn %.3f ms (%.1f us/row)" % (os.environ.get("TS_PIX_PIPE","default"), fused, B, best, best*1000/T)); sys.stdout.flush()
    torch.cuda.synchronize(); e.close()
for B in ((64, 8) if len(res) > 2 else ()):
    print("B=%d fused == plain codes:" % B, torch.equal(res[(1,B)], res[(0,B)]), "mismatches", (res[(1,B)] != res[(0,B)]).sum().item())
