"""per-stage time stamps of the cluster-resident PixelCNN executor (cluster 0, rank 0, thread 0), one latent row."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import torch
from talkshow_b200 import synth
from talkshow_b200.engine import Engine
torch.set_grad_enabled(False)
ck = synth.body_pixel_checkpoint(0)
e = Engine(0); e.load_pixelcnn(ck["generator"]); e.load_audioenc(ck["audioencoder"]); e.set_pixelcnn_mode(2)
T, row = 75, 40
for B in [int(v) for v in (sys.argv[1:] or ["8", "64"])]:
    mfcc = synth.synth_mfcc(B, 4 * T).cuda(); label = (torch.arange(B) % 4).cuda()
    noise = torch.empty(2 * T, B, 2048, device='cuda').exponential_(1)
    a = e.audio_encode(mfcc)
    e.pixelcnn_generate(a, label, noise); torch.cuda.synchronize()
    e.pixelcnn_trace(row)
    e.pixelcnn_generate(a, label, noise); torch.cuda.synchronize()
    tr = e.pixelcnn_trace_read().reshape(-1)
    ns = 52
    t = tr[: ns * 8].view(ns, 8).double()
    e.pixelcnn_trace(-1)
    print("B=%d row %d: stage  wait_sync  stage_x  work  arrive  (weights_wait)  [us]" % (B, row))
    tot = dict(wait=0., x=0., work=0., arr=0., ww=0.)
    for s in range(ns):
        t0, t1, t2, t3, ww, t5 = [t[s, i].item() for i in range(6)]
        samp = t5 < t1
        x = 0. if samp else (t5 - t1) / 1e3
        work = (t2 - t1) / 1e3 - x
        print("  %2d  %6.2f %6.2f %6.2f %6.2f  (%5.2f)%s" % (s, (t1 - t0) / 1e3, x, work, (t3 - t2) / 1e3, ww / 1e3, "  sample" if samp else ""))
        tot["wait"] += (t1 - t0) / 1e3; tot["x"] += x; tot["work"] += work; tot["arr"] += (t3 - t2) / 1e3; tot["ww"] += ww / 1e3
    print("  row total %.1f us: wait_sync %.1f, stage_x %.1f, work %.1f (of which waiting for weights %.1f), arrive %.1f"
          % ((t[ns - 1, 3] - t[0, 0]).item() / 1e3, tot["wait"], tot["x"], tot["work"], tot["ww"], tot["arr"]))
