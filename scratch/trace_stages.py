"""Per-stage skew / cost table of the persistent PixelCNN kernel from the debug trace (ts_pixelcnn_trace).
Usage (GPU): python scratch/trace_stages.py [row]  ->  for each stage of that latent row: task mix, time from the
first CTA leaving the grid barrier to the last CTA arriving, and the spread of per-CTA task times."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from talkshow_b200 import synth, _lib
from talkshow_b200.engine import Engine
torch.set_grad_enabled(False)
row = int(sys.argv[1]) if len(sys.argv) > 1 else 40
ck = synth.body_pixel_checkpoint(0)
e = Engine(0); e.load_pixelcnn(ck["generator"]); e.load_audioenc(ck["audioencoder"])
B, T = (int(sys.argv[2]) if len(sys.argv) > 2 else 64), 75
mfcc = synth.synth_mfcc(B, 4 * T).cuda(); label = (torch.arange(B) % 4).cuda()
noise = torch.empty(2 * T, B, 2048, device='cuda').exponential_(1)
a = e.audio_encode(mfcc)
e.pixelcnn_generate(a, label, noise)
e.pixelcnn_trace(row)
e.pixelcnn_generate(a, label, noise)
tr = e.pixelcnn_trace_read().double()            # [stages, ctas, 4] ns
table, _ = _lib.plan_to_numpy(e.h)
ns, nc = tr.shape[0], tr.shape[1]
tab = torch.tensor(table[32:].reshape(ns, nc, 8))
names = {0: "idle", 1: "VERT0", 2: "VERT", 3: "V2H", 4: "FUSEV", 5: "HGATE", 6: "HRES", 7: "FUSEH", 8: "OUT1", 9: "OUT2", 10: "SAMPLE",
         11: "HRESF", 12: "HGATE2", 13: "OUT1F"}
t0 = tr[0, :, 1].min()
print("stage  jobs                                   span_us  task_us(min/med/max)  wait_us(med)")
tot = 0.0
for s in range(ns):
    leave, done, arr = tr[s, :, 1], tr[s, :, 2], tr[s, :, 3]
    span = (arr.max() - leave.min()) / 1e3
    task = (done - leave) / 1e3
    nxt = tr[s + 1, :, 1] if s + 1 < ns else None
    wait = ((nxt - arr) / 1e3).median().item() if nxt is not None else float('nan')
    jobs = {}
    for c in range(nc):
        ep, K, nr = int(tab[s, c, 0]), int(tab[s, c, 6]), int(tab[s, c, 4])
        if ep: jobs.setdefault((names.get(ep, str(ep)), K), []).append(nr)
    desc = " ".join("%s/K%d:%dx%d-%d" % (n, K, len(v), min(v), max(v)) for (n, K), v in jobs.items())
    print("%3d  %-40s %7.2f   %5.2f/%5.2f/%5.2f   %6.2f" % (s, desc[:40], span, task.min(), task.median(), task.max(), wait))
    tot += span
print("sum of spans: %.1f us (row %d)" % (tot, row))
