"""Sampler time vs the per-row cost term of the CTA split (TS_PIX_ROWCOST), schedule and plan size; codes compared with the default plan."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
LOG = open(os.path.join(ROOT, "gpurun_out", "sweep_rowcost.log"), "a")
T0 = time.time()


def say(msg):
    line = "[%6.1f s] %s" % (time.time() - T0, msg)
    print(line, flush=True)
    LOG.write(line + "\n")
    LOG.flush()


import torch  # noqa: E402

from talkshow_b200 import synth  # noqa: E402
from talkshow_b200.engine import Engine  # noqa: E402

torch.set_grad_enabled(False)
ck = synth.body_pixel_checkpoint(0)
T = 75


def ev():
    x = torch.cuda.Event(enable_timing=True)
    x.record()
    return x


def inputs(B):
    mfcc = synth.synth_mfcc(B, 4 * T).cuda()
    label = (torch.arange(B) % 4).cuda()
    noise = torch.empty(2 * T, B, 2048, device="cuda").exponential_(1, generator=torch.Generator(device="cuda").manual_seed(5))
    return mfcc, label, noise


e0 = Engine(0)
e0.load_pixelcnn(ck["generator"])
e0.load_audioenc(ck["audioencoder"])
INP = {B: inputs(B) for B in (64, 8)}
AUD = {B: e0.audio_encode(INP[B][0]) for B in INP}
REF = {}


def measure(e, B):
    a, (_, label, noise) = AUD[B], INP[B]
    c = e.pixelcnn_generate(a, label, noise)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        t1 = ev()
        e.pixelcnn_generate(a, label, noise)
        t2 = ev()
        torch.cuda.synchronize()
        best = min(best, t1.elapsed_time(t2))
    return (c[0] if isinstance(c, tuple) else c), best


for B in INP:
    REF[B], t = measure(e0, B)
    say("default plan (sched 1, 148 CTAs, rowcost 0) B=%d: %.3f ms (%.1f us/row)" % (B, t, t * 1000 / T))

CONFIGS = [(1, 0), (2, 0), (1, 96), (2, 96)]
ALPHAS = [256, 512, 1024, 2048]
if len(sys.argv) > 1:
    ALPHAS = [float(x) for x in sys.argv[1].split(",")]
for sched, ctas in CONFIGS:
    for alpha in ([0] if (sched, ctas) != (1, 0) else []) + ALPHAS:
        os.environ["TS_PIX_ROWCOST"] = str(alpha)
        e = Engine(0)
        try:
            e.set_pixelcnn_fusion(sched)
            if ctas:
                e.set_pixelcnn_ctas(ctas)
            e.load_pixelcnn(ck["generator"])
            msg = []
            for B in ((64, 8) if sched == 1 else (64,)):
                c, t = measure(e, B)
                msg.append("B=%d %.3f ms (%.1f us/row) codes==default: %s" % (B, t, t * 1000 / T, torch.equal(c, REF[B])))
            say("sched %d, %s CTAs, rowcost %g: %s" % (sched, ctas or 148, alpha, "; ".join(msg)))
        except Exception as ex:      # noqa: BLE001
            say("sched %d, %s CTAs, rowcost %g FAILED: %s" % (sched, ctas or 148, alpha, ex))
        finally:
            torch.cuda.synchronize()
            e.close()
say("done")
e0.close()
