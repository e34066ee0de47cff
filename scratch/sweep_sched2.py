"""Overlapped whole-body step at 33..64 clips: sampler plan of the side engine = (fusion level, persistent CTAs).
Every configuration is compared bit-exactly with the sequential order (same kernels on the 148-CTA default plan) and timed;
lines go to gpurun_out/sweep_sched2.log as they are produced."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
LOG = open(os.path.join(ROOT, "gpurun_out", "sweep_sched2.log"), "a")
T0 = time.time()


def say(msg):
    line = "[%6.1f s] %s" % (time.time() - T0, msg)
    print(line, flush=True)
    LOG.write(line + "\n")
    LOG.flush()


import torch  # noqa: E402

from talkshow_b200 import synth  # noqa: E402
from talkshow_b200.engine import Engine  # noqa: E402
from talkshow_b200.pipeline import WholeBody  # noqa: E402

torch.set_grad_enabled(False)
say("torch imported, device %s" % torch.cuda.get_device_name(0))
ck = dict(pixel=synth.body_pixel_checkpoint(0), vq=synth.body_vq_checkpoint(0), face=synth.face_checkpoint(0))
e = Engine(0)
wb = WholeBody(e, overlap_batch=64, overlap_ctas=96)
wb.load(ck["pixel"], ck["vq"], ck["face"])
wb.pixelcnn_timing(True)
say("engines loaded")


def inputs(B):
    mfcc = synth.synth_mfcc(B, 300, seed=1).cuda()
    wave = synth.synth_wave(B, 160000, seed=2).cuda()
    label = (torch.arange(B) % 4).cuda()
    noise = torch.empty(150, B, 2048, device="cuda").exponential_(1, generator=torch.Generator(device="cuda").manual_seed(7))
    return mfcc, wave, label, noise


def run(inp, reps=4):
    for _ in range(2):
        out = wb.generate(inp[0], inp[1], inp[2], noise=inp[3])
    torch.cuda.synchronize()
    ts, ps = [], []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = wb.generate(inp[0], inp[1], inp[2], noise=inp[3])
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
        ps.append(wb.pixelcnn_last_ms())
    return out.clone(), min(ts), min(ps)


def side_engine(fusion, ctas):
    """body-path engine with a sampler plan for ``ctas`` persistent CTAs (what WholeBody.load builds as e2)."""
    x = Engine(0)
    x.set_pixelcnn_fusion(fusion)
    x.set_pixelcnn_ctas(ctas)
    x.load_pixelcnn(ck["pixel"]["generator"])
    x.load_audioenc(ck["pixel"]["audioencoder"])
    x.load_vq(0, ck["vq"]["g_body"])
    x.load_vq(1, ck["vq"]["g_hand"])
    x.pixelcnn_timing(True)
    return x


PLANS = [(2, 80), (2, 88), (2, 96)]
if len(sys.argv) > 1:
    PLANS = [tuple(int(x) for x in a.split(",")) for a in sys.argv[1:]]
inp = inputs(64)
ob, wb.overlap_batch = wb.overlap_batch, 0
ref, t_seq, p_seq = run(inp)
wb.overlap_batch = ob
say("B=64 sequential (148 CTAs): %.2f ms per step, sampler %.2f ms" % (t_seq, p_seq))
out, t, p = run(inp)
say("B=64 side plan (1, 96) [default]: %.2f ms per step, sampler in step %.2f ms, bit-identical to sequential: %s" % (t, p, torch.equal(out, ref)))
results = {}
default_side = wb.e2
for plan in PLANS:
    try:
        wb.e2 = side_engine(plan[0], plan[1])
        out, t, p = run(inp)
        same = torch.equal(out, ref)
        results[plan] = (t, same)
        say("B=64 side plan %s: %.2f ms per step, sampler in step %.2f ms, bit-identical to sequential: %s" % (plan, t, p, same))
        for B in (33, 48):
            i2 = inputs(B)
            wb.overlap_batch = 0
            r2, ts2, _ = run(i2, reps=2)
            wb.overlap_batch = ob
            o2, t2, p2 = run(i2, reps=2)
            say("B=%d side plan %s: %.2f ms per step (sequential %.2f), sampler in step %.2f ms, bit-identical: %s"
                % (B, plan, t2, ts2, p2, torch.equal(o2, r2)))
    except Exception as ex:      # noqa: BLE001
        say("side plan %s FAILED: %s" % (plan, ex))
    finally:
        torch.cuda.synchronize()
        if wb.e2 is not default_side:
            wb.e2.close()
        wb.e2 = default_side
say("done: " + ", ".join("%s %.2f ms %s" % (k, v[0], "ok" if v[1] else "MISMATCH") for k, v in results.items()))
wb.close()
e.close()
