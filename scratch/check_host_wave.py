"""Host-buffer step with every input copied before generate() vs the public generate_host(); body-path breakdown at 8 / 1 clips.
(profiles/r02c_host_copy_order_and_body_breakdown.log was taken with a pipeline variant that enqueued the waveform copy behind the body
path's launch inside generate(): no measurable gain, not kept — with the committed pipeline both orders are the same code path.)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
LOG = open(os.path.join(ROOT, "gpurun_out", "check_host_wave.log"), "a")
def say(m):
    print(m, flush=True); LOG.write(m + "\n"); LOG.flush()
import torch
from talkshow_b200 import synth
from talkshow_b200.engine import Engine
from talkshow_b200.pipeline import WholeBody
torch.set_grad_enabled(False)
ck = dict(pixel=synth.body_pixel_checkpoint(0), vq=synth.body_vq_checkpoint(0), face=synth.face_checkpoint(0))
e = Engine(0); wb = WholeBody(e); wb.load(ck["pixel"], ck["vq"], ck["face"])
for B in (64, 8):
    mfcc_h = synth.synth_mfcc(B, 300, seed=1).pin_memory(); wave_h = synth.synth_wave(B, 160000, seed=2).pin_memory()
    label_h = (torch.arange(B) % 4).pin_memory()
    noise = torch.empty(150, B, 2048, device="cuda").exponential_(1, generator=torch.Generator(device="cuda").manual_seed(7))
    out_p = torch.empty(B, 300, 265).pin_memory()
    def early():       # every input copied before generate()
        p = wb.generate(mfcc_h.to("cuda", non_blocking=True), wave_h.to("cuda", non_blocking=True), label_h.to("cuda", non_blocking=True), noise=noise)
        out_p.copy_(p, non_blocking=True); torch.cuda.current_stream().synchronize(); return out_p.clone()
    def late():        # the public host-buffer call: waveform copy behind the body path's launch
        return wb.generate_host(mfcc_h, wave_h, label_h, out_host=out_p, noise=noise).clone()
    ref = wb.generate(mfcc_h.cuda(), wave_h.cuda(), label_h.cuda(), noise=noise).cpu()
    for name, fn in (("copies first", early), ("waveform copy behind the body launch", late), ("copies first", early), ("waveform copy behind the body launch", late)):
        for _ in range(2): o = fn()
        torch.cuda.synchronize(); ts = []
        for _ in range(5):
            t0 = time.perf_counter(); o = fn(); ts.append((time.perf_counter() - t0) * 1e3)
        say("B=%d %s: %.3f ms (median %.3f), equal to device-input result: %s" % (B, name, min(ts), sorted(ts)[2], torch.equal(o, ref)))
# body path of the 8-clip shard on the 96-CTA plan, alone: where the time outside the sampler goes
def ev():
    x = torch.cuda.Event(enable_timing=True); x.record(); return x
e2 = wb.e2
for B in (8, 1):
    M = 300 if B == 8 else 120
    mfcc = synth.synth_mfcc(B, M, seed=1).cuda(); label = (torch.arange(B) % 4).cuda(); T = e2.latent_rows(M)
    noise = torch.empty(2 * T, B, 2048, device="cuda").exponential_(1)
    for it in range(3):
        t0 = ev(); a = e2.audio_encode(mfcc); t1 = ev(); c = e2.pixelcnn_generate(a, label, noise); t2 = ev()
        pb = e2.vq_decode(0, c[:, :, 0].contiguous()); ph = e2.vq_decode(1, c[:, :, 1].contiguous()); t3 = ev()
        _, full = e2.body_generate(mfcc, label, noise, want_codes=False); t4 = ev(); torch.cuda.synchronize()
    say("B=%d x %d rows body path alone (96-CTA plan): audio encoder %.3f ms, pixelcnn_generate (3 GEMMs + setup + sampler) %.3f ms, two VQ decoders %.3f ms; "
        "fused ts_body_generate %.3f ms" % (B, T, t0.elapsed_time(t1), t1.elapsed_time(t2), t2.elapsed_time(t3), t3.elapsed_time(t4)))
wb.close(); e.close()
