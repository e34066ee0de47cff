import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from talkshow_b200 import synth
from talkshow_b200.engine import Engine
torch.set_grad_enabled(False)
ck, vq = synth.body_pixel_checkpoint(0), synth.body_vq_checkpoint(0)
e = Engine(0)
e.load_pixelcnn(ck["generator"]); e.load_audioenc(ck["audioencoder"]); e.load_vq(0, vq["g_body"]); e.load_vq(1, vq["g_hand"])
def ev(): 
    x = torch.cuda.Event(enable_timing=True); x.record(); return x
for B in (64, 8, 1):
    M = 300; T = 75
    mfcc = synth.synth_mfcc(B, M).cuda()
    label = (torch.arange(B) % 4).cuda()
    noise = torch.empty(2*T, B, 2048, device='cuda').exponential_(1)
    aud = e.audio_encode(mfcc)
    for it in range(2):
        t0 = ev(); a = e.audio_encode(mfcc); t1 = ev()
        codes = e.pixelcnn_generate(a, label, noise); t2 = ev()
        b = e.vq_decode(0, codes[..., 0].contiguous()); h = e.vq_decode(1, codes[..., 1].contiguous()); t3 = ev()
        c2, poses = e.body_generate(mfcc, label, noise); t4 = ev()
        torch.cuda.synchronize()
    print("B=%d audio %.3f ms  pixelcnn %.3f ms (%.1f us/row)  decode %.3f ms  fused %.3f ms -> %.0f frames/s" % (
        B, t0.elapsed_time(t1), t1.elapsed_time(t2), t1.elapsed_time(t2)*1000/T, t2.elapsed_time(t3), t3.elapsed_time(t4), B*300/(t3.elapsed_time(t4)/1e3)))
import sys
e.set_pixelcnn_mode(2)
for B in (64, 8):
    mfcc = synth.synth_mfcc(B, 300).cuda(); label = (torch.arange(B) % 4).cuda(); noise = torch.empty(150, B, 2048, device='cuda').exponential_(1)
    a = e.audio_encode(mfcc)
    for it in range(2):
        t1 = ev(); codes = e.pixelcnn_generate(a, label, noise); t2 = ev(); torch.cuda.synchronize()
    print('v2 cluster mode B=%d: %.3f ms (%.1f us/row)' % (B, t1.elapsed_time(t2), t1.elapsed_time(t2)*1000/75)); sys.stdout.flush()
e.set_pixelcnn_mode(1)
B=64; mfcc = synth.synth_mfcc(B, 300).cuda(); label = (torch.arange(B) % 4).cuda(); noise = torch.empty(150, B, 2048, device='cuda').exponential_(1)
a = e.audio_encode(mfcc)
t1 = ev(); codes = e.pixelcnn_generate(a, label, noise); t2 = ev(); torch.cuda.synchronize()
print("stage-launch mode B=64: %.3f ms" % t1.elapsed_time(t2))
