import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from talkshow_b200 import synth
from talkshow_b200.engine import Engine
torch.set_grad_enabled(False)
e = Engine(0); e.load_face(synth.face_state(4))
for B in (8, 64):
    wave = synth.synth_wave(B, 160000).cuda(); ids = torch.zeros(B,4).cuda()
    for it in range(2):
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record(); out = e.face_forward(wave, ids, 300); t1.record(); torch.cuda.synchronize()
    print("face B=%d: %.2f ms -> %.1f TFLOP/s (106 GFLOP/clip)" % (B, t0.elapsed_time(t1), B*106e9/(t0.elapsed_time(t1)*1e-3)/1e12))
