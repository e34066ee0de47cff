import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from talkshow_b200 import synth
from talkshow_b200.engine import Engine
torch.set_grad_enabled(False)
e = Engine(0); e.load_face(synth.face_state(4))
B=64; wave = synth.synth_wave(B, 160000).cuda(); ids = torch.zeros(B,4).cuda()
for rep in range(3):
    for mode in (1, 4, 2, 0):
        e.set_tensor_cores(mode)
        e.face_forward(wave, ids, 300)
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record(); e.face_forward(wave, ids, 300); t1.record(); torch.cuda.synchronize()
        print("rep %d mode %d (%s): %.2f ms" % (rep, mode, {0:'ffma',1:'tc pair 256x256',2:'tc 128x256 multicast',4:'tc 128x256'}[mode], t0.elapsed_time(t1)), flush=True)
