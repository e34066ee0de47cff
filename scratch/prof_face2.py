import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from talkshow_b200 import synth
from talkshow_b200.engine import Engine
torch.set_grad_enabled(False)
B = int(sys.argv[1]); mode = int(sys.argv[2])
e = Engine(0); e.set_tensor_cores(mode); e.load_face(synth.face_state(4))
wave = synth.synth_wave(B, 160000).cuda(); ids = torch.zeros(B,4).cuda()
for it in range(2): out = e.face_forward(wave, ids, 300)
torch.cuda.synchronize()
