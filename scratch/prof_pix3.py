import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from talkshow_b200 import synth
from talkshow_b200.engine import Engine
torch.set_grad_enabled(False)
T, B = int(sys.argv[1]), int(sys.argv[2])
ck = synth.body_pixel_checkpoint(0)
e = Engine(0); e.load_pixelcnn(ck["generator"]); e.load_audioenc(ck["audioencoder"]); e.set_pixelcnn_mode(2)
mfcc = synth.synth_mfcc(B, 4 * T).cuda(); label = (torch.arange(B) % 4).cuda()
noise = torch.empty(2 * T, B, 2048, device='cuda').exponential_(1)
a = e.audio_encode(mfcc)
for _ in range(3):
    e.pixelcnn_generate(a, label, noise)
torch.cuda.synchronize()
