"""face forward time and error for the tensor-core GEMM variants: ts_set_tensor_cores 1 (pre-split, default) vs 5 (plain operands, split on chip)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from talkshow_b200 import synth
from talkshow_b200.engine import Engine
torch.set_grad_enabled(False)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
wave = synth.synth_wave(B, 160000, seed=1).cuda(); idz = torch.zeros(B, 4).cuda()
outs = {}
for mode in (1, 5):
    e = Engine(0); e.set_tensor_cores(mode); e.load_face(synth.face_checkpoint(0)["generator"])
    for _ in range(2): out = e.face_forward(wave, idz, 300)
    torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); out = e.face_forward(wave, idz, 300); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    outs[mode] = out.clone()
    print("tensor-core mode %d: face B=%d %.2f ms (min of 3)" % (mode, B, min(ts))); sys.stdout.flush()
    e.close()
print("max-abs diff between the variants: %.3e" % (outs[5] - outs[1]).abs().max().item())
