"""face forward time and error for the tensor-core GEMM variants: ts_set_tensor_cores 1 (pre-split, default) vs 6 (fp16-split planes, kind::f16)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from talkshow_b200 import synth
from talkshow_b200.engine import Engine
torch.set_grad_enabled(False)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
wave = synth.synth_wave(B, 160000, seed=1).cuda(); idz = torch.zeros(B, 4).cuda()
outs = {}
MODES = [int(m) for m in sys.argv[2].split(',')] if len(sys.argv) > 2 else [1, 6]
REPS = int(sys.argv[3]) if len(sys.argv) > 3 else 3
for mode in MODES:
    e = Engine(0); e.set_tensor_cores(mode); e.load_face(synth.face_checkpoint(0)["generator"])
    for _ in range(2): out = e.face_forward(wave, idz, 300)
    torch.cuda.synchronize()
    ts = []
    for _ in range(REPS):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); out = e.face_forward(wave, idz, 300); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    outs[mode] = out.clone()
    print("tensor-core mode %d: face B=%d %.2f ms (min of 3)" % (mode, B, min(ts))); sys.stdout.flush()
    e.close()
if len(MODES) > 1: print("max-abs diff between the variants: %.3e" % (outs[MODES[1]] - outs[MODES[0]]).abs().max().item())
