import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from talkshow_b200 import _lib, synth
from talkshow_b200.engine import Engine
e = Engine(0); e.set_tensor_cores(3)
def run(mode, A, W, bias, act=0):
    M,K = A.shape; N = W.shape[0]
    out = torch.empty(M, N, device='cuda')
    rc = e.L.ts_debug_gemm(e.h, mode, _lib.ptr(A), _lib.ptr(W), _lib.ptr(bias), _lib.ptr(out), M, N, K, act, _lib.stream_ptr(e.device))
    assert rc == 0, e.L.ts_last_error(e.h).decode()
    torch.cuda.synchronize(); return out
torch.manual_seed(0)
for (M,N,K) in [(256,256,32),(256,256,64),(300,200,96),(4096,512,1536),(1000,64,512),(777,3072,768)]:
    A = torch.randn(M,K,device='cuda'); W = torch.randn(N,K,device='cuda')/K**0.5; b = torch.randn(N,device='cuda')
    ref = (A.double() @ W.double().t() + b.double())
    out = run(1, A, W, b)
    print("pair M=%d N=%d K=%d max-abs err %.3e" % (M,N,K,(out.double()-ref).abs().max().item()), flush=True)
e.load_face(synth.face_state(4))
for B in (8, 64):
    wave = synth.synth_wave(B, 160000).cuda(); ids = torch.zeros(B,4).cuda()
    for it in range(2):
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record(); out = e.face_forward(wave, ids, 300); t1.record(); torch.cuda.synchronize()
    print("pair face B=%d: %.2f ms" % (B, t0.elapsed_time(t1)), flush=True)
