import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from talkshow_b200 import _lib
from talkshow_b200.engine import Engine
e = Engine(0)
def run(mode, A, W, bias, act=0):
    M,K = A.shape; N = W.shape[0]
    out = torch.empty(M, N, device='cuda')
    rc = e.L.ts_debug_gemm(e.h, mode, _lib.ptr(A), _lib.ptr(W), _lib.ptr(bias), _lib.ptr(out), M, N, K, act, _lib.stream_ptr(e.device))
    assert rc == 0, e.L.ts_last_error(e.h).decode()
    torch.cuda.synchronize(); return out
torch.manual_seed(0)
M,N,K = 2048,512,3072
A = torch.randn(M,K,device='cuda'); W = torch.randn(N,K,device='cuda')/K**0.5; b = torch.zeros(N,device='cuda')
ref = A.double() @ W.double().t()
for mode in (0,1):
    out = run(mode,A,W,b); d = (out.double()-ref)
    print("full K=%d mode %d: max %.3e mean signed %.3e rms %.3e" % (K, mode, d.abs().max().item(), d.mean().item(), d.pow(2).mean().sqrt().item()))
for ch in (1024, 512, 256, 128):
    acc = torch.zeros(M,N,device='cuda')
    for k0 in range(0,K,ch):
        acc += run(1, A[:,k0:k0+ch].contiguous(), W[:,k0:k0+ch].contiguous(), b)
    d = (acc.double()-ref)
    print("chunk %4d mode 1: max %.3e mean signed %.3e rms %.3e" % (ch, d.abs().max().item(), d.mean().item(), d.pow(2).mean().sqrt().item()))
# positive-only data to expose truncation bias
A2 = A.abs(); W2 = W.abs(); ref2 = A2.double() @ W2.double().t()
for mode in (0,1):
    d = run(mode,A2,W2,b).double()-ref2
    print("positive data mode %d: max %.3e mean signed %.3e (ref mean %.1f)" % (mode, d.abs().max().item(), d.mean().item(), ref2.mean().item()))
