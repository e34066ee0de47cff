import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ctypes as C
from talkshow_b200 import _lib
from talkshow_b200.engine import Engine
e = Engine(0)
e.set_tensor_cores(5)   # mode 2 of ts_debug_gemm = on-chip split; mode 1 always runs the pre-split kernel
def run(mode, A, W, bias, act=0):
    M,K = A.shape; N = W.shape[0]
    out = torch.empty(M, N, device='cuda')
    rc = e.L.ts_debug_gemm(e.h, mode, _lib.ptr(A), _lib.ptr(W), _lib.ptr(bias), _lib.ptr(out), M, N, K, act, _lib.stream_ptr(e.device))
    if rc: raise RuntimeError(e.L.ts_last_error(e.h).decode())
    torch.cuda.synchronize()
    return out
torch.manual_seed(0)
for (M,N,K) in [(128,128,64),(300,200,192),(4096,512,1536),(1000,64,512),(777,3072,768),(19200,768,3072)]:
    A = torch.randn(M,K,device='cuda'); W = torch.randn(N,K,device='cuda')/K**0.5; b = torch.randn(N,device='cuda')
    ref = (A.double() @ W.double().t() + b.double())
    for mode in (0,1,2,3):
        try:
            out = run(mode, A, W, b)
            err = (out.double()-ref).abs().max().item()
            print("M=%d N=%d K=%d mode=%d max-abs err %.3e (ref absmax %.2f)" % (M,N,K,mode,err,ref.abs().max().item()), flush=True)
        except Exception as ex:
            print("M=%d N=%d K=%d mode=%d FAILED: %s" % (M,N,K,mode,ex), flush=True)
# timing
M,N,K = 65536, 768, 3072
A = torch.randn(M,K,device='cuda'); W = torch.randn(N,K,device='cuda')/K**0.5; b = torch.zeros(N,device='cuda')
for mode in (0,1,2,3):
    out = run(mode, A, W, b)
    t0=torch.cuda.Event(enable_timing=True); t1=torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); t0.record(); out = run(mode, A, W, b); t1.record(); torch.cuda.synchronize()
    print("mode %d: %.2f ms (incl. split+alloc for mode 1) -> %.1f TFLOP/s" % (mode, t0.elapsed_time(t1), 2*M*N*K/t0.elapsed_time(t1)/1e9))
