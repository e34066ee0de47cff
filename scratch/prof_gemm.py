import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from talkshow_b200 import _lib
from talkshow_b200.engine import Engine
e = Engine(0); e.set_tensor_cores(5)
M, N, K = 19200, 3072, 768
if len(sys.argv) > 3: M, N, K = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
A = torch.randn(M, K, device='cuda'); W = torch.randn(N, K, device='cuda') / K ** 0.5; b = torch.zeros(N, device='cuda')
out = torch.empty(M, N, device='cuda')
for mode in (1, 2, 1, 2):
    rc = e.L.ts_debug_gemm(e.h, mode, _lib.ptr(A), _lib.ptr(W), _lib.ptr(b), _lib.ptr(out), M, N, K, 0, _lib.stream_ptr(e.device))
    assert rc == 0, e.L.ts_last_error(e.h)
    torch.cuda.synchronize()
