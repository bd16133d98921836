"""Single-shape GEMM launcher for ncu captures: python tools/gpu_prof_gemm.py M N K [flags] [bn]"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sliders_b200 import _cabi  # noqa: E402

M, N, K = (int(a) for a in sys.argv[1:4])
flags = int(sys.argv[4]) if len(sys.argv) > 4 else 0
bn = int(sys.argv[5]) if len(sys.argv) > 5 else 0
lib = _cabi.load()
h = _cabi.handle(0)
dev = torch.device("cuda:0")
x = torch.randn(M, K, device=dev).to(torch.bfloat16)
w = torch.randn(N, K, device=dev).to(torch.bfloat16)
nout = N // 2 if flags & 8 else N
out = torch.empty(M, nout, device=dev, dtype=torch.bfloat16)
bias = torch.zeros(N, device=dev, dtype=torch.bfloat16)
s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for _ in range(3):
    _cabi.check(lib.sb200_gemm(h, s, x.data_ptr(), K, None, 0, K, w.data_ptr(), K, out.data_ptr(), nout, M, N, K,
                               flags, bias.data_ptr(), None, 1, None, 0, None, bn))
torch.cuda.synchronize()
