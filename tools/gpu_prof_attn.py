"""Single-shape attention launcher for ncu captures: python tools/gpu_prof_attn.py B heads Sq Skv"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sliders_b200 import _cabi  # noqa: E402

B, heads, Sq, Skv = (int(a) for a in sys.argv[1:5])
lib = _cabi.load()
h = _cabi.handle(0)
dev = torch.device("cuda:0")
Cc = heads * 64
qkv = torch.randn(B * Sq, 3 * Cc, device=dev).to(torch.bfloat16)
kv = torch.randn(B * Skv, 2 * Cc, device=dev).to(torch.bfloat16)
o = torch.empty(B * Sq, Cc, device=dev, dtype=torch.bfloat16)
s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
if Sq == Skv:
    q, k, v, ld = qkv, qkv[:, Cc:], qkv[:, 2 * Cc:], 3 * Cc
    ldk = ld
else:
    q, k, v, ld, ldk = qkv, kv, kv[:, Cc:], 3 * Cc, 2 * Cc
for _ in range(3):
    _cabi.check(lib.sb200_attention(h, s, q.data_ptr(), ld, k.data_ptr(), ldk, v.data_ptr(), ldk, o.data_ptr(), Cc, B,
                                    heads, Sq, Skv, 64, 0.125))
torch.cuda.synchronize()
