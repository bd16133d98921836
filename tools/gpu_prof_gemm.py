"""Single-shape GEMM launcher for ncu captures: python tools/gpu_prof_gemm.py M N K [flags] [bn]
flags: SB200_EPI_* bits (1 bias, 4 residual, 8 GEGLU, 16 LoRA rank 4 with one adaptor per 1280 (or N) columns)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sliders_b200 import ops  # noqa: E402

M, N, K = (int(a) for a in sys.argv[1:4])
flags = int(sys.argv[4]) if len(sys.argv) > 4 else 0
bn = int(sys.argv[5]) if len(sys.argv) > 5 else 0
dev = torch.device("cuda:0")
x = torch.randn(M, K, device=dev).to(torch.bfloat16)
w = (torch.randn(N, K, device=dev) / K ** 0.5).to(torch.bfloat16)
geglu = bool(flags & 8)
nout = N // 2 if geglu else N
bias = torch.zeros(N, device=dev, dtype=torch.bfloat16) if flags & 1 else None
resid = torch.randn(M, nout, device=dev).to(torch.bfloat16) if flags & 4 else None
lora = None
if flags & 16:
    group_n = 1280 if N % 1280 == 0 else N
    groups = N // group_n
    rt = 16 if groups * 4 <= 16 else 32
    down = torch.zeros(rt, K, device=dev, dtype=torch.bfloat16)
    down[: groups * 4] = (torch.randn(groups * 4, K, device=dev) / K ** 0.5).to(torch.bfloat16)
    up = torch.randn(N, 4, device=dev) * 0.1
    lora = ops.Lora(down, up, 4, group_n, 0.25)
for _ in range(3):
    ops.gemm(x, w, bias=bias, resid=resid, geglu=geglu, lora=lora, bn=bn)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    ops.gemm(x, w, bias=bias, resid=resid, geglu=geglu, lora=lora, bn=bn)
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) / 20 * 1e3
print(f"GEMM M{M} N{N} K{K} flags {flags}: {us:.1f} us, {2.0 * M * N * K / us / 1e6:.0f} TFLOP/s")
