"""Single-shape attention forward + backward launcher for ncu captures:
    ncu --set full --clock-control none --import-source on -k regex:attention -o gpurun_out/attn python tools/gpu_prof_attn.py B heads Sq Skv [head_dim]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sliders_b200 import ops  # noqa: E402

B, heads, Sq, Skv = (int(a) for a in sys.argv[1:5])
d = int(sys.argv[5]) if len(sys.argv) > 5 else 64
dev = torch.device("cuda:0")
Cc = heads * d
qkv = torch.randn(B * Sq, 3 * Cc, device=dev).to(torch.bfloat16)
kv = torch.randn(B * Skv, 2 * Cc, device=dev).to(torch.bfloat16)
if Sq == Skv:
    q, k, v = qkv[:, :Cc], qkv[:, Cc:2 * Cc], qkv[:, 2 * Cc:]
else:
    q, k, v = qkv[:, :Cc], kv[:, :Cc], kv[:, Cc:]
dout = torch.randn(B * Sq, Cc, device=dev).to(torch.bfloat16)
lse = torch.empty(B, heads, Sq, device=dev, dtype=torch.float32)
dqkv = torch.empty(B * Sq, 3 * Cc, device=dev, dtype=torch.bfloat16)
dkv = torch.empty(B * Skv, 2 * Cc, device=dev, dtype=torch.bfloat16)
for _ in range(2):
    o = ops.attention(q, k, v, B, heads, Sq, Skv, d ** -0.5, head_dim=d, lse=lse)
    if Sq == Skv:
        ops.attention_bwd(q, k, v, o, dout, lse, B, heads, Sq, Skv, d ** -0.5, d, dqkv[:, :Cc], dqkv[:, Cc:2 * Cc],
                          dqkv[:, 2 * Cc:])
    else:
        ops.attention_bwd(q, k, v, o, dout, lse, B, heads, Sq, Skv, d ** -0.5, d, dqkv[:, :Cc])
torch.cuda.synchronize()
e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
n = 10
e0.record()
for _ in range(n):
    o = ops.attention(q, k, v, B, heads, Sq, Skv, d ** -0.5, head_dim=d, lse=lse)
e1.record()
for _ in range(n):
    if Sq == Skv:
        ops.attention_bwd(q, k, v, o, dout, lse, B, heads, Sq, Skv, d ** -0.5, d, dqkv[:, :Cc], dqkv[:, Cc:2 * Cc],
                          dqkv[:, 2 * Cc:])
    else:
        ops.attention_bwd(q, k, v, o, dout, lse, B, heads, Sq, Skv, d ** -0.5, d, dqkv[:, :Cc])
e2.record()
torch.cuda.synchronize()
fl = 4.0 * B * heads * Sq * Skv * d
tf, tb = e0.elapsed_time(e1) / n, e1.elapsed_time(e2) / n
print(f"ATTN B{B} h{heads} Sq{Sq} Skv{Skv} d{d}: fwd {tf * 1e3:.1f} us ({fl / tf / 1e9:.0f} TFLOP/s) | bwd {tb * 1e3:.1f} us "
      f"({(2.5 if Sq == Skv else 1.5) * fl / tb / 1e9:.0f} TFLOP/s of {'10' if Sq == Skv else '6'} S*S*d MACs)")
