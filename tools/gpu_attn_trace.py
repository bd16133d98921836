"""Phase timeline of CTA (0,0,0) of the ping-pong attention kernel (trace build: SB200_ATTN_POLY=1).
    SB200_ATTN_POLY=1 python tools/gpu_attn_trace.py [B heads S]"""
import ctypes
import os
import sys

os.environ["SB200_ATTN_POLY"] = "1"
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sliders_b200 import _cabi, ops  # noqa: E402

B, heads, S = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (8, 10, 4096)
dev = torch.device("cuda:0")
Cc = heads * 64
qkv = torch.randn(B * S, 3 * Cc, device=dev).to(torch.bfloat16)
q, k, v = qkv[:, :Cc], qkv[:, Cc:2 * Cc], qkv[:, 2 * Cc:]
for _ in range(3):
    ops.attention(q, k, v, B, heads, S, S, 0.125)
torch.cuda.synchronize()
L = 2048
buf = (ctypes.c_longlong * (4 * L))()
lib = _cabi.load()
rc = lib.sb200_debug_attention_trace(ctypes.cast(buf, ctypes.c_void_p), 4 * L)
assert rc == 0, rc
n = S // 128
soft = [[buf[r * L + 6 * j + e] for e in range(6)] for r in (0, 1) for j in range(n)]
s0, s1 = soft[:n], soft[n:]
qk = [(buf[2 * L + 2 * e], buf[2 * L + 2 * e + 1]) for e in range(2 * n)]
pv = [(buf[3 * L + 2 * e], buf[3 * L + 2 * e + 1]) for e in range(2 * n)]
t0 = min(s0[0][0], qk[0][0])
names = ["S full", "S in regs", "max done", "32 exp done", "prev PV done", "P stored"]
print(f"trace of CTA (0,0,0): B{B} h{heads} S{S}, clocks relative to the first Q K^T issue")
for j in range(4, min(n, 10)):
    for i, st in ((0, s0), (1, s1)):
        row = st[j]
        d = [row[e] - row[e - 1] for e in range(1, 6)]
        print(f"tile {i} key tile {j:2d}: S full @{row[0] - t0:7d} | load {d[0]:4d} max {d[1]:4d} exp32 {d[2]:4d} "
              f"pv-wait {d[3]:4d} exp96+store {d[4]:5d} | P stored @{row[5] - t0:7d} | "
              f"step {row[5] - st[j - 1][5]:5d} wait-for-S {row[0] - st[j - 1][5]:5d}")
# issue order of the Q K^T warp: qk(0,0), qk(0,1), qk(1,0), then qk(1,j), qk(0,j+1)
print("Q K^T issuer: (ready, issued) relative clocks, first 14:", [(a - t0, b - a) for a, b in qk[:14]])
print("P V issuer:   (P full seen, issue time) first 14:", [(a - t0, b - a) for a, b in pv[:14]])
per = [s0[j][5] - s0[j - 1][5] for j in range(4, n)]
print(f"tile 0 mean step {sum(per) / len(per):.0f} clk; tile 1 mean step "
      f"{sum(s1[j][5] - s1[j - 1][5] for j in range(4, n)) / (n - 4):.0f} clk; "
      f"offset tile1 - tile0 at key tile 8: {s1[8][5] - s0[8][5]} clk")
