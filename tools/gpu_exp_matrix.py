"""Same-box matrix: {single, pair} x {full, no TMA loads, no loads + no epilogue} over the GEMM shapes that dominate the
SDXL forward, with torch.matmul (cuBLAS) beside it for calibration only.  Timing only (suppressed variants compute garbage)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gpu_exp_fill import t  # noqa: E402

shapes = ((8192, 1280, 1280, 256), (8192, 10240, 1280, 256), (8192, 1280, 5120, 256),
          (8192, 3840, 1280, 256), (32768, 640, 640, 256), (32768, 5120, 640, 256), (2048, 1280, 1280, 256),
          (2048, 10240, 1280, 256), (2048, 1280, 5120, 256), (1024, 1280, 1280, 256))
if len(sys.argv) > 1:
    shapes = shapes[: int(sys.argv[1])]


def cublas(M, N, K, iters=20):
    a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    b = torch.randn(N, K, device="cuda").to(torch.bfloat16)
    for _ in range(3):
        a @ b.t()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        a @ b.t()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / iters * 1e3
    return us, 2.0 * M * N * K / us / 1e6


for (M, N, K, bn) in shapes:
    us, tf = cublas(M, N, K)
    print(f"M{M} N{N} K{K} cuBLAS (context): {us:7.1f}us {tf:5.0f}TF", flush=True)
    for force, kname in ((0x2000, "single"), (0x1000, "pair  ")):
        row = []
        for b in (256, 224, 192, 160, 128, 96):
            if b > bn:
                continue
            try:
                us, tf = t(M, N, K, b, 0, force=force)
                row.append(f"bn{b} {us:6.1f}us {tf:5.0f}TF")
            except Exception as e:  # noqa: BLE001
                row.append(f"bn{b} ERR {str(e)[:30]}")
        print(f"M{M} N{N} K{K} {kname}: " + " | ".join(row), flush=True)
