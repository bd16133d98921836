"""Experiment: how much of the GEMM main loop is shared-memory fill?  Times the single-CTA kernel with the W
and/or A TMA loads suppressed (results are garbage; timing only)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sliders_b200 import _cabi  # noqa: E402

lib = _cabi.load()
h = _cabi.handle(0)
dev = torch.device("cuda:0")
s = C.c_void_p(torch.cuda.current_stream().cuda_stream)


def t(M, N, K, bn, debug, iters=20, force=0x2000):
    x = torch.randn(M, K, device=dev).to(torch.bfloat16)
    w = torch.randn(N, K, device=dev).to(torch.bfloat16)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    arg = bn | force | (debug << 14)
    call = lambda: lib.sb200_gemm(h, s, x.data_ptr(), K, None, 0, K, w.data_ptr(), K, out.data_ptr(), N, M, N, K, 0,
                                  None, None, 1, None, 0, None, arg)
    for _ in range(3):
        _cabi.check(call())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        call()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / iters * 1e3
    return us, 2.0 * M * N * K / us / 1e6


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "single"
    FORCE = 0x1000 if mode == "pair" else 0x2000
    shapes = ((8192, 8192, 8192), (8192, 1280, 1280)) if mode == "pair" else (
        (8192, 8192, 8192), (8192, 1280, 1280), (8192, 10240, 1280), (8192, 1280, 5120))
    for (M, N, K) in shapes:
        for bn in ((64, 128, 160, 192, 256) if mode == "pair" else (256, 128)):
            row = []
            for debug, name in ((0, "lane0-issue"), (32, "elect-issue"), (0, "lane0 again"), (32, "elect again")):
                us, tf = t(M, N, K, bn, debug, force=FORCE)
                row.append(f"{name} {us:8.1f}us {tf:6.0f}TF")
            print(f"M{M} N{N} K{K} bn{bn}: " + " | ".join(row), flush=True)


if __name__ == "__main__":
    main()
