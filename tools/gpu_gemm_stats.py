"""Where do the producer and the issuer of gemm_kernel wait?  Runs the profiling build (debug bit 5) with the epilogue
suppressed and prints, per configuration, the share of its lifetime each role spent blocked on its mbarrier
(producer: `empty`, issuer: `full`), averaged over CTAs, and the cycles per k-block."""
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


def stats(M, N, K, bn, force, debug):
    x = torch.randn(M, K, device=dev).to(torch.bfloat16)
    w = torch.randn(N, K, device=dev).to(torch.bfloat16)
    out = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
    arg = bn | force | ((debug | 32 | 8) << 14)
    for _ in range(2):
        _cabi.check(lib.sb200_gemm(h, s, x.data_ptr(), K, None, 0, K, w.data_ptr(), K, out.data_ptr(), N, M, N, K, 0,
                                   None, None, 1, None, 0, None, arg))
    torch.cuda.synchronize()
    ctas = 2 if force == 0x1000 else 1
    grid = 148
    v = out.view(-1)[: grid * 16].view(torch.int64).view(grid, 4).double().cpu()
    tiles = -(-M // (128 * ctas)) * -(-N // bn)
    kb_per_unit = tiles / (148 // ctas) * (K // 64)
    lead = v[::ctas]
    return (f"producer waits {100 * (v[:, 0] / v[:, 1]).mean():5.1f}% of {v[:, 1].mean() / kb_per_unit:6.0f} clk/kb | "
            f"issuer waits {100 * (lead[:, 2] / lead[:, 3]).mean():5.1f}% of {lead[:, 3].mean() / kb_per_unit:6.0f} clk/kb")


for (M, N, K) in ((8192, 10240, 1280), (8192, 1280, 1280)):
    for force, name in ((0x2000, "single"), (0x1000, "pair")):
        for bn in (256, 128, 64):
            for debug, dn in ((0, "loads"), (3, "noload")):
                if force == 0x1000 and debug:
                    continue
                print(f"M{M} N{N} K{K} bn{bn} {name} {dn}: {stats(M, N, K, bn, force, debug)}", flush=True)
