"""Where do the roles of gemm_kernel wait?  Runs the profiling build (debug bit 5) of the production configuration of a
few UNet call sites and prints, per role, the share of its lifetime spent blocked on its mbarrier:
  producer 0 on `empty`, the issuer on `full` and on the accumulator hand-back, epilogue warp 0 on `tfull`.
An epilogue that (almost) never waits is the bottleneck of the launch."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sliders_b200 import _cabi  # noqa: E402
from sliders_b200.ops import Lora  # noqa: E402

lib = _cabi.load()
h = _cabi.handle(0)
dev = torch.device("cuda:0")
s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
BF = torch.bfloat16


def run(M, N, K, flags, bn=0, force=0, debug=0, iters=10):
    """flags: string of b (bias) R (residual) G (GEGLU) L (LoRA rank 4)."""
    x = torch.randn(M, K, device=dev).to(BF)
    w = (torch.randn(N, K, device=dev) / K ** 0.5).to(BF)
    geglu = "G" in flags
    nout = N // 2 if geglu else N
    out = torch.zeros(M, nout, device=dev, dtype=BF)
    bias = torch.randn(N, device=dev).to(BF) if "b" in flags else None
    resid = torch.randn(M, nout, device=dev).to(BF) if "R" in flags else None
    la = None
    if "L" in flags:
        group_n = 1280 if N % 1280 == 0 else N
        groups = N // group_n
        rt = 16 if groups * 4 <= 16 else 32
        down = torch.zeros(rt, K, device=dev, dtype=BF)
        down[: groups * 4] = (torch.randn(groups * 4, K, device=dev) / K ** 0.5).to(BF)
        la = Lora(down, torch.randn(N, 4, device=dev) * 0.1, 4, group_n, 0.25)
    fl = (1 if bias is not None else 0) | (4 if resid is not None else 0) | (8 if geglu else 0) | (16 if la else 0)
    stats = torch.zeros(148 * 8, device=dev, dtype=torch.int64)
    p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)

    def call(dbg, sp):
        arg = (bn | force | (dbg << 14)) if (bn or force or dbg) else 0
        return lib.sb200_gemm(h, s, p(x), K, None, 0, K, p(w), K, p(out), nout, M, N, K, fl, p(bias), sp, 1, p(resid),
                              nout if resid is not None else 0, la.ref() if la else None, arg)

    for _ in range(3):
        _cabi.check(call(debug, C.c_void_p(0)))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        call(debug, C.c_void_p(0))
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / iters * 1e3
    _cabi.check(call(debug | 32, p(stats)))     # bn / force 0 with a debug bit: encode needs a nonzero low part
    torch.cuda.synchronize()
    v = stats.view(148, 8).double().cpu()
    act = v[v[:, 5] > 0]
    lead = v[v[:, 3] > 0]
    pct = lambda a, b: 100 * (a / b).mean().item()
    return (f"{us:7.1f}us {2.0 * M * N * K / us / 1e6:5.0f}TF | producer waits {pct(act[:, 0], act[:, 1]):4.1f}% | issuer waits "
            f"{pct(lead[:, 2], lead[:, 3]):4.1f}% on full, {pct(lead[:, 6], lead[:, 3]):4.1f}% on the accumulator | epilogue waits "
            f"{pct(act[:, 4], act[:, 5]):4.1f}%")


cases = [(8192, 1280, 1280, ""), (8192, 1280, 1280, "bR"), (8192, 1280, 1280, "bRL"), (8192, 3840, 1280, "L"),
         (8192, 10240, 1280, "bG"), (8192, 1280, 5120, "bR"), (32768, 5120, 640, "bG"), (32768, 640, 640, "bR"),
         (32768, 640, 640, "bRL"), (32768, 1920, 640, "L"), (616, 153600, 2048, ""), (2048, 1280, 1280, "bRL")]
for (M, N, K, fl) in cases:
    print(f"M{M} N{N} K{K} {fl or '-':4s} auto tile: {run(M, N, K, fl, bn=0, force=0, debug=0)}", flush=True)
