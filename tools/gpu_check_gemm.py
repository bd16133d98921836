"""Bring-up check of the tcgen05 GEMM / implicit-GEMM conv kernels on a B200 (run under gpurun).

Compares libsb200 against torch fp32 math on the same bf16 inputs and prints one line per case.
Not part of the pytest suite (tests/test_gpu_kernels.py is); this is the verbose developer harness.
"""
import ctypes as C
import json
import os
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sliders_b200 import _cabi  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False

lib = C.CDLL(_cabi.lib_path())
for name in ("sb200_create", "sb200_gemm", "sb200_conv3x3", "sb200_last_error"):
    fn = getattr(lib, name)
    fn.argtypes = _cabi.SIGNATURES[name]
    fn.restype = _cabi._RESTYPE.get(name, C.c_int)

dev = torch.device("cuda:0")
h = C.c_void_p()
assert lib.sb200_create(0, C.byref(h)) == 0, lib.sb200_last_error()
results = []
FORCE = 0  # 0 auto, 0x2000 single-CTA kernel, 0x1000 CTA-pair (cta_group::2) kernel


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def chk(st):
    if st != 0:
        raise RuntimeError(lib.sb200_last_error().decode())


def report(name, got, ref, extra=""):
    got = got.float()
    ref = ref.float()
    err = (got - ref).abs().max().item()
    rel = ((got - ref).norm() / (ref.norm() + 1e-12)).item()
    ok = rel < 1e-2 and bool(torch.isfinite(got).all())
    print(f"{'OK ' if ok else 'BAD'} {name}: max_abs {err:.4g} rel_rms {rel:.3g} {extra}", flush=True)
    results.append({"name": name, "ok": ok, "max_abs": err, "rel": rel})
    return ok


def lora_args(K, N, r, rt, group_n, scale, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    groups = (N + group_n - 1) // group_n
    down = torch.zeros(rt, K)
    down[: groups * r] = torch.randn(groups * r, K, generator=g) / K ** 0.5
    up = torch.randn(N, r, generator=g) * 0.5
    down = down.to(dev, torch.bfloat16)
    up = up.to(torch.bfloat16).float().to(dev)  # sb200_lora.up is fp32 (bf16-representable values)
    la = _cabi.LoraArgs(down.data_ptr(), up.data_ptr(), r, rt, group_n, scale)
    return la, down, up


def lora_ref(x, down, up, r, group_n, scale):
    # x [M, K] fp32; returns [M, N]
    t = x @ down.float().t()  # [M, rt]
    N = up.shape[0]
    out = torch.zeros(x.shape[0], N, device=x.device)
    for g0 in range(0, N, group_n):
        grp = g0 // group_n
        out[:, g0:g0 + group_n] = t[:, grp * r:(grp + 1) * r] @ up[g0:g0 + group_n].float().t()
    return out * scale


def run_gemm(M, N, K, flags=0, split=0, lora=None, bn=0, seed=0, name=None):
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = (torch.randn(M, K, generator=g)).to(dev, torch.bfloat16)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(dev, torch.bfloat16)
    bias = torch.randn(N, generator=g).to(dev, torch.bfloat16)
    rpb = max(1, M // 4)
    nb = (M + rpb - 1) // rpb
    nout = N // 2 if flags & _cabi.EPI_GEGLU else N
    rowbias = torch.randn(nb, nout, generator=g).to(dev, torch.bfloat16)
    resid = torch.randn(M, nout, generator=g).to(dev, torch.bfloat16)
    out = torch.full((M, nout), float("nan"), device=dev, dtype=torch.bfloat16)
    la = None
    if lora:
        r, rt, group_n, scale = lora
        la, down, up = lora_args(K, N, r, rt, group_n, scale, seed + 1)
    if split:
        x0 = x[:, :split].contiguous()
        x1 = x[:, split:].contiguous()
        a0, l0, a1, l1, k0 = x0, x0.stride(0), x1, x1.stride(0), split
    else:
        a0, l0, a1, l1, k0 = x, K, None, 0, K
    st = lib.sb200_gemm(h, stream(), ptr(a0), l0, ptr(a1), l1, k0, ptr(w), K, ptr(out), nout, M, N, K, flags,
                        ptr(bias), ptr(rowbias), rpb, ptr(resid), nout, C.byref(la) if la else None, bn | FORCE)
    chk(st)
    torch.cuda.synchronize()
    ref = x.float() @ w.float().t()
    if lora:
        ref = ref + lora_ref(x.float(), down, up, lora[0], lora[2], lora[3])
    if flags & _cabi.EPI_BIAS:
        ref = ref + bias.float()
    if flags & _cabi.EPI_GEGLU:
        a, gte = ref.chunk(2, dim=-1)
        ref = a * F.gelu(gte)
    if flags & _cabi.EPI_ROWBIAS:
        idx = torch.arange(M, device=dev) // rpb
        ref = ref + rowbias.float()[idx]
    if flags & _cabi.EPI_RESID:
        ref = ref + resid.float()
    return report((name or f"gemm M{M} N{N} K{K} flags{flags} split{split} lora{lora} bn{bn}") + f" force{FORCE:#x}", out, ref)


def run_conv(B, H, W, C0, C1, Cout, stride=1, flags=0, lora=None, bn=0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    Cin = C0 + C1
    x = torch.randn(B, H, W, Cin, generator=g).to(dev, torch.bfloat16)
    w = (torch.randn(Cout, 3, 3, Cin, generator=g) / (9 * Cin) ** 0.5).to(dev, torch.bfloat16)
    bias = torch.randn(Cout, generator=g).to(dev, torch.bfloat16)
    Ho, Wo = H // stride, W // stride
    rowbias = torch.randn(B, Cout, generator=g).to(dev, torch.bfloat16)
    resid = torch.randn(B, Ho, Wo, Cout, generator=g).to(dev, torch.bfloat16)
    out = torch.full((B, Ho, Wo, Cout), float("nan"), device=dev, dtype=torch.bfloat16)
    x0 = x[..., :C0].contiguous()
    x1 = x[..., C0:].contiguous() if C1 else None
    la = None
    if lora:
        r, rt, group_n, scale = lora
        la, down, up = lora_args(9 * Cin, Cout, r, rt, group_n, scale, seed + 1)
    st = lib.sb200_conv3x3(h, stream(), ptr(x0), C0, ptr(x1), C1, C0, C1, ptr(w), ptr(out), Cout, B, H, W, Cout,
                           stride, flags, ptr(bias), ptr(rowbias), ptr(resid), Cout,
                           C.byref(la) if la else None, bn | FORCE)
    chk(st)
    torch.cuda.synchronize()
    xn = x.float().permute(0, 3, 1, 2)
    wn = w.float().permute(0, 3, 1, 2)
    ref = F.conv2d(xn, wn, None, stride=stride, padding=1)
    if lora:
        dn = down.float().view(-1, 3, 3, Cin).permute(0, 3, 1, 2)
        t = F.conv2d(xn, dn, None, stride=stride, padding=1)  # [B, rt, Ho, Wo]
        r = lora[0]
        ref = ref + torch.einsum("brhw,or->bohw", t[:, :r], up.float()) * lora[3]
    ref = ref.permute(0, 2, 3, 1)
    if flags & _cabi.EPI_BIAS:
        ref = ref + bias.float()
    if flags & _cabi.EPI_ROWBIAS:
        ref = ref + rowbias.float()[:, None, None, :]
    if flags & _cabi.EPI_RESID:
        ref = ref + resid.float()
    return report(f"conv B{B} {H}x{W} C{C0}+{C1}->{Cout} s{stride} flags{flags} lora{lora} bn{bn} force{FORCE:#x}", out, ref)


def bench_gemm(M, N, K, flags=0, bn=0, iters=20):
    x = torch.randn(M, K, device=dev).to(torch.bfloat16)
    w = torch.randn(N, K, device=dev).to(torch.bfloat16)
    nout = N // 2 if flags & _cabi.EPI_GEGLU else N
    out = torch.empty(M, nout, device=dev, dtype=torch.bfloat16)
    bias = torch.zeros(N, device=dev, dtype=torch.bfloat16)
    resid = torch.randn(M, nout, device=dev).to(torch.bfloat16) if flags & _cabi.EPI_RESID else None
    args = (h, stream(), ptr(x), K, None, 0, K, ptr(w), K, ptr(out), nout, M, N, K, flags, ptr(bias), None, 1,
            ptr(resid), nout, None, bn | FORCE)
    for _ in range(3):
        chk(lib.sb200_gemm(*args))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        lib.sb200_gemm(*args)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    tf = 2.0 * M * N * K / ms / 1e9
    t0 = time.time()
    for _ in range(3):
        y = x @ w.t()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        y = x @ w.t()
    e1.record()
    torch.cuda.synchronize()
    ms_t = e0.elapsed_time(e1) / iters
    print(f"PERF gemm M{M} N{N} K{K} flags{flags} bn{bn} force{FORCE:#x}: {ms * 1e3:.1f} us {tf:.0f} TFLOP/s | cuBLAS "
          f"{ms_t * 1e3:.1f} us {2.0 * M * N * K / ms_t / 1e9:.0f} TFLOP/s", flush=True)
    results.append({"name": f"perf M{M} N{N} K{K} f{flags} bn{bn}", "us": ms * 1e3, "tflops": tf,
                    "cublas_us": ms_t * 1e3})


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    print(torch.cuda.get_device_name(0), flush=True)
    B_, R_, RB_, G_, L_ = _cabi.EPI_BIAS, _cabi.EPI_RESID, _cabi.EPI_ROWBIAS, _cabi.EPI_GEGLU, _cabi.EPI_LORA
    ok = True
    for FORCE in ((0x2000, 0x1000, 0) if which != 'pair' else (0x1000,)):
        globals()['FORCE'] = FORCE
        print(f'---- force {FORCE:#x}', flush=True)
        # 1. smallest possible: one tile, one k-block
        ok &= run_gemm(128, 64, 64, bn=64, name="gemm 1 tile 1 kblock")
        ok &= run_gemm(128, 128, 256, bn=128, name="gemm 1 tile 4 kblocks")
        ok &= run_gemm(256, 256, 512, bn=128)
        ok &= run_gemm(1024, 1280, 1280, flags=B_)
        ok &= run_gemm(1000, 640, 640, flags=B_ | R_)             # M tail
        ok &= run_gemm(77 * 2, 1280, 2048)                        # cross-attn K/V shape
        ok &= run_gemm(2048, 2560, 640, flags=B_ | G_)            # GEGLU
        ok &= run_gemm(1024, 640, 1920, flags=B_, split=1280)     # concat shortcut
        ok &= run_gemm(1024, 1280, 1280, flags=B_ | L_, lora=(4, 16, 1280, 0.25))
        ok &= run_gemm(1024, 1920, 640, flags=L_, lora=(4, 16, 640, 1.0))  # fused qkv
        ok &= run_gemm(512, 640, 640, flags=L_ | B_ | R_, lora=(8, 16, 640, -2.0))
        ok &= run_gemm(4096, 1280, 1280, flags=B_ | RB_ | R_)
        if which in ("all", "conv"):
            ok &= run_conv(1, 32, 32, 64, 0, 64, bn=64)
            ok &= run_conv(2, 32, 32, 128, 0, 128, flags=B_)
            ok &= run_conv(1, 64, 64, 320, 0, 320, flags=B_ | RB_)
            ok &= run_conv(1, 128, 128, 320, 0, 320, flags=B_ | R_)
            ok &= run_conv(2, 32, 32, 1280, 640, 1280, flags=B_)  # concat
            ok &= run_conv(2, 64, 64, 320, 0, 320, stride=2, flags=B_)
            ok &= run_conv(2, 32, 32, 640, 0, 640, flags=B_ | L_, lora=(4, 16, 640, 0.5))
            ok &= run_conv(3, 8, 8, 128, 0, 128, flags=B_)       # SD1.x smallest level (bb=2, M tail)
            ok &= run_conv(1, 16, 16, 128, 64, 64, flags=B_)
    if which in ("all", "perf"):
      for FORCE in (0x2000, 0x1000, 0):
        globals()['FORCE'] = FORCE
        bench_gemm(8192, 1280, 1280)
        bench_gemm(8192, 10240, 1280, flags=G_ | B_)
        bench_gemm(8192, 1280, 5120, flags=B_)
        bench_gemm(8192, 3840, 1280)
        bench_gemm(32768, 640, 640)
        bench_gemm(8192, 8192, 8192)
        bench_gemm(8192, 8192, 8192, bn=128)
        bench_gemm(8192, 1280, 1280, flags=B_ | R_)
        bench_gemm(32768, 640, 640, flags=B_ | R_)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/gemm_check.json", "w") as f:
        json.dump(results, f, indent=1)
    print("ALL OK" if ok else "SOME BAD", flush=True)
    sys.exit(0 if ok else 1)
