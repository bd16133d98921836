"""Bring-up check of attention / norm / elementwise kernels on a B200 (run under gpurun)."""
import ctypes as C
import math
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sliders_b200 import _cabi  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
lib = _cabi.load()
dev = torch.device("cuda:0")
h = _cabi.handle(0)
BF = torch.bfloat16
all_ok = True


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def report(name, got, ref, tol=1e-2):
    global all_ok
    got, ref = got.float(), ref.float()
    err = (got - ref).abs().max().item()
    rel = ((got - ref).norm() / (ref.norm() + 1e-12)).item()
    ok = rel < tol and bool(torch.isfinite(got).all())
    all_ok &= ok
    print(f"{'OK ' if ok else 'BAD'} {name}: max_abs {err:.4g} rel_rms {rel:.3g}", flush=True)


def attn_case(B, heads, Sq, Skv, seed=0, fused_qkv=False, perf=False):
    g = torch.Generator().manual_seed(seed)
    Cc = heads * 64
    if fused_qkv:
        qkv = torch.randn(B * Sq, 3 * Cc, generator=g).to(dev, BF)
        q, k, v = qkv[:, :Cc], qkv[:, Cc:2 * Cc], qkv[:, 2 * Cc:]
        ldq = ldk = ldv = 3 * Cc
    else:
        q = torch.randn(B * Sq, Cc, generator=g).to(dev, BF)
        kv = torch.randn(B * Skv, 2 * Cc, generator=g).to(dev, BF)
        k, v = kv[:, :Cc], kv[:, Cc:]
        ldq, ldk, ldv = Cc, 2 * Cc, 2 * Cc
    o = torch.full((B * Sq, Cc), float("nan"), device=dev, dtype=BF)
    scale = 1.0 / 8.0
    args = (h, stream(), ptr(q), ldq, ptr(k), ldk, ptr(v), ldv, ptr(o), Cc, B, heads, Sq, Skv, 64, scale, None)
    _cabi.check(lib.sb200_attention(*args))
    torch.cuda.synchronize()
    qf = q.float().reshape(B, Sq, heads, 64).transpose(1, 2)
    kf = k.float().reshape(B, Skv, heads, 64).transpose(1, 2)
    vf = v.float().reshape(B, Skv, heads, 64).transpose(1, 2)
    ref = torch.softmax(qf @ kf.transpose(-1, -2) * scale, dim=-1) @ vf
    ref = ref.transpose(1, 2).reshape(B * Sq, Cc)
    report(f"attention B{B} h{heads} Sq{Sq} Skv{Skv} fused{fused_qkv}", o, ref, tol=2e-2)
    if perf:
        for _ in range(3):
            lib.sb200_attention(*args)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            lib.sb200_attention(*args)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        fl = 4.0 * B * heads * Sq * Skv * 64
        qq, kk, vv = (t.reshape(B, -1, heads, 64).transpose(1, 2) for t in (q, k, v))
        for _ in range(3):
            F.scaled_dot_product_attention(qq, kk, vv)
        e0.record()
        for _ in range(20):
            F.scaled_dot_product_attention(qq, kk, vv)
        e1.record()
        torch.cuda.synchronize()
        ms2 = e0.elapsed_time(e1) / 20
        print(f"PERF attention B{B} h{heads} Sq{Sq} Skv{Skv}: {ms * 1e3:.1f} us {fl / ms / 1e9:.0f} TFLOP/s | "
              f"torch SDPA {ms2 * 1e3:.1f} us {fl / ms2 / 1e9:.0f} TFLOP/s", flush=True)


def gn_case(B, HW, C0, C1, silu, eps=1e-5, seed=0):
    g = torch.Generator().manual_seed(seed)
    Cc = C0 + C1
    x = (torch.randn(B, HW, Cc, generator=g) * 2 + 0.5).to(dev, BF)
    gamma = (1 + 0.1 * torch.randn(Cc, generator=g)).to(dev, BF)
    beta = (0.1 * torch.randn(Cc, generator=g)).to(dev, BF)
    x0 = x[..., :C0].contiguous()
    x1 = x[..., C0:].contiguous() if C1 else None
    out = torch.full((B, HW, Cc), float("nan"), device=dev, dtype=BF)
    ws = torch.empty(B * 32 * 2 * 129, device=dev, dtype=torch.float32)
    _cabi.check(lib.sb200_groupnorm(h, stream(), ptr(x0), C0, C0, ptr(x1), C1, C1, ptr(gamma), ptr(beta), ptr(out),
                                    Cc, B, HW, 32, eps, int(silu), ptr(ws)))
    torch.cuda.synchronize()
    ref = F.group_norm(x.float().transpose(1, 2), 32, gamma.float(), beta.float(), eps).transpose(1, 2)
    if silu:
        ref = F.silu(ref)
    report(f"groupnorm B{B} HW{HW} C{C0}+{C1} silu{silu}", out, ref)


def ln_case(M, Cc, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = (torch.randn(M, Cc, generator=g) * 3 - 1).to(dev, BF)
    gamma = (1 + 0.1 * torch.randn(Cc, generator=g)).to(dev, BF)
    beta = (0.1 * torch.randn(Cc, generator=g)).to(dev, BF)
    out = torch.full((M, Cc), float("nan"), device=dev, dtype=BF)
    _cabi.check(lib.sb200_layernorm(h, stream(), ptr(x), Cc, ptr(gamma), ptr(beta), ptr(out), Cc, M, Cc, 1e-5))
    torch.cuda.synchronize()
    ref = F.layer_norm(x.float(), (Cc,), gamma.float(), beta.float(), 1e-5)
    report(f"layernorm M{M} C{Cc}", out, ref)


def small_linear_case(M, N, K, act_in, act_out, lora=None, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(M, K, generator=g).to(dev, BF)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(dev, BF)
    b = torch.randn(N, generator=g).to(dev, BF)
    out = torch.full((M, N), float("nan"), device=dev, dtype=BF)
    la = None
    if lora:
        r, scale = lora
        down = torch.zeros(16, K)
        down[:r] = torch.randn(r, K, generator=g) / K ** 0.5
        up = torch.randn(N, r, generator=g) * 0.5
        down, up = down.to(dev, BF), up.to(BF).float().to(dev)  # up is fp32 in the ABI
        la = _cabi.LoraArgs(down.data_ptr(), up.data_ptr(), r, 16, N, scale)
    _cabi.check(lib.sb200_small_linear(h, stream(), ptr(x), K, ptr(w), K, ptr(b), ptr(out), N, M, N, K, act_in,
                                       act_out, C.byref(la) if la else None, None))
    torch.cuda.synchronize()
    xin = F.silu(x.float()).to(BF).float() if act_in else x.float()
    ref = xin @ w.float().t() + b.float()
    if lora:
        ref = ref + (xin @ down.float()[:r].t()) @ up.float().t() * scale
    if act_out:
        ref = F.silu(ref)
    report(f"small_linear M{M} N{N} K{K} act{act_in}{act_out} lora{lora}", out, ref)


def misc_cases():
    # sinusoid
    vals = torch.tensor([0.0, 1.0, 500.0, 999.0, 1024.0], device=dev)
    out = torch.empty(5, 320, device=dev, dtype=BF)
    _cabi.check(lib.sb200_sinusoid(h, stream(), ptr(vals), 5, 320, ptr(out), 320))
    half = 160
    freqs = torch.exp(-math.log(10000.0) * torch.arange(half, device=dev, dtype=torch.float32) / half)
    arg = vals[:, None] * freqs[None]
    ref = torch.cat([torch.cos(arg), torch.sin(arg)], -1)
    report("sinusoid", out, ref)
    # conv_in
    g = torch.Generator().manual_seed(1)
    lat = torch.randn(2, 4, 32, 32, generator=g).to(dev)
    w = (torch.randn(320, 3, 3, 4, generator=g) / 6).to(dev, BF)
    b = torch.randn(320, generator=g).to(dev, BF)
    out = torch.empty(2, 32, 32, 320, device=dev, dtype=BF)
    _cabi.check(lib.sb200_conv_in(h, stream(), ptr(lat), 1, ptr(w), ptr(b), ptr(out), 2, 32, 32, 320))
    ref = F.conv2d(lat.to(BF).float(), w.float().permute(0, 3, 1, 2), b.float(), padding=1).permute(0, 2, 3, 1)
    report("conv_in f32 latent", out, ref)
    latb = lat.to(BF)
    _cabi.check(lib.sb200_conv_in(h, stream(), ptr(latb), 0, ptr(w), ptr(b), ptr(out), 2, 32, 32, 320))
    report("conv_in bf16 latent", out, ref)
    # conv_out
    x = torch.randn(2, 32, 32, 320, generator=g).to(dev, BF)
    w = (torch.randn(4, 3, 3, 320, generator=g) / 54).to(dev, BF)
    b = torch.randn(4, generator=g).to(dev, BF)
    out = torch.empty(2, 4, 32, 32, device=dev, dtype=torch.float32)
    _cabi.check(lib.sb200_conv_out(h, stream(), ptr(x), ptr(w), ptr(b), ptr(out), 1, 2, 32, 32, 320))
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float().permute(0, 3, 1, 2), b.float(), padding=1)
    report("conv_out f32", out, ref)
    outb = torch.empty(2, 4, 32, 32, device=dev, dtype=BF)
    _cabi.check(lib.sb200_conv_out(h, stream(), ptr(x), ptr(w), ptr(b), ptr(outb), 0, 2, 32, 32, 320))
    report("conv_out bf16", outb, ref)
    # upsample
    x = torch.randn(2, 8, 8, 64, generator=g).to(dev, BF)
    out = torch.empty(2, 16, 16, 64, device=dev, dtype=BF)
    _cabi.check(lib.sb200_upsample2x(h, stream(), ptr(x), ptr(out), 2, 8, 8, 64))
    ref = x.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2)
    report("upsample2x", out, ref, tol=1e-6)
    # cfg + ddim
    n = 2 * 4 * 32 * 32
    eps2 = torch.randn(2 * n, generator=g).to(dev, BF)
    xx = torch.randn(n, generator=g).to(dev)
    xp = torch.empty(n, device=dev)
    eo = torch.empty(n, device=dev)
    a_t, a_p, gs = 0.3, 0.5, 3.0
    _cabi.check(lib.sb200_cfg_ddim(h, stream(), ptr(eps2), 0, gs, ptr(xx), a_t, a_p, ptr(xp), ptr(eo), 1, n))
    eu, ec = eps2[:n].float(), eps2[n:].float()
    e = eu + gs * (ec - eu)
    x0 = (xx - (1 - a_t) ** 0.5 * e) / a_t ** 0.5
    report("cfg eps", eo, e, tol=1e-5)
    report("ddim step", xp, a_p ** 0.5 * x0 + (1 - a_p) ** 0.5 * e, tol=1e-5)


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), flush=True)
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "attn"):
        attn_case(1, 1, 128, 128)
        attn_case(1, 2, 256, 256, fused_qkv=True)
        attn_case(2, 5, 1024, 1024, fused_qkv=True)
        attn_case(2, 4, 1024, 77)
        attn_case(1, 2, 64, 64, fused_qkv=True)       # Sq, Skv < tile
        attn_case(1, 3, 320, 200)                     # ragged both
        attn_case(2, 10, 4096, 4096, fused_qkv=True, perf=True)
        attn_case(4, 20, 1024, 1024, fused_qkv=True, perf=True)
        attn_case(4, 20, 1024, 77, perf=True)
    if which in ("all", "norm"):
        gn_case(2, 1024, 320, 0, True)
        gn_case(2, 4096, 640, 320, True)
        gn_case(1, 16384, 320, 0, False, eps=1e-6)
        gn_case(2, 1024, 1280, 1280, True)
        gn_case(3, 64, 64, 64, True)
        ln_case(2048, 640)
        ln_case(1000, 1280)
        ln_case(77, 128)
        small_linear_case(2, 1280, 320, 0, 1)
        small_linear_case(2, 1280, 1280, 0, 0)
        small_linear_case(8, 1280, 2816, 0, 1)
        small_linear_case(4, 640, 1280, 1, 0)
        small_linear_case(16, 320, 1280, 1, 0, lora=(4, 0.5))
        small_linear_case(3, 1280, 1280, 1, 0, lora=(8, -1.5))
        misc_cases()
    print("ALL OK" if all_ok else "SOME BAD", flush=True)
    sys.exit(0 if all_ok else 1)
