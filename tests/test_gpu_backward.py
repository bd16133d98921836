"""GPU parity tests of the backward-to-LoRA kernels (csrc/backward.cu, csrc/attention_bwd.cu), called through the C
ABI and compared with torch autograd in fp32 on the same bf16 inputs.  Tolerances: gradients are bf16 (2^-9 relative
rounding) -> rel-RMS <= 1e-2 for the elementwise / reduction kernels; attention backward additionally rounds P and dS
to bf16 (like flash-attention) -> 3e-2.  AdamW must match torch.optim.AdamW on bf16 parameters bit for bit."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

BF = torch.bfloat16


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    return torch.device("cuda:0")


def rel_rms(got, ref):
    got, ref = got.float(), ref.float()
    assert torch.isfinite(got).all()
    return ((got - ref).norm() / (ref.norm() + 1e-12)).item()


@pytest.mark.parametrize("B,heads,Sq,Skv,d,cross", [
    (2, 5, 1024, 1024, 64, False),
    (1, 10, 4096, 4096, 64, False),   # SDXL 64x64 level
    (2, 4, 1024, 77, 64, True),       # cross-attention: dq only
    (2, 4, 1024, 77, 64, False),      # ... and with dk / dv (train methods that adapt attn2.to_k / to_v)
    (1, 3, 320, 200, 64, False),      # ragged
    (1, 2, 64, 64, 64, False),
    (2, 8, 1024, 1024, 80, False),    # SD1.x
    (2, 8, 256, 256, 160, False),
    (1, 8, 4096, 4096, 40, False),
    (3, 8, 64, 64, 160, False),
    (1, 5, 320, 200, 24, False),
])
def test_attention_bwd(dev, B, heads, Sq, Skv, d, cross):
    from sliders_b200 import ops

    g = torch.Generator().manual_seed(Sq + Skv + d)
    Cc = heads * d
    qkv = torch.randn(B * Sq, Cc, generator=g).to(dev, BF)
    kv = torch.randn(B * Skv, 2 * Cc, generator=g).to(dev, BF)
    q, k, v = qkv, kv[:, :Cc], kv[:, Cc:]
    dout = torch.randn(B * Sq, Cc, generator=g).to(dev, BF)
    scale = d ** -0.5
    lse = torch.empty(B, heads, Sq, device=dev, dtype=torch.float32)
    o = ops.attention(q, k, v, B, heads, Sq, Skv, scale, head_dim=d, lse=lse)
    dq = torch.empty_like(q)
    dkv = torch.empty_like(kv)
    if cross:
        ops.attention_bwd(q, k, v, o, dout, lse, B, heads, Sq, Skv, scale, d, dq)
    else:
        ops.attention_bwd(q, k, v, o, dout, lse, B, heads, Sq, Skv, scale, d, dq, dkv[:, :Cc], dkv[:, Cc:])
    torch.cuda.synchronize()
    qf = q.float().reshape(B, Sq, heads, d).transpose(1, 2).requires_grad_()
    kf = k.float().reshape(B, Skv, heads, d).transpose(1, 2).requires_grad_()
    vf = v.float().reshape(B, Skv, heads, d).transpose(1, 2).requires_grad_()
    ref = (torch.softmax(qf @ kf.transpose(-1, -2) * scale, dim=-1) @ vf).transpose(1, 2).reshape(B * Sq, Cc)
    ref.backward(dout.float())
    back = lambda t, S: t.transpose(1, 2).reshape(B * S, Cc)
    assert rel_rms(dq, back(qf.grad, Sq)) < 3e-2
    if not cross:
        assert rel_rms(dkv[:, :Cc], back(kf.grad, Skv)) < 3e-2
        assert rel_rms(dkv[:, Cc:], back(vf.grad, Skv)) < 3e-2


@pytest.mark.parametrize("B,H,W,C0,C1,groups,silu,with_add", [
    (2, 32, 32, 320, 0, 32, True, False),
    (2, 16, 16, 1280, 640, 32, True, True),   # up-block resnet norm1: concat input, + shortcut-path gradient
    (1, 64, 64, 640, 0, 32, False, True),     # Transformer2DModel.norm
    (3, 8, 8, 128, 0, 32, True, False),
    (1, 128, 128, 320, 0, 32, True, False),
])
def test_groupnorm_bwd(dev, B, H, W, C0, C1, groups, silu, with_add):
    from sliders_b200 import ops

    g = torch.Generator().manual_seed(H + C0)
    C = C0 + C1
    x = (torch.randn(B, H, W, C, generator=g) * 1.5 + 0.3).to(dev, BF)
    gamma = (1 + 0.2 * torch.randn(C, generator=g)).to(dev, BF)
    beta = (0.2 * torch.randn(C, generator=g)).to(dev, BF)
    dy = torch.randn(B, H, W, C, generator=g).to(dev, BF)
    add = torch.randn(B, H, W, C, generator=g).to(dev, BF) if with_add else None
    x0 = x[..., :C0].contiguous()
    x1 = x[..., C0:].contiguous() if C1 else None
    ws = ops.gn_ws(B, groups, dev)
    ops.groupnorm(x0, gamma, beta, groups, 1e-5, silu, x1=x1, stats_ws=ws)
    dx = ops.groupnorm_bwd(x0, gamma, beta, groups, silu, dy, ws, x1=x1, add=add)
    torch.cuda.synchronize()
    xf = x.float().permute(0, 3, 1, 2).requires_grad_()
    y = F.group_norm(xf, groups, gamma.float(), beta.float(), 1e-5)
    if silu:
        y = F.silu(y)
    y.backward(dy.float().permute(0, 3, 1, 2))
    ref = xf.grad.permute(0, 2, 3, 1)
    if add is not None:
        ref = ref + add.float()
    assert rel_rms(dx, ref) < 1e-2


@pytest.mark.parametrize("M,C,with_add", [(1000, 640, True), (2048, 1280, False), (64, 320, True)])
def test_layernorm_bwd(dev, M, C, with_add):
    from sliders_b200 import ops

    g = torch.Generator().manual_seed(M + C)
    x = (torch.randn(M, C, generator=g) * 2 + 0.5).to(dev, BF)
    gamma = (1 + 0.2 * torch.randn(C, generator=g)).to(dev, BF)
    dy = torch.randn(M, C, generator=g).to(dev, BF)
    add = torch.randn(M, C, generator=g).to(dev, BF) if with_add else None
    dx = ops.layernorm_bwd(x, gamma, dy, 1e-5, add=add)
    torch.cuda.synchronize()
    xf = x.float().requires_grad_()
    F.layer_norm(xf, (C,), gamma.float(), torch.zeros(C, device=dev), 1e-5).backward(dy.float())
    ref = xf.grad + (add.float() if add is not None else 0)
    assert rel_rms(dx, ref) < 1e-2


def test_geglu_fwd_bwd(dev):
    from sliders_b200 import ops

    g = torch.Generator().manual_seed(3)
    M, Fh = 1000, 2560
    pre = (torch.randn(M, 2 * Fh, generator=g) * 1.5).to(dev, BF)
    dout = torch.randn(M, Fh, generator=g).to(dev, BF)
    out = ops.geglu(pre)
    dpre = ops.geglu_bwd(pre, dout)
    torch.cuda.synchronize()
    pf = pre.float().requires_grad_()
    a, gate = pf.chunk(2, dim=-1)
    ref = a * F.gelu(gate)
    ref.backward(dout.float())
    assert rel_rms(out, ref) < 1e-2
    assert rel_rms(dpre, pf.grad) < 1e-2


def test_add_upsample_zero_stuff_colsum(dev):
    from sliders_b200 import ops

    g = torch.Generator().manual_seed(4)
    a = torch.randn(300, 1920, generator=g).to(dev, BF)
    b = torch.randn(300, 640, generator=g).to(dev, BF)
    c = torch.randn(300, 640, generator=g).to(dev, BF)
    out = ops.add(a[:, 640:1280], b, c)
    assert rel_rms(out, a[:, 640:1280].float() + b.float() + c.float()) < 4e-3
    dy = torch.randn(2, 16, 24, 64, generator=g).to(dev, BF)
    dx = ops.upsample2x_bwd(dy)
    ref = dy.float().view(2, 8, 2, 12, 2, 64).sum(dim=(2, 4))
    assert rel_rms(dx, ref) < 4e-3
    z = ops.zero_stuff(dy)
    zr = torch.zeros(2, 32, 48, 64, device=dev)
    zr[:, ::2, ::2] = dy.float()
    assert torch.equal(z.float(), zr)
    big = torch.randn(3, 40, 40, 320, generator=g).to(dev, BF)
    cs = ops.colsum(big)
    assert rel_rms(cs, big.float().sum(dim=(1, 2))) < 1e-5


@pytest.mark.parametrize("stride,C0,C1,Cout", [(1, 128, 0, 192), (1, 128, 64, 128), (2, 128, 0, 128)])
def test_conv3x3_input_grad_via_forward_kernel(dev, stride, C0, C1, Cout):
    """dX of a 3x3 conv = the same implicit-GEMM kernel on dY (zero-stuffed for stride 2) with the flipped-tap,
    channel-transposed weight (UNet2DConditionModel._w_dgrad)."""
    from sliders_b200 import ops

    g = torch.Generator().manual_seed(stride + C1)
    B, H, W = 2, 16, 16
    Cin = C0 + C1
    x = torch.randn(B, Cin, H, W, generator=g).to(dev, BF)
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) / (9 * Cin) ** 0.5).to(dev, BF)
    dy = torch.randn(B, H // stride, W // stride, Cout, generator=g).to(dev, BF)
    wd = w.permute(0, 2, 3, 1).flip(1, 2).permute(3, 1, 2, 0).contiguous()  # [Cin,3,3,Cout]
    src = ops.zero_stuff(dy) if stride == 2 else dy
    dx = ops.conv3x3(src, wd)
    torch.cuda.synchronize()
    xf = x.float().requires_grad_()
    F.conv2d(xf, w.float(), None, stride=stride, padding=1).backward(dy.float().permute(0, 3, 1, 2))
    assert rel_rms(dx, xf.grad.permute(0, 2, 3, 1)) < 1e-2


def test_conv_out_bwd(dev):
    from sliders_b200 import ops

    g = torch.Generator().manual_seed(5)
    B, H, W, C = 2, 24, 16, 320
    w = (torch.randn(4, C, 3, 3, generator=g) / (9 * C) ** 0.5).to(dev, BF)
    for dt in (torch.float32, BF):
        deps = torch.randn(B, 4, H, W, generator=g).to(dev, dt)
        dx = ops.conv_out_bwd(deps, w.permute(0, 2, 3, 1).contiguous())
        torch.cuda.synchronize()
        xf = torch.zeros(B, C, H, W, device=dev, requires_grad=True)
        F.conv2d(xf, w.float(), None, padding=1).backward(deps.float())
        assert rel_rms(dx, xf.grad.permute(0, 2, 3, 1)) < 1e-2


@pytest.mark.parametrize("M,K,N,r", [(4096, 1280, 1280, 4), (1000, 640, 5120, 8), (154, 2048, 640, 4)])
def test_lora_linear_grads(dev, M, K, N, r):
    """d_down, d_up and the rank-r part of dX for y = x W^T + s * (x down^T) up^T (lora.py:108-112)."""
    from sliders_b200 import ops

    g = torch.Generator().manual_seed(M + r)
    x = torch.randn(M, K, generator=g).to(dev, BF)
    down = (torch.randn(r, K, generator=g) / K ** 0.5).to(dev, BF)
    up = (torch.randn(N, r, generator=g) * 0.1).to(dev, BF)
    dy = torch.randn(M, N, generator=g).to(dev, BF)
    s = 0.75
    t = ops.lora_proj(x, down)
    u = ops.lora_proj(dy, up.t().contiguous())
    d_up = torch.empty(N, r, device=dev, dtype=torch.float32)
    d_down = torch.empty(r, K, device=dev, dtype=torch.float32)
    ops.lora_wgrad(dy, t, d_up, False, s)
    ops.lora_wgrad(x, u, d_down, True, s)
    dx = torch.zeros(M, K, device=dev, dtype=BF)
    ops.lora_rank_update(dx, u, down, s)
    torch.cuda.synchronize()
    xf, df, uf = x.float().requires_grad_(), down.float().requires_grad_(), up.float().requires_grad_()
    ((xf @ df.t()) @ uf.t() * s).backward(dy.float())
    assert rel_rms(d_up, uf.grad) < 2e-3
    assert rel_rms(d_down, df.grad) < 2e-3
    assert rel_rms(dx, xf.grad) < 1e-2
    # accumulate flag
    ops.lora_wgrad(dy, t, d_up, False, s, accumulate=True)
    torch.cuda.synchronize()
    assert rel_rms(d_up, 2 * uf.grad) < 2e-3


@pytest.mark.parametrize("stride,C0,C1,r", [(1, 320, 0, 4), (1, 256, 128, 8), (2, 128, 0, 4)])
def test_lora_conv_grads(dev, stride, C0, C1, r):
    from sliders_b200 import ops

    g = torch.Generator().manual_seed(stride + C0 + r)
    B, H, W, Cout = 2, 16, 24, 192
    Cin = C0 + C1
    x = torch.randn(B, H, W, Cin, generator=g).to(dev, BF)
    down = (torch.randn(r, Cin, 3, 3, generator=g) / (9 * Cin) ** 0.5).to(dev, BF)
    up = (torch.randn(Cout, r, generator=g) * 0.1).to(dev, BF)
    Ho, Wo = H // stride, W // stride
    dy = torch.randn(B, Ho, Wo, Cout, generator=g).to(dev, BF)
    s = -1.5
    D = down.permute(0, 2, 3, 1).contiguous()  # [r,3,3,Cin]
    x0 = x[..., :C0].contiguous()
    x1 = x[..., C0:].contiguous() if C1 else None
    t = ops.lora_conv_proj(x0, D, stride, x1=x1)
    dy2 = dy.view(-1, Cout)
    u = ops.lora_proj(dy2, up.t().contiguous())
    d_up = torch.empty(Cout, r, device=dev, dtype=torch.float32)
    ops.lora_wgrad(dy2, t, d_up, False, s)
    d_down = torch.empty(r, 3, 3, Cin, device=dev, dtype=torch.float32)
    ops.lora_conv_wgrad(x0, u, d_down, s, stride, x1=x1)
    dx = torch.zeros(B, H, W, Cin, device=dev, dtype=BF)
    ops.lora_conv_rank_update(dx, u, D, s, stride)
    torch.cuda.synchronize()
    xf = x.float().permute(0, 3, 1, 2).requires_grad_()
    df, uf = down.float().requires_grad_(), up.float().requires_grad_()
    y = F.conv2d(F.conv2d(xf, df, None, stride=stride, padding=1), uf[:, :, None, None]) * s
    y.backward(dy.float().permute(0, 3, 1, 2))
    assert rel_rms(t, F.conv2d(xf, df, None, stride=stride, padding=1).permute(0, 2, 3, 1).reshape(-1, r)) < 2e-3
    assert rel_rms(d_up, uf.grad) < 2e-3
    assert rel_rms(d_down, df.grad.permute(0, 2, 3, 1)) < 2e-3
    assert rel_rms(dx, xf.grad.permute(0, 2, 3, 1)) < 1e-2


def test_adamw_matches_torch_bf16(dev):
    """train_util.py:362-363 builds torch.optim.AdamW over bf16 LoRA weights (config-xl.yaml: lr 2e-4)."""
    from sliders_b200.optim import AdamW

    g = torch.Generator().manual_seed(11)
    shapes = [(4, 1280), (1280, 4), (4, 320, 3, 3), (640, 4, 1, 1), (8, 77)]
    ref_p = [torch.nn.Parameter((torch.randn(*s, generator=g) * 0.1).to(dev, BF)) for s in shapes]
    my_p = [torch.nn.Parameter(p.detach().clone()) for p in ref_p]
    ref_opt = torch.optim.AdamW(ref_p, lr=2e-4)
    my_opt = AdamW(my_p, lr=2e-4)
    for step in range(5):
        for a, b in zip(ref_p, my_p):
            gr = (torch.randn(a.shape, generator=g) * (0.01 if step % 2 else 1.0)).to(dev, BF)
            a.grad = gr.clone()
            b.grad = gr.clone()
        ref_opt.step()
        my_opt.step()
        torch.cuda.synchronize()
        for a, b in zip(ref_p, my_p):
            assert torch.equal(a.detach(), b.detach()), f"step {step}: parameters differ"
    st = ref_opt.state[ref_p[0]]
    assert torch.equal(st["exp_avg"], my_opt.state[my_p[0]]["exp_avg"])
    assert torch.equal(st["exp_avg_sq"], my_opt.state[my_p[0]]["exp_avg_sq"])
