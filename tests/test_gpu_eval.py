"""GPU test of the CSV-driven slider sweep (sliders_b200/eval_sweep.py; eval-scripts/generate_images_xl.py:445-508) on
the tiny SDXL-topology fixture model: slider checkpoint round trip through its file name, the output tree, the same
start noise for every scale of a case, scale 0 == adaptors off, and the `t > start_noise` gating."""
import os

import pytest
import torch

from test_gpu_unet import GOLDEN, build_product, dev  # noqa: F401  (fixture re-export)

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def test_csv_sweep_on_tiny_model(dev, tmp_path):
    from oracle import unet as ounet
    from sliders_b200 import eval_sweep as es
    from sliders_b200 import synthetic
    from sliders_b200.scheduler import create_noise_scheduler
    from sliders_b200.unet import UNet2DConditionModel, UNetConfig

    fx = torch.load(os.path.join(GOLDEN, "tiny_xl.pt"))
    _, trained = build_product(fx, dev)
    slider = tmp_path / "smile_alpha1.0_rank4_noxattn_last.pt"
    trained.save_weights(str(slider), dtype=BF)
    pm = UNet2DConditionModel(UNetConfig.from_dict(ounet.UNetConfig.tiny_xl().__dict__))
    synthetic.init_synthetic_(pm, seed=fx["weight_seed"])
    pm = pm.to(dev, BF).eval().requires_grad_(False)
    pm.use_cuda_graph = True
    net = es.build_network(pm, str(slider), dev, BF)
    assert len(net.unet_loras) == fx["n_lora"]
    assert torch.equal(net.state_dict()[net.unet_loras[0].lora_name + ".lora_up.weight"],
                       trained.state_dict()[net.unet_loras[0].lora_name + ".lora_up.weight"])
    csv_path = tmp_path / "p.csv"
    csv_path.write_text("case_number,prompt,evaluation_seed,concept\n3,a,11,x\n4,b,12,x\n")
    rows = es.read_prompts_csv(str(csv_path))
    g = torch.Generator().manual_seed(0)
    table = {p: (torch.randn(2, 77, 256, generator=g), torch.randn(2, 128, generator=g)) for p in ("a", "b")}
    out_dir = tmp_path / "out"
    sched = create_noise_scheduler("euler")
    n = es.sweep(pm, net, sched, rows, lambda p: table[p], str(out_dir), scales=(-2, 0, 2), num_samples=2,
                 num_inference_steps=6, guidance_scale=5.0, start_noise=750, image_size=256)
    assert n == 2 * 3 * 2
    lat = {s: torch.load(out_dir / str(s) / "3_0.pt") for s in (-2, 0, 2)}
    assert all(torch.isfinite(v).all() and v.shape == (4, 32, 32) for v in lat.values())
    assert not torch.equal(lat[-2], lat[2]) and not torch.equal(lat[0], lat[2])
    assert not torch.equal(lat[0], torch.load(out_dir / "0" / "3_1.pt"))          # second sample: other noise
    # scale 0 is the un-adapted model: the same loop with every adaptor inert
    from sliders_b200 import generate
    sched.set_timesteps(6, device=dev)
    x = es.initial_latents(11, 2, 256, 256, sched.init_noise_sigma, dev, BF)
    ids = torch.tensor([[256., 256, 0, 0, 256, 256]], device=dev).repeat(4, 1)
    pe, ae = table["a"]
    ref = generate.denoise_loop(pm, net, sched, x, pe.to(dev, BF).repeat_interleave(2, 0),
                                ae.to(dev, BF).repeat_interleave(2, 0), ids, num_inference_steps=6, guidance_scale=5.0,
                                scale=2.0, start_noise=-1)          # start_noise below every t: slider never switched on
    assert torch.equal(ref[0].float().cpu(), lat[0])
