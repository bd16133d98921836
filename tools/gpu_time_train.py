"""Device timings of the training path at BASELINE size (SDXL, 1024 px, rank-4 LoRA on 346 leaves):
grad-carrying prediction forward, its backward, AdamW, and whole text-slider iterations.

    python tools/gpu_time_train.py [iters]
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sliders_b200 import lora as plora, ops, synthetic, train_util, trainer  # noqa: E402
from sliders_b200.scheduler import create_noise_scheduler  # noqa: E402
from sliders_b200.unet import UNet2DConditionModel, UNetConfig  # noqa: E402

BF = torch.bfloat16


def ev():
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    dev = torch.device("cuda:0")
    with torch.device(dev):
        pm = UNet2DConditionModel(UNetConfig.sdxl()).to(BF)
    synthetic.init_synthetic_(pm, seed=1)
    pm.requires_grad_(False)
    saved = list(plora.DEFAULT_TARGET_REPLACE)
    plora.DEFAULT_TARGET_REPLACE += plora.UNET_TARGET_REPLACE_MODULE_CONV
    net = plora.LoRANetwork(pm, rank=4, multiplier=1.0, alpha=1.0, train_method="noxattn").to(dev, BF)
    del plora.DEFAULT_TARGET_REPLACE[len(saved):]
    synthetic.init_lora_nonzero_(net, seed=2, up_std=0.02)
    net.requires_grad_(True)
    opt = train_util.get_optimizer("AdamW")(net.prepare_optimizer_params(), lr=2e-4)
    sched = create_noise_scheduler("ddim")
    g = torch.Generator().manual_seed(0)
    mk = lambda: trainer.PromptEmbedsXL(torch.randn(1, 77, 2048, generator=g).to(dev, BF),
                                        torch.randn(1, 1280, generator=g).to(dev, BF))
    unc, tgt, pos = mk(), mk(), mk()
    pair = trainer.PromptEmbedsPair(torch.nn.MSELoss(), tgt, pos, unc, unc,
                                    trainer.PromptSettings(guidance_scale=4.0, resolution=1024, batch_size=1,
                                                           action="enhance"))
    # ---- pieces
    sched.set_timesteps(1000)
    lat = torch.randn(1, 4, 128, 128, generator=g).to(dev, BF)
    ehs = train_util.concat_embeddings(unc.text_embeds, tgt.text_embeds, 1)
    pooled = train_util.concat_embeddings(unc.pooled_embeds, tgt.pooled_embeds, 1)
    tids = train_util.get_add_time_ids(1024, 1024, dtype=BF).to(dev).repeat(2, 1)
    goal = torch.randn(1, 4, 128, 128, generator=g).to(dev, BF)
    for rep in range(3):
        torch.cuda.synchronize()
        n0 = ops.launch_count
        e0 = ev()
        with net:
            pred = train_util.predict_noise_xl(pm, sched, 500, lat, ehs, pooled, tids, guidance_scale=1)
        e1 = ev()
        n1 = ops.launch_count
        loss = torch.nn.functional.mse_loss(pred, goal)
        loss.backward()
        e2 = ev()
        n2 = ops.launch_count
        opt.step()
        e3 = ev()
        opt.zero_grad()
        torch.cuda.synchronize()
        print(f"PIECES rep{rep}: train-forward (CFG pair) {e0.elapsed_time(e1):.2f} ms [{n1 - n0} launches] | backward "
              f"{e1.elapsed_time(e2):.2f} ms [{n2 - n1} launches] | AdamW(692 tensors) {e2.elapsed_time(e3):.3f} ms | "
              f"peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
    with torch.no_grad():
        for rep in range(2):
            e0 = ev()
            train_util.predict_noise_xl(pm, sched, 500, lat, ehs, pooled, tids, guidance_scale=1)
            e1 = ev()
            torch.cuda.synchronize()
            print(f"PIECES inference forward (CFG pair, eager): {e0.elapsed_time(e1):.2f} ms")
    # ---- whole iterations (train_lora_xl.py:162-356), fixed timesteps_to = 25 (the mean of randint(1, 50))
    pm.use_cuda_graph = os.environ.get("SB200_TRAIN_GRAPH", "1") == "1"
    for it in range(iters):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0 = ev()
        loss = trainer.text_slider_step_xl(pm, net, sched, opt, None, pair, timesteps_to=25, device=dev,
                                           weight_dtype=BF)
        e1 = ev()
        torch.cuda.synchronize()
        print(f"ITER {it}: text-slider iteration (25 denoise steps + 4 predictions + backward + AdamW) "
              f"{e0.elapsed_time(e1):.1f} ms device, {1e3 * (time.perf_counter() - t0):.1f} ms wall, "
              f"loss {float(loss):.5f}")


if __name__ == "__main__" and not (len(sys.argv) > 1 and sys.argv[1] in ("prof", "ncu")):
    main()


def profile_iteration():
    """Host-side profile of one iteration + device time of the denoise segment alone."""
    import cProfile
    import pstats

    dev = torch.device("cuda:0")
    with torch.device(dev):
        pm = UNet2DConditionModel(UNetConfig.sdxl()).to(BF)
    synthetic.init_synthetic_(pm, seed=1)
    pm.requires_grad_(False)
    saved = list(plora.DEFAULT_TARGET_REPLACE)
    plora.DEFAULT_TARGET_REPLACE += plora.UNET_TARGET_REPLACE_MODULE_CONV
    net = plora.LoRANetwork(pm, rank=4, multiplier=1.0, alpha=1.0, train_method="noxattn").to(dev, BF)
    del plora.DEFAULT_TARGET_REPLACE[len(saved):]
    synthetic.init_lora_nonzero_(net, seed=2, up_std=0.02)
    sched = create_noise_scheduler("ddim")
    g = torch.Generator().manual_seed(0)
    lat = torch.randn(1, 4, 128, 128, generator=g).to(dev, BF)
    ehs = torch.randn(2, 77, 2048, generator=g).to(dev, BF)
    pooled = torch.randn(2, 1280, generator=g).to(dev, BF)
    tids = train_util.get_add_time_ids(1024, 1024, dtype=BF).to(dev).repeat(2, 1)
    pm.use_cuda_graph = True
    sched.set_timesteps(50, device=dev)
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0 = ev()
        with net:
            train_util.diffusion_xl(pm, sched, lat, ehs, pooled, tids, guidance_scale=3, total_timesteps=25)
        e1 = ev()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        print(f"DENOISE rep{rep}: 25 CFG-pair steps {e0.elapsed_time(e1):.1f} ms device ({e0.elapsed_time(e1) / 25:.2f} "
              f"ms/step), host enqueue {1e3 * (t1 - t0):.1f} ms")
    pr = cProfile.Profile()
    pr.enable()
    with net:
        train_util.diffusion_xl(pm, sched, lat, ehs, pooled, tids, guidance_scale=3, total_timesteps=25)
    torch.cuda.synchronize()
    pr.disable()
    import io
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(18)
    print("\n".join("PROF " + l for l in s.getvalue().splitlines()[:40]))


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "prof":
    profile_iteration()


def ncu_mode():
    """One training forward + backward (SDXL CFG pair) between cudaProfilerStart/Stop:
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv ... python tools/gpu_time_train.py ncu"""
    dev = torch.device("cuda:0")
    with torch.device(dev):
        pm = UNet2DConditionModel(UNetConfig.sdxl()).to(BF)
    synthetic.init_synthetic_(pm, seed=1)
    pm.requires_grad_(False)
    saved = list(plora.DEFAULT_TARGET_REPLACE)
    plora.DEFAULT_TARGET_REPLACE += plora.UNET_TARGET_REPLACE_MODULE_CONV
    net = plora.LoRANetwork(pm, rank=4, multiplier=1.0, alpha=1.0, train_method="noxattn").to(dev, BF)
    del plora.DEFAULT_TARGET_REPLACE[len(saved):]
    synthetic.init_lora_nonzero_(net, seed=2, up_std=0.02)
    net.requires_grad_(True)
    sched = create_noise_scheduler("ddim")
    sched.set_timesteps(1000)
    g = torch.Generator().manual_seed(0)
    lat = torch.randn(1, 4, 128, 128, generator=g).to(dev, BF)
    ehs = torch.randn(2, 77, 2048, generator=g).to(dev, BF)
    pooled = torch.randn(2, 1280, generator=g).to(dev, BF)
    tids = train_util.get_add_time_ids(1024, 1024, dtype=BF).to(dev).repeat(2, 1)
    goal = torch.randn(1, 4, 128, 128, generator=g).to(dev, BF)
    for rep in range(2):
        if rep == 1:
            torch.cuda.synchronize()
            torch.cuda.profiler.start()
        with net:
            pred = train_util.predict_noise_xl(pm, sched, 500, lat, ehs, pooled, tids, guidance_scale=1)
        torch.nn.functional.mse_loss(pred, goal).backward()
        torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "ncu":
    ncu_mode()
