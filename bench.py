#!/usr/bin/env python
"""bench.py — SDXL UNet conditioned-forward passes/sec (rank-4 LoRA, 1024 px) on 1..8 B200s.

Workload (BASELINE.json metric): SDXL-base UNet, synthetic seeded weights (no checkpoint is
reachable offline), rank-4 alpha-1 LoRA on the `noxattn` + `c3lier` leaf set (346 adaptors, lora_up != 0,
multiplier 1), latents [B,4,128,128] (1024 px), text embeddings [B,77,2048], pooled [B,1280], time ids
[1024,1024,0,0,1024,1024], timestep 500.  One *step* = one UNet forward over B conditioned passes per GPU
(B = 8: the reference's per-iteration fan-out of 4 predictions x CFG pair, train_lora_xl.py:236-322).

  value     passes/s over all ranks, inputs resident in HBM, CUDA-graph replay of the whole forward,
            timed with CUDA events, max over ranks.
  e2e       the same metric through the public call (`sliders_b200.train_util.predict_noise_xl`, the
            reference's `unet(...)` call site + CFG combine) with HOST (pinned) input buffers: H2D of latents and
            embeddings and D2H of the guided eps are inside the timed region, every step.
  roofline  tensor-bound: algorithmic FLOPs of the dominant kernel (`gemm_kernel`: every Linear / conv as
            tcgen05 GEMM) / its device time, measured live with CUDA events around each launch of one
            instrumented forward; peak = MEASURED_PEAKS.json bf16_tflops_sustained.
  cpu_baseline  the oracle (fp32 PyTorch restatement of the diffusers UNet, oracle/unet.py) on the host cores,
            rank 0, N=1 only, bounded sample.
`--impl reference`: the reference's own (CPU) implementation cannot run here (diffusers is not installable and
the trainers hard-code CUDA/xformers), so this arm times the oracle port on all host cores (kind "port").
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOPS_PER_PASS = 6.761e12          # SURVEY.md §8d, SDXL @128x128 (LoRA r=4 adds 0.0189e12)
FLOPS_LORA_R4 = 0.0189e12
METRIC = "SDXL UNet conditioned-fwd passes/sec (rank-4 LoRA, 1024px)"
UNIT = "passes/s"
LATENT = 128


def host_threads() -> int:
    """CPU threads this process can really use: min(affinity mask, cgroup CPU quota).  os.cpu_count() alone
    reports the host's cores (128 on the GPU boxes) even when the container is throttled to a fraction of them, and
    oversubscribing a quota makes the fp32 oracle several times slower."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            with open(path) as f:
                parts = f.read().split()
            if path.endswith("cpu.max"):
                if parts and parts[0] != "max":
                    n = min(n, max(1, int(float(parts[0]) / float(parts[1]) + 0.5)))
            else:
                quota = int(parts[0])
                with open("/sys/fs/cgroup/cpu/cpu.cfs_period_us") as f2:
                    period = int(f2.read().split()[0])
                if quota > 0:
                    n = min(n, max(1, int(quota / period + 0.5)))
            break
        except (OSError, ValueError, IndexError):
            continue
    env = os.environ.get("SLIDERS_CPU_THREADS")
    if env:
        n = int(env)
    return max(1, min(n, 64))  # MKL/oneDNN fp32 convs stop scaling (and start thrashing) well before 64 threads


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"sustained": float(p["bf16_tflops_sustained"]), "burst": float(p["bf16_tflops"]),
                "hbm": float(p["hbm_gbs"]), "source": "measured (MEASURED_PEAKS.json)"}
    return {"sustained": 1400.0, "burst": 1590.0, "hbm": 6650.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clock / throttle-reason sampler running during the timed region."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) >= 7:
                self.samples.append(parts)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for p in self.samples:
            try:
                sm.append(float(p[0]))
                mx = float(p[1])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------
def build_product(dev, batch, seed=0):
    from sliders_b200 import lora as plora
    from sliders_b200 import synthetic
    from sliders_b200.unet import UNet2DConditionModel, UNetConfig

    with torch.device(dev):
        unet = UNet2DConditionModel(UNetConfig.sdxl()).to(torch.bfloat16)
    synthetic.init_synthetic_(unet, seed=seed + 1)
    unet.requires_grad_(False)
    unet.eval()
    saved = list(plora.DEFAULT_TARGET_REPLACE)
    plora.DEFAULT_TARGET_REPLACE += plora.UNET_TARGET_REPLACE_MODULE_CONV  # c3lier (train_lora_xl.py:50-52)
    try:
        net = plora.LoRANetwork(unet, rank=4, multiplier=1.0, alpha=1.0, train_method="noxattn").to(dev, torch.bfloat16)
    finally:
        del plora.DEFAULT_TARGET_REPLACE[len(saved):]
    synthetic.init_lora_nonzero_(net, seed=seed + 2, up_std=0.02)
    return unet, net


def make_host_inputs(batch, seed=0, pin=True):
    g = torch.Generator().manual_seed(seed)
    lat = torch.randn(batch, 4, LATENT, LATENT, generator=g)
    ehs = torch.randn(batch, 77, 2048, generator=g).to(torch.bfloat16)
    pooled = torch.randn(batch, 1280, generator=g).to(torch.bfloat16)
    tids = torch.tensor([[1024., 1024., 0., 0., 1024., 1024.]] * batch)
    ts = [lat, ehs, pooled, tids]
    if pin and torch.cuda.is_available():
        ts = [t.pin_memory() for t in ts]
    return ts


def run_cpu_oracle(state_dict_cpu_f32, lora_fold, n_timed, threads, batch=1, seed=0):
    """Times the fp32 oracle UNet forward on the host cores.  Returns (passes_per_s, eps) for parity."""
    from oracle import unet as ounet

    torch.set_num_threads(threads)
    with torch.device("meta"):
        om = ounet.UNet2DConditionModel(ounet.UNetConfig.sdxl())
    om = om.to_empty(device="cpu")
    om.load_state_dict(state_dict_cpu_f32, assign=True)
    om.eval()
    if lora_fold is not None:
        lora_fold(om)
    lat, ehs, pooled, tids = make_host_inputs(batch, seed=seed, pin=False)
    added = {"text_embeds": pooled.float(), "time_ids": tids}
    with torch.no_grad():
        t0 = time.time()
        eps = om(lat.to(torch.bfloat16).float(), 500, ehs.float(), added_cond_kwargs=added).sample  # warm-up
        warm = time.time() - t0
        # n_timed is a time budget in seconds when negative: run as many timed calls as fit (1..4)
        if n_timed < 0:
            n_timed = int(max(1, min(4, (-n_timed) // max(warm, 1e-3))))
        times = []
        for _ in range(n_timed):
            t0 = time.time()
            eps = om(lat.to(torch.bfloat16).float(), 500, ehs.float(), added_cond_kwargs=added).sample
            times.append(time.time() - t0)
    dt = statistics.mean(times) if times else warm
    return batch / dt, eps, len(times)


def fold_lora_into(net_state, scales):
    """Returns f(oracle_model) adding s * up @ down to every adapted leaf (algebraically the LoRA hook)."""

    def fold(om):
        mods = {("lora_unet_" + n.replace(".", "_")): m for n, m in om.named_modules()}
        with torch.no_grad():
            for name, s in scales.items():
                up = net_state[name + ".lora_up.weight"].float()
                down = net_state[name + ".lora_down.weight"].float()
                delta = torch.einsum("or,rikl->oikl", up[:, :, 0, 0], down) if down.dim() == 4 else up @ down
                mods[name].weight.add_(delta * s)

    return fold


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=8, help="conditioned passes per GPU per step")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=25.0)
    ap.add_argument("--no-train", action="store_true", help="skip the text-slider training-iteration timing")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    peaks = load_peaks()
    config = {"workload": f"sdxl_unet_fwd_{args.batch}passes_per_gpu_1024px_lora_r4_noxattn_c3lier",
              "latent": [args.batch, 4, LATENT, LATENT], "timestep": 500, "lora": "rank4 alpha1 noxattn+c3lier (346)",
              "parallelism": f"dp{world} (independent passes, weights replicated, no data-path collective)",
              "l2": "working set (5.1 GB bf16 weights + activations) >> 126 MB L2; no flush needed"}

    # ------------------------------------------------------------------ reference arm (CPU, rank 0 only)
    if args.impl == "reference":
        if rank != 0:
            return
        threads = host_threads()
        from sliders_b200 import synthetic
        from oracle import unet as ounet
        with torch.device("meta"):
            om = ounet.UNet2DConditionModel(ounet.UNetConfig.sdxl())
        sd = {k: synthetic.synthetic_tensor(k, p.shape, 1, "cpu") for k, p in om.named_parameters()}
        # bounded sample: one conditioned pass per step (same model / resolution / inputs as the GPU arm)
        torch.set_num_threads(threads)
        om = om.to_empty(device="cpu")
        om.load_state_dict(sd, assign=True)
        om.eval()
        lat, ehs, pooled, tids = make_host_inputs(1, pin=False)
        added = {"text_embeds": pooled.float(), "time_ids": tids}
        times = []
        budget_s = 200.0
        t_start = time.time()
        with torch.no_grad():
            for i in range(args.warmup + args.steps):
                t0 = time.time()
                om(lat.to(torch.bfloat16).float(), 500, ehs.float(), added_cond_kwargs=added)
                dt = time.time() - t0
                if i >= args.warmup:
                    times.append(dt)
                if time.time() - t_start > budget_s and len(times) >= 1:
                    break
        ms = 1e3 * statistics.mean(times)
        v = 1e3 / ms
        sample = (f"{len(times)} timed single-pass forwards (B=1) of the same SDXL@128x128 workload, fp32, "
                  f"{threads} threads; LoRA delta omitted (0.3% of FLOPs)")
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
                          "steps": len(times), "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                          "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                          "config": config,
                          "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                           "sample": sample},
                          "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    # ------------------------------------------------------------------ our arm
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the sliders_b200 path has no CPU fallback "
                         "(use --impl reference for the CPU oracle timing)")
    import torch.distributed as dist
    from sliders_b200 import ops, train_util
    from sliders_b200.scheduler import create_noise_scheduler

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch
    unet, net = build_product(dev, B)
    lat_h, ehs_h, pooled_h, tids_h = make_host_inputs(B, seed=rank)
    lat = lat_h.to(dev)
    ehs = ehs_h.to(dev)
    added = {"text_embeds": pooled_h.to(dev), "time_ids": tids_h.to(dev)}

    # launches per forward (eager, counted by the op wrappers) + per-kernel roofline pass
    net.__enter__()  # multiplier = 1 (lora.py:252-254)
    with torch.no_grad():
        unet(lat, 500, ehs, added_cond_kwargs=added)  # packs weights, fills TMA descriptor cache
        torch.cuda.synchronize()
        ops.launch_count = 0
        ops.profile_log = []
        unet(lat, 500, ehs, added_cond_kwargs=added)
        torch.cuda.synchronize()
        log, ops.profile_log = ops.profile_log, None
        launches_per_fwd = ops.launch_count
    per_kind = {}
    for name, fl, e0, e1 in log:
        d = per_kind.setdefault(name, [0.0, 0.0, 0])
        d[0] += e0.elapsed_time(e1)
        d[1] += fl
        d[2] += 1
    gemm_ms = per_kind.get("gemm", [0, 0, 0])[0] + per_kind.get("conv3x3", [0, 0, 0])[0]
    gemm_fl = per_kind.get("gemm", [0, 0, 0])[1] + per_kind.get("conv3x3", [0, 0, 0])[1]
    gemm_n = per_kind.get("gemm", [0, 0, 0])[2] + per_kind.get("conv3x3", [0, 0, 0])[2]
    total_ms_eager = sum(v[0] for v in per_kind.values())
    achieved = gemm_fl / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    # DRAM traffic of the dominant kernel: from the committed ncu capture of the same workload (profiles/), per launch
    traffic, traffic_note = None, None
    try:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r01_dram_b8.json")) as f:
            dj = json.load(f)["per_kernel"]["gemm_kernel"]
        if B == 8:
            traffic = dj["dram_read_bytes_per_launch"] + dj["dram_write_bytes_per_launch"]
            traffic_note = ("bytes per gemm_kernel launch (mean over the 493 launches of one 8-pass forward), ncu "
                            "dram__bytes_read.sum + dram__bytes_write.sum, profiles/r01_dram_b8_ncu.csv; algorithmic "
                            "A + W + out (+ residual) bytes per launch: 105.5e6")
    except (OSError, KeyError, ValueError):
        pass
    roofline = {"bound": "tensor", "kernel": "gemm_kernel (tcgen05 GEMM / implicit-GEMM conv)",
                "achieved": achieved, "peak": peaks["sustained"], "unit": "TFLOP/s",
                "frac": achieved / peaks["sustained"], "traffic": traffic, "traffic_note": traffic_note,
                "peak_source": peaks["source"],
                "launches": gemm_n, "avg_launch_us": 1e3 * gemm_ms / max(gemm_n, 1),
                "share_of_step": gemm_ms / total_ms_eager if total_ms_eager else None,
                "breakdown_ms": {k: round(v[0], 3) for k, v in sorted(per_kind.items())},
                "whole_forward_frac": None}

    # ---- value: graph replay, inputs resident
    unet.use_cuda_graph = True
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.no_grad():
        for _ in range(args.warmup):
            out = unet(lat, 500, ehs, added_cond_kwargs=added).sample
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        sampler = ClockSampler(local_rank)
        sampler.start()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(args.steps):
            out = unet(lat, 500, ehs, added_cond_kwargs=added).sample
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        clocks = sampler.stop()
    ms_total = e0.elapsed_time(e1)
    t = torch.tensor([ms_total], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = t.item() / args.steps
    value = world * B / (ms_step * 1e-3)
    roofline["whole_forward_frac"] = (value / world) * (FLOPS_PER_PASS + FLOPS_LORA_R4) / 1e12 / peaks["sustained"]

    # ---- e2e: public API, host buffers, H2D + D2H inside the timed region
    sched = create_noise_scheduler("ddim")
    sched.set_timesteps(1000)
    eps_host = torch.empty(B // 2 if B > 1 else 1, 4, LATENT, LATENT, dtype=torch.float32).pin_memory()
    half = max(B // 2, 1)

    def e2e_step():
        # CFG-pair call exactly like the trainers': latents [half], embeddings [2*half] (uncond ; cond)
        l_d = lat_h[:half].to(dev, non_blocking=True)
        e_d = ehs_h[:2 * half].to(dev, non_blocking=True)
        p_d = pooled_h[:2 * half].to(dev, non_blocking=True)
        t_d = tids_h[:2 * half].to(dev, non_blocking=True)
        eps = train_util.predict_noise_xl(unet, sched, 500, l_d, e_d, p_d, t_d, guidance_scale=3.0)
        eps_host[:half].copy_(eps, non_blocking=True)

    with torch.no_grad():
        for _ in range(args.warmup):
            e2e_step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0.record()
        for _ in range(args.steps):
            e2e_step()
        e1.record()
        torch.cuda.synchronize()
    t2 = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_ms = t2.item() / args.steps
    e2e_passes = 2 * half
    h2d = (lat_h[:half].numel() * 4 + ehs_h[:2 * half].numel() * 2 + pooled_h[:2 * half].numel() * 2
           + tids_h[:2 * half].numel() * 4)
    d2h = eps_host[:half].numel() * 4
    e2e = {"value": world * e2e_passes / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms,
           "api": "sliders_b200.train_util.predict_noise_xl (CFG pair, guidance 3) with pinned host buffers"}
    net.__exit__(None, None, None)

    # ---- training path: whole text-slider iterations (train_lora_xl.py:162-356) through sliders_b200.trainer —
    # partial denoise (25 of 50 DDIM steps, CFG pair, graph replays) + 3 frozen predictions + the grad-carrying one +
    # backward-to-LoRA + fused AdamW.  Under torchrun the four predictions are sharded one per rank (BASELINE config 3).
    train = None
    if not args.no_train:
        from sliders_b200 import trainer
        net.requires_grad_(True)
        opt = train_util.get_optimizer("AdamW")(net.prepare_optimizer_params(), lr=2e-4)
        gtr = torch.Generator().manual_seed(77)
        mk = lambda: trainer.PromptEmbedsXL(torch.randn(1, 77, 2048, generator=gtr).to(dev, torch.bfloat16),
                                            torch.randn(1, 1280, generator=gtr).to(dev, torch.bfloat16))
        unc, tgt, pos = mk(), mk(), mk()
        pair = trainer.PromptEmbedsPair(torch.nn.MSELoss(), tgt, pos, unc, unc,
                                        trainer.PromptSettings(guidance_scale=4.0, resolution=1024, batch_size=1,
                                                               action="enhance"))
        tsched = create_noise_scheduler("ddim")
        n_it = 2
        ops.launch_count = 0
        for it in range(1 + n_it):
            if it == 1:
                torch.cuda.synchronize()
                if world > 1:
                    dist.barrier()
                ops.launch_count = 0
                e0.record()
            loss = trainer.text_slider_step_xl(unet, net, tsched, opt, None, pair, timesteps_to=25, device=dev,
                                               weight_dtype=torch.bfloat16,
                                               generator=torch.Generator().manual_seed(1000 + it))
        e1.record()
        torch.cuda.synchronize()
        t3 = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t3, op=dist.ReduceOp.MAX)
        it_ms = t3.item() / n_it
        passes = 2 * (25 + 4)  # CFG pairs: 25 denoise steps + positive / neutral / unconditional / target
        train = {"what": "text-slider iteration, SDXL 1024 px, batch 1, rank-4 LoRA (train_lora_xl.py:162-356): 25 DDIM "
                         "denoise steps (guidance 3) + 4 CFG-pair predictions + backward-to-LoRA + AdamW(692 tensors)",
                 "ms_per_iteration": it_ms, "iterations_timed": n_it, "passes_per_iteration": passes,
                 "passes_per_s": passes / (it_ms * 1e-3), "loss": float(loss),
                 "eager_launches_per_iteration": ops.launch_count // n_it,
                 "sharding": ("single GPU" if world == 1 else
                              f"denoise CFG-split over rank parity (1 all-gather of 64 KiB per step), target prediction on rank "
                              f"{world - 1}, frozen predictions over ranks 0..{max(world - 2, 0)}, 1 LoRA-grad all-reduce")}
        net.requires_grad_(False)
        opt = None
    unet.use_cuda_graph = False

    # ---- CPU baseline (rank 0, N == 1 only)
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = host_threads()
        sd = {k: v.detach().float().cpu() for k, v in unet.state_dict().items()}
        nsd = {k: v.detach().float().cpu() for k, v in net.state_dict().items()}
        scales = {l.lora_name: float(l.scale) for l in net.unet_loras}
        t0 = time.time()
        # one untimed warm-up, then as many timed single-pass forwards as fit args.cpu_seconds (1..4)
        pps, eps_cpu, n_timed = run_cpu_oracle(sd, fold_lora_into(nsd, scales), -args.cpu_seconds, threads,
                                               batch=1, seed=0)
        # parity of the kernel path against this very oracle run (same weights, same inputs)
        with torch.no_grad(), net:
            lat0, ehs0, pooled0, tids0 = make_host_inputs(1, seed=0, pin=False)
            got = unet(lat0.to(dev).to(torch.bfloat16), 500, ehs0.to(dev),
                       added_cond_kwargs={"text_embeds": pooled0.to(dev), "time_ids": tids0.to(dev)}).sample
        rel = ((got.float().cpu() - eps_cpu).norm() / eps_cpu.norm()).item()
        cpu_baseline = {"value": pps, "unit": UNIT, "cores": threads, "kind": "port",
                        "sample": f"{n_timed} timed (+1 warm-up) single-pass fp32 forwards of the same SDXL@128x128 LoRA-r4 workload "
                                  f"(oracle/unet.py), {threads} torch threads, {time.time() - t0:.0f}s wall incl. build",
                        "eps_rel_rms_kernels_vs_this_oracle": rel}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": config,
                "clocks": clocks, "e2e": e2e, "gpu_launches": launches_per_fwd * args.steps,
                "launches_per_step": launches_per_fwd, "roofline": roofline, "cpu_baseline": cpu_baseline,
                "train": train}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
