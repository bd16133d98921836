"""One optimisation step of the reference trainers, with the UNet forward / backward in sliders_b200 kernels:

  text_slider_step_xl   trainscripts/textsliders/train_lora_xl.py:162-356   
  text_slider_step      trainscripts/textsliders/train_lora.py:155-309      (SD1.x)
  image_slider_step_xl  trainscripts/imagesliders/train_lora-scale-xl.py:178-384 (a9; latents in place of VAE-encoded
                        image pairs — the VAE and the text encoders are off the denoise path)
  image_slider_step     trainscripts/imagesliders/train_lora-scale.py:185-330  (SD1.x)

Prompt embeddings are inputs (the text encoders run once, before the loop: train_lora_xl.py:100-151).  The dataflow,
the order of the UNet calls, where `with network:` is open, what carries grad, and the DDIM bookkeeping
(`set_timesteps(max_denoising_steps)` -> partial denoise -> `set_timesteps(1000)` -> `current_timestep`) follow the
reference line by line; `flush()` (gc + empty_cache, :356) is deliberately not replicated — the caching allocator
keeping its blocks is what makes the next iteration launch-bound rather than malloc-bound.
The two unused predictions of the image-slider loop (`high_latents`, `low_latents`, dead code in the reference,
SURVEY.md §8 a9) are skipped unless `reference_dead_code=True`.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch

import torch.distributed as dist

from . import parallel, train_util


@dataclass
class PromptEmbedsXL:
    """prompt_util.py:28-35 — (text_embeds [1,77,2048], pooled_embeds [1,1280])."""
    text_embeds: torch.Tensor
    pooled_embeds: torch.Tensor


@dataclass
class PromptSettings:
    """prompt_util.py:38-68 (fields used inside the training loop)."""
    guidance_scale: float = 1.0
    resolution: int = 512
    dynamic_resolution: bool = False
    batch_size: int = 1
    dynamic_crops: bool = False
    action: str = "erase"


class PromptEmbedsPair:
    """prompt_util.py:71-148: the four embeddings of one slider prompt plus the erase / enhance objective."""

    def __init__(self, loss_fn, target, positive, unconditional, neutral, settings: PromptSettings):
        self.loss_fn = loss_fn
        self.target, self.positive, self.unconditional, self.neutral = target, positive, unconditional, neutral
        self.guidance_scale = settings.guidance_scale
        self.resolution = settings.resolution
        self.dynamic_resolution = settings.dynamic_resolution
        self.batch_size = settings.batch_size
        self.dynamic_crops = settings.dynamic_crops
        self.action = settings.action

    def loss(self, target_latents, positive_latents, unconditional_latents, neutral_latents):
        if self.action == "erase":      # prompt_util.py:108-121
            return self.loss_fn(target_latents,
                                neutral_latents - self.guidance_scale * (positive_latents - unconditional_latents))
        if self.action == "enhance":    # :123-135
            return self.loss_fn(target_latents,
                                neutral_latents + self.guidance_scale * (positive_latents - unconditional_latents))
        raise ValueError("action must be erase or enhance")


def _world(group=None):
    """(world size, rank) of `group` (None = default group); `group=False` forces the single-process path."""
    if group is not False and dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def _global_rank(group_rank: int, group=None) -> int:
    return dist.get_global_rank(group, group_rank) if group is not None else group_rank


def _xl_inputs(pair: PromptEmbedsPair, which: PromptEmbedsXL, add_time_ids):
    bs = pair.batch_size
    return dict(
        text_embeddings=train_util.concat_embeddings(pair.unconditional.text_embeds, which.text_embeds, bs),
        add_text_embeddings=train_util.concat_embeddings(pair.unconditional.pooled_embeds, which.pooled_embeds, bs),
        add_time_ids=train_util.concat_embeddings(add_time_ids, add_time_ids, bs))


def text_slider_step_xl(unet, network, noise_scheduler, optimizer, lr_scheduler, prompt_pair: PromptEmbedsPair, *,
                        max_denoising_steps: int = 50, timesteps_to: Optional[int] = None, device=None,
                        weight_dtype=torch.bfloat16, generator: Optional[torch.Generator] = None,
                        group=None) -> torch.Tensor:
    """train_lora_xl.py:162-347.  Returns the (detached) loss.

    Under torch.distributed (one process per GPU) the four conditioned predictions are sharded one per rank
    (BASELINE config 3): the three frozen
    predictions are broadcast to the rank that owns the grad-carrying `target` prediction, that rank back-propagates,
    and ONE all-reduce of the flat LoRA gradient buffer makes every replica take the same AdamW step.  The serial
    partial denoise is CFG-split: even ranks run the unconditional sample of each step, odd ranks the conditional one,
    with one 64 KiB all-gather per step (train_util._denoise_loop)."""
    device = device or unet.device
    world, rank = _world(group)
    with torch.no_grad():
        noise_scheduler.set_timesteps(max_denoising_steps, device=device)
        optimizer.zero_grad()
        # host-RNG draws in the reference's order (:177-203): step count, bucketed resolution, noise, crop ids
        if timesteps_to is None:
            timesteps_to = torch.randint(1, max_denoising_steps, (1,)).item()          # :177-179
        height = width = prompt_pair.resolution
        if prompt_pair.dynamic_resolution:
            height, width = train_util.get_random_resolution_in_bucket(prompt_pair.resolution)
        latents = train_util.get_initial_latents(noise_scheduler, prompt_pair.batch_size, height, width, 1,
                                                 generator=generator).to(device, dtype=weight_dtype)
        add_time_ids = train_util.get_add_time_ids(height, width, dynamic_crops=prompt_pair.dynamic_crops,
                                                   dtype=torch.float32)
        if world > 1:
            # replicas follow group rank 0's draws, whatever their own RNG state: one small broadcast carries the step
            # count, the resolution and the six time ids, a second one the noise (its shape depends on the first)
            vals = parallel.sync_draws([timesteps_to, height, width] + add_time_ids.flatten().tolist(), device, group)
            timesteps_to, sh, sw = int(vals[0]), int(vals[1]), int(vals[2])
            add_time_ids = torch.tensor([vals[3:9]], dtype=torch.float32)
            if (sh, sw) != (height, width):
                height, width = sh, sw
                latents = torch.empty(prompt_pair.batch_size, train_util.UNET_IN_CHANNELS,
                                      height // train_util.VAE_SCALE_FACTOR, width // train_util.VAE_SCALE_FACTOR,
                                      device=device, dtype=weight_dtype)
            dist.broadcast(latents, src=_global_rank(0, group), group=group)
        add_time_ids = add_time_ids.to(device, dtype=weight_dtype)
        with network:                                                                   # :205-227
            denoised_latents = train_util.diffusion_xl(
                unet, noise_scheduler, latents, **_xl_inputs(prompt_pair, prompt_pair.target, add_time_ids),
                start_timesteps=0, total_timesteps=timesteps_to, guidance_scale=3,
                **({"cfg_split_group": group} if _world(group)[0] > 1 else {}))
        noise_scheduler.set_timesteps(1000)
        current_timestep = noise_scheduler.timesteps[int(timesteps_to * 1000 / max_denoising_steps)]
        # outside `with network:` the adaptors are inert (:236-297)
        # the grad-carrying prediction (forward + backward, ~4x a frozen one) gets a rank of its own
        owner = {"target": world - 1}
        for i, name in enumerate(("positive", "neutral", "unconditional")):
            owner[name] = i % max(world - 1, 1)
        preds = {}
        for name in ("positive", "neutral", "unconditional"):
            if owner[name] == rank:
                preds[name] = train_util.predict_noise_xl(
                    unet, noise_scheduler, current_timestep, denoised_latents,
                    **_xl_inputs(prompt_pair, getattr(prompt_pair, name), add_time_ids),
                    guidance_scale=1).to(device, dtype=weight_dtype)
        if world > 1:  # one condition per GPU: the three frozen predictions travel to the target rank
            for name in ("positive", "neutral", "unconditional"):
                if owner[name] != rank:
                    preds[name] = torch.empty_like(denoised_latents, dtype=weight_dtype)
                dist.broadcast(preds[name], src=_global_rank(owner[name], group), group=group)
    loss = torch.zeros((), device=device)
    if owner["target"] == rank:
        with network:                                                                   # :299-322, grad on
            target_latents = train_util.predict_noise_xl(
                unet, noise_scheduler, current_timestep, denoised_latents,
                **_xl_inputs(prompt_pair, prompt_pair.target, add_time_ids),
                guidance_scale=1).to(device, dtype=weight_dtype)
        loss = prompt_pair.loss(target_latents=target_latents, positive_latents=preds["positive"],
                                neutral_latents=preds["neutral"], unconditional_latents=preds["unconditional"])
        loss.backward()                                                                 # :345
    if world > 1:
        # one all-reduce of the flat LoRA-gradient buffer (ranks without the graph contribute zeros), so every
        # replica applies the identical optimizer step
        parallel.allreduce_lora_grads([p for g in optimizer.param_groups for p in g["params"]], group=group)
        loss = loss.detach().float().clone()
        dist.broadcast(loss, src=_global_rank(owner["target"], group), group=group)
    optimizer.step()
    if lr_scheduler is not None:
        lr_scheduler.step()
    return loss.detach()


def text_slider_step(unet, network, noise_scheduler, optimizer, lr_scheduler, prompt_pair: PromptEmbedsPair, *,
                     max_denoising_steps: int = 50, timesteps_to: Optional[int] = None, device=None,
                     weight_dtype=torch.bfloat16, generator: Optional[torch.Generator] = None) -> torch.Tensor:
    """train_lora.py:155-300 (SD1.x: embeddings are plain tensors [1,77,768], no added conditioning)."""
    device = device or unet.device
    bs = prompt_pair.batch_size
    emb = lambda which: train_util.concat_embeddings(prompt_pair.unconditional, which, bs)
    with torch.no_grad():
        noise_scheduler.set_timesteps(max_denoising_steps, device=device)
        optimizer.zero_grad()
        if timesteps_to is None:
            timesteps_to = torch.randint(1, max_denoising_steps, (1,)).item()
        height = width = prompt_pair.resolution
        latents = train_util.get_initial_latents(noise_scheduler, bs, height, width, 1,
                                                 generator=generator).to(device, dtype=weight_dtype)
        with network:
            denoised_latents = train_util.diffusion(unet, noise_scheduler, latents, emb(prompt_pair.target),
                                                    start_timesteps=0, total_timesteps=timesteps_to, guidance_scale=3)
        noise_scheduler.set_timesteps(1000)
        current_timestep = noise_scheduler.timesteps[int(timesteps_to * 1000 / max_denoising_steps)]
        preds = {name: train_util.predict_noise(unet, noise_scheduler, current_timestep, denoised_latents,
                                                emb(getattr(prompt_pair, name)), guidance_scale=1
                                                ).to(device, dtype=weight_dtype)
                 for name in ("positive", "neutral", "unconditional")}
    with network:
        target_latents = train_util.predict_noise(unet, noise_scheduler, current_timestep, denoised_latents,
                                                  emb(prompt_pair.target), guidance_scale=1
                                                  ).to(device, dtype=weight_dtype)
    loss = prompt_pair.loss(target_latents=target_latents, positive_latents=preds["positive"],
                            neutral_latents=preds["neutral"], unconditional_latents=preds["unconditional"])
    loss.backward()
    optimizer.step()
    if lr_scheduler is not None:
        lr_scheduler.step()
    return loss.detach()


def image_slider_step_xl(unet, network, noise_scheduler, optimizer, lr_scheduler, prompt_pair: PromptEmbedsPair,
                         latents_low: torch.Tensor, latents_high: torch.Tensor, scale_to_look: float, *,
                         max_denoising_steps: int = 50, timesteps_to: Optional[int] = None, device=None,
                         weight_dtype=torch.bfloat16, seed: Optional[int] = None,
                         reference_dead_code: bool = False, group=None):
    """train_lora-scale-xl.py:178-375 with the image pair already in latent space ([bs,4,h,w], scaled by the VAE
    factor as `get_noisy_image` does, imagesliders/train_util.py:201-235).  Two grad-carrying predictions (+scale on
    the `high` sample, -scale on the `low` one), two `backward()` calls accumulating into .grad, one optimizer step.

    Under torch.distributed (BASELINE config 4) the two grad-carrying predictions run on rank parity (even ranks: high /
    +scale, odd ranks: low / -scale) and, when the batch divides, each parity group splits the batch; every rank
    back-propagates its share of `loss_high + loss_low` and ONE all-reduce of the flat LoRA gradient precedes the
    identical AdamW step."""
    return _image_slider_step(True, unet, network, noise_scheduler, optimizer, lr_scheduler, prompt_pair, latents_low,
                              latents_high, scale_to_look, max_denoising_steps, timesteps_to, device, weight_dtype, seed,
                              reference_dead_code, group)


def image_slider_step(unet, network, noise_scheduler, optimizer, lr_scheduler, prompt_pair: PromptEmbedsPair,
                      latents_low: torch.Tensor, latents_high: torch.Tensor, scale_to_look: float, *,
                      max_denoising_steps: int = 50, timesteps_to: Optional[int] = None, device=None,
                      weight_dtype=torch.bfloat16, seed: Optional[int] = None, reference_dead_code: bool = False,
                      group=None):
    """SD1.x image slider, trainscripts/imagesliders/train_lora-scale.py:185-330: the same step as `image_slider_step_xl`
    with `predict_noise` and plain [1,77,768] embeddings (`prompt_pair.{unconditional,positive,neutral}` are tensors);
    the reference resizes the image pair to 256 px there, i.e. latents [bs,4,32,32]."""
    return _image_slider_step(False, unet, network, noise_scheduler, optimizer, lr_scheduler, prompt_pair, latents_low,
                              latents_high, scale_to_look, max_denoising_steps, timesteps_to, device, weight_dtype, seed,
                              reference_dead_code, group)


def _image_slider_step(xl: bool, unet, network, noise_scheduler, optimizer, lr_scheduler, prompt_pair, latents_low,
                       latents_high, scale_to_look, max_denoising_steps, timesteps_to, device, weight_dtype, seed,
                       reference_dead_code, group):
    device = device or unet.device
    criteria = torch.nn.MSELoss()
    world, rank = _world(group)
    with torch.no_grad():
        noise_scheduler.set_timesteps(max_denoising_steps, device=device)
        optimizer.zero_grad()
        if timesteps_to is None:
            # train_lora-scale-xl.py:191-193 draws from [1, max); the SD1.x script from [1, max - 1) (:186-189)
            timesteps_to = torch.randint(1, max_denoising_steps if xl else max_denoising_steps - 1, (1,)).item()
        if seed is None:
            seed = int(torch.randint(0, 2 ** 15, (1,)).item())
        h, w = latents_low.shape[-2:]
        height, width = h * train_util.VAE_SCALE_FACTOR, w * train_util.VAE_SCALE_FACTOR
        ids = train_util.get_add_time_ids(height, width, dynamic_crops=prompt_pair.dynamic_crops,
                                          dtype=torch.float32) if xl else torch.zeros(1, 6)
        if world > 1:  # replicas must agree on the step count, the noise seed and the (random-crop) time ids
            vals = parallel.sync_draws([timesteps_to, seed] + ids.flatten().tolist(), device, group)
            timesteps_to, seed = int(vals[0]), int(vals[1])
            ids = torch.tensor([vals[2:8]], dtype=torch.float32)
        timestep = noise_scheduler.timesteps[timesteps_to]                              # get_noisy_image :224-231
        noise = torch.randn(latents_low.shape, generator=torch.Generator().manual_seed(seed)).to(device)
        noise_w = noise.to(weight_dtype)   # the loss target is the noise as stored in weight_dtype (:235, :251)
        ts = torch.as_tensor(timestep).reshape(1)
        noisy_low = noise_scheduler.add_noise(latents_low.to(device).float(), noise, ts).to(weight_dtype)
        noisy_high = noise_scheduler.add_noise(latents_high.to(device).float(), noise, ts).to(weight_dtype)
        noise_scheduler.set_timesteps(1000)
        add_time_ids = ids.to(device, dtype=weight_dtype) if xl else None
        current_timestep = noise_scheduler.timesteps[int(timesteps_to * 1000 / max_denoising_steps)]

        def predict(noisy, which, lo=0, hi=None):
            """CFG-pair prediction at guidance 1 for samples [lo, hi) of the batch."""
            n = noisy.shape[0]
            hi = n if hi is None else hi
            if xl:
                inputs = _xl_inputs(prompt_pair, which, add_time_ids)
                if (lo, hi) != (0, n):  # [uncond x n ; cond x n] -> this shard's samples of both halves
                    inputs = {k: torch.cat([v[lo:hi], v[n + lo:n + hi]]) for k, v in inputs.items()}
                return train_util.predict_noise_xl(unet, noise_scheduler, current_timestep, noisy[lo:hi], **inputs,
                                                   guidance_scale=1)
            emb = train_util.concat_embeddings(prompt_pair.unconditional, which, n)
            if (lo, hi) != (0, n):
                emb = torch.cat([emb[lo:hi], emb[n + lo:n + hi]])
            return train_util.predict_noise(unet, noise_scheduler, current_timestep, noisy[lo:hi], emb, guidance_scale=1)

        if reference_dead_code:  # train_lora-scale-xl.py:258-306 / train_lora-scale.py:252-283, unused by the loss
            predict(noisy_high, prompt_pair.positive)
            predict(noisy_low, prompt_pair.neutral if xl else prompt_pair.unconditional)
    bs = noisy_low.shape[0]
    groups = world // 2 if world > 1 else 1            # batch shards per sign
    if world > 1 and (bs % groups != 0 or groups == 0):
        groups = 1                                      # batch does not divide: one rank per sign, the rest idle
    losses = [torch.zeros((), device=device), torch.zeros((), device=device)]
    for i, (sign, noisy, which) in enumerate(((+1.0, noisy_high, prompt_pair.positive),
                                              (-1.0, noisy_low, prompt_pair.neutral))):
        lo, hi = 0, bs
        if world > 1:
            if rank % 2 != i or rank // 2 >= groups:
                continue
            lo, hi = parallel.shard_range(bs, rank // 2, groups)
        network.set_lora_slider(scale=sign * scale_to_look)                             # :311, :343
        with network:
            pred = predict(noisy, which, lo, hi).to(device, dtype=torch.float32)
        loss = criteria(pred, noise_w[lo:hi].to(torch.float32))                         # :338, :370
        # MSE is a mean over the batch: a shard contributes (its mean) / (number of shards)
        (loss / groups if world > 1 else loss).backward()
        losses[i] = loss.detach() / (groups if world > 1 else 1)
    if world > 1:
        parallel.allreduce_lora_grads([p for g in optimizer.param_groups for p in g["params"]], group=group)
        both = torch.stack(losses).float()
        dist.all_reduce(both, group=group)
        losses = [both[0], both[1]]
    optimizer.step()
    if lr_scheduler is not None:
        lr_scheduler.step()
    return losses
