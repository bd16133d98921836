"""ORACLE (test / benchmark infrastructure) — CPU port of the reference's OWN code on the denoise path, for the places
where /root/reference cannot be imported (the GPU box): bench.py's `--impl reference` arm and `cpu_baseline` leg fall
back to this when `oracle.reference_bridge.available()` is False, and say so (`cpu_baseline.kind = "port"`).

  LoRAHook / attach_lora      trainscripts/textsliders/lora.py:50-112, :164-258 (forward = org(x) + up(down(x)) * multiplier
                              * alpha/rank on the `noxattn` + `c3lier` leaf set; context manager sets multiplier 1 / 0)
  predict_noise(_xl)          trainscripts/textsliders/train_util.py:145-171, :220-260 (CFG pair batched, guidance after)
  diffusion(_xl)              :175-196, :263-294
  text_slider_iteration       trainscripts/textsliders/train_lora.py:155-300 / train_lora_xl.py:162-347, one iteration

PARITY STATUS: pinned against the reference's real modules where they exist — tests/test_oracle.py runs this port and
the reference's lora.py / train_util.py (through reference_bridge) on the same tiny oracle UNet and requires identical
outputs and gradients.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn as nn

CONV_TARGETS = ("ResnetBlock2D", "Downsample2D", "Upsample2D", "DownBlock2D", "UpBlock2D")
ATTN_TARGETS = ("Attention",)


class LoRAHook(nn.Module):
    def __init__(self, name: str, org: nn.Module, rank: int, alpha: float):
        super().__init__()
        self.lora_name = name
        if isinstance(org, nn.Linear):
            self.lora_down = nn.Linear(org.in_features, rank, bias=False)
            self.lora_up = nn.Linear(rank, org.out_features, bias=False)
        else:
            rank = min(rank, org.in_channels, org.out_channels)
            self.lora_down = nn.Conv2d(org.in_channels, rank, org.kernel_size, org.stride, org.padding, bias=False)
            self.lora_up = nn.Conv2d(rank, org.out_channels, (1, 1), (1, 1), bias=False)
        self.scale = alpha / rank
        self.multiplier = 1.0
        nn.init.kaiming_uniform_(self.lora_down.weight, a=1)
        nn.init.zeros_(self.lora_up.weight)
        self.org_forward = org.forward
        org.forward = self.forward

    def forward(self, x):
        return self.org_forward(x) + self.lora_up(self.lora_down(x)) * self.multiplier * self.scale


class LoRAHooks(nn.Module):
    """The `noxattn` network of the shipped configs: every Linear / Conv2d leaf under an Attention (and, with c3lier,
    under the conv container classes), skipping module paths that contain `attn2` or `time_embed`."""

    def __init__(self, unet: nn.Module, rank: int = 4, alpha: float = 1.0, c3lier: bool = True):
        super().__init__()
        targets = ATTN_TARGETS + (CONV_TARGETS if c3lier else ())
        self.unet_loras: List[LoRAHook] = []
        seen = set()
        for name, module in unet.named_modules():
            if "attn2" in name or "time_embed" in name or module.__class__.__name__ not in targets:
                continue
            for child_name, child in module.named_modules():
                if not isinstance(child, (nn.Linear, nn.Conv2d)):
                    continue
                lora_name = ("lora_unet." + name + "." + child_name).replace(".", "_")
                if lora_name in seen:
                    continue
                seen.add(lora_name)
                hook = LoRAHook(lora_name, child, rank, alpha)
                self.unet_loras.append(hook)
                self.add_module(lora_name, hook)
        self.lora_scale = 1.0

    def set_lora_slider(self, scale):
        self.lora_scale = scale

    def __enter__(self):
        for h in self.unet_loras:
            h.multiplier = 1.0 * self.lora_scale

    def __exit__(self, *exc):
        for h in self.unet_loras:
            h.multiplier = 0

    def prepare_optimizer_params(self):
        return [{"params": [p for h in self.unet_loras for p in h.parameters()]}]

    def load_weights(self, state: Dict[str, torch.Tensor]) -> None:
        own = self.state_dict()
        self.load_state_dict({k: state[k].to(own[k].dtype) for k in own if k in state}, strict=False)


def concat_embeddings(unconditional, conditional, n_imgs: int):
    return torch.cat([unconditional, conditional]).repeat_interleave(n_imgs, dim=0)


def predict_noise(unet, scheduler, timestep, latents, text_embeddings, guidance_scale=7.5):
    x = scheduler.scale_model_input(torch.cat([latents] * 2), timestep)
    eps = unet(x, timestep, encoder_hidden_states=text_embeddings).sample
    u, c = eps.chunk(2)
    return u + guidance_scale * (c - u)


def predict_noise_xl(unet, scheduler, timestep, latents, text_embeddings, add_text_embeddings, add_time_ids,
                     guidance_scale=7.5):
    x = scheduler.scale_model_input(torch.cat([latents] * 2), timestep)
    eps = unet(x, timestep, encoder_hidden_states=text_embeddings,
               added_cond_kwargs={"text_embeds": add_text_embeddings, "time_ids": add_time_ids}).sample
    u, c = eps.chunk(2)
    return u + guidance_scale * (c - u)


@torch.no_grad()
def diffusion(unet, scheduler, latents, text_embeddings, total_timesteps=1000, start_timesteps=0, guidance_scale=7.5,
              added=None):
    for t in scheduler.timesteps[start_timesteps:total_timesteps]:
        if added is None:
            eps = predict_noise(unet, scheduler, t, latents, text_embeddings, guidance_scale)
        else:
            eps = predict_noise_xl(unet, scheduler, t, latents, text_embeddings, added[0], added[1], guidance_scale)
        latents = scheduler.step(eps, t, latents).prev_sample
    return latents


def text_slider_iteration(unet, network: LoRAHooks, scheduler, optimizer, emb: Dict[str, torch.Tensor], latents,
                          timesteps_to: int, guidance_scale: float = 4.0, action: str = "enhance",
                          max_denoising_steps: int = 50, added: Optional[Dict[str, tuple]] = None, batch_size: int = 1):
    """One iteration of the text-slider loop with the draws (timesteps_to, latents) given.  `emb[name]` are the four
    prompt embeddings ([1,77,D]); `added[name] = (pooled, time_ids)` for SDXL."""
    pair = lambda name: concat_embeddings(emb["unconditional"], emb[name], batch_size)
    add = lambda name: None if added is None else (concat_embeddings(added["unconditional"][0], added[name][0], batch_size),
                                                   concat_embeddings(added[name][1], added[name][1], batch_size))

    def predict(name):
        if added is None:
            return predict_noise(unet, scheduler, current_timestep, denoised, pair(name), guidance_scale=1)
        a = add(name)
        return predict_noise_xl(unet, scheduler, current_timestep, denoised, pair(name), a[0], a[1], guidance_scale=1)

    with torch.no_grad():
        scheduler.set_timesteps(max_denoising_steps)
        optimizer.zero_grad()
        with network:
            denoised = diffusion(unet, scheduler, latents, pair("target"), total_timesteps=timesteps_to,
                                 guidance_scale=3, added=add("target"))
        scheduler.set_timesteps(1000)
        current_timestep = scheduler.timesteps[int(timesteps_to * 1000 / max_denoising_steps)]
        positive, neutral, unconditional = predict("positive"), predict("neutral"), predict("unconditional")
    with network:
        target = predict("target")
    sign = 1.0 if action == "enhance" else -1.0
    loss = torch.nn.functional.mse_loss(target, neutral + sign * guidance_scale * (positive - unconditional))
    loss.backward()
    optimizer.step()
    return loss.detach()
