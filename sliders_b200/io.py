"""Checkpoint ingestion for the denoise path.

* `load_unet(config, path)`: a Hugging Face `unet/diffusion_pytorch_model.safetensors` (or `.bin` / `.pt`) state dict
  loads with plain `load_state_dict(strict=True)` because the module tree keeps diffusers' key names and tensor shapes
  (trainscripts/textsliders/model_util.py:200-227 `load_models_xl` -> `UNet2DConditionModel.from_pretrained`); kernel
  layouts (tap-major conv weights, fused QKV, transposed copies) are derived lazily from these tensors.
* `load_slider(network, path)`: a slider checkpoint written by either implementation (`LoRANetwork.save_weights`,
  trainscripts/textsliders/lora.py:231-248: `.pt` via torch.save or `.safetensors`), key layout
  `lora_unet_<module path>.{alpha, lora_down.weight, lora_up.weight}`.
"""
from __future__ import annotations

import os
from typing import Dict

import torch

from .unet import UNet2DConditionModel, UNetConfig


def read_state_dict(path: str) -> Dict[str, torch.Tensor]:
    if os.path.splitext(path)[1] == ".safetensors":
        from safetensors.torch import load_file

        return load_file(path)
    sd = torch.load(path, map_location="cpu")
    return sd.get("state_dict", sd) if isinstance(sd, dict) else sd


def load_unet(config: UNetConfig, path: str, device=None, dtype=torch.bfloat16) -> UNet2DConditionModel:
    unet = UNet2DConditionModel(config)
    try:
        missing, unexpected = unet.load_state_dict(read_state_dict(path), strict=False)
    except RuntimeError as e:  # torch reports shape mismatches this way
        raise RuntimeError(f"{path}: not a UNet2DConditionModel state dict for this config "
                           f"({str(e).splitlines()[1].strip() if len(str(e).splitlines()) > 1 else e})") from None
    if missing or unexpected:
        raise RuntimeError(f"{path}: not a UNet2DConditionModel state dict for this config "
                           f"(missing {len(missing)} e.g. {missing[:3]}, unexpected {len(unexpected)} e.g. {unexpected[:3]})")
    unet.requires_grad_(False)
    unet.eval()
    return unet.to(device=device, dtype=dtype) if device is not None else unet.to(dtype=dtype)


def load_slider(network, path: str) -> None:
    """Loads a slider checkpoint into an already injected `LoRANetwork` (same rank / train_method)."""
    sd = read_state_dict(path)
    own = network.state_dict()
    extra = sorted(set(sd) - set(own))
    lacking = sorted(set(own) - set(sd))
    if extra or lacking:
        raise RuntimeError(f"{path}: adaptor keys do not match the network (unknown {extra[:3]}, missing {lacking[:3]}); "
                           "was it trained with another train_method or without the conv (c3lier) targets?")
    for k, v in sd.items():
        if tuple(v.shape) != tuple(own[k].shape):
            raise RuntimeError(f"{path}: {k} has shape {tuple(v.shape)}, the network expects {tuple(own[k].shape)} (rank?)")
    network.load_state_dict({k: v.to(own[k].dtype) for k, v in sd.items()})
