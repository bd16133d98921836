"""Model loading for the trainers — the call surface of trainscripts/textsliders/model_util.py (`load_models` :103-130,
`load_models_xl` :200-227, `create_noise_scheduler` :230-278) on top of the sliders_b200 UNet.

The reference resolves `pretrained_model.name_or_path` through diffusers (hub id, local diffusers directory or a single
.ckpt / .safetensors file).  Here the UNet is `sliders_b200.unet.UNet2DConditionModel`, loaded by HF key name
(`sliders_b200.io.load_unet`), and `name_or_path` may be

  * a local diffusers directory (`unet/diffusion_pytorch_model.safetensors` [+ `text_encoder*/`, `tokenizer*/`]),
  * a UNet state-dict file (`.safetensors` / `.bin` / `.pt`) in diffusers key layout,
  * `synthetic` or `synthetic:<seed>` — the seeded synthetic weights of SURVEY.md A.7 (no checkpoint is reachable in
    an offline container; the architecture, shapes and key names are the real ones),
  * a hub id: resolved from the local Hugging Face cache only (there is no network path in this package).

Tokenizers / text encoders are stock `transformers` modules (they run once, before the loop: train_lora_xl.py:100-151,
off the denoise path) and are returned as `None` when the source holds none — the trainers then need `--embeds`.
Single-file `.ckpt` in the original CompVis layout (`StableDiffusionPipeline.from_ckpt`, model_util.py:75-100) is
not converted here: its key map lives in diffusers.
"""
from __future__ import annotations

import os
from typing import List, Optional, Tuple

import torch

from . import io as sio
from . import synthetic
from .scheduler import create_noise_scheduler  # noqa: F401  (model_util.create_noise_scheduler in the reference)
from .unet import UNet2DConditionModel, UNetConfig

AVAILABLE_SCHEDULERS = ("ddim", "ddpm", "lms", "euler_a")
_UNET_FILES = ("diffusion_pytorch_model.safetensors", "diffusion_pytorch_model.fp16.safetensors",
               "diffusion_pytorch_model.bin")


def _resolve(name_or_path: str) -> Tuple[str, Optional[str]]:
    """-> (kind, path): kind in {"synthetic", "dir", "file"}."""
    if name_or_path.startswith("synthetic"):
        return "synthetic", name_or_path
    if os.path.isdir(name_or_path):
        return "dir", name_or_path
    if os.path.isfile(name_or_path):
        if name_or_path.endswith(".ckpt"):
            raise NotImplementedError(f"{name_or_path}: CompVis-layout .ckpt files need diffusers' key conversion; "
                                      "pass a diffusers directory or the UNet's diffusion_pytorch_model.safetensors")
        return "file", name_or_path
    try:  # hub id: local cache only
        from huggingface_hub import snapshot_download

        return "dir", snapshot_download(name_or_path, local_files_only=True)
    except Exception as e:  # noqa: BLE001
        raise FileNotFoundError(f"{name_or_path}: not a local path and not in the Hugging Face cache ({type(e).__name__}); "
                                "this package never downloads — use a local diffusers directory, a UNet state-dict "
                                "file, or 'synthetic'") from None


def _unet_file(root: str) -> str:
    for sub in ("unet", ""):
        for f in _UNET_FILES:
            p = os.path.join(root, sub, f)
            if os.path.isfile(p):
                return p
    raise FileNotFoundError(f"{root}: no unet/diffusion_pytorch_model.(safetensors|bin)")


def load_unet(name_or_path: str, config: UNetConfig, weight_dtype: torch.dtype = torch.bfloat16,
              device=None) -> UNet2DConditionModel:
    kind, path = _resolve(name_or_path)
    if kind == "synthetic":
        seed = int(path.split(":", 1)[1]) if ":" in path else 1
        if device is not None:
            with torch.device(device):
                unet = UNet2DConditionModel(config).to(weight_dtype)
        else:
            unet = UNet2DConditionModel(config).to(weight_dtype)
        synthetic.init_synthetic_(unet, seed=seed)
        unet.requires_grad_(False)
        return unet.eval()
    return sio.load_unet(config, _unet_file(path) if kind == "dir" else path, device=device, dtype=weight_dtype)


def _text_stack(root: str, subs: List[Tuple[str, str, str]], weight_dtype):
    """[(tokenizer subfolder, encoder subfolder, encoder class name)] -> (tokenizers, encoders) or (None, None)."""
    if not all(os.path.isdir(os.path.join(root, s)) for pair in subs for s in pair[:2]):
        return None, None
    import transformers

    toks, encs = [], []
    for tok_sub, enc_sub, cls in subs:
        kw = {"pad_token_id": 0} if tok_sub == "tokenizer_2" else {}   # model_util.py:146-152 ("same as open clip")
        toks.append(transformers.CLIPTokenizer.from_pretrained(root, subfolder=tok_sub, **kw))
        encs.append(getattr(transformers, cls).from_pretrained(root, subfolder=enc_sub, torch_dtype=weight_dtype))
    return toks, encs


def load_models(pretrained_model_name_or_path: str, scheduler_name: str, v2: bool = False, v_pred: bool = False,
                weight_dtype: torch.dtype = torch.float32, device=None):
    """-> (tokenizer, text_encoder, unet, noise_scheduler) for SD1.x (model_util.py:103-130)."""
    if v2 or v_pred:
        raise NotImplementedError("SD 2.x / v-prediction models are outside the slider path (SURVEY.md A.4)")
    unet = load_unet(pretrained_model_name_or_path, UNetConfig.sd15(), weight_dtype, device)
    kind, path = _resolve(pretrained_model_name_or_path)
    toks, encs = (None, None)
    if kind == "dir":
        toks, encs = _text_stack(path, [("tokenizer", "text_encoder", "CLIPTextModel")], weight_dtype)
    return (toks[0] if toks else None), (encs[0] if encs else None), unet, create_noise_scheduler(scheduler_name)


def load_models_xl(pretrained_model_name_or_path: str, scheduler_name: str,
                   weight_dtype: torch.dtype = torch.float32, device=None):
    """-> (tokenizers, text_encoders, unet, noise_scheduler) for SDXL (model_util.py:200-227)."""
    unet = load_unet(pretrained_model_name_or_path, UNetConfig.sdxl(), weight_dtype, device)
    kind, path = _resolve(pretrained_model_name_or_path)
    toks, encs = (None, None)
    if kind == "dir":
        toks, encs = _text_stack(path, [("tokenizer", "text_encoder", "CLIPTextModel"),
                                        ("tokenizer_2", "text_encoder_2", "CLIPTextModelWithProjection")], weight_dtype)
    return toks, encs, unet, create_noise_scheduler(scheduler_name)
