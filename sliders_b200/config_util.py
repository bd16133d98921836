"""Training configuration — the YAML schema of the reference trainers (trainscripts/textsliders/config_util.py:12-104;
imagesliders/config_util.py is the same file), so `data/config.yaml` / `data/config-xl.yaml` written for the reference
parse unchanged and `--config_file` keeps its meaning.

Sections and defaults are the reference's: `prompts_file`, `pretrained_model{name_or_path, v2, v_pred, clip_skip}`,
`network{type, rank, alpha, training_method}`, `train{precision, noise_scheduler, iterations, lr, optimizer,
optimizer_args, lr_scheduler, max_denoising_steps}`, `save{name, path, per_steps, precision}`,
`logging{use_wandb, verbose}`, `other{use_xformers}`; absent optional sections are filled with their defaults
(config_util.py:86-104).  `other.use_xformers` is accepted and ignored: attention always runs in the sb200 kernel.
"""
from __future__ import annotations

from typing import Literal, Optional

import torch
import yaml
from pydantic import BaseModel

from .lora import TRAINING_METHODS

PRECISION_TYPES = Literal["fp32", "fp16", "bf16", "float32", "float16", "bfloat16"]
NETWORK_TYPES = Literal["lierla", "c3lier"]
_DTYPES = {"fp32": torch.float32, "float32": torch.float32, "fp16": torch.float16, "float16": torch.float16,
           "bf16": torch.bfloat16, "bfloat16": torch.bfloat16}


class _Section(BaseModel):
    def json(self, **kw):  # the trainers log `config.json()` (train_lora_xl.py:43-46)
        return self.model_dump_json(**kw)


class PretrainedModelConfig(_Section):
    name_or_path: str
    v2: bool = False
    v_pred: bool = False
    clip_skip: Optional[int] = None


class NetworkConfig(_Section):
    type: NETWORK_TYPES = "lierla"
    rank: int = 4
    alpha: float = 1.0
    training_method: TRAINING_METHODS = "full"


class TrainConfig(_Section):
    precision: PRECISION_TYPES = "bfloat16"
    noise_scheduler: Literal["ddim", "ddpm", "lms", "euler_a"] = "ddim"
    iterations: int = 500
    lr: float = 1e-4
    optimizer: str = "adamw"
    optimizer_args: str = ""
    lr_scheduler: str = "constant"
    max_denoising_steps: int = 50


class SaveConfig(_Section):
    name: str = "untitled"
    path: str = "./output"
    per_steps: int = 200
    precision: PRECISION_TYPES = "float32"


class LoggingConfig(_Section):
    use_wandb: bool = False
    verbose: bool = False


class OtherConfig(_Section):
    use_xformers: bool = False


class RootConfig(_Section):
    prompts_file: str
    pretrained_model: PretrainedModelConfig
    network: NetworkConfig
    train: Optional[TrainConfig] = None
    save: Optional[SaveConfig] = None
    logging: Optional[LoggingConfig] = None
    other: Optional[OtherConfig] = None


def parse_precision(precision: str) -> torch.dtype:
    try:
        return _DTYPES[precision]
    except KeyError:
        raise ValueError(f"Invalid precision type: {precision}") from None


def load_config_from_yaml(config_path: str) -> RootConfig:
    with open(config_path, "r") as f:
        root = RootConfig(**yaml.safe_load(f))
    for name, cls in (("train", TrainConfig), ("save", SaveConfig), ("logging", LoggingConfig), ("other", OtherConfig)):
        if getattr(root, name) is None:
            setattr(root, name, cls())
    return root
