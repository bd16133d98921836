"""Prompt files — the YAML schema and helpers of trainscripts/textsliders/prompt_util.py (imagesliders: same file).

  PromptSettings          :38-68   one entry of `prompts*.yaml`; `positive` defaults to `target`, `neutral` to
                                   `unconditional`, unknown keys (the GPT-written files carry `guidance:` / `rank:`,
                                   SURVEY.md C.7) are ignored
  PromptEmbedsCache       :26-35   prompt -> embedding memo used while the text encoders are still loaded
  PromptEmbedsXL / Pair   :16-23, :71-148   re-exported from sliders_b200.trainer (the loop-side classes)
  load_prompts_from_yaml  :151-174 `--attributes a,b` multiplies every entry by the attribute list, prefixing
                                   target / positive / neutral / unconditional with "<attribute> "
"""
from __future__ import annotations

import copy
from typing import Dict, List, Literal, Optional, Sequence

import yaml
from pydantic import BaseModel, ConfigDict, model_validator

from .trainer import PromptEmbedsPair, PromptEmbedsXL  # noqa: F401  (same names as the reference module exports)

ACTION_TYPES = Literal["erase", "enhance"]
_PROMPT_KEYS = ("target", "positive", "neutral", "unconditional")


class PromptSettings(BaseModel):
    model_config = ConfigDict(extra="ignore")

    target: str
    positive: Optional[str] = None      # None -> target
    unconditional: str = ""
    neutral: Optional[str] = None       # None -> unconditional
    action: ACTION_TYPES = "erase"
    guidance_scale: float = 1.0
    resolution: int = 512
    dynamic_resolution: bool = False
    batch_size: int = 1
    dynamic_crops: bool = False         # XL only

    @model_validator(mode="before")
    @classmethod
    def _fill_prompts(cls, values):
        values = dict(values)
        if "target" not in values:
            raise ValueError("target must be specified")
        values.setdefault("positive", values["target"])
        values.setdefault("unconditional", "")
        values.setdefault("neutral", values["unconditional"])
        return values

    def json(self, **kw):  # the trainers log `prompt.json()` (train_lora_xl.py:44)
        return self.model_dump_json(**kw)


class PromptEmbedsCache:
    """`cache[prompt]` is None until the prompt has been encoded (train_lora_xl.py:113-131 relies on that)."""

    def __init__(self) -> None:
        self.prompts: Dict[str, object] = {}

    def __setitem__(self, name: str, value) -> None:
        self.prompts[name] = value

    def __getitem__(self, name: str):
        return self.prompts.get(name)


def expand_attributes(prompts: Sequence[dict], attributes: Sequence[str]) -> List[dict]:
    """prompt_util.py:157-168: one copy of every entry per disentanglement attribute, entry-major."""
    if not attributes:
        return copy.deepcopy(list(prompts))
    out = []
    for entry in prompts:
        filled = PromptSettings._fill_prompts(entry)
        for att in attributes:
            e = dict(filled)
            for key in _PROMPT_KEYS:
                e[key] = f"{att} {filled[key]}"
            out.append(e)
    return out


def load_prompts_from_yaml(path, attributes: Sequence[str] = ()) -> List[PromptSettings]:
    with open(path, "r") as f:
        prompts = yaml.safe_load(f)
    if not prompts:
        raise ValueError("prompts file is empty")
    expanded = expand_attributes(prompts, list(attributes))
    print(len(prompts), len(expanded))
    return [PromptSettings(**p) for p in expanded]
