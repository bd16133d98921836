"""ORACLE (test infrastructure) — import the reference's own Python modules *verbatim* from
/root/reference/trainscripts/textsliders (lora.py, train_util.py, prompt_util.py, config_util.py, model_util.py)
and drive the oracle UNet with them.

The reference depends on `diffusers` (requirements.txt:3), which is not installed and not installable here.
Its own modules only import *names* from it (type annotations and scheduler constructors), so a stub module
named `diffusers` that exposes the oracle restatements under those names is enough for them to import and run
unmodified.  Nothing is copied: the files are loaded from where they lie.  /root/reference does not exist on the
GPU box — callers must check `available()`; golden vectors produced through this bridge are committed under
tests/golden/ by tests/golden/make_golden.py.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("SLIDERS_REFERENCE_ROOT", "/root/reference")
_TEXT = os.path.join(REFERENCE_ROOT, "trainscripts", "textsliders")
_IMAGE = os.path.join(REFERENCE_ROOT, "trainscripts", "imagesliders")


def available() -> bool:
    return os.path.isfile(os.path.join(_TEXT, "lora.py"))


def install_diffusers_stub() -> None:
    """Register stub `diffusers` / `diffusers.schedulers` modules backed by the oracle (idempotent)."""
    if "diffusers" in sys.modules and not getattr(sys.modules["diffusers"], "_sb200_stub", False):
        return  # a real diffusers is importable: use it
    from . import ddim, unet

    d = types.ModuleType("diffusers")
    d._sb200_stub = True
    d.UNet2DConditionModel = unet.UNet2DConditionModel
    d.SchedulerMixin = ddim.SchedulerMixin
    for name in ("StableDiffusionPipeline", "StableDiffusionXLPipeline", "AutoencoderKL"):
        setattr(d, name, type(name, (), {}))
    s = types.ModuleType("diffusers.schedulers")
    s.DDIMScheduler = ddim.DDIMScheduler
    for name in ("DDPMScheduler", "LMSDiscreteScheduler", "EulerAncestralDiscreteScheduler"):
        setattr(s, name, type(name, (), {"__init__": lambda self, *a, **k: (_ for _ in ()).throw(
            NotImplementedError("only DDIM is restated in the oracle"))}))
    d.schedulers = s
    # imagesliders/train_util.py:7-9 additionally imports the VAE image pre-processor and `randn_tensor`
    ip = types.ModuleType("diffusers.image_processor")
    ip.VaeImageProcessor = VaeImageProcessor
    ut = types.ModuleType("diffusers.utils")
    ut.randn_tensor = randn_tensor
    d.image_processor, d.utils = ip, ut
    d.__path__ = []  # lets `import diffusers.x` resolve the registered sub-modules
    sys.modules["diffusers"] = d
    sys.modules["diffusers.schedulers"] = s
    sys.modules["diffusers.image_processor"] = ip
    sys.modules["diffusers.utils"] = ut


class VaeImageProcessor:
    """Restatement of diffusers 0.20.2 `VaeImageProcessor.preprocess` for PIL input with the default flags
    (do_resize, do_normalize, no RGB conversion): round the size down to a multiple of the VAE scale factor, scale to
    [0, 1], NHWC -> NCHW, then to [-1, 1].  Called by imagesliders/train_util.py:212-214 (`get_noisy_image`)."""

    def __init__(self, vae_scale_factor: int = 8):
        self.vae_scale_factor = vae_scale_factor

    def preprocess(self, image):
        import numpy as np
        import torch

        images = image if isinstance(image, (list, tuple)) else [image]
        out = []
        for im in images:
            w, h = im.size
            w, h = w - w % self.vae_scale_factor, h - h % self.vae_scale_factor
            if (w, h) != im.size:
                im = im.resize((w, h))
            a = np.array(im).astype(np.float32) / 255.0
            if a.ndim == 2:
                a = a[..., None]
            out.append(a)
        t = torch.from_numpy(np.stack(out, axis=0).transpose(0, 3, 1, 2))
        return 2.0 * t - 1.0


def randn_tensor(shape, generator=None, device=None, dtype=None, layout=None):
    """diffusers.utils.randn_tensor for a single (CPU) generator: draw on the generator's device, then move."""
    import torch

    rand_device = generator.device if generator is not None else (device or "cpu")
    t = torch.randn(shape, generator=generator, device=rand_device, dtype=dtype, layout=layout or torch.strided)
    return t.to(device) if device is not None else t


def load(module: str, flavour: str = "text"):
    """Import `module` (e.g. 'lora', 'train_util') from the reference tree, unmodified."""
    if not available():
        raise FileNotFoundError(f"reference tree not found at {REFERENCE_ROOT}")
    install_diffusers_stub()
    path = _TEXT if flavour == "text" else _IMAGE
    # the reference modules import each other by bare name (`from model_util import …`), so the directory has to
    # be on sys.path while they load; modules are cached under a flavour-specific alias afterwards
    names = ("lora", "train_util", "prompt_util", "config_util", "model_util", "debug_util", "flush")
    alias = f"_ref_{flavour}_{module}"
    if alias in sys.modules:
        return sys.modules[alias]
    saved = {k: sys.modules.pop(k) for k in names if k in sys.modules}
    for k in names:  # modules of this flavour that are already loaded resolve by their bare names again
        if f"_ref_{flavour}_{k}" in sys.modules:
            sys.modules[k] = sys.modules[f"_ref_{flavour}_{k}"]
    sys.path.insert(0, path)
    try:
        mod = importlib.import_module(module)
    finally:
        sys.path.remove(path)
        for k in names:
            if k in sys.modules:
                sys.modules[f"_ref_{flavour}_{k}"] = sys.modules.pop(k)
        sys.modules.update(saved)
    return mod
