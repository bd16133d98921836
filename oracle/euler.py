"""ORACLE (test infrastructure) — restatement of diffusers==0.20.2 `schedulers/scheduling_euler_discrete.py`
`EulerDiscreteScheduler`, the scheduler the SDXL pipeline carries in eval-scripts/generate_images_xl.py (the script
never replaces `self.scheduler`: :267 `set_timesteps`, :334 `scale_model_input`, :358 `step`), with the
stabilityai/stable-diffusion-xl-base-1.0 scheduler config (scaled_linear betas 0.00085..0.012, 1000 train steps,
epsilon prediction, interpolation "linear", timestep_spacing "leading", steps_offset 1, no Karras sigmas).

    sigma_t        = sqrt((1 - alpha_bar_t) / alpha_bar_t)           (interpolated at the inference timesteps; 0 appended)
    scale_model_input(x, t) = x / sqrt(sigma_t^2 + 1)
    step (s_churn = 0):  x0 = x - sigma_t eps;  dx = (x - x0) / sigma_t = eps;  x_next = x + dx (sigma_next - sigma_t)
    init_noise_sigma = sigma_max ("linspace"/"trailing" spacing) or sqrt(sigma_max^2 + 1) ("leading")

PARITY STATUS: unpinned by the reference (it has no tests and diffusers is not installable here).  Pinned by known
answers (sigma_max = 14.6146, sigma_min = 0.0292 for these betas) and by the exact equivalence with DDIM (eta = 0) in
variance-preserving coordinates, x_vp = x / sqrt(1 + sigma^2), on a shared timestep grid (tests/test_oracle.py).
Only tests/, smoke() and bench.py's CPU legs may import this.
"""
from __future__ import annotations

from types import SimpleNamespace

import numpy as np
import torch


class EulerDiscreteScheduler:
    order = 1

    def __init__(self, num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear",
                 prediction_type="epsilon", interpolation_type="linear", use_karras_sigmas=False,
                 timestep_spacing="linspace", steps_offset=0, **unused):
        if beta_schedule == "linear":
            betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        elif beta_schedule == "scaled_linear":
            betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        else:
            raise NotImplementedError(beta_schedule)
        assert prediction_type == "epsilon" and interpolation_type == "linear" and not use_karras_sigmas
        self.config = SimpleNamespace(num_train_timesteps=num_train_timesteps, prediction_type=prediction_type,
                                      timestep_spacing=timestep_spacing, steps_offset=steps_offset)
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        sigmas = (((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5).numpy()
        self.sigmas = torch.from_numpy(np.concatenate([sigmas[::-1], [0.0]]).astype(np.float32))
        self.timesteps = torch.from_numpy(np.linspace(0, num_train_timesteps - 1, num_train_timesteps,
                                                      dtype=float)[::-1].copy())
        self.num_inference_steps = None

    @property
    def init_noise_sigma(self):
        if self.config.timestep_spacing in ("linspace", "trailing"):
            return self.sigmas.max()
        return (self.sigmas.max() ** 2 + 1) ** 0.5

    def _index(self, timestep):
        return int((self.timesteps == float(timestep)).nonzero()[0].item())

    def scale_model_input(self, sample, timestep):
        sigma = self.sigmas[self._index(timestep)]
        return sample / ((sigma ** 2 + 1) ** 0.5)

    def set_timesteps(self, num_inference_steps, device=None):
        T = self.config.num_train_timesteps
        self.num_inference_steps = num_inference_steps
        sp = self.config.timestep_spacing
        if sp == "linspace":
            timesteps = np.linspace(0, T - 1, num_inference_steps, dtype=float)[::-1].copy()
        elif sp == "leading":
            step_ratio = T // num_inference_steps
            timesteps = (np.arange(0, num_inference_steps) * step_ratio).round()[::-1].copy().astype(float)
            timesteps += self.config.steps_offset
        elif sp == "trailing":
            step_ratio = T / num_inference_steps
            timesteps = (np.arange(T, 0, -step_ratio)).round().copy().astype(float) - 1
        else:
            raise ValueError(sp)
        sigmas = (((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5).numpy()
        sigmas = np.interp(timesteps, np.arange(0, len(sigmas)), sigmas)
        self.sigmas = torch.from_numpy(np.concatenate([sigmas, [0.0]]).astype(np.float32)).to(device)
        self.timesteps = torch.from_numpy(timesteps).to(device)

    def step(self, model_output, timestep, sample, return_dict=True, **unused):
        i = self._index(timestep)
        sigma = self.sigmas[i]
        pred_original_sample = sample - sigma * model_output
        derivative = (sample - pred_original_sample) / sigma
        dt = self.sigmas[i + 1] - sigma
        prev = sample + derivative * dt
        return SimpleNamespace(prev_sample=prev) if return_dict else (prev,)
