"""ORACLE (test infrastructure) — restatements of the two stochastic samplers the reference's scheduler factory offers
besides DDIM / LMS (trainscripts/textsliders/model_util.py:247-278, `config.train.noise_scheduler: "ddpm" | "euler_a"`):
diffusers==0.20.2 `schedulers/scheduling_ddpm.py::DDPMScheduler` and
`schedulers/scheduling_euler_ancestral_discrete.py::EulerAncestralDiscreteScheduler`, with the arguments the factory
passes (scaled_linear betas 0.00085..0.012, 1000 train steps, epsilon prediction; DDPM: clip_sample=False) and the
library defaults otherwise (DDPM: variance_type "fixed_small", timestep_spacing "leading"; Euler-a: "linspace").

  DDPM    x0 = (x - sqrt(1-abar_t) eps) / sqrt(abar_t);  a_t = abar_t / abar_prev;  b_t = 1 - a_t
          x_prev = sqrt(abar_prev) b_t / (1-abar_t) x0 + sqrt(a_t) (1-abar_prev) / (1-abar_t) x
                   + [t > 0] sqrt(clamp((1-abar_prev)/(1-abar_t) b_t, 1e-20)) z
  Euler-a sigma_up = sqrt(s_to^2 (s_from^2 - s_to^2) / s_from^2);  sigma_down = sqrt(s_to^2 - sigma_up^2)
          x_next = x + eps (sigma_down - s_from) + sigma_up z          (x0 = x - s_from eps, derivative = eps)
  z = randn_tensor(model_output.shape, generator=generator, device=..., dtype=...)

PARITY STATUS: unpinned by the reference (no tests; diffusers not installable).  Pinned by (tests/test_oracle.py): the
DDPM posterior mean/variance identities (one step from x_t built by add_noise(x0, eps) has mean
posterior_mean(x0, x_t) and the closed-form variance), sigma_up^2 + sigma_down^2 = sigma_to^2 for Euler-a, the shared
sigma grid with EulerDiscrete, and eta-free limits.  Only tests/, smoke() and bench.py's CPU legs may import this.
"""
from __future__ import annotations

from types import SimpleNamespace

import numpy as np
import torch

from .euler import EulerDiscreteScheduler


def _randn_like_output(model_output, generator):
    dev = generator.device if generator is not None else model_output.device
    return torch.randn(model_output.shape, generator=generator, device=dev, dtype=model_output.dtype).to(model_output.device)


class DDPMScheduler:
    order = 1

    def __init__(self, num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear",
                 variance_type="fixed_small", clip_sample=True, prediction_type="epsilon", timestep_spacing="leading",
                 steps_offset=0, **unused):
        if beta_schedule == "linear":
            betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        elif beta_schedule == "scaled_linear":
            betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        else:
            raise NotImplementedError(beta_schedule)
        assert prediction_type == "epsilon" and variance_type == "fixed_small" and not clip_sample
        self.config = SimpleNamespace(num_train_timesteps=num_train_timesteps, timestep_spacing=timestep_spacing,
                                      steps_offset=steps_offset, prediction_type=prediction_type)
        self.betas = betas
        self.alphas = 1.0 - betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.one = torch.tensor(1.0)
        self.init_noise_sigma = 1.0
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy())

    def scale_model_input(self, sample, timestep=None):
        return sample

    def set_timesteps(self, num_inference_steps, device=None):
        self.num_inference_steps = num_inference_steps
        step_ratio = self.config.num_train_timesteps // num_inference_steps   # "leading"
        ts = (np.arange(0, num_inference_steps) * step_ratio).round()[::-1].copy().astype(np.int64)
        ts += self.config.steps_offset
        self.timesteps = torch.from_numpy(ts).to(device)

    def previous_timestep(self, timestep):
        n = self.num_inference_steps if self.num_inference_steps else self.config.num_train_timesteps
        return timestep - self.config.num_train_timesteps // n

    def _get_variance(self, t):
        prev_t = self.previous_timestep(t)
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.one
        current_beta_t = 1 - a_t / a_prev
        return torch.clamp((1 - a_prev) / (1 - a_t) * current_beta_t, min=1e-20)

    def step(self, model_output, timestep, sample, generator=None, return_dict=True):
        t = int(timestep)
        prev_t = self.previous_timestep(t)
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.one
        b_t, b_prev = 1 - a_t, 1 - a_prev
        cur_a = a_t / a_prev
        cur_b = 1 - cur_a
        x0 = (sample - b_t ** 0.5 * model_output) / a_t ** 0.5
        prev = (a_prev ** 0.5 * cur_b) / b_t * x0 + cur_a ** 0.5 * b_prev / b_t * sample
        if t > 0:
            prev = prev + self._get_variance(t) ** 0.5 * _randn_like_output(model_output, generator)
        return SimpleNamespace(prev_sample=prev, pred_original_sample=x0) if return_dict else (prev,)

    def add_noise(self, original_samples, noise, timesteps):
        acp = self.alphas_cumprod.to(device=original_samples.device, dtype=original_samples.dtype)
        timesteps = torch.as_tensor(timesteps, device=original_samples.device).reshape(-1)
        sa, sb = acp[timesteps] ** 0.5, (1 - acp[timesteps]) ** 0.5
        while sa.dim() < original_samples.dim():
            sa, sb = sa.unsqueeze(-1), sb.unsqueeze(-1)
        return sa * original_samples + sb * noise


class EulerAncestralDiscreteScheduler(EulerDiscreteScheduler):
    def __init__(self, num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear",
                 prediction_type="epsilon", timestep_spacing="linspace", steps_offset=0, **unused):
        super().__init__(num_train_timesteps, beta_start, beta_end, beta_schedule, prediction_type, "linear", False,
                         timestep_spacing, steps_offset)

    def step(self, model_output, timestep, sample, generator=None, return_dict=True, **unused):
        i = self._index(timestep)
        sigma = self.sigmas[i]
        pred_original_sample = sample - sigma * model_output
        sigma_from, sigma_to = self.sigmas[i], self.sigmas[i + 1]
        sigma_up = (sigma_to ** 2 * (sigma_from ** 2 - sigma_to ** 2) / sigma_from ** 2) ** 0.5
        sigma_down = (sigma_to ** 2 - sigma_up ** 2) ** 0.5
        derivative = (sample - pred_original_sample) / sigma
        prev = sample + derivative * (sigma_down - sigma)
        prev = prev + _randn_like_output(model_output, generator) * sigma_up
        return SimpleNamespace(prev_sample=prev, pred_original_sample=pred_original_sample) if return_dict else (prev,)
