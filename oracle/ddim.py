"""ORACLE (test infrastructure) — restatement of diffusers==0.20.2 `schedulers/scheduling_ddim.py`
`DDIMScheduler` as the reference constructs it (trainscripts/textsliders/model_util.py:237-246) and calls it
(train_util.py:156,193,234,291; train_lora_xl.py:164,229-233; imagesliders/train_util.py:201-235).
Tensor-in / tensor-out, fp32 coefficient tables, exactly the published update rule (eta = 0, no clipping):

    alpha_prod_t      = alphas_cumprod[t]
    alpha_prod_t_prev = alphas_cumprod[t - T/N] if t - T/N >= 0 else final_alpha_cumprod (= 1)
    x0     = (x_t - sqrt(1 - alpha_prod_t) * eps) / sqrt(alpha_prod_t)
    x_prev = sqrt(alpha_prod_t_prev) * x0 + sqrt(1 - alpha_prod_t_prev) * eps

PARITY STATUS: unpinned by the reference (no tests there); pinned by the known answers of SURVEY.md §4
(alphas_cumprod[0] = 0.99915, [980] = 0.0058438, [999] = 0.0046601; 50-step grid 980, 960, …, 0) in
tests/test_oracle.py.  Only tests/, smoke() and bench.py's CPU legs may import this.
"""
from __future__ import annotations

from types import SimpleNamespace

import numpy as np
import torch


class SchedulerMixin:  # name imported by the reference (train_util.py:6), type annotation only
    pass


class DDIMScheduler(SchedulerMixin):
    order = 1

    def __init__(self, num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear",
                 clip_sample=True, set_alpha_to_one=True, steps_offset=0, prediction_type="epsilon", **unused):
        if beta_schedule == "linear":
            self.betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        elif beta_schedule == "scaled_linear":
            self.betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps,
                                        dtype=torch.float32) ** 2
        else:
            raise NotImplementedError(beta_schedule)
        self.config = SimpleNamespace(num_train_timesteps=num_train_timesteps, clip_sample=clip_sample,
                                      steps_offset=steps_offset, prediction_type=prediction_type,
                                      timestep_spacing="leading")
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.init_noise_sigma = 1.0
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))

    def scale_model_input(self, sample, timestep=None):
        return sample

    def set_timesteps(self, num_inference_steps, device=None):
        self.num_inference_steps = num_inference_steps
        step_ratio = self.config.num_train_timesteps // self.num_inference_steps
        timesteps = (np.arange(0, num_inference_steps) * step_ratio).round()[::-1].copy().astype(np.int64)
        timesteps += self.config.steps_offset
        self.timesteps = torch.from_numpy(timesteps).to(device)

    def step(self, model_output, timestep, sample, eta=0.0, return_dict=True, **unused):
        assert self.config.prediction_type == "epsilon" and eta == 0.0 and not self.config.clip_sample
        prev_timestep = timestep - self.config.num_train_timesteps // self.num_inference_steps
        alpha_prod_t = self.alphas_cumprod[timestep]
        alpha_prod_t_prev = self.alphas_cumprod[prev_timestep] if prev_timestep >= 0 else self.final_alpha_cumprod
        beta_prod_t = 1 - alpha_prod_t
        pred_original_sample = (sample - beta_prod_t ** 0.5 * model_output) / alpha_prod_t ** 0.5
        pred_epsilon = model_output
        pred_sample_direction = (1 - alpha_prod_t_prev) ** 0.5 * pred_epsilon
        prev_sample = alpha_prod_t_prev ** 0.5 * pred_original_sample + pred_sample_direction
        if not return_dict:
            return (prev_sample,)
        return SimpleNamespace(prev_sample=prev_sample, pred_original_sample=pred_original_sample)

    def add_noise(self, original_samples, noise, timesteps):
        alphas_cumprod = self.alphas_cumprod.to(device=original_samples.device, dtype=original_samples.dtype)
        timesteps = timesteps.to(original_samples.device)
        sqrt_alpha_prod = alphas_cumprod[timesteps] ** 0.5
        sqrt_alpha_prod = sqrt_alpha_prod.flatten()
        while len(sqrt_alpha_prod.shape) < len(original_samples.shape):
            sqrt_alpha_prod = sqrt_alpha_prod.unsqueeze(-1)
        sqrt_one_minus_alpha_prod = (1 - alphas_cumprod[timesteps]) ** 0.5
        sqrt_one_minus_alpha_prod = sqrt_one_minus_alpha_prod.flatten()
        while len(sqrt_one_minus_alpha_prod.shape) < len(original_samples.shape):
            sqrt_one_minus_alpha_prod = sqrt_one_minus_alpha_prod.unsqueeze(-1)
        return sqrt_alpha_prod * original_samples + sqrt_one_minus_alpha_prod * noise
