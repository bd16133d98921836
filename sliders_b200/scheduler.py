"""DDIM scheduler with the call surface the reference uses on diffusers' `DDIMScheduler`
(constructed at trainscripts/textsliders/model_util.py:237-246: scaled_linear betas 0.00085..0.012, 1000 train
steps, clip_sample=False, epsilon prediction; defaults otherwise: set_alpha_to_one=True, steps_offset=0,
timestep_spacing="leading", eta=0).

Used as: `scheduler.set_timesteps(n, device=…)`, `scheduler.timesteps[i]`, `scheduler.init_noise_sigma`,
`scheduler.scale_model_input(x, t)`, `scheduler.step(eps, t, x).prev_sample`, `scheduler.add_noise(x0, n, t)`
(train_lora_xl.py:164-233, train_util.py:156,193,234,291; imagesliders/train_util.py:201-235).
On CUDA tensors `step` runs the fused `cfg_ddim_kernel` (ops.cfg_ddim); scalar coefficient look-ups stay on the host.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Optional, Union

import numpy as np
import torch

from . import ops


class DDIMSchedulerOutput(SimpleNamespace):
    pass


class DDIMScheduler:
    order = 1
    # how the denoise loops fuse guidance + step into `cfg_ddim_kernel`: "ddim" (alpha-bar pair), "affine"
    # (x' = cx x + ce eps), "affine+noise" (+ std * z), "generic" (scheduler.step on the guided eps)
    step_kind = "ddim"

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.00085, beta_end: float = 0.012,
                 beta_schedule: str = "scaled_linear", clip_sample: bool = False, set_alpha_to_one: bool = True,
                 steps_offset: int = 0, prediction_type: str = "epsilon"):
        if beta_schedule == "scaled_linear":
            betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        elif beta_schedule == "linear":
            betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        else:
            raise NotImplementedError(beta_schedule)
        if prediction_type != "epsilon":
            raise NotImplementedError("only epsilon prediction is on the slider path (model_util.py:126)")
        if clip_sample:
            raise NotImplementedError("clip_sample=True is not used by the reference (model_util.py:243)")
        self.config = SimpleNamespace(num_train_timesteps=num_train_timesteps, beta_start=beta_start,
                                      beta_end=beta_end, beta_schedule=beta_schedule, clip_sample=clip_sample,
                                      set_alpha_to_one=set_alpha_to_one, steps_offset=steps_offset,
                                      prediction_type=prediction_type)
        self.betas = betas
        self.alphas = 1.0 - betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self._acp = self.alphas_cumprod.double().tolist()  # host copy for scalar look-ups
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.init_noise_sigma = 1.0
        self.num_inference_steps: Optional[int] = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))

    def scale_model_input(self, sample: torch.Tensor, timestep=None) -> torch.Tensor:
        return sample

    def set_timesteps(self, num_inference_steps: int, device: Union[str, torch.device, None] = None):
        if num_inference_steps > self.config.num_train_timesteps:
            raise ValueError("num_inference_steps exceeds num_train_timesteps")
        self.num_inference_steps = num_inference_steps
        step_ratio = self.config.num_train_timesteps // num_inference_steps  # "leading" spacing
        ts = (np.arange(0, num_inference_steps) * step_ratio).round()[::-1].copy().astype(np.int64)
        ts += self.config.steps_offset
        self.timesteps = torch.from_numpy(ts).to(device)
        # host mirror: the denoise loops index it so that no step has to read a CUDA scalar back (a device sync per
        # step; diffusers' scheduler.step does exactly that when `set_timesteps(..., device=cuda)` is used)
        self.timesteps_host = [int(v) for v in ts]

    def _alphas_for(self, timestep) -> tuple:
        t = int(timestep)
        prev = t - self.config.num_train_timesteps // self.num_inference_steps
        a_t = self._acp[t]
        a_prev = self._acp[prev] if prev >= 0 else float(self.final_alpha_cumprod)
        return a_t, a_prev

    def step(self, model_output: torch.Tensor, timestep, sample: torch.Tensor, eta: float = 0.0,
             return_dict: bool = True, **unused):
        if self.num_inference_steps is None:
            raise ValueError("call set_timesteps first")
        if eta != 0.0:
            raise NotImplementedError("eta > 0 is not used by the reference")
        a_t, a_prev = self._alphas_for(timestep)
        if model_output.is_cuda:
            _, prev = ops.cfg_ddim(model_output.contiguous(), 0.0, sample.contiguous(), a_t, a_prev,
                                   out_dtype=sample.dtype if sample.dtype in (torch.float32, torch.bfloat16)
                                   else torch.float32, single=True)
            prev = prev.to(sample.dtype)
        else:  # host tensors (scheduler unit tests); not a model path
            x0 = (sample - (1 - a_t) ** 0.5 * model_output) / a_t ** 0.5
            prev = a_prev ** 0.5 * x0 + (1 - a_prev) ** 0.5 * model_output
        if not return_dict:
            return (prev,)
        return DDIMSchedulerOutput(prev_sample=prev)

    def add_noise(self, original_samples: torch.Tensor, noise: torch.Tensor, timesteps) -> torch.Tensor:
        acp = self.alphas_cumprod.to(device=original_samples.device, dtype=original_samples.dtype)
        timesteps = torch.as_tensor(timesteps, device=original_samples.device).reshape(-1)
        sa = acp[timesteps] ** 0.5
        sb = (1 - acp[timesteps]) ** 0.5
        while sa.dim() < original_samples.dim():
            sa, sb = sa.unsqueeze(-1), sb.unsqueeze(-1)
        return sa * original_samples + sb * noise

    def __len__(self):
        return self.config.num_train_timesteps


class EulerDiscreteScheduler:
    """diffusers' `EulerDiscreteScheduler` (epsilon prediction, linear interpolation, s_churn = 0) — the scheduler the
    SDXL pipeline carries through eval-scripts/generate_images_xl.py (:267 set_timesteps, :334 scale_model_input,
    :358 step); defaults are the stabilityai/stable-diffusion-xl-base-1.0 scheduler config.  Coefficients stay on the
    host; on CUDA tensors `step` is one `cfg_ddim_kernel` launch in its affine mode (x_next = x + (sigma' - sigma) eps),
    and `generate.denoise_loop` fuses the guidance into the same launch through `_step_coeffs`."""
    order = 1
    step_kind = "affine"  # x_next = cx x + ce eps with host-side (cx, ce): eligible for the fused guidance + step kernel

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.00085, beta_end: float = 0.012,
                 beta_schedule: str = "scaled_linear", prediction_type: str = "epsilon",
                 interpolation_type: str = "linear", use_karras_sigmas: bool = False,
                 timestep_spacing: str = "leading", steps_offset: int = 1):
        if beta_schedule == "scaled_linear":
            betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        elif beta_schedule == "linear":
            betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        else:
            raise NotImplementedError(beta_schedule)
        if prediction_type != "epsilon" or interpolation_type != "linear" or use_karras_sigmas:
            raise NotImplementedError("EulerDiscreteScheduler: epsilon prediction, linear interpolation, no Karras sigmas")
        if timestep_spacing not in ("linspace", "leading", "trailing"):
            raise ValueError(f"timestep_spacing {timestep_spacing}")
        self.config = SimpleNamespace(num_train_timesteps=num_train_timesteps, beta_start=beta_start, beta_end=beta_end,
                                      beta_schedule=beta_schedule, prediction_type=prediction_type,
                                      timestep_spacing=timestep_spacing, steps_offset=steps_offset)
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self._train_sigmas = (((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5).numpy()
        self.num_inference_steps: Optional[int] = None
        self._set(np.linspace(0, num_train_timesteps - 1, num_train_timesteps, dtype=float)[::-1].copy(),
                  self._train_sigmas[::-1].copy(), None)

    def _set(self, timesteps: np.ndarray, sigmas: np.ndarray, device):
        sig = np.concatenate([sigmas, [0.0]]).astype(np.float32)
        self.sigmas = torch.from_numpy(sig).to(device)
        self.timesteps = torch.from_numpy(timesteps).to(device)
        self.timesteps_host = [float(v) for v in timesteps]   # host mirrors: no device read-back inside the loops
        self._sigmas_host = [float(v) for v in sig]

    @property
    def init_noise_sigma(self) -> float:
        m = max(self._sigmas_host)
        return m if self.config.timestep_spacing in ("linspace", "trailing") else (m * m + 1.0) ** 0.5

    def set_timesteps(self, num_inference_steps: int, device: Union[str, torch.device, None] = None):
        T = self.config.num_train_timesteps
        self.num_inference_steps = num_inference_steps
        sp = self.config.timestep_spacing
        if sp == "linspace":
            timesteps = np.linspace(0, T - 1, num_inference_steps, dtype=float)[::-1].copy()
        elif sp == "leading":
            timesteps = (np.arange(0, num_inference_steps) * (T // num_inference_steps)).round()[::-1].copy().astype(float)
            timesteps += self.config.steps_offset
        else:
            timesteps = (np.arange(T, 0, -T / num_inference_steps)).round().copy().astype(float) - 1
        sigmas = np.interp(timesteps, np.arange(0, len(self._train_sigmas)), self._train_sigmas)
        self._set(timesteps, sigmas, device)

    def _index(self, timestep) -> int:
        return self.timesteps_host.index(float(timestep))

    def _step_coeffs(self, timestep) -> tuple:
        """(cx, ce) of x_next = cx x + ce eps."""
        i = self._index(timestep)
        return 1.0, self._sigmas_host[i + 1] - self._sigmas_host[i]

    def scale_model_input(self, sample: torch.Tensor, timestep) -> torch.Tensor:
        sigma = self._sigmas_host[self._index(timestep)]
        return sample / ((sigma * sigma + 1.0) ** 0.5)

    def step(self, model_output: torch.Tensor, timestep, sample: torch.Tensor, return_dict: bool = True, **unused):
        if self.num_inference_steps is None:
            raise ValueError("call set_timesteps first")
        cx, ce = self._step_coeffs(timestep)
        if model_output.is_cuda:
            out_dtype = sample.dtype if sample.dtype in (torch.float32, torch.bfloat16) else torch.float32
            _, prev = ops.cfg_ddim(model_output.contiguous(), 0.0, sample.contiguous(), cx, ce, out_dtype=out_dtype,
                                   single=True, affine=True)
            prev = prev.to(sample.dtype)
        else:  # host tensors (scheduler unit tests); not a model path
            prev = cx * sample + ce * model_output
        if not return_dict:
            return (prev,)
        return DDIMSchedulerOutput(prev_sample=prev)

    def __len__(self):
        return self.config.num_train_timesteps


class LMSDiscreteScheduler(EulerDiscreteScheduler):
    """diffusers' `LMSDiscreteScheduler` as eval-scripts/generate_images_sd1.py:51 builds it (scaled_linear betas
    0.00085..0.012, 1000 train steps; linspace timesteps) and drives it (:169-192): linear multistep (order 4) on the
    same sigma grid as Euler.  x_next = x + sum_k c_k d_{i-k} with d = eps (epsilon prediction) and c_k the integral of
    the k-th Lagrange basis polynomial over [sigma_i, sigma_{i+1}] (scipy.integrate.quad, epsrel 1e-4, as diffusers).
    The combine is a 4-term linear combination of [N,4,64,64] latents: plain torch ops on whatever device they live on."""

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.00085, beta_end: float = 0.012,
                 beta_schedule: str = "scaled_linear", prediction_type: str = "epsilon",
                 timestep_spacing: str = "linspace", steps_offset: int = 0):
        # (the class attribute `step_kind = "generic"` below keeps the denoise loops on their scheduler.step branch)
        super().__init__(num_train_timesteps, beta_start, beta_end, beta_schedule, prediction_type, "linear", False,
                         timestep_spacing, steps_offset)
        self.derivatives = []

    def set_timesteps(self, num_inference_steps: int, device: Union[str, torch.device, None] = None):
        super().set_timesteps(num_inference_steps, device)
        self.derivatives = []

    def get_lms_coefficient(self, order: int, t: int, current_order: int) -> float:
        from scipy import integrate

        sig = self._sigmas_host

        def lms_derivative(tau):
            prod = 1.0
            for k in range(order):
                if current_order == k:
                    continue
                prod *= (tau - sig[t - k]) / (sig[t - current_order] - sig[t - k])
            return prod

        return integrate.quad(lms_derivative, sig[t], sig[t + 1], epsrel=1e-4)[0]

    step_kind = "generic"  # multistep: not an affine function of (x, eps) alone

    def step(self, model_output: torch.Tensor, timestep, sample: torch.Tensor, order: int = 4,
             return_dict: bool = True, **unused):
        if self.num_inference_steps is None:
            raise ValueError("call set_timesteps first")
        i = self._index(timestep)
        sigma = self._sigmas_host[i]
        pred_original_sample = sample - sigma * model_output
        self.derivatives.append((sample - pred_original_sample) / sigma)
        if len(self.derivatives) > order:
            self.derivatives.pop(0)
        order = min(i + 1, order)
        coeffs = [self.get_lms_coefficient(order, i, k) for k in range(order)]
        prev = sample + sum(c * d for c, d in zip(coeffs, reversed(self.derivatives)))
        if not return_dict:
            return (prev,)
        return DDIMSchedulerOutput(prev_sample=prev)


def _noise_like(model_output: torch.Tensor, generator: Optional[torch.Generator]) -> torch.Tensor:
    """diffusers' `randn_tensor(model_output.shape, generator=…, device=…, dtype=…)`: drawn on the generator's device
    (or the tensor's, from the global RNG), then moved."""
    dev = generator.device if generator is not None else model_output.device
    return torch.randn(model_output.shape, generator=generator, device=dev, dtype=model_output.dtype).to(model_output.device)


class DDPMScheduler(DDIMScheduler):
    """diffusers' `DDPMScheduler` as model_util.py:247-256 builds it (scaled_linear betas, 1000 train steps,
    clip_sample=False, epsilon prediction; library defaults: variance_type "fixed_small", "leading" spacing) — the
    ancestral sampler behind `train.noise_scheduler: "ddpm"`.  Same grid, `add_noise`, `init_noise_sigma` and
    `scale_model_input` as DDIM; the step is the posterior mean, affine in (x, eps) with host-side coefficients (one
    `cfg_ddim_kernel` launch in its affine mode), plus sqrt(posterior variance) * z for t > 0."""
    step_kind = "affine+noise"

    def _abar(self, timestep):
        t = int(timestep)
        prev = t - self.config.num_train_timesteps // self.num_inference_steps
        return t, self._acp[t], (self._acp[prev] if prev >= 0 else 1.0)

    def _step_coeffs(self, timestep) -> tuple:
        """(cx, ce) of the posterior mean  cx x + ce eps."""
        _, a_t, a_prev = self._abar(timestep)
        b_t, b_prev = 1.0 - a_t, 1.0 - a_prev
        cur_a = a_t / a_prev
        c0 = (a_prev ** 0.5) * (1.0 - cur_a) / b_t          # weight of x0 = (x - sqrt(b_t) eps) / sqrt(a_t)
        return c0 / a_t ** 0.5 + cur_a ** 0.5 * b_prev / b_t, -c0 * (b_t / a_t) ** 0.5

    def _noise_std(self, timestep) -> float:
        t, a_t, a_prev = self._abar(timestep)
        if t <= 0:
            return 0.0
        return max((1.0 - a_prev) / (1.0 - a_t) * (1.0 - a_t / a_prev), 1e-20) ** 0.5

    def step(self, model_output: torch.Tensor, timestep, sample: torch.Tensor, generator: Optional[torch.Generator] = None,
             return_dict: bool = True, **unused):
        if self.num_inference_steps is None:
            raise ValueError("call set_timesteps first")
        cx, ce = self._step_coeffs(timestep)
        if model_output.is_cuda:
            out_dtype = sample.dtype if sample.dtype in (torch.float32, torch.bfloat16) else torch.float32
            _, prev = ops.cfg_ddim(model_output.contiguous(), 0.0, sample.contiguous(), cx, ce, out_dtype=out_dtype,
                                   single=True, affine=True)
            prev = prev.to(sample.dtype)
        else:  # host tensors (scheduler unit tests); not a model path
            prev = cx * sample + ce * model_output
        std = self._noise_std(timestep)
        if std > 0.0:
            prev = prev + std * _noise_like(model_output, generator).to(prev.dtype)
        if not return_dict:
            return (prev,)
        return DDIMSchedulerOutput(prev_sample=prev)


class EulerAncestralDiscreteScheduler(EulerDiscreteScheduler):
    """diffusers' `EulerAncestralDiscreteScheduler` as model_util.py:266-274 builds it (scaled_linear betas, "linspace"
    spacing): Euler step to sigma_down plus sigma_up * z.  The deterministic part is the affine mode of `cfg_ddim_kernel`."""
    step_kind = "affine+noise"

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.00085, beta_end: float = 0.012,
                 beta_schedule: str = "scaled_linear", prediction_type: str = "epsilon",
                 timestep_spacing: str = "linspace", steps_offset: int = 0):
        super().__init__(num_train_timesteps, beta_start, beta_end, beta_schedule, prediction_type, "linear", False,
                         timestep_spacing, steps_offset)

    def _sigmas_at(self, timestep):
        i = self._index(timestep)
        s_from, s_to = self._sigmas_host[i], self._sigmas_host[i + 1]
        up = (s_to * s_to * (s_from * s_from - s_to * s_to) / (s_from * s_from)) ** 0.5
        return s_from, (s_to * s_to - up * up) ** 0.5, up

    def _step_coeffs(self, timestep) -> tuple:
        s_from, s_down, _ = self._sigmas_at(timestep)
        return 1.0, s_down - s_from

    def _noise_std(self, timestep) -> float:
        return self._sigmas_at(timestep)[2]

    def step(self, model_output: torch.Tensor, timestep, sample: torch.Tensor, generator: Optional[torch.Generator] = None,
             return_dict: bool = True, **unused):
        out = super().step(model_output, timestep, sample, return_dict=False)[0]
        std = self._noise_std(timestep)
        if std > 0.0:
            out = out + std * _noise_like(model_output, generator).to(out.dtype)
        if not return_dict:
            return (out,)
        return DDIMSchedulerOutput(prev_sample=out)


def create_noise_scheduler(scheduler_name: str = "ddim", prediction_type: str = "epsilon"):
    """model_util.create_noise_scheduler (model_util.py:230-278): "ddim", "ddpm", "lms", "euler_a" as the reference's
    factory builds them, plus "euler" (EulerDiscrete, what the SDXL eval pipeline ships with)."""
    name = scheduler_name.lower().replace(" ", "_")
    common = dict(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", num_train_timesteps=1000,
                  prediction_type=prediction_type)
    if name == "ddim":
        return DDIMScheduler(clip_sample=False, **common)
    if name == "ddpm":
        return DDPMScheduler(clip_sample=False, **common)
    if name == "lms":
        return LMSDiscreteScheduler(**common)
    if name == "euler_a":
        return EulerAncestralDiscreteScheduler(**common)
    if name == "euler":
        return EulerDiscreteScheduler(**common)
    raise ValueError(f"Unknown scheduler name: {name}")
