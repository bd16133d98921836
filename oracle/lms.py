"""ORACLE (test infrastructure) — restatement of diffusers==0.20.2 `schedulers/scheduling_lms_discrete.py`
`LMSDiscreteScheduler` as eval-scripts/generate_images_sd1.py:51 constructs it (scaled_linear betas 0.00085..0.012, 1000
train steps; defaults otherwise: epsilon prediction, timestep_spacing "linspace", no Karras sigmas) and calls it
(:169 set_timesteps, :171 init_noise_sigma, :182 scale_model_input, :192 step):

    sigmas, timesteps, scale_model_input: as EulerDiscreteScheduler (oracle/euler.py), linspace grid
    step: d_i = (x - (x - sigma_i eps)) / sigma_i;  order = min(i + 1, 4)
          x_next = x + sum_{k < order} c_k d_{i-k},  c_k = integral_{sigma_i}^{sigma_{i+1}} prod_{j != k} (tau - sigma_{i-j}) / (sigma_{i-k} - sigma_{i-j}) dtau
          (scipy.integrate.quad, epsrel = 1e-4)

PARITY STATUS: unpinned by the reference (no tests there; diffusers not installable here); pinned by properties: the first
step is the Euler step, the coefficients of every step sum to sigma_{i+1} - sigma_i (the Lagrange basis sums to one), and
a derivative that is a cubic polynomial of sigma is integrated exactly from the fourth step on (tests/test_oracle.py).
Only tests/, smoke() and bench.py's CPU legs may import this.
"""
from __future__ import annotations

from types import SimpleNamespace

from scipy import integrate

from .euler import EulerDiscreteScheduler


class LMSDiscreteScheduler(EulerDiscreteScheduler):
    def __init__(self, num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear",
                 prediction_type="epsilon", timestep_spacing="linspace", steps_offset=0, **unused):
        super().__init__(num_train_timesteps, beta_start, beta_end, beta_schedule, prediction_type, "linear", False,
                         timestep_spacing, steps_offset)
        self.derivatives = []

    def set_timesteps(self, num_inference_steps, device=None):
        super().set_timesteps(num_inference_steps, device)
        self.derivatives = []

    def get_lms_coefficient(self, order, t, current_order):
        def lms_derivative(tau):
            prod = 1.0
            for k in range(order):
                if current_order == k:
                    continue
                prod *= (tau - self.sigmas[t - k]) / (self.sigmas[t - current_order] - self.sigmas[t - k])
            return float(prod)

        return integrate.quad(lms_derivative, float(self.sigmas[t]), float(self.sigmas[t + 1]), epsrel=1e-4)[0]

    def step(self, model_output, timestep, sample, order=4, return_dict=True, **unused):
        i = self._index(timestep)
        sigma = self.sigmas[i]
        pred_original_sample = sample - sigma * model_output
        derivative = (sample - pred_original_sample) / sigma
        self.derivatives.append(derivative)
        if len(self.derivatives) > order:
            self.derivatives.pop(0)
        order = min(i + 1, order)
        lms_coeffs = [self.get_lms_coefficient(order, i, curr_order) for curr_order in range(order)]
        prev = sample + sum(coeff * d for coeff, d in zip(lms_coeffs, reversed(self.derivatives)))
        return SimpleNamespace(prev_sample=prev) if return_dict else (prev,)
