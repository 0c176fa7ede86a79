"""DDIMScheduler restated (TEST INFRASTRUCTURE, see oracle/__init__.py).

Third-party: diffusers==0.21.4 `DDIMScheduler` (requirements.txt:13), configured at generate.py:68-76:
num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", clip_sample=False,
set_alpha_to_one=False, steps_offset=1 (timestep_spacing="leading", prediction_type="epsilon", eta=0).
Restated from the published algorithm; anchored on the call sites models/pipelines.py:382,410,447,
598,630,830.  Unpinned by any reference test; pinned here by closed-form properties (tests/test_oracle_cpu.py).
"""
from __future__ import annotations

import numpy as np
import torch


class DDIM:
    init_noise_sigma = 1.0

    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, steps_offset=1):
        self.num_train_timesteps = num_train_timesteps
        self.betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        self.alphas_cumprod = torch.cumprod(1.0 - self.betas, dim=0)
        self.final_alpha_cumprod = self.alphas_cumprod[0]  # set_alpha_to_one=False
        self.steps_offset = steps_offset
        self.timesteps = None
        self.num_inference_steps = None

    def set_timesteps(self, n, device=None):
        self.num_inference_steps = n
        step_ratio = self.num_train_timesteps // n
        ts = (np.arange(0, n) * step_ratio).round()[::-1].copy().astype(np.int64) + self.steps_offset
        self.timesteps = torch.from_numpy(ts)
        return self.timesteps

    def scale_model_input(self, sample, timestep=None):
        return sample

    def _alphas(self, t):
        t = int(t)
        prev = t - self.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev] if prev >= 0 else self.final_alpha_cumprod
        return a_t, a_prev

    def step(self, eps, t, sample):
        """x0 = (x - sqrt(1-a_t) eps)/sqrt(a_t);  x' = sqrt(a_prev) x0 + sqrt(1-a_prev) eps  (eta = 0).
        Scalars are fp32 0-d CPU tensors -> each op keeps the tensor dtype, exactly like diffusers."""
        a_t, a_prev = self._alphas(t)
        a_t, a_prev = a_t.item(), a_prev.item()
        x0 = (sample - (1 - a_t) ** 0.5 * eps) / a_t ** 0.5
        return a_prev ** 0.5 * x0 + (1 - a_prev) ** 0.5 * eps

    def add_noise(self, x, noise, timesteps):
        a = self.alphas_cumprod.to(x.device)[timesteps.to(x.device)].to(x.dtype)
        sa, s1a = a ** 0.5, (1 - a) ** 0.5
        while sa.dim() < x.dim():
            sa, s1a = sa.unsqueeze(-1), s1a.unsqueeze(-1)
        return sa * x + s1a * noise


def cfg(eps, guidance_scale):
    """models/pipelines.py:441-442 / 820-821"""
    u, c = eps.chunk(2)
    return u + guidance_scale * (c - u)


def stage1_loop(unet_fn, sched, latents, ctx, num_steps=50, guidance_scale=7.5):
    """models/pipelines.py:406-453: cat x2 -> unet -> CFG -> DDIM step; returns (latents, latents_all[T+1])."""
    sched.set_timesteps(num_steps)
    trace = [latents]
    for t in sched.timesteps:
        eps = unet_fn(torch.cat([latents] * 2), int(t), ctx)
        latents = sched.step(cfg(eps, guidance_scale), t, latents)
        trace.append(latents)
    return latents, torch.stack(trace)


def stage2_loop(unet_fn, sched, latents, ctx, frozen_all, frozen_mask, frozen_steps, num_steps=50, guidance_scale=7.5):
    """models/pipelines.py:742-835 (SD-1.5 branch) without the ControlNet call: the per-step masked composite
    latents = latents_all[index+1] * frozen_mask + latents * (1 - frozen_mask)   (:833-834)."""
    sched.set_timesteps(num_steps)
    for index, t in enumerate(sched.timesteps):
        eps = unet_fn(torch.cat([latents] * 2), int(t), ctx)
        latents = sched.step(cfg(eps, guidance_scale), t, latents)
        if index < frozen_steps:
            latents = frozen_all[index + 1] * frozen_mask + latents * (1.0 - frozen_mask)
    return latents
