"""DDIM scheduler with the surface the reference uses (generate.py:68-76; models/pipelines.py:382,410,447,598,
630,830): set_timesteps / timesteps / scale_model_input / step / add_noise / init_noise_sigma / alphas_cumprod.

Host side it is plain arithmetic; for the fused CUDA step it also exports per-step coefficient rows
{guidance, sqrt(a_t), sqrt(1-a_t), sqrt(a_prev), sqrt(1-a_prev)} that `tg_cfg_ddim_composite` reads from device
memory, so a captured step graph can be replayed for every timestep.
"""
from __future__ import annotations

import numpy as np
import torch


class _Cfg(dict):
    __getattr__ = dict.__getitem__


class _StepOutput:
    def __init__(self, prev_sample):
        self.prev_sample = prev_sample

    def __getitem__(self, i):
        return (self.prev_sample,)[i]


class DDIMScheduler:
    init_noise_sigma = 1.0

    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                 clip_sample=False, set_alpha_to_one=False, steps_offset=1):
        assert beta_schedule == "scaled_linear" and not clip_sample
        self.config = _Cfg(num_train_timesteps=num_train_timesteps, beta_start=beta_start, beta_end=beta_end,
                           beta_schedule=beta_schedule, clip_sample=clip_sample, set_alpha_to_one=set_alpha_to_one,
                           steps_offset=steps_offset)
        self.betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        self.alphas_cumprod = torch.cumprod(1.0 - self.betas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))

    def set_timesteps(self, num_inference_steps, device=None):
        self.num_inference_steps = num_inference_steps
        ratio = self.config.num_train_timesteps // num_inference_steps
        ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64)
        self.timesteps = torch.from_numpy(ts + self.config.steps_offset)
        if device is not None:
            self.timesteps = self.timesteps.to(device)

    def scale_model_input(self, sample, timestep=None):
        return sample

    def alphas_for(self, t):
        t = int(t)
        prev = t - self.config.num_train_timesteps // self.num_inference_steps
        a_t = float(self.alphas_cumprod[t])
        a_prev = float(self.alphas_cumprod[prev]) if prev >= 0 else float(self.final_alpha_cumprod)
        return a_t, a_prev

    def coef_row(self, t, guidance_scale):
        a_t, a_prev = self.alphas_for(t)
        return [float(guidance_scale), a_t ** 0.5, (1 - a_t) ** 0.5, a_prev ** 0.5, (1 - a_prev) ** 0.5]

    def coef_table(self, guidance_scale, device):
        """[num_inference_steps, 8] float32 rows (5 used) for tg_cfg_ddim_composite."""
        rows = [self.coef_row(int(t), guidance_scale) + [0.0, 0.0, 0.0] for t in self.timesteps]
        return torch.tensor(rows, dtype=torch.float32, device=device)

    def step(self, model_output, timestep, sample, **kwargs):
        """Eager fallback-free reference arithmetic on whatever device the tensors live on (used only outside the
        fused loop, e.g. by callers that drive the scheduler themselves)."""
        a_t, a_prev = self.alphas_for(timestep)
        x0 = (sample - (1 - a_t) ** 0.5 * model_output) / a_t ** 0.5
        return _StepOutput(a_prev ** 0.5 * x0 + (1 - a_prev) ** 0.5 * model_output)

    def add_noise(self, original_samples, noise, timesteps):
        a = self.alphas_cumprod.to(original_samples.device)[timesteps.to(original_samples.device).long()]
        a = a.to(original_samples.dtype)
        sa, s1a = a ** 0.5, (1 - a) ** 0.5
        while sa.dim() < original_samples.dim():
            sa, s1a = sa.unsqueeze(-1), s1a.unsqueeze(-1)
        return sa * original_samples + s1a * noise
