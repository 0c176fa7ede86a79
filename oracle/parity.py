"""Full-trajectory parity checker (TEST INFRASTRUCTURE, see oracle/__init__.py): the reference's stage-1 / stage-2
denoise loops (models/pipelines.py:406-453, 742-835) run by the oracle on the GPU, in fp32 and in fp16 (= the
reference's own numerics), for BASELINE.json configs 2 and 3.

Only `tests/` and the parity / cpu_baseline / gpu_eager_reference legs of `bench.py` call this, and only as the
checker: it never produces anything the product path returns.  It takes plain tensors and state dicts with the
diffusers key names; it does not import the product package.

The north_star tolerance ("<= 1e-3 max-abs latent diff vs reference") is reported, not redefined: `summarize`
returns the measured max-abs drift of a trajectory against the fp32 oracle next to the drift of the fp16 oracle (what
diffusers would execute on this GPU) against the same fp32 oracle -- the noise floor of the reference itself.
"""
from __future__ import annotations

import torch

from . import controlnet as OC
from . import ddim as OD
from . import guidance as OG
from . import unet as OU

KEYS = [("mid", 0, 0, 0), ("up", 1, 0, 0), ("up", 1, 1, 0), ("up", 1, 2, 0)]


def _cast(sd, dtype):
    # both arms see the fp16-rounded weights the CUDA path consumes
    return {k: v.half().to(dtype) for k, v in sd.items()}


@torch.no_grad()
def stage1_traces(sd_unet, latents, ctx, ip_scales, steps, dtype, guidance_scale=7.5):
    """K independent characters (the reference runs them sequentially, theatergen.py:214-271).
    latents [K,4,64,64], ctx [2K,T,768] rows [0,K) unconditional / [K,2K) conditional, ip_scales list[K] of floats.
    -> trace [steps+1, K, 4, 64, 64] in `dtype`."""
    sd = _cast(sd_unet, dtype)
    K = latents.shape[0]
    out = []
    for k in range(K):
        c = torch.stack([ctx[k], ctx[K + k]]).to(dtype)
        fn = lambda x, t, c_, s=float(ip_scales[k]): OU.unet_forward(sd, x, t, c_, ip_scale=s)  # noqa: E731
        _, tr = OD.stage1_loop(fn, OD.DDIM(), latents[k:k + 1].to(dtype), c, num_steps=steps,
                               guidance_scale=guidance_scale)
        out.append(tr[:, 0])
    return torch.stack(out, 1)


@torch.no_grad()
def stage2_trace(sd_unet, sd_cnet, latents, ctx, ctrl_ctx, hint, frozen_all, frozen_mask, frozen_steps, steps, dtype,
                 ip_scale=0.1, guidance_scale=7.5, energy=None):
    """Stage 2 of one turn: ControlNet (77 text tokens, hint image) -> UNet with its residuals -> CFG -> DDIM ->
    masked composite for index < frozen_steps (models/pipelines.py:770-834).
    latents [1,4,64,64], ctx [2,T,768] (uncond, cond), ctrl_ctx [2,77,768], hint [1,3,512,512], frozen_all
    [steps+1,1,4,64,64], frozen_mask [64,64].  energy: optional dict(bboxes, positions, kwargs, upto) -> the guidance
    energy of the conditional half's captured maps is evaluated for index < upto.
    -> (trace [steps+1,1,4,64,64], energies list)"""
    sdu = _cast(sd_unet, dtype)
    sdc = _cast(sd_cnet, dtype) if sd_cnet is not None else None
    sched = OD.DDIM()
    sched.set_timesteps(steps)
    lat = latents.to(dtype)
    ctx, ctrl_ctx = ctx.to(dtype), ctrl_ctx.to(dtype)
    cond = hint.to(dtype).repeat(2, 1, 1, 1)
    fa, fm = frozen_all.to(dtype), frozen_mask.to(dtype)
    trace, energies = [lat], []
    for index, t in enumerate(sched.timesteps):
        x = torch.cat([lat] * 2)
        d = m = None
        if sdc is not None:
            d, m = OC.controlnet_forward(sdc, x, int(t), ctrl_ctx, cond)
        cap = {} if (energy is not None and index < energy["upto"]) else None
        eps = OU.unet_forward(sdu, x, int(t), ctx, ip_scale=ip_scale, down_residuals=d, mid_residual=m, capture=cap,
                              save_keys=KEYS)
        if cap is not None:
            saved = {k: v[1:2].float() for k, v in cap.items()}  # conditional half (return_cond_ca_only)
            energies.append(OG.compute_ca_lossv3(saved, energy["bboxes"], energy["positions"], KEYS,
                                                 **energy["kwargs"]).item())
        lat = sched.step(OD.cfg(eps, guidance_scale), t, lat)
        if index < frozen_steps:
            lat = fa[index + 1] * fm + lat * (1.0 - fm)
        trace.append(lat)
    return torch.stack(trace), energies


def summarize(ours, ref32, ref16):
    """ours / ref32 / ref16: traces [steps+1, K, 4, 64, 64].  Per-step max-abs drift against the fp32 oracle."""
    e = (ours.float() - ref32.float()).abs().amax(dim=(2, 3, 4))          # [steps+1, K]
    e16 = (ref16.float() - ref32.float()).abs().amax(dim=(2, 3, 4))
    rms = (ours.float() - ref32.float()).pow(2).mean(dim=(2, 3, 4)).sqrt()
    rms16 = (ref16.float() - ref32.float()).pow(2).mean(dim=(2, 3, 4)).sqrt()
    return {
        "steps": int(ours.shape[0] - 1), "trajectories": int(ours.shape[1]),
        "max_abs_latent_vs_fp32_oracle": float(e.max()), "fp16_oracle_gap": float(e16.max()),
        "final_max_abs": float(e[-1].max()), "final_fp16_oracle_gap": float(e16[-1].max()),
        "first_step_max_abs": float(e[1].max()), "first_step_fp16_oracle_gap": float(e16[1].max()),
        "rms_final": float(rms[-1].max()), "rms_final_fp16_oracle": float(rms16[-1].max()),
        "per_step_max_abs": [round(float(v), 5) for v in e.amax(1)],
        "per_step_fp16_oracle_gap": [round(float(v), 5) for v in e16.amax(1)],
    }
