"""Seeded synthetic SD-1.5-shaped weights generated on the device (no checkpoints exist offline).
Used by bench.py / smoke paths that must not touch oracle/.  Same key names as a diffusers state dict."""
from __future__ import annotations

import math

import torch

from .unet import BLOCK_OUT, _resnet_specs, _transformer_specs


def unet_shapes(with_ip=True, cross=768):
    s = {"time_embedding.linear_1.weight": (1280, 320), "time_embedding.linear_1.bias": (1280,),
         "time_embedding.linear_2.weight": (1280, 1280), "time_embedding.linear_2.bias": (1280,),
         "conv_in.weight": (320, 4, 3, 3), "conv_in.bias": (320,), "conv_norm_out.weight": (320,),
         "conv_norm_out.bias": (320,), "conv_out.weight": (4, 320, 3, 3), "conv_out.bias": (4,)}
    for p, c1, c2, co in _resnet_specs():
        cin = c1 + c2
        s[p + ".norm1.weight"], s[p + ".norm1.bias"] = (cin,), (cin,)
        s[p + ".conv1.weight"], s[p + ".conv1.bias"] = (co, cin, 3, 3), (co,)
        s[p + ".time_emb_proj.weight"], s[p + ".time_emb_proj.bias"] = (co, 1280), (co,)
        s[p + ".norm2.weight"], s[p + ".norm2.bias"] = (co,), (co,)
        s[p + ".conv2.weight"], s[p + ".conv2.bias"] = (co, co, 3, 3), (co,)
        if cin != co:
            s[p + ".conv_shortcut.weight"], s[p + ".conv_shortcut.bias"] = (co, cin, 1, 1), (co,)
    for p, c, _ in _transformer_specs():
        s[p + ".norm.weight"], s[p + ".norm.bias"] = (c,), (c,)
        s[p + ".proj_in.weight"], s[p + ".proj_in.bias"] = (c, c, 1, 1), (c,)
        s[p + ".proj_out.weight"], s[p + ".proj_out.bias"] = (c, c, 1, 1), (c,)
        tb = p + ".transformer_blocks.0"
        for n in ("norm1", "norm2", "norm3"):
            s[f"{tb}.{n}.weight"], s[f"{tb}.{n}.bias"] = (c,), (c,)
        for a, kd in (("attn1", c), ("attn2", cross)):
            s[f"{tb}.{a}.to_q.weight"] = (c, c)
            s[f"{tb}.{a}.to_k.weight"] = (c, kd)
            s[f"{tb}.{a}.to_v.weight"] = (c, kd)
            s[f"{tb}.{a}.to_out.0.weight"], s[f"{tb}.{a}.to_out.0.bias"] = (c, c), (c,)
        if with_ip:
            s[f"{tb}.attn2.processor.to_k_ip.weight"] = (c, cross)
            s[f"{tb}.attn2.processor.to_v_ip.weight"] = (c, cross)
        s[f"{tb}.ff.net.0.proj.weight"], s[f"{tb}.ff.net.0.proj.bias"] = (8 * c, c), (8 * c,)
        s[f"{tb}.ff.net.2.weight"], s[f"{tb}.ff.net.2.bias"] = (c, 4 * c), (c,)
    for i in range(3):
        c = BLOCK_OUT[i]
        s[f"down_blocks.{i}.downsamplers.0.conv.weight"] = (c, c, 3, 3)
        s[f"down_blocks.{i}.downsamplers.0.conv.bias"] = (c,)
    rev = list(reversed(BLOCK_OUT))
    for i in range(3):
        s[f"up_blocks.{i}.upsamplers.0.conv.weight"] = (rev[i], rev[i], 3, 3)
        s[f"up_blocks.{i}.upsamplers.0.conv.bias"] = (rev[i],)
    return s


def controlnet_shapes(cross=768):
    """diffusers ControlNetModel (SD-1.5) key set: the UNet's encoder half + conditioning embedder + zero convs."""
    u = unet_shapes(with_ip=False, cross=cross)
    s = {k: v for k, v in u.items() if not (k.startswith("up_blocks") or k.startswith("conv_norm_out") or k.startswith("conv_out"))}
    ch = (16, 32, 96, 256)
    p = "controlnet_cond_embedding"
    s[p + ".conv_in.weight"], s[p + ".conv_in.bias"] = (ch[0], 3, 3, 3), (ch[0],)
    for i in range(3):
        s[f"{p}.blocks.{2 * i}.weight"], s[f"{p}.blocks.{2 * i}.bias"] = (ch[i], ch[i], 3, 3), (ch[i],)
        s[f"{p}.blocks.{2 * i + 1}.weight"], s[f"{p}.blocks.{2 * i + 1}.bias"] = (ch[i + 1], ch[i], 3, 3), (ch[i + 1],)
    s[p + ".conv_out.weight"], s[p + ".conv_out.bias"] = (320, ch[3], 3, 3), (320,)
    for i, c in enumerate((320, 320, 320, 320, 640, 640, 640, 1280, 1280, 1280, 1280, 1280)):
        s[f"controlnet_down_blocks.{i}.weight"], s[f"controlnet_down_blocks.{i}.bias"] = (c, c, 1, 1), (c,)
    s["controlnet_mid_block.weight"], s["controlnet_mid_block.bias"] = (1280, 1280, 1, 1), (1280,)
    return s


def synth_controlnet_state_dict(device="cuda", seed=4321, fill=True):
    return synth_state_dict(device, seed, fill=fill, shapes=controlnet_shapes())


def synth_state_dict(device="cuda", seed=1234, with_ip=True, fill=True, shapes=None):
    """fill=False returns uninitialised tensors of the right shapes (ranks that receive the NCCL broadcast)."""
    g = torch.Generator(device=device).manual_seed(seed)
    out = {}
    for name, shp in (shapes or unet_shapes(with_ip)).items():
        if not fill:
            out[name] = torch.empty(shp, dtype=torch.float16, device=device)
            continue
        if len(shp) >= 2:
            fan_in = 1
            for d in shp[1:]:
                fan_in *= d
            scale = 0.5 if name.endswith(("to_out.0.weight", "ff.net.2.weight")) or name.startswith("controlnet_") else 1.0
            t = torch.randn(shp, generator=g, device=device) * (scale / math.sqrt(fan_in))
        elif name.endswith("weight") and ".norm" in name or name == "conv_norm_out.weight":
            t = 1.0 + 0.1 * torch.randn(shp, generator=g, device=device)
        else:
            t = 0.02 * torch.randn(shp, generator=g, device=device)
        out[name] = t.half()
    return out
