"""Seeded synthetic parameters shared by tests/golden/make_golden.py and the tests (TEST INFRASTRUCTURE).

`fill(shapes, seed)` is the single source of truth: the golden generator loads its output into the REFERENCE
modules, the tests feed the same tensors to the oracle / the CUDA path, so only seeds + outputs are committed.
"""
from __future__ import annotations

import math

import torch


def fill(shapes, seed):
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name in sorted(shapes):
        shp = tuple(shapes[name])
        leaf = name.split(".")[-1]
        if len(shp) >= 2:
            fan_in = 1
            for s in shp[1:]:
                fan_in *= s
            if name == "latents":
                out[name] = torch.randn(shp, generator=g) / math.sqrt(shp[-1])
            else:
                out[name] = torch.randn(shp, generator=g) / math.sqrt(fan_in)
        elif leaf == "weight":  # norm scale
            out[name] = 1.0 + 0.1 * torch.randn(shp, generator=g)
        else:
            out[name] = 0.05 * torch.randn(shp, generator=g)
    return out


def attention_shapes(C, cross=768):
    """state-dict shapes of one self-attention + one IP cross-attention (ip_adapter/attention_processor.py:113-128,
    418-419)."""
    s = {}
    for a, kd in (("attn1", C), ("attn2", cross)):
        s[f"{a}.to_q.weight"] = (C, C)
        s[f"{a}.to_k.weight"] = (C, kd)
        s[f"{a}.to_v.weight"] = (C, kd)
        s[f"{a}.to_out.0.weight"] = (C, C)
        s[f"{a}.to_out.0.bias"] = (C,)
    s["attn2.processor.to_k_ip.weight"] = (C, cross)
    s["attn2.processor.to_v_ip.weight"] = (C, cross)
    return s


def resampler_shapes(dim=768, depth=4, dim_head=64, heads=12, num_queries=16, embedding_dim=1280, output_dim=768,
                     ff_mult=4):
    """ip_adapter/resampler.py:81-125"""
    inner = dim_head * heads
    s = {"latents": (1, num_queries, dim), "proj_in.weight": (dim, embedding_dim), "proj_in.bias": (dim,),
         "proj_out.weight": (output_dim, dim), "proj_out.bias": (output_dim,), "norm_out.weight": (output_dim,),
         "norm_out.bias": (output_dim,)}
    for i in range(depth):
        p = f"layers.{i}"
        for n in ("norm1", "norm2"):
            s[f"{p}.0.{n}.weight"], s[f"{p}.0.{n}.bias"] = (dim,), (dim,)
        s[f"{p}.0.to_q.weight"] = (inner, dim)
        s[f"{p}.0.to_kv.weight"] = (2 * inner, dim)
        s[f"{p}.0.to_out.weight"] = (dim, inner)
        s[f"{p}.1.0.weight"], s[f"{p}.1.0.bias"] = (dim,), (dim,)
        s[f"{p}.1.1.weight"] = (dim * ff_mult, dim)
        s[f"{p}.1.3.weight"] = (dim, dim * ff_mult)
    return s


def image_proj_shapes(clip_dim=1024, cross=768, tokens=4):
    """ip_adapter/ip_adapter.py:30-39"""
    return {"proj.weight": (tokens * cross, clip_dim), "proj.bias": (tokens * cross,), "norm.weight": (cross,),
            "norm.bias": (cross,)}
