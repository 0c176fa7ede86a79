"""Image-token projectors restated functionally (TEST INFRASTRUCTURE, see oracle/__init__.py).

  ImageProjModel  /root/reference/ip_adapter/ip_adapter.py:30-47
  Resampler       /root/reference/ip_adapter/resampler.py:81-147 (PerceiverAttention :34-78, FeedForward :13-20)
State-dict keys are those of the reference nn.Modules (image_proj.* of the IP-Adapter checkpoints).
Pinned against the reference classes by tests/golden/make_golden.py.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def image_proj(sd, image_embeds, num_tokens=4, cross_dim=768):
    x = F.linear(image_embeds, sd["proj.weight"], sd["proj.bias"]).reshape(-1, num_tokens, cross_dim)
    return F.layer_norm(x, (cross_dim,), sd["norm.weight"], sd["norm.bias"], 1e-5)


def _ln(x, sd, name):
    return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"], 1e-5)


def perceiver_attention(sd, p, x, latents, heads, dim_head):
    x = _ln(x, sd, p + ".norm1")
    latents = _ln(latents, sd, p + ".norm2")
    b, l, _ = latents.shape
    q = F.linear(latents, sd[p + ".to_q.weight"])
    kv = F.linear(torch.cat((x, latents), dim=-2), sd[p + ".to_kv.weight"])
    k, v = kv.chunk(2, dim=-1)

    def rs(t):
        return t.view(b, t.shape[1], heads, -1).transpose(1, 2)

    q, k, v = rs(q), rs(k), rs(v)
    scale = 1 / math.sqrt(math.sqrt(dim_head))
    w = (q * scale) @ (k * scale).transpose(-2, -1)
    w = torch.softmax(w.float(), dim=-1).type(w.dtype)
    out = (w @ v).permute(0, 2, 1, 3).reshape(b, l, -1)
    return F.linear(out, sd[p + ".to_out.weight"])


def resampler(sd, x, depth=4, heads=12, dim_head=64):
    latents = sd["latents"].repeat(x.size(0), 1, 1)
    x = F.linear(x, sd["proj_in.weight"], sd["proj_in.bias"])
    for i in range(depth):
        latents = perceiver_attention(sd, f"layers.{i}.0", x, latents, heads, dim_head) + latents
        h = _ln(latents, sd, f"layers.{i}.1.0")
        h = F.linear(F.gelu(F.linear(h, sd[f"layers.{i}.1.1.weight"])), sd[f"layers.{i}.1.3.weight"])
        latents = h + latents
    latents = F.linear(latents, sd["proj_out.weight"], sd["proj_out.bias"])
    return _ln(latents, sd, "norm_out")
