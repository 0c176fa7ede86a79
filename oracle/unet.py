"""Functional restatement of the SD-1.5 UNet forward (TEST INFRASTRUCTURE, see oracle/__init__.py).

Weights live in a flat dict keyed by the diffusers state-dict names (plus
`<...>.attn2.processor.to_{k,v}_ip.weight` for the IP-Adapter branch), so real checkpoints drop in.

Wiring follows the vendored definition in /root/reference:
  UNet2DConditionModel.forward       ip_adapter/unet_2d_condition.py:799-1125
  CrossAttnDownBlock2D.forward       ip_adapter/unet_2d_blocks.py:1045-1116
  DownBlock2D.forward                ip_adapter/unet_2d_blocks.py:1171-1202
  UNetMidBlock2DCrossAttn.forward    ip_adapter/unet_2d_blocks.py:658-712
  CrossAttnUpBlock2D.forward         ip_adapter/unet_2d_blocks.py:2190-2254
  UpBlock2D.forward                  ip_adapter/unet_2d_blocks.py:2304-2336
  Transformer2DModel.forward         ip_adapter/transformer_2d.py:216-373
  BasicTransformerBlock.forward      ip_adapter/attention.py:174-266, FeedForward/GEGLU :269-367
Third-party arithmetic restated from diffusers 0.21.4 (requirements.txt:13), NOT present under
/root/reference and therefore unpinned: ResnetBlock2D, Downsample2D, Upsample2D, Timesteps,
TimestepEmbedding.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from . import attention as A

BLOCK_OUT = (320, 640, 1280, 1280)
LAYERS_PER_BLOCK = 2
GROUPS = 32


# ---------------------------------------------------------------------------------------------------------
# structure helpers (shared with tests / synthetic weight generation)
# ---------------------------------------------------------------------------------------------------------
def resnet_specs():
    """[(prefix, cin, cout)] for the 22 ResnetBlock2D in execution order."""
    out = []
    ch = BLOCK_OUT[0]
    skips = [ch]
    for i, co in enumerate(BLOCK_OUT):
        for j in range(LAYERS_PER_BLOCK):
            out.append((f"down_blocks.{i}.resnets.{j}", ch, co))
            ch = co
            skips.append(ch)
        if i < 3:
            skips.append(ch)
    out.append(("mid_block.resnets.0", ch, ch))
    out.append(("mid_block.resnets.1", ch, ch))
    rev = list(reversed(BLOCK_OUT))
    for i, co in enumerate(rev):
        for j in range(LAYERS_PER_BLOCK + 1):
            skip = skips.pop()
            out.append((f"up_blocks.{i}.resnets.{j}", ch + skip, co))
            ch = co
    return out


def transformer_specs():
    """[(prefix, C, attn_key)] for the 16 Transformer2DModel in *processor order*
    (named_children: down_blocks, up_blocks, mid_block -- ip_adapter/unet_2d_condition.py:460-461,532;
    the order ip_adapter.{1,3,..,31} weights are loaded in, ip_adapter/ip_adapter.py:139-140)."""
    out = []
    for i in range(3):
        for j in range(2):
            out.append((f"down_blocks.{i}.attentions.{j}", BLOCK_OUT[i], ("down", i, j, 0)))
    rev = list(reversed(BLOCK_OUT))
    for i in range(1, 4):
        for j in range(3):
            out.append((f"up_blocks.{i}.attentions.{j}", rev[i], ("up", i, j, 0)))
    out.append(("mid_block.attentions.0", 1280, ("mid", 0, 0, 0)))
    return out


# ---------------------------------------------------------------------------------------------------------
# third-party blocks (diffusers 0.21.4), restated
# ---------------------------------------------------------------------------------------------------------
def timestep_embedding(t, dim=320, dtype=torch.float32):
    """Timesteps(320, flip_sin_to_cos=True, downscale_freq_shift=0): fp32 sinusoid, cos first, cast to dtype
    (ip_adapter/unet_2d_condition.py:897-902)."""
    half = dim // 2
    exponent = -math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device) / half
    emb = t.float()[:, None] * torch.exp(exponent)[None, :]
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
    emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)  # flip_sin_to_cos
    return emb.to(dtype)


def _conv(x, sd, name, stride=1, padding=1):
    return F.conv2d(x, sd[name + ".weight"], sd[name + ".bias"], stride=stride, padding=padding)


def _lin(x, sd, name):
    return F.linear(x, sd[name + ".weight"], sd.get(name + ".bias"))


def _gn(x, sd, name, eps):
    return F.group_norm(x, GROUPS, sd[name + ".weight"], sd[name + ".bias"], eps)


def resnet_block(sd, p, x, emb):
    """ResnetBlock2D(groups=32, eps=1e-5, silu, time_embedding_norm='default', output_scale_factor=1)."""
    h = _conv(F.silu(_gn(x, sd, p + ".norm1", 1e-5)), sd, p + ".conv1")
    h = h + _lin(F.silu(emb), sd, p + ".time_emb_proj")[:, :, None, None]
    h = _conv(F.silu(_gn(h, sd, p + ".norm2", 1e-5)), sd, p + ".conv2")
    if (p + ".conv_shortcut.weight") in sd:
        x = _conv(x, sd, p + ".conv_shortcut", padding=0)
    return x + h


def downsample(sd, p, x):
    return _conv(x, sd, p + ".conv", stride=2, padding=1)


def upsample(sd, p, x):
    return _conv(F.interpolate(x, scale_factor=2.0, mode="nearest"), sd, p + ".conv")


# ---------------------------------------------------------------------------------------------------------
# vendored blocks (ip_adapter/attention.py, transformer_2d.py)
# ---------------------------------------------------------------------------------------------------------
def _ln(x, sd, name):
    return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"], 1e-5)


def basic_transformer_block(sd, p, x, ctx, ip_scale, num_ip_tokens, capture, attn_key, save_keys):
    x = A.attn_self(sd, p + ".attn1", _ln(x, sd, p + ".norm1")) + x
    want = capture is not None and (save_keys is None or attn_key in save_keys)
    r = A.attn_cross_ip(sd, p + ".attn2", _ln(x, sd, p + ".norm2"), ctx, ip_scale, num_ip_tokens,
                        return_probs=want)
    if want:
        r, probs = r
        capture[attn_key] = probs
    x = r + x
    h = _lin(_ln(x, sd, p + ".norm3"), sd, p + ".ff.net.0.proj")
    val, gate = h.chunk(2, dim=-1)
    x = _lin(val * F.gelu(gate), sd, p + ".ff.net.2") + x
    return x


def transformer_2d(sd, p, x, ctx, ip_scale, num_ip_tokens, capture, attn_key, save_keys):
    b, c, hh, ww = x.shape
    res = x
    h = _gn(x, sd, p + ".norm", 1e-6)
    h = _conv(h, sd, p + ".proj_in", padding=0)
    h = h.permute(0, 2, 3, 1).reshape(b, hh * ww, c)
    h = basic_transformer_block(sd, p + ".transformer_blocks.0", h, ctx, ip_scale, num_ip_tokens, capture,
                                attn_key, save_keys)
    h = h.reshape(b, hh, ww, c).permute(0, 3, 1, 2).contiguous()
    return _conv(h, sd, p + ".proj_out", padding=0) + res


# ---------------------------------------------------------------------------------------------------------
# the UNet
# ---------------------------------------------------------------------------------------------------------
def unet_forward(sd, sample, t, ctx, ip_scale=1.0, num_ip_tokens=4, down_residuals=None, mid_residual=None,
                 capture=None, save_keys=None):
    """sample [B,4,H,W], t scalar (int / 0-d tensor), ctx [B, 77+num_ip_tokens, 768] -> eps [B,4,H,W].

    capture: optional dict filled with {attn_key: probs[B,8,N,77]} for keys in save_keys
    (the reference's save_attn_to_dict protocol, ip_adapter/attention_processor.py:532-549)."""
    dtype = sample.dtype
    b = sample.shape[0]
    tt = torch.as_tensor(t, device=sample.device).reshape(-1).expand(b)
    emb = timestep_embedding(tt, 320, dtype)
    emb = _lin(F.silu(_lin(emb, sd, "time_embedding.linear_1")), sd, "time_embedding.linear_2")

    def tf(prefix, x, key):
        return transformer_2d(sd, prefix, x, ctx, ip_scale, num_ip_tokens, capture, key, save_keys)

    x = _conv(sample, sd, "conv_in")
    skips = [x]
    for i in range(4):
        for j in range(LAYERS_PER_BLOCK):
            x = resnet_block(sd, f"down_blocks.{i}.resnets.{j}", x, emb)
            if i < 3:
                x = tf(f"down_blocks.{i}.attentions.{j}", x, ("down", i, j, 0))
            skips.append(x)
        if i < 3:
            x = downsample(sd, f"down_blocks.{i}.downsamplers.0", x)
            skips.append(x)
    if down_residuals is not None:  # ControlNet (unet_2d_condition.py:1044-1053)
        skips = [s + r for s, r in zip(skips, down_residuals)]

    x = resnet_block(sd, "mid_block.resnets.0", x, emb)
    x = tf("mid_block.attentions.0", x, ("mid", 0, 0, 0))
    x = resnet_block(sd, "mid_block.resnets.1", x, emb)
    if mid_residual is not None:
        x = x + mid_residual

    for i in range(4):
        for j in range(LAYERS_PER_BLOCK + 1):
            x = torch.cat([x, skips.pop()], dim=1)
            x = resnet_block(sd, f"up_blocks.{i}.resnets.{j}", x, emb)
            if i > 0:
                x = tf(f"up_blocks.{i}.attentions.{j}", x, ("up", i, j, 0))
        if i < 3:
            x = upsample(sd, f"up_blocks.{i}.upsamplers.0", x)

    x = F.silu(_gn(x, sd, "conv_norm_out", 1e-5))
    return _conv(x, sd, "conv_out")


# ---------------------------------------------------------------------------------------------------------
# seeded synthetic SD-1.5-shaped weights (no checkpoints exist offline)
# ---------------------------------------------------------------------------------------------------------
def synth_state_dict(seed=1234, dtype=torch.float32, with_ip=True, cross_dim=768, gain=1.0):
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def lin(name, out_f, in_f, bias=True, scale=1.0):
        sd[name + ".weight"] = (torch.randn(out_f, in_f, generator=g) * (scale * gain / math.sqrt(in_f))).to(dtype)
        if bias:
            sd[name + ".bias"] = (torch.randn(out_f, generator=g) * 0.02).to(dtype)

    def conv(name, out_c, in_c, k):
        sd[name + ".weight"] = (torch.randn(out_c, in_c, k, k, generator=g) * (gain / math.sqrt(in_c * k * k))).to(dtype)
        sd[name + ".bias"] = (torch.randn(out_c, generator=g) * 0.02).to(dtype)

    def norm(name, c):
        sd[name + ".weight"] = (1.0 + 0.1 * torch.randn(c, generator=g)).to(dtype)
        sd[name + ".bias"] = (0.05 * torch.randn(c, generator=g)).to(dtype)

    lin("time_embedding.linear_1", 1280, 320)
    lin("time_embedding.linear_2", 1280, 1280)
    conv("conv_in", 320, 4, 3)
    for p, cin, cout in resnet_specs():
        norm(p + ".norm1", cin)
        conv(p + ".conv1", cout, cin, 3)
        lin(p + ".time_emb_proj", cout, 1280)
        norm(p + ".norm2", cout)
        conv(p + ".conv2", cout, cout, 3)
        if cin != cout:
            conv(p + ".conv_shortcut", cout, cin, 1)
    for p, c, _ in transformer_specs():
        norm(p + ".norm", c)
        conv(p + ".proj_in", c, c, 1)
        conv(p + ".proj_out", c, c, 1)
        tb = p + ".transformer_blocks.0"
        for n in ("norm1", "norm2", "norm3"):
            norm(f"{tb}.{n}", c)
        for a, kd in (("attn1", c), ("attn2", cross_dim)):
            lin(f"{tb}.{a}.to_q", c, c, bias=False)
            lin(f"{tb}.{a}.to_k", c, kd, bias=False)
            lin(f"{tb}.{a}.to_v", c, kd, bias=False)
            lin(f"{tb}.{a}.to_out.0", c, c, scale=0.5)
        if with_ip:
            lin(f"{tb}.attn2.processor.to_k_ip", c, cross_dim, bias=False)
            lin(f"{tb}.attn2.processor.to_v_ip", c, cross_dim, bias=False)
        lin(f"{tb}.ff.net.0.proj", 8 * c, c)
        lin(f"{tb}.ff.net.2", c, 4 * c, scale=0.5)
    for i in range(3):
        conv(f"down_blocks.{i}.downsamplers.0.conv", BLOCK_OUT[i], BLOCK_OUT[i], 3)
    rev = list(reversed(BLOCK_OUT))
    for i in range(3):
        conv(f"up_blocks.{i}.upsamplers.0.conv", rev[i], rev[i], 3)
    norm("conv_norm_out", 320)
    conv("conv_out", 4, 320, 3)
    return sd


def count_params(sd):
    return sum(v.numel() for v in sd.values())
