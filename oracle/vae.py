"""Functional restatement of the SD-1.5 VAE, decoder and encoder (TEST INFRASTRUCTURE, see oracle/__init__.py).

The reference builds the stock diffusers `AutoencoderKL` (`stabilityai/sd-vae-ft-mse`, generate.py:59,77) and decodes
K + 1 latents per turn: `vae.decode(latents / scaling_factor, return_dict=False)[0]` at the end of stage 1
(models/pipelines.py:468) and `decode(vae, latents)` at the end of stage 2 (:162-173, :849-854).  SURVEY.md 8(f) row 2.
`diffusers==0.21.4` is NOT under /root/reference, so this restates the published algorithm of
`AutoencoderKL.decode` / `Decoder.forward` -- PARITY UNPINNED, anchored on those call sites:

  z    = post_quant_conv(z)                                  (1x1, 4 -> 4)
  h    = conv_in(z)                                          (3x3, 4 -> 512)
  mid  : ResnetBlock2D(512) -> Attention(1 head of 512, GroupNorm(32, eps 1e-6) in front, residual) -> ResnetBlock2D(512)
  up i : 3 x ResnetBlock2D (512,512 | 512,512 | 512->256,256 | 256->128,128), nearest x2 + conv3x3 after blocks 0..2
  out  = conv_out(silu(GroupNorm(32, eps 1e-6)(h)))          (3x3, 128 -> 3)
ResnetBlock2D here has no time embedding (temb_channels=None), GroupNorm eps 1e-6, 1x1 shortcut when channels change.
Encoder (`vae.encode(image).latent_dist.sample(generator)`, models/pipelines.py:624-626, one call per turn):
  h    = conv_in(x)                                          (3x3, 3 -> 128)
  down i: 2 x ResnetBlock2D (128 | 128->256 | 256->512 | 512), then F.pad(h, (0,1,0,1)) + conv3x3 stride 2 for i < 3
  mid  : as in the decoder;  moments = quant_conv(conv_out(silu(GroupNorm(h))))   (3x3 512 -> 8, then 1x1 8 -> 8)
  DiagonalGaussianDistribution: mean, logvar = chunk(moments, 2); logvar clamped to [-30, 20]; sample = mean + exp(0.5 logvar) * noise
State-dict keys are diffusers' (`post_quant_conv`, `decoder.conv_in`, `decoder.mid_block.{resnets,attentions}`,
`decoder.up_blocks.N.{resnets,upsamplers}`, `decoder.conv_norm_out`, `decoder.conv_out`; attention projections
`to_q/to_k/to_v/to_out.0` with `group_norm`), so a real checkpoint drops in.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

GROUPS = 32
UP_CHANNELS = (512, 512, 256, 128)  # reversed block_out_channels (128, 256, 512, 512)


def resnet_specs():
    """[(prefix, cin, cout)] of the decoder's 14 ResnetBlock2D in execution order."""
    out = [("decoder.mid_block.resnets.0", 512, 512), ("decoder.mid_block.resnets.1", 512, 512)]
    ch = 512
    for i, co in enumerate(UP_CHANNELS):
        for j in range(3):
            out.append((f"decoder.up_blocks.{i}.resnets.{j}", ch, co))
            ch = co
    return out


def _conv(x, sd, name, stride=1, padding=1):
    return F.conv2d(x, sd[name + ".weight"], sd[name + ".bias"], stride=stride, padding=padding)


def _gn(x, sd, name):
    return F.group_norm(x, GROUPS, sd[name + ".weight"], sd[name + ".bias"], 1e-6)


def resnet_block(sd, p, x):
    h = _conv(F.silu(_gn(x, sd, p + ".norm1")), sd, p + ".conv1")
    h = _conv(F.silu(_gn(h, sd, p + ".norm2")), sd, p + ".conv2")
    if (p + ".conv_shortcut.weight") in sd:
        x = _conv(x, sd, p + ".conv_shortcut", padding=0)
    return x + h


def attention_block(sd, p, x):
    """diffusers Attention(heads=1, dim_head=512, norm_num_groups=32, residual_connection=True, bias=True)"""
    b, c, hh, ww = x.shape
    h = _gn(x, sd, p + ".group_norm").reshape(b, c, hh * ww).transpose(1, 2)
    q = F.linear(h, sd[p + ".to_q.weight"], sd[p + ".to_q.bias"])
    k = F.linear(h, sd[p + ".to_k.weight"], sd[p + ".to_k.bias"])
    v = F.linear(h, sd[p + ".to_v.weight"], sd[p + ".to_v.bias"])
    probs = torch.baddbmm(torch.empty(b, hh * ww, hh * ww, dtype=q.dtype, device=q.device), q, k.transpose(1, 2), beta=0,
                          alpha=c ** -0.5).float().softmax(dim=-1).to(q.dtype)  # upcast_softmax=True
    o = F.linear(torch.bmm(probs, v), sd[p + ".to_out.0.weight"], sd[p + ".to_out.0.bias"])
    return o.transpose(1, 2).reshape(b, c, hh, ww) + x


def decode(sd, z):
    """z [B,4,64,64] (already divided by the scaling factor) -> image [B,3,512,512] in about [-1, 1]."""
    h = _conv(_conv(z, sd, "post_quant_conv", padding=0), sd, "decoder.conv_in")
    h = resnet_block(sd, "decoder.mid_block.resnets.0", h)
    h = attention_block(sd, "decoder.mid_block.attentions.0", h)
    h = resnet_block(sd, "decoder.mid_block.resnets.1", h)
    for i in range(4):
        for j in range(3):
            h = resnet_block(sd, f"decoder.up_blocks.{i}.resnets.{j}", h)
        if i < 3:
            h = _conv(F.interpolate(h, scale_factor=2.0, mode="nearest"), sd, f"decoder.up_blocks.{i}.upsamplers.0.conv")
    return _conv(F.silu(_gn(h, sd, "decoder.conv_norm_out")), sd, "decoder.conv_out")


def synth_state_dict(seed=777, dtype=torch.float32):
    """Seeded synthetic decoder weights with the diffusers AutoencoderKL key set (decoder half + post_quant_conv)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def conv(name, out_c, in_c, k, scale=1.0):
        sd[name + ".weight"] = (torch.randn(out_c, in_c, k, k, generator=g) * (scale / math.sqrt(in_c * k * k))).to(dtype)
        sd[name + ".bias"] = (torch.randn(out_c, generator=g) * 0.02).to(dtype)

    def lin(name, out_f, in_f, scale=1.0):
        sd[name + ".weight"] = (torch.randn(out_f, in_f, generator=g) * (scale / math.sqrt(in_f))).to(dtype)
        sd[name + ".bias"] = (torch.randn(out_f, generator=g) * 0.02).to(dtype)

    def norm(name, c):
        sd[name + ".weight"] = (1.0 + 0.1 * torch.randn(c, generator=g)).to(dtype)
        sd[name + ".bias"] = (0.05 * torch.randn(c, generator=g)).to(dtype)

    conv("post_quant_conv", 4, 4, 1, scale=2.0)
    conv("decoder.conv_in", 512, 4, 3)
    for p, cin, cout in resnet_specs():
        norm(p + ".norm1", cin)
        conv(p + ".conv1", cout, cin, 3)
        norm(p + ".norm2", cout)
        conv(p + ".conv2", cout, cout, 3, scale=0.5)
        if cin != cout:
            conv(p + ".conv_shortcut", cout, cin, 1)
    a = "decoder.mid_block.attentions.0"
    norm(a + ".group_norm", 512)
    for n in ("to_q", "to_k", "to_v"):
        lin(f"{a}.{n}", 512, 512, scale=2.0)
    lin(a + ".to_out.0", 512, 512, scale=0.5)
    for i in range(3):
        conv(f"decoder.up_blocks.{i}.upsamplers.0.conv", UP_CHANNELS[i], UP_CHANNELS[i], 3)
    norm("decoder.conv_norm_out", 128)
    conv("decoder.conv_out", 3, 128, 3)
    return sd


ENC_CHANNELS = (128, 256, 512, 512)


def enc_resnet_specs():
    out, ch = [], 128
    for i, co in enumerate(ENC_CHANNELS):
        for j in range(2):
            out.append((f"encoder.down_blocks.{i}.resnets.{j}", ch, co))
            ch = co
    return out + [("encoder.mid_block.resnets.0", 512, 512), ("encoder.mid_block.resnets.1", 512, 512)]


def encode_moments(sd, x):
    """x [B,3,512,512] in [-1, 1] -> moments [B,8,64,64] (mean | logvar before clamping)."""
    h = _conv(x, sd, "encoder.conv_in")
    for i in range(4):
        for j in range(2):
            h = resnet_block(sd, f"encoder.down_blocks.{i}.resnets.{j}", h)
        if i < 3:
            h = _conv(F.pad(h, (0, 1, 0, 1)), sd, f"encoder.down_blocks.{i}.downsamplers.0.conv", stride=2, padding=0)
    h = resnet_block(sd, "encoder.mid_block.resnets.0", h)
    h = attention_block(sd, "encoder.mid_block.attentions.0", h)
    h = resnet_block(sd, "encoder.mid_block.resnets.1", h)
    h = _conv(F.silu(_gn(h, sd, "encoder.conv_norm_out")), sd, "encoder.conv_out")
    return _conv(h, sd, "quant_conv", padding=0)


def sample_posterior(moments, noise):
    """DiagonalGaussianDistribution(moments).sample() with the noise supplied by the caller."""
    mean, logvar = moments.chunk(2, dim=1)
    std = torch.exp(0.5 * logvar.clamp(-30.0, 20.0))
    return mean + std * noise


def synth_encoder_state_dict(seed=778, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def conv(name, out_c, in_c, k, scale=1.0):
        sd[name + ".weight"] = (torch.randn(out_c, in_c, k, k, generator=g) * (scale / math.sqrt(in_c * k * k))).to(dtype)
        sd[name + ".bias"] = (torch.randn(out_c, generator=g) * 0.02).to(dtype)

    def lin(name, out_f, in_f, scale=1.0):
        sd[name + ".weight"] = (torch.randn(out_f, in_f, generator=g) * (scale / math.sqrt(in_f))).to(dtype)
        sd[name + ".bias"] = (torch.randn(out_f, generator=g) * 0.02).to(dtype)

    def norm(name, c):
        sd[name + ".weight"] = (1.0 + 0.1 * torch.randn(c, generator=g)).to(dtype)
        sd[name + ".bias"] = (0.05 * torch.randn(c, generator=g)).to(dtype)

    conv("encoder.conv_in", 128, 3, 3, scale=2.0)
    for p, cin, cout in enc_resnet_specs():
        norm(p + ".norm1", cin)
        conv(p + ".conv1", cout, cin, 3)
        norm(p + ".norm2", cout)
        conv(p + ".conv2", cout, cout, 3, scale=0.5)
        if cin != cout:
            conv(p + ".conv_shortcut", cout, cin, 1)
    for i in range(3):
        conv(f"encoder.down_blocks.{i}.downsamplers.0.conv", ENC_CHANNELS[i], ENC_CHANNELS[i], 3)
    a = "encoder.mid_block.attentions.0"
    norm(a + ".group_norm", 512)
    for n in ("to_q", "to_k", "to_v"):
        lin(f"{a}.{n}", 512, 512, scale=2.0)
    lin(a + ".to_out.0", 512, 512, scale=0.5)
    norm("encoder.conv_norm_out", 512)
    conv("encoder.conv_out", 8, 512, 3)
    conv("quant_conv", 8, 8, 1, scale=1.5)
    return sd
