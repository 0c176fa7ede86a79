"""Functional restatement of the SD-1.5 ControlNet forward used by stage 2 (TEST INFRASTRUCTURE, see
oracle/__init__.py).

The reference builds the STOCK diffusers `ControlNetModel` (generate.py:91, checkpoint
`ControlNet-1-1-preview/control_v11p_sd15_lineart`) and calls it once per stage-2 step
(models/pipelines.py:770-778: `controlnet(control_model_input, t, encoder_hidden_states=text_embeddings[2,77,768],
controlnet_cond=imagetight[2,3,512,512], conditioning_scale=1.0, guess_mode=False, return_dict=False)`); the 12 down
residuals and the mid residual feed the UNet (pipelines.py:802-816 -> ip_adapter/unet_2d_condition.py:1044-1053,
1077-1078).  `diffusers==0.21.4` (requirements.txt:13) is NOT under /root/reference and cannot be installed here, so
this file restates the published algorithm of `diffusers.models.controlnet.ControlNetModel.forward` -- PARITY
UNPINNED, anchored only on the reference's call site above:

  t_emb  = Timesteps(320, flip_sin_to_cos=True, shift 0)(t) -> TimestepEmbedding(320 -> 1280 -> 1280, SiLU)
  sample = conv_in(sample) + controlnet_cond_embedding(cond)
  cond embedding: conv3x3(3->16) SiLU; [conv3x3(c->c) SiLU, conv3x3(c->c', stride 2) SiLU] for (16,32,96,256);
                  conv3x3(256->320) (zero-initialised in a fresh ControlNet)
  encoder = the UNet's down blocks + mid block (text-only cross-attention: plain AttnProcessor on the 77 tokens)
  residual_i = controlnet_down_blocks[i] (1x1 conv) of the i-th of the 12 skip tensors; mid = controlnet_mid_block(mid)
  all residuals * conditioning_scale (guess_mode False)
State-dict key names are diffusers' (`controlnet_cond_embedding.*`, `controlnet_down_blocks.N`,
`controlnet_mid_block`, `down_blocks.*`, `mid_block.*`), so a real checkpoint drops in.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from . import unet as U

COND_CHANNELS = (16, 32, 96, 256)
SKIP_CHANNELS = (320, 320, 320, 320, 640, 640, 640, 1280, 1280, 1280, 1280, 1280)


def resnet_specs():
    """the 10 ResnetBlock2D of the encoder half: [(prefix, cin, cout)]"""
    return [s for s in U.resnet_specs() if not s[0].startswith("up_blocks")]


def transformer_specs():
    """the 7 Transformer2DModel of the encoder half in execution order: [(prefix, C)]"""
    out = [(f"down_blocks.{i}.attentions.{j}", U.BLOCK_OUT[i]) for i in range(3) for j in range(2)]
    return out + [("mid_block.attentions.0", 1280)]


def cond_embedding(sd, cond):
    """ControlNetConditioningEmbedding.forward: [B,3,512,512] in [0,1] -> [B,320,64,64]."""
    p = "controlnet_cond_embedding"
    h = F.silu(U._conv(cond, sd, p + ".conv_in"))
    for i in range(6):
        h = F.silu(U._conv(h, sd, f"{p}.blocks.{i}", stride=2 if i % 2 else 1))
    return U._conv(h, sd, p + ".conv_out")


def controlnet_forward(sd, sample, t, ctx, cond, conditioning_scale=1.0):
    """sample [B,4,64,64], t scalar, ctx [B,77,768] (text only), cond [B,3,512,512]
    -> (12 down residuals [B,C_i,H_i,W_i], mid residual [B,1280,8,8])."""
    dtype = sample.dtype
    b = sample.shape[0]
    tt = torch.as_tensor(t, device=sample.device).reshape(-1).expand(b)
    emb = U.timestep_embedding(tt, 320, dtype)
    emb = U._lin(F.silu(U._lin(emb, sd, "time_embedding.linear_1")), sd, "time_embedding.linear_2")

    def tf(prefix, x):
        return U.transformer_2d(sd, prefix, x, ctx, 0.0, 0, None, None, None)

    x = U._conv(sample, sd, "conv_in") + cond_embedding(sd, cond)
    skips = [x]
    for i in range(4):
        for j in range(U.LAYERS_PER_BLOCK):
            x = U.resnet_block(sd, f"down_blocks.{i}.resnets.{j}", x, emb)
            if i < 3:
                x = tf(f"down_blocks.{i}.attentions.{j}", x)
            skips.append(x)
        if i < 3:
            x = U.downsample(sd, f"down_blocks.{i}.downsamplers.0", x)
            skips.append(x)
    x = U.resnet_block(sd, "mid_block.resnets.0", x, emb)
    x = tf("mid_block.attentions.0", x)
    x = U.resnet_block(sd, "mid_block.resnets.1", x, emb)
    down = [U._conv(s, sd, f"controlnet_down_blocks.{i}", padding=0) * conditioning_scale for i, s in enumerate(skips)]
    mid = U._conv(x, sd, "controlnet_mid_block", padding=0) * conditioning_scale
    return down, mid


def synth_state_dict(seed=4321, dtype=torch.float32, cross_dim=768, zero_gain=0.5):
    """Seeded synthetic weights with the diffusers ControlNetModel key set.  The "zero" convolutions of a trained
    ControlNet are not zero; they get N(0, zero_gain^2 / fan_in) so the residuals exercise the whole path."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def lin(name, out_f, in_f, bias=True, scale=1.0):
        sd[name + ".weight"] = (torch.randn(out_f, in_f, generator=g) * (scale / math.sqrt(in_f))).to(dtype)
        if bias:
            sd[name + ".bias"] = (torch.randn(out_f, generator=g) * 0.02).to(dtype)

    def conv(name, out_c, in_c, k, scale=1.0):
        sd[name + ".weight"] = (torch.randn(out_c, in_c, k, k, generator=g) * (scale / math.sqrt(in_c * k * k))).to(dtype)
        sd[name + ".bias"] = (torch.randn(out_c, generator=g) * 0.02).to(dtype)

    def norm(name, c):
        sd[name + ".weight"] = (1.0 + 0.1 * torch.randn(c, generator=g)).to(dtype)
        sd[name + ".bias"] = (0.05 * torch.randn(c, generator=g)).to(dtype)

    lin("time_embedding.linear_1", 1280, 320)
    lin("time_embedding.linear_2", 1280, 1280)
    conv("conv_in", 320, 4, 3)
    p = "controlnet_cond_embedding"
    conv(p + ".conv_in", COND_CHANNELS[0], 3, 3, scale=2.0)
    for i in range(3):
        conv(f"{p}.blocks.{2 * i}", COND_CHANNELS[i], COND_CHANNELS[i], 3, scale=1.5)
        conv(f"{p}.blocks.{2 * i + 1}", COND_CHANNELS[i + 1], COND_CHANNELS[i], 3, scale=1.5)
    conv(p + ".conv_out", 320, COND_CHANNELS[3], 3, scale=zero_gain)
    for pr, cin, cout in resnet_specs():
        norm(pr + ".norm1", cin)
        conv(pr + ".conv1", cout, cin, 3)
        lin(pr + ".time_emb_proj", cout, 1280)
        norm(pr + ".norm2", cout)
        conv(pr + ".conv2", cout, cout, 3)
        if cin != cout:
            conv(pr + ".conv_shortcut", cout, cin, 1)
    for pr, c in transformer_specs():
        norm(pr + ".norm", c)
        conv(pr + ".proj_in", c, c, 1)
        conv(pr + ".proj_out", c, c, 1)
        tb = pr + ".transformer_blocks.0"
        for n in ("norm1", "norm2", "norm3"):
            norm(f"{tb}.{n}", c)
        for a, kd in (("attn1", c), ("attn2", cross_dim)):
            lin(f"{tb}.{a}.to_q", c, c, bias=False)
            lin(f"{tb}.{a}.to_k", c, kd, bias=False)
            lin(f"{tb}.{a}.to_v", c, kd, bias=False)
            lin(f"{tb}.{a}.to_out.0", c, c, scale=0.5)
        lin(f"{tb}.ff.net.0.proj", 8 * c, c)
        lin(f"{tb}.ff.net.2", c, 4 * c, scale=0.5)
    for i in range(3):
        conv(f"down_blocks.{i}.downsamplers.0.conv", U.BLOCK_OUT[i], U.BLOCK_OUT[i], 3)
    for i, c in enumerate(SKIP_CHANNELS):
        conv(f"controlnet_down_blocks.{i}", c, c, 1, scale=zero_gain)
    conv("controlnet_mid_block", 1280, 1280, 1, scale=zero_gain)
    return sd
