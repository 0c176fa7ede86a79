"""TGVAEDecoder / TGVAEEncoder -- the SD-1.5 VAE (latents <-> 512x512 RGB) on libtheatergen.so kernels.

The reference holds a stock diffusers `AutoencoderKL` (`stabilityai/sd-vae-ft-mse`, generate.py:59,77) and decodes
K + 1 latents per turn: `adapter.pipe.vae.decode(latents / scaling_factor, return_dict=False)[0]` after stage 1
(models/pipelines.py:468; the image becomes the character's reference PNG that later turns wait for) and
`decode(vae, latents)` after stage 2 (:162-173, :849-854).  SURVEY.md section 8(f) row 2.

Everything reuses the hot-path kernels: conv_in = im2col + GEMM, ResnetBlock2D = GroupNorm+SiLU -> implicit-GEMM conv3x3
(residual in the epilogue; no time embedding here), nearest-x2 + conv3x3, conv_out (Cout = 3).  The mid block's attention
is ONE head of 512 channels over 4096 tokens -- outside the flash kernels' head dims -- so it runs as
GEMM (Q K^T) -> fp32 row softmax -> GEMM (P V): 33 MB of probabilities per image, once per decode.
Activations are channels-last fp16 like everywhere else; the module takes / returns NCHW like diffusers.

The encoder (one call per turn: the pasted collage of stage 2, `vae.encode(image).latent_dist.sample(generator)`,
models/pipelines.py:624-626) is the mirror image: conv_in (3 -> 128) as im2col + GEMM, two ResnetBlock2D per level,
stride-2 convs with the VAE's one-sided padding as im2col + GEMM, the same mid block, conv_out (512 -> 8, padded to 64
GEMM columns), quant_conv and the posterior sample in two tiny kernels.

`TGAutoencoderKL` is the duck-typed stand-in for `adapter.pipe.vae`: `.config.scaling_factor`, `.decode()`,
`.encode(image).latent_dist.sample(generator)`.
"""
from __future__ import annotations

import torch

from . import ops
from .unet import _getter, downsample_conv, upsample_conv
from .weights import pack_conv3x3, pack_upsample_conv3x3

UP_CHANNELS = (512, 512, 256, 128)


def _resnet_specs():
    out = [("decoder.mid_block.resnets.0", 512, 512), ("decoder.mid_block.resnets.1", 512, 512)]
    ch = 512
    for i, co in enumerate(UP_CHANNELS):
        for j in range(3):
            out.append((f"decoder.up_blocks.{i}.resnets.{j}", ch, co))
            ch = co
    return out


class DecoderOutput:
    def __init__(self, sample):
        self.sample = sample

    def __getitem__(self, i):
        return (self.sample,)[i]


def _pack_resnet(g, sd, p, cin, co):
    r = dict(cin=cin, cout=co)
    r["n1g"], r["n1b"] = g(p + ".norm1.weight"), g(p + ".norm1.bias")
    r["c1w"], r["c1b"] = pack_conv3x3(g(p + ".conv1.weight")), g(p + ".conv1.bias")
    r["n2g"], r["n2b"] = g(p + ".norm2.weight"), g(p + ".norm2.bias")
    r["c2w"], r["c2b"] = pack_conv3x3(g(p + ".conv2.weight")), g(p + ".conv2.bias")
    if (p + ".conv_shortcut.weight") in sd:
        r["scw"] = g(p + ".conv_shortcut.weight").reshape(co, cin).contiguous()
        r["scb"] = g(p + ".conv_shortcut.bias")
    return r


def _pack_attention(g, w, a):
    w["a.ng"], w["a.nb"] = g(a + ".group_norm.weight"), g(a + ".group_norm.bias")
    w["a.qkvw"] = torch.cat([g(f"{a}.to_{n}.weight") for n in "qkv"], 0).contiguous()  # one GEMM for q, k, v
    w["a.qkvb"] = torch.cat([g(f"{a}.to_{n}.bias") for n in "qkv"], 0).contiguous()
    w["a.ow"], w["a.ob"] = g(a + ".to_out.0.weight"), g(a + ".to_out.0.bias")


class _VAEBlocks:
    """ResnetBlock2D (no time embedding, eps 1e-6) and the single-head mid-block attention shared by both halves."""
    dtype = torch.float16

    # diffusers ResnetBlock2D(temb_channels=None, groups=32, eps=1e-6)
    def _resnet(self, p, x, B, H, W):
        r = self.resnets[p]
        cin, co, HW = r["cin"], r["cout"], H * W
        n1 = ops.groupnorm(x.view(B, HW, cin), r["n1g"], r["n1b"], 1e-6, silu=True)
        h = ops.conv3x3(n1.view(B, H, W, cin), r["c1w"], r["c1b"])
        n2 = ops.groupnorm(h.view(B, HW, co), r["n2g"], r["n2b"], 1e-6, silu=True)
        sc = x
        if "scw" in r:
            sc = ops.gemm(x.view(B * HW, cin), r["scw"], r["scb"]).view(B, H, W, co)
        return ops.conv3x3(n2.view(B, H, W, co), r["c2w"], r["c2b"], residual=sc)

    # diffusers Attention(heads=1, dim_head=512, norm_num_groups=32, residual_connection=True, upcast_softmax=True)
    def _attention(self, x, B, H, W):
        w, C, N = self.w, 512, H * W
        n = ops.groupnorm(x.view(B, N, C), w["a.ng"], w["a.nb"], 1e-6, silu=False)
        qkv = ops.gemm(n.view(B * N, C), w["a.qkvw"], w["a.qkvb"])  # [B*N, 3C]
        o = torch.empty((B * N, C), dtype=torch.float16, device=x.device)
        for b in range(B):
            r = slice(b * N, (b + 1) * N)
            s = ops.gemm(qkv[r, :C], qkv[r, C:2 * C].contiguous())         # Q K^T  [N, N] (the B operand must be dense)
            p = ops.softmax_rows(s, C ** -0.5, out=s)                      # in place, fp32 math
            vt = ops.nchw_to_nhwc(qkv[r, 2 * C:].contiguous().view(1, N, C, 1)).view(C, N)  # V^T [C, N]
            ops.gemm(p, vt, out=o[r])                                      # P V    [N, C]
        return ops.gemm(o, w["a.ow"], w["a.ob"], residual=x.view(B * N, C)).view(B, H, W, C)


class TGVAEDecoder(_VAEBlocks):
    def __init__(self, state_dict, device="cuda"):
        self.device = torch.device(device)
        g = _getter(state_dict, self.device)
        w = {}
        w["pq.w"] = g("post_quant_conv.weight").reshape(4, 4).contiguous()
        w["pq.b"] = g("post_quant_conv.bias")
        w["in.w"], w["in.b"] = ops.pack_conv_in(g("decoder.conv_in.weight")), g("decoder.conv_in.bias")
        self.resnets = {p: _pack_resnet(g, state_dict, p, cin, co) for p, cin, co in _resnet_specs()}
        _pack_attention(g, w, "decoder.mid_block.attentions.0")
        for i in range(3):
            w[f"up{i}.w"] = pack_conv3x3(g(f"decoder.up_blocks.{i}.upsamplers.0.conv.weight"))
            w[f"up{i}.w4"] = pack_upsample_conv3x3(g(f"decoder.up_blocks.{i}.upsamplers.0.conv.weight"))
            w[f"up{i}.b"] = g(f"decoder.up_blocks.{i}.upsamplers.0.conv.bias")
        w["no.g"], w["no.b"] = g("decoder.conv_norm_out.weight"), g("decoder.conv_norm_out.bias")
        w["out.w"], w["out.b"] = pack_conv3x3(g("decoder.conv_out.weight")), g("decoder.conv_out.bias")
        self.w = w

    @torch.no_grad()
    def decode_raw(self, z):
        """z NCHW fp16 [B,4,h,w] (already divided by the scaling factor) -> NCHW fp16 [B,3,8h,8w]."""
        w = self.w
        B, _, H, W = z.shape
        x = ops.conv_in_gemm(ops.pointwise_nchw(z, w["pq.w"], w["pq.b"]), w["in.w"], w["in.b"])
        x = self._resnet("decoder.mid_block.resnets.0", x, B, H, W)
        x = self._attention(x, B, H, W)
        x = self._resnet("decoder.mid_block.resnets.1", x, B, H, W)
        for i in range(4):
            for j in range(3):
                x = self._resnet(f"decoder.up_blocks.{i}.resnets.{j}", x, B, H, W)
            if i < 3:
                x = upsample_conv(x, w, f"up{i}")
                H, W = 2 * H, 2 * W
        n = ops.groupnorm(x.view(B, H * W, 128), w["no.g"], w["no.b"], 1e-6, silu=True)
        return ops.conv_out(n.view(B, H, W, 128), w["out.w"], w["out.b"])


ENC_CHANNELS = (128, 256, 512, 512)


def _enc_resnet_specs():
    out, ch = [], 128
    for i, co in enumerate(ENC_CHANNELS):
        for j in range(2):
            out.append((f"encoder.down_blocks.{i}.resnets.{j}", ch, co))
            ch = co
    return out + [("encoder.mid_block.resnets.0", 512, 512), ("encoder.mid_block.resnets.1", 512, 512)]


class TGVAEEncoder(_VAEBlocks):
    def __init__(self, state_dict, device="cuda"):
        self.device = torch.device(device)
        g = _getter(state_dict, self.device)
        w = {}
        w["in.w"], w["in.b"] = ops.pack_conv_in(g("encoder.conv_in.weight")), g("encoder.conv_in.bias")
        self.resnets = {p: _pack_resnet(g, state_dict, p, cin, co) for p, cin, co in _enc_resnet_specs()}
        for i in range(3):
            c = ENC_CHANNELS[i]
            w[f"down{i}.w"] = pack_conv3x3(g(f"encoder.down_blocks.{i}.downsamplers.0.conv.weight")).reshape(c, 9 * c)
            w[f"down{i}.b"] = g(f"encoder.down_blocks.{i}.downsamplers.0.conv.bias")
        _pack_attention(g, w, "encoder.mid_block.attentions.0")
        w["no.g"], w["no.b"] = g("encoder.conv_norm_out.weight"), g("encoder.conv_norm_out.bias")
        co = g("encoder.conv_out.weight")  # [8,512,3,3] -> 64 GEMM columns (zero rows beyond 8)
        wp = torch.zeros((64, 512, 3, 3), dtype=torch.float16, device=self.device)
        wp[:8] = co
        bp = torch.zeros((64,), dtype=torch.float16, device=self.device)
        bp[:8] = g("encoder.conv_out.bias")
        w["out.w"], w["out.b"] = pack_conv3x3(wp), bp
        w["q.w"] = g("quant_conv.weight").reshape(8, 8).contiguous()
        w["q.b"] = g("quant_conv.bias")
        self.w = w

    @torch.no_grad()
    def moments(self, x):
        """x NCHW fp16 [B,3,H,W] in [-1,1] -> moments NCHW fp16 [B,8,H/8,W/8] (mean | logvar)."""
        w = self.w
        B, _, H, W = x.shape
        h = ops.conv_in_gemm(x, w["in.w"], w["in.b"])
        for i in range(4):
            for j in range(2):
                h = self._resnet(f"encoder.down_blocks.{i}.resnets.{j}", h, B, H, W)
            if i < 3:  # F.pad(h, (0,1,0,1)) + conv3x3 stride 2 pad 0
                c = ENC_CHANNELS[i]
                h = downsample_conv(h, w[f"down{i}.w"], w[f"down{i}.b"], c, pad_lo=0)
                H, W = H // 2, W // 2
        h = self._resnet("encoder.mid_block.resnets.0", h, B, H, W)
        h = self._attention(h, B, H, W)
        h = self._resnet("encoder.mid_block.resnets.1", h, B, H, W)
        n = ops.groupnorm(h.view(B, H * W, 512), w["no.g"], w["no.b"], 1e-6, silu=True)
        m = ops.conv3x3(n.view(B, H, W, 512), w["out.w"], w["out.b"])                 # [B,H,W,64], 8 live channels
        m = ops.nchw_to_nhwc(m.view(B, H * W, 64, 1)).view(B, 64, H, W)[:, :8].contiguous()  # -> NCHW [B,8,H,W]
        return ops.pointwise_nchw(m, w["q.w"], w["q.b"])


class LatentDist:
    """diffusers DiagonalGaussianDistribution surface: `.sample(generator)`, `.mode()`, `.mean`."""

    def __init__(self, moments):
        self.moments = moments
        self.mean = moments[:, :moments.shape[1] // 2]

    def sample(self, generator=None, noise=None):
        B, L2, H, W = self.moments.shape
        if noise is None:  # the reference draws it with a CUDA generator (models/pipelines.py:594, 624-626)
            noise = torch.randn((B, L2 // 2, H, W), generator=generator, device=self.moments.device, dtype=torch.float16)
        return ops.gaussian_sample(self.moments, noise.to(torch.float16).contiguous())

    def mode(self):
        return ops.gaussian_sample(self.moments, None)


class EncoderOutput:
    def __init__(self, latent_dist):
        self.latent_dist = latent_dist

    def __getitem__(self, i):
        return (self.latent_dist,)[i]


class _Config(dict):
    __getattr__ = dict.__getitem__


class TGAutoencoderKL:
    """Stand-in for `adapter.pipe.vae` (diffusers AutoencoderKL) with both halves on our kernels.  `state_dict` uses the
    diffusers key names; the encoder half is optional (a decoder-only dict makes `.encode` raise)."""
    dtype = torch.float16

    def __init__(self, state_dict, device="cuda", scaling_factor=0.18215):
        self.decoder = TGVAEDecoder(state_dict, device)
        self.encoder = TGVAEEncoder(state_dict, device) if "encoder.conv_in.weight" in state_dict else None
        self.config = _Config(scaling_factor=scaling_factor)
        self.device = self.decoder.device

    def to(self, *a, **k):
        return self

    def decode(self, z, return_dict=True, **unused):
        if not z.is_cuda:
            raise RuntimeError("TGAutoencoderKL.decode runs on a CUDA device only (no CPU fallback)")
        img = self.decoder.decode_raw(z.to(torch.float16).contiguous())
        return DecoderOutput(img) if return_dict else (img,)

    def encode(self, image, return_dict=True, **unused):
        if self.encoder is None:
            raise RuntimeError("this TGAutoencoderKL was built from a decoder-only state dict")
        if not image.is_cuda:
            raise RuntimeError("TGAutoencoderKL.encode runs on a CUDA device only (no CPU fallback)")
        dist = LatentDist(self.encoder.moments(image.to(torch.float16).contiguous()))
        return EncoderOutput(dist) if return_dict else (dist,)
