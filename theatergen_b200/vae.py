"""TGVAEDecoder -- the SD-1.5 VAE decode (latents -> 512x512 RGB) on libtheatergen.so kernels.

The reference holds a stock diffusers `AutoencoderKL` (`stabilityai/sd-vae-ft-mse`, generate.py:59,77) and decodes
K + 1 latents per turn: `adapter.pipe.vae.decode(latents / scaling_factor, return_dict=False)[0]` after stage 1
(models/pipelines.py:468; the image becomes the character's reference PNG that later turns wait for) and
`decode(vae, latents)` after stage 2 (:162-173, :849-854).  SURVEY.md section 8(f) row 2.

Everything reuses the hot-path kernels: conv_in = im2col + GEMM, ResnetBlock2D = GroupNorm+SiLU -> implicit-GEMM conv3x3
(residual in the epilogue; no time embedding here), nearest-x2 + conv3x3, conv_out (Cout = 3).  The mid block's attention
is ONE head of 512 channels over 4096 tokens -- outside the flash kernels' head dims -- so it runs as
GEMM (Q K^T) -> fp32 row softmax -> GEMM (P V): 33 MB of probabilities per image, once per decode.
Activations are channels-last fp16 like everywhere else; the module takes / returns NCHW like diffusers.

`TGAutoencoderKL` is the duck-typed stand-in for `adapter.pipe.vae`: `.config.scaling_factor`, `.decode()`;
`.encode()` (one call per turn, the collage of stage 2) is not built yet and forwards to a caller-supplied encoder.
"""
from __future__ import annotations

import torch

from . import ops
from .unet import _getter
from .weights import pack_conv3x3

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


class TGVAEDecoder:
    dtype = torch.float16

    def __init__(self, state_dict, device="cuda"):
        self.device = torch.device(device)
        g = _getter(state_dict, self.device)
        w = {}
        w["pq.w"] = g("post_quant_conv.weight").reshape(4, 4).contiguous()
        w["pq.b"] = g("post_quant_conv.bias")
        w["in.w"], w["in.b"] = ops.pack_conv_in(g("decoder.conv_in.weight")), g("decoder.conv_in.bias")
        self.resnets = {}
        for p, cin, co in _resnet_specs():
            r = dict(cin=cin, cout=co)
            r["n1g"], r["n1b"] = g(p + ".norm1.weight"), g(p + ".norm1.bias")
            r["c1w"], r["c1b"] = pack_conv3x3(g(p + ".conv1.weight")), g(p + ".conv1.bias")
            r["n2g"], r["n2b"] = g(p + ".norm2.weight"), g(p + ".norm2.bias")
            r["c2w"], r["c2b"] = pack_conv3x3(g(p + ".conv2.weight")), g(p + ".conv2.bias")
            if (p + ".conv_shortcut.weight") in state_dict:
                r["scw"] = g(p + ".conv_shortcut.weight").reshape(co, cin).contiguous()
                r["scb"] = g(p + ".conv_shortcut.bias")
            self.resnets[p] = r
        a = "decoder.mid_block.attentions.0"
        w["a.ng"], w["a.nb"] = g(a + ".group_norm.weight"), g(a + ".group_norm.bias")
        w["a.qkvw"] = torch.cat([g(f"{a}.to_{n}.weight") for n in "qkv"], 0).contiguous()  # one GEMM for q, k, v
        w["a.qkvb"] = torch.cat([g(f"{a}.to_{n}.bias") for n in "qkv"], 0).contiguous()
        w["a.ow"], w["a.ob"] = g(a + ".to_out.0.weight"), g(a + ".to_out.0.bias")
        for i in range(3):
            w[f"up{i}.w"] = pack_conv3x3(g(f"decoder.up_blocks.{i}.upsamplers.0.conv.weight"))
            w[f"up{i}.b"] = g(f"decoder.up_blocks.{i}.upsamplers.0.conv.bias")
        w["no.g"], w["no.b"] = g("decoder.conv_norm_out.weight"), g("decoder.conv_norm_out.bias")
        w["out.w"], w["out.b"] = pack_conv3x3(g("decoder.conv_out.weight")), g("decoder.conv_out.bias")
        self.w = w

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
                x = ops.conv3x3(ops.upsample2x(x), w[f"up{i}.w"], w[f"up{i}.b"])
                H, W = 2 * H, 2 * W
        n = ops.groupnorm(x.view(B, H * W, 128), w["no.g"], w["no.b"], 1e-6, silu=True)
        return ops.conv_out(n.view(B, H, W, 128), w["out.w"], w["out.b"])


class _Config(dict):
    __getattr__ = dict.__getitem__


class TGAutoencoderKL:
    """Stand-in for `adapter.pipe.vae` (diffusers AutoencoderKL): decode on our kernels, encode through `encoder`
    (any object with the diffusers `.encode(image).latent_dist.sample(generator)` surface) until it is built."""
    dtype = torch.float16

    def __init__(self, state_dict, device="cuda", scaling_factor=0.18215, encoder=None):
        self.decoder = TGVAEDecoder(state_dict, device)
        self.config = _Config(scaling_factor=scaling_factor)
        self.device = self.decoder.device
        self._encoder = encoder

    def to(self, *a, **k):
        return self

    def decode(self, z, return_dict=True, **unused):
        if not z.is_cuda:
            raise RuntimeError("TGAutoencoderKL.decode runs on a CUDA device only (no CPU fallback)")
        img = self.decoder.decode_raw(z.to(torch.float16).contiguous())
        return DecoderOutput(img) if return_dict else (img,)

    def encode(self, image, **kw):
        if self._encoder is None:
            raise NotImplementedError("VAE encode is not on the kernel path yet: pass encoder=<diffusers AutoencoderKL>")
        return self._encoder.encode(image, **kw)
