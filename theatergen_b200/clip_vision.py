"""TGCLIPVision -- the CLIP ViT-H/14 image encoder in front of the IP-Adapter projectors, on libtheatergen.so kernels.

The reference loads `CLIPVisionModelWithProjection` (transformers) in `IPAdapter.__init__` (ip_adapter/ip_adapter.py:78-81)
and calls it once per character / turn: `.image_embeds` for `ImageProjModel` (ip_adapter.py:142-153) and
`output_hidden_states=True ... .hidden_states[-2]` for the `Resampler` of IPAdapterPlus (ip_adapter.py:301-317).
SURVEY.md section 8(f) row 3.  Parity is checked against the transformers class itself (it is importable here and on the GPU
box; weights are seeded random -- no checkpoint exists offline).

Structure (transformers CLIPVisionTransformer): patch embedding (conv 14x14 stride 14, no bias) = im2col + GEMM with the
position embeddings added in the GEMM epilogue; class token; pre-LayerNorm; 32 x [LN -> fused QKV GEMM -> flash attention
(16 heads of 80, 257 tokens) -> out-proj GEMM (+residual) -> LN -> fc1 GEMM + GELU(erf) -> fc2 GEMM (+residual)];
post-LayerNorm of the class token; visual projection.  State-dict keys are the transformers ones.
"""
from __future__ import annotations

import torch

from . import ops
from .unet import _getter


class CLIPVisionOutput:
    def __init__(self, image_embeds, last_hidden_state, hidden_states):
        self.image_embeds, self.last_hidden_state, self.hidden_states = image_embeds, last_hidden_state, hidden_states


class TGCLIPVision:
    dtype = torch.float16

    def __init__(self, state_dict, device="cuda", num_heads=16, patch=14, eps=1e-5):
        self.device = torch.device(device)
        g = _getter(state_dict, self.device)
        p = "vision_model."
        w = {}
        pe = g(p + "embeddings.patch_embedding.weight")  # [width, 3, P, P]
        self.width, self.patch, self.heads, self.eps = pe.shape[0], patch, num_heads, eps
        kk = pe.shape[1] * patch * patch
        self.kpad = (kk + 63) // 64 * 64
        wp = torch.zeros((self.width, self.kpad), dtype=torch.float16, device=self.device)
        wp[:, :kk] = pe.reshape(self.width, kk)
        w["patch.w"] = wp
        pos = g(p + "embeddings.position_embedding.weight")  # [1 + n_patches, width]
        w["pos.patches"] = pos[1:].contiguous()
        w["cls"] = (g(p + "embeddings.class_embedding").float() + pos[0].float()).half().contiguous()
        w["pre.g"], w["pre.b"] = g(p + "pre_layrnorm.weight"), g(p + "pre_layrnorm.bias")
        w["post.g"], w["post.b"] = g(p + "post_layernorm.weight"), g(p + "post_layernorm.bias")
        w["proj.w"] = g("visual_projection.weight")
        self.layers = []
        i = 0
        while f"{p}encoder.layers.{i}.layer_norm1.weight" in state_dict:
            q = f"{p}encoder.layers.{i}."
            L = {}
            L["ln1g"], L["ln1b"] = g(q + "layer_norm1.weight"), g(q + "layer_norm1.bias")
            L["ln2g"], L["ln2b"] = g(q + "layer_norm2.weight"), g(q + "layer_norm2.bias")
            L["qkvw"] = torch.cat([g(f"{q}self_attn.{n}_proj.weight") for n in "qkv"], 0).contiguous()
            L["qkvb"] = torch.cat([g(f"{q}self_attn.{n}_proj.bias") for n in "qkv"], 0).contiguous()
            L["ow"], L["ob"] = g(q + "self_attn.out_proj.weight"), g(q + "self_attn.out_proj.bias")
            L["f1w"], L["f1b"] = g(q + "mlp.fc1.weight"), g(q + "mlp.fc1.bias")
            L["f2w"], L["f2b"] = g(q + "mlp.fc2.weight"), g(q + "mlp.fc2.bias")
            self.layers.append(L)
            i += 1
        self.w = w

    @torch.no_grad()
    def __call__(self, pixel_values, output_hidden_states=False, **unused):
        """pixel_values NCHW [B,3,224,224] (CLIPImageProcessor output) -> .image_embeds [B,1024] (+ .hidden_states)."""
        if not pixel_values.is_cuda:
            raise RuntimeError("TGCLIPVision runs on a CUDA device only (no CPU fallback)")
        w, C = self.w, self.width
        x_img = pixel_values.to(torch.float16).contiguous()
        B = x_img.shape[0]
        npatch = (x_img.shape[2] // self.patch) * (x_img.shape[3] // self.patch)
        T = npatch + 1
        cols = ops.im2col_patch(x_img, self.patch, self.kpad)          # [B*npatch, kpad]
        x = torch.empty((B, T, C), dtype=torch.float16, device=self.device)
        for b in range(B):  # patch rows land behind the class token; position embeddings ride in the epilogue
            ops.gemm(cols[b * npatch:(b + 1) * npatch], w["patch.w"], residual=w["pos.patches"], out=x[b, 1:])
            x[b, 0].copy_(w["cls"])
        x = x.view(B * T, C)
        h = ops.layernorm(x, w["pre.g"], w["pre.b"], self.eps)
        hidden = [h] if output_hidden_states else None
        d = C // self.heads
        for L in self.layers:
            ln = ops.layernorm(h, L["ln1g"], L["ln1b"], self.eps)
            qkv = ops.gemm(ln, L["qkvw"], L["qkvb"])
            a = ops.attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], B, self.heads, scale=d ** -0.5)
            h = ops.gemm(a, L["ow"], L["ob"], residual=h)
            ln = ops.layernorm(h, L["ln2g"], L["ln2b"], self.eps)
            f = ops.gemm(ln, L["f1w"], L["f1b"], act=ops.TG_ACT_GELU)
            h = ops.gemm(f, L["f2w"], L["f2b"], residual=h)
            if output_hidden_states:
                hidden.append(h)
        pooled = h.view(B, T, C)[:, 0].contiguous()
        pooled = ops.layernorm(pooled, w["post.g"], w["post.b"], self.eps)
        embeds = ops.gemm(pooled, w["proj.w"])
        hs = tuple(t.view(B, T, C) for t in hidden) if output_hidden_states else None
        return CLIPVisionOutput(embeds, h.view(B, T, C), hs)

    forward = __call__
