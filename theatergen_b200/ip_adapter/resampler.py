"""Perceiver resampler (IPAdapterPlus image-token projector) on libtheatergen.so kernels.

Same constructor / forward contract and state-dict keys as the reference's `Resampler`
(ip_adapter/resampler.py:81-147; PerceiverAttention :34-78; FeedForward :13-20):
    proj_in -> depth x [ latents += to_out(softmax(q k^T / sqrt(d)) v) ; latents += FFN(latents) ] -> proj_out -> LN
Small tensor-core GEMMs (M = 16 / 257 rows, weight-bandwidth bound) + one flash-attention launch per layer
(16 queries over 257+16 keys, 12 heads x 64).  `apply_pos_emb` / `num_latents_mean_pooled` are never enabled by
IPAdapterPlus (ip_adapter/ip_adapter.py:292-302) and are not implemented.
"""
from __future__ import annotations

import torch

from .. import ops


class Resampler:
    def __init__(self, dim=1024, depth=8, dim_head=64, heads=16, num_queries=8, embedding_dim=768, output_dim=1024,
                 ff_mult=4, max_seq_len=257, apply_pos_emb=False, num_latents_mean_pooled=0):
        if apply_pos_emb or num_latents_mean_pooled:
            raise NotImplementedError("pos-emb / mean-pooled latents are unused by IPAdapterPlus")
        self.dim, self.depth, self.dim_head, self.heads = dim, depth, dim_head, heads
        self.num_queries, self.embedding_dim, self.output_dim, self.ff_mult = num_queries, embedding_dim, output_dim, ff_mult
        self.w = None

    def to(self, *a, **k):
        return self

    def load_state_dict(self, sd, device="cuda"):
        self.w = {k: v.detach().to(device=device, dtype=torch.float16).contiguous() for k, v in sd.items()}
        return self

    def __call__(self, x):
        w = self.w
        if w is None:
            raise RuntimeError("Resampler: load_state_dict first")
        if not x.is_cuda:
            raise RuntimeError("Resampler runs on CUDA only (no CPU fallback)")
        x = x.to(torch.float16).contiguous()
        B, n1, E = x.shape
        dim, H, dh, nq = self.dim, self.heads, self.dim_head, self.num_queries
        inner = H * dh
        T = n1 + nq
        xp = ops.gemm(x.view(B * n1, E), w["proj_in.weight"], w["proj_in.bias"])                  # [B*n1, dim]
        lat = w["latents"].reshape(nq, dim).repeat(B, 1).contiguous()                              # [B*nq, dim]
        kv = torch.empty((B * T, 2 * inner), dtype=torch.float16, device=x.device)
        for i in range(self.depth):
            p = f"layers.{i}"
            xn = ops.layernorm(xp, w[f"{p}.0.norm1.weight"], w[f"{p}.0.norm1.bias"])
            ln = ops.layernorm(lat, w[f"{p}.0.norm2.weight"], w[f"{p}.0.norm2.bias"])
            q = ops.gemm(ln, w[f"{p}.0.to_q.weight"])
            for b in range(B):  # kv_input = cat((x, latents), dim=-2) per batch element, written in place
                ops.gemm(xn[b * n1:(b + 1) * n1], w[f"{p}.0.to_kv.weight"], out=kv[b * T:b * T + n1])
                ops.gemm(ln[b * nq:(b + 1) * nq], w[f"{p}.0.to_kv.weight"], out=kv[b * T + n1:(b + 1) * T])
            # (q s)(k s)^T with s = d^-1/4  ==  q k^T d^-1/2 ; softmax in fp32 (resampler.py:71-73)
            a = ops.attention(q, kv[:, :inner], kv[:, inner:], B, H, scale=dh ** -0.5)
            lat = ops.gemm(a, w[f"{p}.0.to_out.weight"], residual=lat)
            h = ops.layernorm(lat, w[f"{p}.1.0.weight"], w[f"{p}.1.0.bias"])
            h = ops.gemm(h, w[f"{p}.1.1.weight"], act=ops.TG_ACT_GELU)
            lat = ops.gemm(h, w[f"{p}.1.3.weight"], residual=lat)
        out = ops.gemm(lat, w["proj_out.weight"], w["proj_out.bias"])
        out = ops.layernorm(out, w["norm_out.weight"], w["norm_out.bias"])
        return out.view(B, nq, self.output_dim)

    forward = __call__
