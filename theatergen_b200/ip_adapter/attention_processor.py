"""Processor-level drop-in: `AttnProcessor` / `IPAttnProcessor` with the diffusers attention-processor protocol of
the reference (ip_adapter/attention_processor.py:282-393, 396-553), executing on libtheatergen.so kernels.

`processor(attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None, ...)`: `attn` is any
module exposing to_q / to_k / to_v / to_out[0] (nn.Linear, fp16, CUDA) and `heads` -- a diffusers `Attention`, or the
minimal `Attention` below.  Install with `unet.set_attn_processor({...})` exactly like
IPAdapter.set_ip_adapter (ip_adapter/ip_adapter.py:95-119).  The whole-UNet fast path is theatergen_b200.unet.TGUNet;
this file is the per-layer boundary, and what the golden-vector parity tests call.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops


def _cuda16(t, what):
    if not t.is_cuda:
        raise RuntimeError(f"theatergen_b200 attention processors run on CUDA only ({what} is on {t.device}); "
                           "there is no CPU fallback")
    return t.to(torch.float16).contiguous()


def _w(linear):
    return linear.weight.detach().to(torch.float16).contiguous()


def _b(linear):
    return None if linear.bias is None else linear.bias.detach().to(torch.float16).contiguous()


class Attention(nn.Module):
    """Minimal stand-in for the reference's `Attention` module (attention_processor.py:12-167): just the
    projections the SD-1.5 layers use (no bias on q/k/v, bias on to_out[0], dropout 0)."""

    def __init__(self, query_dim, cross_attention_dim=None, heads=8, dim_head=64, processor=None):
        super().__init__()
        inner = dim_head * heads
        kd = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.heads, self.scale = heads, dim_head ** -0.5
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(kd, inner, bias=False)
        self.to_v = nn.Linear(kd, inner, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim), nn.Dropout(0.0)])
        self.processor = processor if processor is not None else AttnProcessor()

    def set_processor(self, processor):
        self.processor = processor

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **cross_attention_kwargs):
        return self.processor(self, hidden_states, encoder_hidden_states=encoder_hidden_states,
                              attention_mask=attention_mask, **cross_attention_kwargs)


def _save_probs(probs, batch_size, attn_key, return_token_ca_only, return_cond_ca_only, offload, save_attn_to_dict,
                save_keys):
    """attention_processor.py:532-549 (probs arrives already restricted to the cond half when requested)."""
    if return_token_ca_only is not None:
        if isinstance(return_token_ca_only, int):
            probs = probs[:, :, :, return_token_ca_only:return_token_ca_only + 1]
        else:
            probs = probs[:, :, :, return_token_ca_only]
    if offload:
        probs = probs.cpu()
    if save_attn_to_dict is not None and (save_keys is None or (tuple(attn_key) in [tuple(k) for k in save_keys])):
        save_attn_to_dict[tuple(attn_key)] = probs
    return probs


class AttnProcessor(nn.Module):
    """Self- (or plain cross-) attention: q/k/v GEMMs -> flash attention -> out GEMM."""

    def __init__(self, hidden_size=None, cross_attention_dim=None):
        super().__init__()

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None,
                 return_attntion_probs=False, attn_key=None, attn_process_fn=None, return_cond_ca_only=False,
                 return_token_ca_only=None, offload_cross_attn_to_cpu=False, save_attn_to_dict=None, save_keys=None,
                 enable_flash_attn=True):
        if attn_process_fn is not None:
            raise NotImplementedError("attn_process_fn needs materialised probabilities; generate.py never sets it")
        x = _cuda16(hidden_states, "hidden_states")
        B, N, Cc = x.shape
        x2 = x.view(B * N, Cc)
        q = ops.gemm(x2, _w(attn.to_q))
        if encoder_hidden_states is None:
            k, v = ops.gemm(x2, _w(attn.to_k)), ops.gemm(x2, _w(attn.to_v))
            a = ops.attention(q, k, v, B, attn.heads)
        else:
            ctx = _cuda16(encoder_hidden_states, "encoder_hidden_states")
            T = ctx.shape[1]
            c2 = ctx.view(B * T, -1)
            kv = torch.empty((B * T, 2 * Cc), dtype=torch.float16, device=x.device)
            ops.gemm(c2, _w(attn.to_k), out=kv[:, :Cc])
            ops.gemm(c2, _w(attn.to_v), out=kv[:, Cc:])
            mask = None if attention_mask is None else _cuda16(attention_mask, "attention_mask").reshape(B, T)
            probs, b0 = None, 0
            if save_attn_to_dict is not None:
                b0 = B // 2 if return_cond_ca_only else 0
                probs = torch.empty((B - b0, attn.heads, N, T), dtype=torch.float16, device=x.device)
            a = ops.attention_cross_ip(q, kv[:, :Cc], kv[:, Cc:], None, None, B, attn.heads, T, 0, add_mask=mask,
                                       probs=probs, probs_b0=b0)
            if probs is not None:
                # the reference offloads to the CPU whenever the dict is non-empty (:386-387)
                _save_probs(probs, B, attn_key, return_token_ca_only, return_cond_ca_only, bool(save_attn_to_dict),
                            save_attn_to_dict, save_keys)
        out = ops.gemm(a, _w(attn.to_out[0]), _b(attn.to_out[0]))
        return out.view(B, N, Cc)


class IPAttnProcessor(nn.Module):
    """Text cross-attention + decoupled image-token cross-attention (one fused kernel for both softmaxes)."""

    def __init__(self, hidden_size, cross_attention_dim=None, scale=1.0, num_tokens=4):
        super().__init__()
        self.hidden_size, self.cross_attention_dim = hidden_size, cross_attention_dim
        self.scale, self.num_tokens = scale, num_tokens
        self.to_k_ip = nn.Linear(cross_attention_dim or hidden_size, hidden_size, bias=False)
        self.to_v_ip = nn.Linear(cross_attention_dim or hidden_size, hidden_size, bias=False)

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None,
                 return_attntion_probs=False, attn_key=None, attn_process_fn=None, return_cond_ca_only=False,
                 return_token_ca_only=None, offload_cross_attn_to_cpu=False, save_attn_to_dict=None, save_keys=None,
                 enable_flash_attn=True):
        if attn_process_fn is not None:
            raise NotImplementedError("attn_process_fn needs materialised probabilities; generate.py never sets it")
        x = _cuda16(hidden_states, "hidden_states")
        B, N, Cc = x.shape
        ctx = _cuda16(encoder_hidden_states, "encoder_hidden_states")
        T = ctx.shape[1]
        n_ip, n_text = self.num_tokens, T - self.num_tokens
        q = ops.gemm(x.view(B * N, Cc), _w(attn.to_q))
        c2 = ctx.view(B * T, -1)
        kv = torch.empty((B * T, 4 * Cc), dtype=torch.float16, device=x.device)
        for i, lin in enumerate((attn.to_k, attn.to_v, self.to_k_ip, self.to_v_ip)):
            ops.gemm(c2, _w(lin), out=kv[:, i * Cc:(i + 1) * Cc])
        if torch.is_tensor(self.scale):
            scale = self.scale.to(device=x.device, dtype=torch.float32).reshape(B)
        else:
            scale = torch.full((B,), float(self.scale), dtype=torch.float32, device=x.device)
        mask = None if attention_mask is None else _cuda16(attention_mask, "attention_mask").reshape(B, n_text)
        want = return_attntion_probs or save_attn_to_dict is not None
        probs, b0 = None, 0
        if want:
            b0 = B // 2 if return_cond_ca_only else 0
            probs = torch.empty((B - b0, attn.heads, N, n_text), dtype=torch.float16, device=x.device)
        a = ops.attention_cross_ip(q, kv[:, :Cc], kv[:, Cc:2 * Cc], kv[:, 2 * Cc:3 * Cc], kv[:, 3 * Cc:], B, attn.heads,
                                   n_text, n_ip, ip_scale=scale, add_mask=mask, probs=probs, probs_b0=b0)
        out = ops.gemm(a, _w(attn.to_out[0]), _b(attn.to_out[0])).view(B, N, Cc)
        if want:
            probs = _save_probs(probs, B, attn_key, return_token_ca_only, return_cond_ca_only,
                                offload_cross_attn_to_cpu, save_attn_to_dict, save_keys)
            if return_attntion_probs:
                return out, probs
        return out
