"""Restatement of the reference attention processors (TEST INFRASTRUCTURE, see oracle/__init__.py).

Follows /root/reference/ip_adapter/attention_processor.py:
  head_to_batch_dim / batch_to_head_dim   :169-185
  get_attention_scores                    :187-219   (baddbmm(alpha=d^-0.5) -> softmax in the input dtype)
  AttnProcessor.__call__                  :294-393
  IPAttnProcessor.__call__                :421-553
Pinned against the reference classes themselves by tests/golden/make_golden.py.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

HEADS = 8


def head_to_batch(x, heads=HEADS):
    b, n, c = x.shape
    return x.reshape(b, n, heads, c // heads).permute(0, 2, 1, 3).reshape(b * heads, n, c // heads)


def batch_to_head(x, heads=HEADS):
    bh, n, d = x.shape
    return x.reshape(bh // heads, heads, n, d).permute(0, 2, 1, 3).reshape(bh // heads, n, d * heads)


def attention_probs(q, k, scale, mask=None):
    """attention_processor.py:187-219; the softmax runs in q.dtype (upcast_softmax=False for SD-1.5)."""
    if mask is None:
        empty = torch.empty(q.shape[0], q.shape[1], k.shape[1], dtype=q.dtype, device=q.device)
        scores = torch.baddbmm(empty, q, k.transpose(-1, -2), beta=0, alpha=scale)
    else:
        scores = torch.baddbmm(mask, q, k.transpose(-1, -2), beta=1, alpha=scale)
    return scores.softmax(dim=-1).to(q.dtype)


def _proj(x, sd, name):
    return F.linear(x, sd[name + ".weight"], sd.get(name + ".bias"))


def attn_self(sd, prefix, x, heads=HEADS):
    """AttnProcessor with encoder_hidden_states=None (attention_processor.py:333-361)."""
    d = sd[prefix + ".to_q.weight"].shape[0] // heads
    q = head_to_batch(_proj(x, sd, prefix + ".to_q"), heads)
    k = head_to_batch(_proj(x, sd, prefix + ".to_k"), heads)
    v = head_to_batch(_proj(x, sd, prefix + ".to_v"), heads)
    p = attention_probs(q, k, d ** -0.5)
    h = batch_to_head(torch.bmm(p, v), heads)
    return _proj(h, sd, prefix + ".to_out.0")


def attn_cross_ip(sd, prefix, x, ctx, ip_scale, num_ip_tokens=4, heads=HEADS, attention_mask=None,
                  return_probs=False):
    """IPAttnProcessor (attention_processor.py:461-530).  `ip_scale` is the processor's `.scale` (a float in
    the reference; a [B] tensor here also allows co-batched characters with different scales).
    num_ip_tokens == 0 degenerates to the plain AttnProcessor cross-attention (:333-361)."""
    b = x.shape[0]
    d = sd[prefix + ".to_q.weight"].shape[0] // heads
    scale = d ** -0.5
    q = head_to_batch(_proj(x, sd, prefix + ".to_q"), heads)
    end = ctx.shape[1] - num_ip_tokens
    txt, ip = ctx[:, :end], ctx[:, end:]
    k = head_to_batch(_proj(txt, sd, prefix + ".to_k"), heads)
    v = head_to_batch(_proj(txt, sd, prefix + ".to_v"), heads)
    mask = None
    if attention_mask is not None:  # prepare_attention_mask (:221-260): [B,1,T] additive -> repeat over heads
        mask = attention_mask.repeat_interleave(heads, dim=0)
    p = attention_probs(q, k, scale, mask)
    h = batch_to_head(torch.bmm(p, v), heads)
    if num_ip_tokens > 0:
        k_ip = head_to_batch(F.linear(ip, sd[prefix + ".processor.to_k_ip.weight"]), heads)
        v_ip = head_to_batch(F.linear(ip, sd[prefix + ".processor.to_v_ip.weight"]), heads)
        p_ip = attention_probs(q, k_ip, scale)
        h_ip = batch_to_head(torch.bmm(p_ip, v_ip), heads)
        if torch.is_tensor(ip_scale):
            ip_scale = ip_scale.to(h.dtype).reshape(b, 1, 1)
        h = h + ip_scale * h_ip
    out = _proj(h, sd, prefix + ".to_out.0")
    if return_probs:
        return out, p.unflatten(0, (b, heads))  # [B, heads, N, T_text]  (:534)
    return out
