"""TGCLIPText -- the SD-1.5 text encoder (CLIP ViT-L/14 text tower) on libtheatergen.so kernels.

The reference loads transformers' `CLIPTextModel` (models/models.py:31, generate.py:60 through the diffusers pipeline) and
calls it as `text_encoder(input_ids)[0]` for every prompt of a turn: K per-character prompts + one shared negative
(theatergen.py:378 -> models/models.py:53-70), the overall and background prompts (theatergen.py:434-435) and the IP-Adapter
pipeline's `_encode_prompt` (models/pipelines.py:283-331).  SURVEY.md section 8(f) row 3.  Parity is checked against the
transformers class itself with seeded random weights (no checkpoint exists offline).

Structure (transformers CLIPTextTransformer): token + position embedding (one gather kernel); 12 x [LN -> fused QKV GEMM
-> causal flash attention (12 heads of 64, 77 tokens) -> out-proj GEMM (+residual) -> LN -> fc1 GEMM + quick_gelu ->
fc2 GEMM (+residual)]; final LayerNorm.  `[0]` is the last hidden state [B,77,768]; `[1]` the hidden state at each
sequence's EOS position (argmax of the ids: the SD-1.5 config keeps transformers' legacy eos_token_id = 2 path).
An `attention_mask` argument is accepted and must be None or all ones: SD-1.5's text-encoder config has no
`use_attention_mask`, so the reference passes None (models/pipelines.py:296-299).
"""
from __future__ import annotations

import torch

from . import ops
from .unet import _getter


class CLIPTextOutput(tuple):
    """(last_hidden_state, pooler_output[, hidden_states]) with attribute access, like transformers' ModelOutput."""

    def __new__(cls, last_hidden_state, pooler_output, hidden_states=None):
        items = (last_hidden_state, pooler_output) + ((hidden_states,) if hidden_states is not None else ())
        self = super().__new__(cls, items)
        self.last_hidden_state, self.pooler_output, self.hidden_states = last_hidden_state, pooler_output, hidden_states
        return self


class _Config:
    """What diffusers' `encode_prompt` probes on `text_encoder.config` (no `use_attention_mask`: SD-1.5 has none)."""

    def __init__(self, **kw):
        self.__dict__.update(kw)


class TGCLIPText:
    dtype = torch.float16

    def __init__(self, state_dict, device="cuda", num_heads=12, eps=1e-5, act="quick_gelu"):
        self.device = torch.device(device)
        g = _getter(state_dict, self.device)
        p = "text_model."
        self.w = {"tok": g(p + "embeddings.token_embedding.weight"), "pos": g(p + "embeddings.position_embedding.weight"),
                  "fin.g": g(p + "final_layer_norm.weight"), "fin.b": g(p + "final_layer_norm.bias")}
        self.width, self.heads, self.eps = self.w["tok"].shape[1], num_heads, eps
        self.act = {"quick_gelu": ops.TG_ACT_QUICK_GELU, "gelu": ops.TG_ACT_GELU}[act]
        self.config = _Config(hidden_size=self.width, num_attention_heads=num_heads, hidden_act=act,
                              max_position_embeddings=self.w["pos"].shape[0], vocab_size=self.w["tok"].shape[0])
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

    def weight_bytes(self):
        return 2 * (sum(v.numel() for v in self.w.values()) + sum(v.numel() for L in self.layers for v in L.values()))

    @torch.no_grad()
    def __call__(self, input_ids, attention_mask=None, output_hidden_states=False, **unused):
        """input_ids int64 [B,T] (T <= 77, CLIPTokenizer output padded to max_length) -> (last_hidden_state [B,T,768],
        pooled [B,768])."""
        if not input_ids.is_cuda:
            raise RuntimeError("TGCLIPText runs on a CUDA device only (no CPU fallback)")
        if attention_mask is not None and not bool(attention_mask.bool().all()):
            raise NotImplementedError("padding masks are not used by SD-1.5's text encoder (models/pipelines.py:296-299)")
        w, C = self.w, self.width
        ids = input_ids.to(torch.int64).contiguous()
        B, T = ids.shape
        h = ops.embed_tokens(ids, w["tok"], w["pos"])
        hidden = [h] if output_hidden_states else None
        d = C // self.heads
        for L in self.layers:
            ln = ops.layernorm(h, L["ln1g"], L["ln1b"], self.eps)
            qkv = ops.gemm(ln, L["qkvw"], L["qkvb"])
            a = ops.attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], B, self.heads, scale=d ** -0.5, causal=True)
            h = ops.gemm(a, L["ow"], L["ob"], residual=h)
            ln = ops.layernorm(h, L["ln2g"], L["ln2b"], self.eps)
            f = ops.gemm(ln, L["f1w"], L["f1b"], act=self.act)
            h = ops.gemm(f, L["f2w"], L["f2b"], residual=h)
            if output_hidden_states:
                hidden.append(h)
        last = ops.layernorm(h, w["fin.g"], w["fin.b"], self.eps).view(B, T, C)
        eos = ids.argmax(dim=-1)
        pooled = last[torch.arange(B, device=ids.device), eos]
        hs = tuple(t.view(B, T, C) for t in hidden) if output_hidden_states else None
        return CLIPTextOutput(last, pooled, hs)

    forward = __call__
