"""Prompt-embedding helpers of the reference's `models/models.py` (encode_prompts :53-79, process_input_embeddings
:81-101) for the kernel-path text encoder (`theatergen_b200.clip_text.TGCLIPText`, or any callable with the
`text_encoder(input_ids)[0]` surface).

Difference from the reference, results unchanged: the negative and the positive prompts go through the encoder as ONE
batch (rows of a transformer batch are independent) instead of two calls -- K + 1 rows for a turn with K characters and
one shared negative prompt (theatergen.py:378).
"""
from __future__ import annotations

import torch


def _device_of(text_encoder):
    dev = getattr(text_encoder, "device", None)
    return torch.device(dev) if dev is not None else torch.device("cuda")


def encode_prompts(tokenizer, text_encoder, prompts, negative_prompt="", return_full_only=False, one_uncond_input_only=False):
    """-> (uncond [1|n,77,768], cond [n,77,768]) when `one_uncond_input_only`, else (cat, uncond, cond) or just cat."""
    if negative_prompt == "":
        print("Note that negative_prompt is an empty string")
    prompts = [prompts] if isinstance(prompts, str) else list(prompts)  # the tokenizer accepts a bare string too
    cond_ids = tokenizer(prompts, padding="max_length", max_length=tokenizer.model_max_length, truncation=True,
                         return_tensors="pt").input_ids
    n_uncond = 1 if one_uncond_input_only else len(prompts)
    uncond_ids = tokenizer([negative_prompt] * n_uncond, padding="max_length", max_length=cond_ids.shape[-1],
                           return_tensors="pt").input_ids
    dev = _device_of(text_encoder)
    # identical negative rows are encoded once and expanded afterwards
    ids = torch.cat([uncond_ids[:1], cond_ids], 0).to(dev)
    with torch.no_grad():
        emb = text_encoder(ids)[0]
    uncond, cond = emb[:1].expand(n_uncond, -1, -1).contiguous(), emb[1:]
    if one_uncond_input_only:
        return uncond, cond
    full = torch.cat([uncond, cond])
    return full if return_full_only else (full, uncond, cond)


def process_input_embeddings(input_embeddings):
    """(uncond, cond) or (full, uncond, cond) -> (full, uncond, cond) with the negative rows first."""
    if not isinstance(input_embeddings, (tuple, list)):
        raise TypeError("input_embeddings must be a tuple or list")
    if len(input_embeddings) == 3:
        _, uncond, cond = input_embeddings
        if uncond.shape[0] != cond.shape[0]:
            raise ValueError(f"{uncond.shape[0]} != {cond.shape[0]}")
        return input_embeddings
    if len(input_embeddings) == 2:
        uncond, cond = input_embeddings
        if uncond.shape[0] == 1:
            uncond = uncond.expand(cond.shape)
        return torch.cat((uncond, cond), dim=0), uncond, cond
    raise ValueError(f"input_embeddings length: {len(input_embeddings)}")
