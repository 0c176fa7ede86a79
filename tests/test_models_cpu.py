"""Host logic of theatergen_b200/models/models.py (reference models/models.py:53-101) with a stub tokenizer / encoder:
the single co-batched encoder call must return what the reference's two calls return."""
import types

import pytest
import torch

from theatergen_b200.models import models


class StubTokenizer:
    model_max_length = 77

    def __call__(self, prompts, padding=None, max_length=None, truncation=False, return_tensors=None):
        rows = []
        for p in prompts:
            ids = [49406] + [(ord(c) * 131) % 49000 for c in p][: max_length - 2] + [49407]
            rows.append(ids + [49407] * (max_length - len(ids)))
        return types.SimpleNamespace(input_ids=torch.tensor(rows, dtype=torch.int64))


class StubEncoder:
    device = "cpu"

    def __init__(self):
        g = torch.Generator().manual_seed(0)
        self.table = torch.randn(49408, 8, generator=g)
        self.calls = 0

    def __call__(self, ids):
        self.calls += 1
        x = self.table[ids]
        return (x + x.cumsum(1) * 0.01,)  # causal-looking: row t depends on rows <= t


def _reference_semantics(tok, enc, prompts, negative, one_uncond):
    cond = enc(tok(prompts, max_length=77).input_ids)[0]
    unc = enc(tok([negative] * (1 if one_uncond else len(prompts)), max_length=77).input_ids)[0]
    return unc, cond


def test_encode_prompts_matches_two_call_semantics():
    tok, enc = StubTokenizer(), StubEncoder()
    prompts = ["full-body picture of a cat", "full-body picture of a dog", "x"]
    unc, cond = models.encode_prompts(tok, enc, prompts, negative_prompt="blurry", one_uncond_input_only=True)
    assert enc.calls == 1 and unc.shape == (1, 77, 8) and cond.shape == (3, 77, 8)
    r_unc, r_cond = _reference_semantics(tok, StubEncoder(), prompts, "blurry", True)
    assert torch.equal(unc, r_unc) and torch.equal(cond, r_cond)
    full, unc, cond = models.encode_prompts(tok, enc, prompts, negative_prompt="blurry")
    r_unc, r_cond = _reference_semantics(tok, StubEncoder(), prompts, "blurry", False)
    assert torch.equal(unc, r_unc) and torch.equal(cond, r_cond) and torch.equal(full, torch.cat([r_unc, r_cond]))
    only = models.encode_prompts(tok, enc, prompts, negative_prompt="blurry", return_full_only=True)
    assert torch.equal(only, full)


def test_process_input_embeddings():
    unc, cond = torch.randn(1, 77, 8), torch.randn(3, 77, 8)
    full, u, c = models.process_input_embeddings((unc, cond))
    assert full.shape == (6, 77, 8) and torch.equal(full[:3], unc.expand(3, -1, -1)) and torch.equal(full[3:], cond)
    trip = (full, u, c)
    assert models.process_input_embeddings(trip) is trip
    with pytest.raises(ValueError):
        models.process_input_embeddings((full, unc, cond))
    with pytest.raises(ValueError):
        models.process_input_embeddings((unc,))
