"""SD-1.5 text encoder (CLIP ViT-L/14 text tower) on the kernel path (reference models/models.py:31,
53-70, models/pipelines.py:283-331) against the transformers implementation ITSELF (CLIPTextModel, random-initialised from
the SD-1.5 text-encoder config with a fixed seed -- no checkpoint exists offline), fp32 on the same GPU, with transformers'
own fp16 run as the noise floor."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _hf(layers=12, seed=0):
    from transformers import CLIPTextConfig, CLIPTextModel
    cfg = CLIPTextConfig(vocab_size=49408, hidden_size=768, intermediate_size=3072, num_hidden_layers=layers,
                         num_attention_heads=12, max_position_embeddings=77, hidden_act="quick_gelu", projection_dim=768,
                         pad_token_id=1, bos_token_id=0, eos_token_id=2)
    torch.manual_seed(seed)
    m = CLIPTextModel(cfg).eval()
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if n.endswith("bias"):
                p.add_(0.02 * torch.randn(p.shape, generator=g))
            elif "norm" in n and n.endswith("weight"):
                p.add_(0.1 * torch.randn(p.shape, generator=g))
            elif "embedding" in n:
                p.mul_(10.0)  # trained CLIP embeddings are O(0.1-1), far above the 0.02 initialisation
    return m


def _ids(B, T=77, seed=5):
    """CLIPTokenizer-shaped rows: BOS 49406, a prompt of random length, EOS 49407, padded with EOS to T."""
    g = torch.Generator().manual_seed(seed)
    ids = torch.full((B, T), 49407, dtype=torch.int64)
    ids[:, 0] = 49406
    for b in range(B):
        n = int(torch.randint(1, T - 2, (1,), generator=g)) if b else 0  # row 0 = the empty (negative) prompt
        ids[b, 1:1 + n] = torch.randint(0, 49000, (n,), generator=g)
    return ids


def _rel(a, ref):
    e = (a.float() - ref).abs()
    return e.max().item(), (e.pow(2).mean().sqrt() / ref.pow(2).mean().sqrt().clamp_min(1e-12)).item()


def test_embed_tokens_exact():
    from theatergen_b200 import ops
    g = torch.Generator().manual_seed(3)
    table = torch.randn(1000, 768, generator=g).half().cuda()
    pos = torch.randn(77, 768, generator=g).half().cuda()
    ids = torch.randint(0, 1000, (3, 77), generator=g).cuda()
    out = ops.embed_tokens(ids, table, pos)
    ref = (table[ids].float() + pos.float()[None]).half().view(-1, 768)
    assert torch.equal(out, ref)
    short = ops.embed_tokens(ids[:, :9].contiguous(), table, pos)  # ragged: T below the table's 77 positions
    assert torch.equal(short, (table[ids[:, :9]].float() + pos.float()[None, :9]).half().view(-1, 768))


@pytest.mark.parametrize("N,d,heads", [(77, 64, 12), (200, 64, 3), (64, 40, 2), (130, 80, 2)])
def test_causal_attention_vs_torch(N, d, heads):
    from theatergen_b200 import ops
    g = torch.Generator().manual_seed(N + d)
    B = 2
    qkv = (torch.randn(B * N, 3 * heads * d, generator=g) * 1.5).half().cuda()
    C = heads * d
    out = ops.attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], B, heads, causal=True)
    q, k, v = (qkv[:, i * C:(i + 1) * C].float().view(B, N, heads, d).transpose(1, 2) for i in range(3))
    ref = torch.nn.functional.scaled_dot_product_attention(q, k, v, is_causal=True).transpose(1, 2).reshape(B * N, C)
    assert torch.isfinite(out).all()
    assert (out.float() - ref).abs().max().item() < 4e-3 * max(1.0, ref.abs().max().item())
    # row 0 attends to key 0 only
    assert torch.allclose(out.view(B, N, C)[:, 0].float(), qkv.view(B, N, 3 * C)[:, 0, 2 * C:].float(), atol=2e-3)


@pytest.mark.parametrize("layers,B", [(2, 3), (12, 5)])
def test_clip_text_parity_vs_transformers(layers, B):
    from theatergen_b200.clip_text import TGCLIPText
    torch.backends.cuda.matmul.allow_tf32 = False
    m = _hf(layers)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    ours = TGCLIPText(sd)
    ids = _ids(B).cuda()
    out = ours(ids, output_hidden_states=True)
    assert out[0].shape == (B, 77, 768) and out[1].shape == (B, 768) and len(out.hidden_states) == layers + 1
    m32 = m.cuda().float()
    for p in m32.parameters():  # both arms see the fp16-rounded weights the kernel path consumes
        p.data = p.data.half().float()
    with torch.no_grad():
        r32 = m32(ids, output_hidden_states=True)
        r16 = m32.half()(ids, output_hidden_states=True)
    for name, a, r, h in (("last_hidden_state", out[0], r32.last_hidden_state, r16.last_hidden_state),
                          ("pooler_output", out.pooler_output, r32.pooler_output, r16.pooler_output),
                          ("hidden_states[-2]", out.hidden_states[-2], r32.hidden_states[-2], r16.hidden_states[-2])):
        mx, rr = _rel(a, r)
        mx16, rr16 = _rel(h, r)
        print(f"CLIP text {layers} layers {name}: max_abs {mx:.3e} rel_rms {rr:.3e} | transformers fp16 {mx16:.3e} {rr16:.3e}")
        assert rr <= max(1.5 * rr16, 3e-3) and mx <= max(2.0 * mx16, 2e-2 * r.abs().max().item())
    if layers == 12:  # same-GPU timing beside transformers' fp16 eager path (reported, not asserted)
        def ms(fn, n=10):
            fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / n
        with torch.no_grad():
            print(f"CLIP text B={B}: kernel path {ms(lambda: ours(ids)):.3f} ms, transformers fp16 eager {ms(lambda: m32(ids)):.3f} ms")


def test_clip_text_rejects_padding_masks_and_cpu():
    from theatergen_b200.clip_text import TGCLIPText
    m = _hf(1)
    ours = TGCLIPText(m.state_dict())
    ids = _ids(2)
    with pytest.raises(RuntimeError):
        ours(ids)
    mask = torch.ones_like(ids)
    mask[1, 50:] = 0
    with pytest.raises(NotImplementedError):
        ours(ids.cuda(), attention_mask=mask.cuda())
    assert ours(ids.cuda(), attention_mask=torch.ones_like(ids).cuda())[0].shape == (2, 77, 768)


def test_encode_prompts_through_kernel_text_encoder():
    """models.encode_prompts (reference models/models.py:53-79) over TGCLIPText == transformers' two separate calls."""
    import types
    from theatergen_b200.clip_text import TGCLIPText
    from theatergen_b200.models import models

    class Tok:
        model_max_length = 77

        def __call__(self, prompts, padding=None, max_length=None, truncation=False, return_tensors=None):
            rows = []
            for p in prompts:
                ids = [49406] + [(ord(c) * 131) % 49000 for c in p][: max_length - 2] + [49407]
                rows.append(ids + [49407] * (max_length - len(ids)))
            return types.SimpleNamespace(input_ids=torch.tensor(rows, dtype=torch.int64))

    m = _hf(2)
    enc = TGCLIPText({k: v.clone() for k, v in m.state_dict().items()})
    prompts = ["full-body picture of a cat", "full-body picture of a small brown dog"]
    unc, cond = models.encode_prompts(Tok(), enc, prompts, negative_prompt="blurry, lowres", one_uncond_input_only=True)
    assert unc.shape == (1, 77, 768) and cond.shape == (2, 77, 768)
    m16 = m.cuda().half()
    with torch.no_grad():
        r_cond = m16(Tok()(prompts, max_length=77).input_ids.cuda())[0].float()
        r_unc = m16(Tok()(["blurry, lowres"], max_length=77).input_ids.cuda())[0].float()
    for a, r in ((unc, r_unc), (cond, r_cond)):
        mx, rr = _rel(a, r)
        assert rr < 5e-3, (mx, rr)
