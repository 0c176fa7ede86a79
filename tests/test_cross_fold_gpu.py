"""Folded cross-attention (tg_cross_scores / tg_cross_apply: attn2.to_q folded into the per-trajectory key operand,
attn2.to_out into the value operand, both softmaxes in the GEMM epilogue) against
  * an fp32 torch restatement of IPAttnProcessor.__call__ (ip_adapter/attention_processor.py:461-553), and
  * the attention-kernel path it replaces (tg_gemm + tg_attention_cross_ip + tg_gemm),
per layer and through the whole UNet."""
import pytest
import torch

pytestmark = pytest.mark.gpu
HEADS = 8


def _rand(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).half().cuda()


def _reference(x, h, ctx, w, B, HW, C, n_text, n_ip, ip_scale):
    """fp32: to_q -> text softmax . V + ip_scale[b] * image softmax . V_ip -> to_out + bias + residual, and the text probabilities"""
    d = C // HEADS
    f = {k: v.float() for k, v in w.items()}
    q = (x.float() @ f["q"].T).view(B, HW, HEADS, d).transpose(1, 2)

    def heads(t):
        return t.view(B, -1, HEADS, d).transpose(1, 2)
    ct, ci = ctx[:, :n_text].float(), ctx[:, n_text:].float()
    pt = torch.softmax(q @ heads(ct @ f["k"].T).transpose(-1, -2) * d ** -0.5, -1)
    a = pt @ heads(ct @ f["v"].T)
    if n_ip:
        pi = torch.softmax(q @ heads(ci @ f["ki"].T).transpose(-1, -2) * d ** -0.5, -1)
        a = a + ip_scale.view(B, 1, 1, 1) * (pi @ heads(ci @ f["vi"].T))
    out = a.transpose(1, 2).reshape(B * HW, C) @ f["o"].T + f["ob"] + h.float()
    return out, pt


@pytest.mark.parametrize("B,HW,C,n_ip", [(2, 1024, 640, 4), (10, 256, 1280, 4), (2, 256, 1280, 0), (4, 1024, 640, 16),
                                         (2, 4096, 320, 4)])
def test_folded_layer_matches_reference_and_attention_kernel(B, HW, C, n_ip):
    from theatergen_b200 import ops
    from theatergen_b200.unet import _fold_pair
    n_text, S = 77, ops.CROSS_SLOTS
    sc = C ** -0.5
    w = dict(q=_rand((C, C), 1, sc), k=_rand((C, 768), 2, 0.03), v=_rand((C, 768), 3, 0.03), ki=_rand((C, 768), 4, 0.03),
             vi=_rand((C, 768), 5, 0.03), o=_rand((C, C), 6, sc), ob=_rand((C,), 7, 0.1))
    x, h = _rand((B * HW, C), 8), _rand((B * HW, C), 9)
    ctx = _rand((B, n_text + n_ip, 768), 10)
    ip_scale = torch.linspace(0.0, 0.9, B, device="cuda", dtype=torch.float32)
    ref, ref_p = _reference(x, h, ctx, w, B, HW, C, n_text, n_ip, ip_scale)

    # the path it replaces
    T = n_text + n_ip
    c2 = ctx.reshape(B * T, 768)
    kv = ops.gemm(c2, torch.cat([w["k"], w["v"], w["ki"], w["vi"]], 0).contiguous())
    q = ops.gemm(x, w["q"])
    probs_old = torch.empty((B, HEADS, HW, n_text), dtype=torch.float16, device="cuda")
    a = ops.attention_cross_ip(q, kv[:, :C], kv[:, C:2 * C], kv[:, 2 * C:3 * C] if n_ip else None,
                               kv[:, 3 * C:] if n_ip else None, B, HEADS, n_text, n_ip,
                               ip_scale=ip_scale if n_ip else None, probs=probs_old)
    old = ops.gemm(a, w["o"], w["ob"], residual=h)

    # folded
    t = dict(C=C, q2w=w["q"], o2w=w["o"])
    fkt = _fold_pair(t, w["k"], w["v"])
    kf = torch.zeros((B, HEADS, S, C), dtype=torch.float16, device="cuda")
    vf = torch.zeros((B, C, HEADS, S), dtype=torch.float16, device="cuda")
    pr = ops.gemm(ctx[:, :n_text].reshape(B * n_text, 768), fkt).view(B, n_text, 2, HEADS, C)
    kf[:, :, :n_text] = pr[:, :, 0].permute(0, 2, 1, 3)
    vf[..., :n_text] = pr[:, :, 1].permute(0, 3, 2, 1)
    if n_ip:
        fki = _fold_pair(t, w["ki"], w["vi"])
        pi = ops.gemm(ctx[:, n_text:].reshape(B * n_ip, 768), fki).view(B, n_ip, 2, HEADS, C)
        kf[:, :, 80:80 + n_ip] = pi[:, :, 0].permute(0, 2, 1, 3)
        vf[..., 80:80 + n_ip] = pi[:, :, 1].permute(0, 3, 2, 1)
    pm = ops.cross_scores(x, kf.view(B, HEADS * S, C), n_text, n_ip, ip_scale if n_ip else None)
    new = ops.cross_apply(pm, vf.view(B, C, HEADS * S), w["ob"], residual=h)
    torch.cuda.synchronize()

    p4 = pm.view(B, HW, HEADS, S).float()
    # structure of the probability rows: text softmax sums to 1, image softmax to ip_scale[b], padding slots are exact zeros
    assert bool((p4[..., n_text:80] == 0).all()) and bool((p4[..., 80 + n_ip:] == 0).all())
    assert (p4[..., :n_text].sum(-1) - 1).abs().max().item() < 5e-3
    if n_ip:
        assert (p4[..., 80:80 + n_ip].sum(-1) - ip_scale.view(B, 1, 1)).abs().max().item() < 5e-3
    probs_new = p4[..., :n_text].permute(0, 2, 1, 3)
    e_p_new = (probs_new - ref_p).abs().max().item()
    e_p_old = (probs_old.float() - ref_p).abs().max().item()
    e_new = (new.float() - ref).abs().max().item()
    e_old = (old.float() - ref).abs().max().item()
    print(f"B={B} HW={HW} C={C} n_ip={n_ip}: out err folded {e_new:.3e} / attention kernel {e_old:.3e}; "
          f"probs err folded {e_p_new:.3e} / attention kernel {e_p_old:.3e}")
    # the folded path must be as close to the fp32 reference as the path it replaces (fp16 operands, fp32 accumulate)
    assert e_new <= max(2.0 * e_old, 8e-3), (e_new, e_old)
    assert e_p_new <= max(2.0 * e_p_old, 2e-3), (e_p_new, e_p_old)


def test_folded_scores_are_batch_independent():
    """An image's probabilities do not depend on its position in the batch or on its neighbours (per-image B operand)."""
    from theatergen_b200 import ops
    B, HW, C, S = 6, 256, 1280, ops.CROSS_SLOTS
    x = _rand((B * HW, C), 1)
    kf = _rand((B, HEADS * S, C), 2, 0.05)
    sc = torch.full((B,), 0.4, device="cuda")
    full = ops.cross_scores(x, kf, 77, 4, sc)
    for b in (0, 3, 5):
        one = ops.cross_scores(x[b * HW:(b + 1) * HW].contiguous(), kf[b:b + 1].contiguous(), 77, 4, sc[b:b + 1].contiguous())
        assert torch.equal(one, full[b * HW:(b + 1) * HW]), b


def test_unet_forward_folded_vs_attention_kernels():
    """Whole UNet: the folded layers (every level but the 8x8 mid block) against the same network on the attention kernels, both against
    the fp32 oracle; captured guidance maps included."""
    from oracle import unet as OU
    from theatergen_b200 import unet as U
    sd = OU.synth_state_dict(1234)
    g = torch.Generator().manual_seed(3)
    lat = torch.randn(2, 4, 64, 64, generator=g).half().cuda()
    ctx = torch.randn(2, 81, 768, generator=g).half().cuda()
    keys = [("down", 2, 1, 0), ("up", 1, 0, 0)]
    outs, caps = {}, {}
    saved = U.CROSS_FOLD_C
    try:
        for name, fold in (("folded", (320, 640, 1280)), ("kernels", ())):
            U.CROSS_FOLD_C = fold
            net = U.TGUNet(sd, "cuda:0")
            for p in net.attn_processors.values():
                if hasattr(p, "scale"):
                    p.scale = 0.4
            cap = {}
            outs[name] = net(lat, 601, encoder_hidden_states=ctx,
                             cross_attention_kwargs=dict(save_attn_to_dict=cap, save_keys=keys)).sample.float()
            caps[name] = cap
            assert any("fkt" in t for t in net.transformers.values()) == bool(fold)
    finally:
        U.CROSS_FOLD_C = saved
    sd32 = {k: v.half().float().cuda() for k, v in sd.items()}
    with torch.no_grad():
        ref = OU.unet_forward(sd32, lat.float(), 601, ctx.float(), ip_scale=0.4)
    e_f = (outs["folded"] - ref).abs().max().item()
    e_k = (outs["kernels"] - ref).abs().max().item()
    print(f"max|eps - fp32 oracle|: folded {e_f:.3e}, attention kernels {e_k:.3e}")
    assert e_f <= max(1.5 * e_k, 4e-3), (e_f, e_k)
    assert set(caps["folded"]) == set(caps["kernels"]) and len(caps["folded"]) > 0
    for k in caps["kernels"]:
        a, b = caps["folded"][k].float(), caps["kernels"][k].float()
        assert a.shape == b.shape
        assert (a - b).abs().max().item() < 2e-3, k


def test_unet_forward_ff2_proj_out_as_one_gemm():
    """ff.net.2 (+h) and proj_out (+x) as ONE GEMM over [h | ff] with the weights [Wpo | Wpo W2] (unet.pack_transformer) against
    the two GEMMs of the reference's op order (ip_adapter/attention.py:339-367, transformer_2d.py:330-342), both vs the oracle."""
    from oracle import unet as OU
    from theatergen_b200 import unet as U
    sd = OU.synth_state_dict(1234)
    g = torch.Generator().manual_seed(5)
    lat = torch.randn(2, 4, 64, 64, generator=g).half().cuda()
    ctx = torch.randn(2, 81, 768, generator=g).half().cuda()
    outs = {}
    saved = U.FOLD_FF2_PROJ_OUT
    try:
        for name, fold in (("one", True), ("two", False)):
            U.FOLD_FF2_PROJ_OUT = fold
            net = U.TGUNet(sd, "cuda:0")
            assert any("fpw" in t for t in net.transformers.values()) == fold
            outs[name] = net(lat, 301, encoder_hidden_states=ctx).sample.float()
    finally:
        U.FOLD_FF2_PROJ_OUT = saved
    sd32 = {k: v.half().float().cuda() for k, v in sd.items()}
    with torch.no_grad():
        ref = OU.unet_forward(sd32, lat.float(), 301, ctx.float(), ip_scale=1.0)
    e1 = (outs["one"] - ref).abs().max().item()
    e2 = (outs["two"] - ref).abs().max().item()
    print(f"max|eps - fp32 oracle|: one GEMM {e1:.3e}, two GEMMs {e2:.3e}")
    assert e1 <= max(1.5 * e2, 4e-3), (e1, e2)
