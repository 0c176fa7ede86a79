"""GPU parity of attention / norm / layout / step kernels against fp32 torch restatements of the same op.
(The whole-path parity against oracle/ lives in test_unet_gpu.py.)"""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rand(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).half().cuda()


def _close(got, ref, atol, rtol, what=""):
    got, ref = got.float(), ref.float()
    err = (got - ref).abs()
    tol = atol + rtol * ref.abs()
    assert bool((err <= tol).all()), f"{what} max err {err.max().item():.4e}, worst excess {(err - tol).max().item():.3e}"


@pytest.mark.parametrize("B,N,C", [(2, 4096, 320), (2, 1024, 640), (2, 256, 1280), (2, 64, 1280), (4, 273, 768)])
def test_self_attention(B, N, C):
    from theatergen_b200 import ops
    heads = 12 if C == 768 else 8
    d = C // heads
    qkv = _rand((B * N, 3 * C), 1)
    out = ops.attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], B, heads)
    q, k, v = [t.float().reshape(B, N, heads, d).transpose(1, 2) for t in (qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:])]
    ref = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B * N, C)
    _close(out, ref, 2e-3, 1e-2, "self-attn")


@pytest.mark.parametrize("B,N,d", [(2, 4096, 40), (8, 256, 40), (2, 1024, 48), (1, 128, 8), (2, 512, 32), (2, 1024, 80),
                                   (10, 1024, 80), (1, 256, 64), (2, 384, 120), (1, 128, 72)])
def test_self_attention_tcgen05(B, N, d):
    """tcgen05/TMEM flash attention on head-padded (64 columns for d <= 56, 128 for d <= 120) Q/K/V vs fp32 SDPA."""
    from theatergen_b200 import ops
    heads = 8
    pad = 64 if d + 1 <= 64 else 128
    qkv = torch.zeros((B * N, 3, heads, pad), dtype=torch.float16, device="cuda")
    qkv[..., :d] = _rand((B * N, 3, heads, d), 1, 1.5)
    flat = qkv.view(B * N, 3 * heads * pad)
    P = heads * pad
    out = ops.attention_tc(flat[:, :P], flat[:, P:2 * P], flat[:, 2 * P:], B, heads, d)
    assert out.shape == (B * N, heads * d)
    q, k, v = [qkv[:, i, :, :d].float().reshape(B, N, heads, d).transpose(1, 2) for i in range(3)]
    ref = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B * N, heads * d)
    _close(out, ref, 2e-3, 1e-2, "tcgen05 self-attn")
    assert torch.equal(out, ops.attention_tc(flat[:, :P], flat[:, P:2 * P], flat[:, 2 * P:], B, heads, d))


@pytest.mark.parametrize("B,N,d,heads", [(10, 256, 160, 8), (2, 128, 160, 8), (1, 1024, 160, 8), (2, 256, 128, 4), (3, 384, 176, 5),
                                         (2, 4096, 40, 8), (10, 1024, 80, 8), (2, 256, 48, 8), (1, 128, 8, 8), (2, 384, 120, 3),
                                         (2, 512, 56, 5), (1, 256, 72, 8)])
def test_self_attention_tcgen05_packed_heads(B, N, d, heads):
    """Heads packed at stride d inside the fused QKV buffer (d = 160: the 16x16 level of SD-1.5, three 64-column chunks per
    head; d = 40 / 80: the 64x64 / 32x32 levels) -- the last box of a head reaches into the next head / past the tensor --
    vs fp32 SDPA and vs the mma.sync kernel.  The NaN columns planted behind V check that nothing outside the tensor map
    reaches an output."""
    from theatergen_b200 import ops
    C = heads * d
    qkv = torch.zeros((B * N, 3 * C + 64), dtype=torch.float16, device="cuda")
    qkv[:, :3 * C] = _rand((B * N, 3 * C), 3, 1.2)
    qkv[:, 3 * C:] = float("nan")  # what a box past the last head of V would read if the tensor map were too wide
    q, k, v = qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:3 * C]
    out = ops.attention_tc(q, k, v, B, heads, d, packed=True)
    assert out.shape == (B * N, C) and bool(torch.isfinite(out).all())
    q4, k4, v4 = [t.float().reshape(B, N, heads, d).transpose(1, 2) for t in (q, k, v)]
    ref = F.scaled_dot_product_attention(q4, k4, v4).transpose(1, 2).reshape(B * N, C)
    _close(out, ref, 2e-3, 1e-2, "tcgen05 self-attn, packed heads")
    if d == 160:  # the mma.sync kernel is instantiated for SD-1.5's head dims only
        old = ops.attention(q, k, v, B, heads)
        assert (out.float() - old.float()).abs().max().item() < 3e-3
    assert torch.equal(out, ops.attention_tc(q, k, v, B, heads, d, packed=True))


def test_self_attention_tcgen05_lazy_rescale():
    """Scores that keep growing along the key axis force the in-TMEM rescale of O (reference exponent raised by > 2^8
    several times per row); fully masked-like rows (huge negative scores early) exercise P underflow."""
    from theatergen_b200 import ops
    B, N, d, heads = 2, 1024, 40, 8
    qkv = torch.zeros((B * N, 3, heads, 64), dtype=torch.float16, device="cuda")
    base = _rand((B * N, 3, heads, d), 7, 1.0)
    ramp = torch.linspace(0.2, 6.0, N, device="cuda").repeat(B)[:, None, None]     # |k| grows with the key index
    base[:, 1] = (base[:, 1].float() * ramp).half()
    qkv[..., :d] = base
    flat = qkv.view(B * N, 3 * heads * 64)
    P = heads * 64
    out = ops.attention_tc(flat[:, :P], flat[:, P:2 * P], flat[:, 2 * P:], B, heads, d)
    q, k, v = [qkv[:, i, :, :d].float().reshape(B, N, heads, d).transpose(1, 2) for i in range(3)]
    ref = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B * N, heads * d)
    _close(out, ref, 3e-3, 1e-2, "tcgen05 self-attn (lazy rescale)")


@pytest.mark.parametrize("variant", [2, 3])
def test_self_attention_tcgen05_alternative_schedules(variant):
    """tg_attn_set_variant: the measured alternatives of the softmax / MMA-issue schedule (P in its own tensor-memory
    columns; two query tiles per CTA with alternating MUFU turns) give the default schedule's results within the
    kernel's tolerance, lazy rescale included."""
    from theatergen_b200 import _lib, ops
    heads = 8
    try:
        for B, N, d, ramp_hi in ((2, 4096, 40, 0.0), (3, 1024, 40, 6.0), (1, 256, 16, 0.0), (2, 384, 48, 4.0), (2, 1024, 80, 0.0)):
            pad = 64 if d + 1 <= 64 else 128
            qkv = torch.zeros((B * N, 3, heads, pad), dtype=torch.float16, device="cuda")
            base = _rand((B * N, 3, heads, d), 11 + N, 1.2)
            if ramp_hi:
                ramp = torch.linspace(0.2, ramp_hi, N, device="cuda").repeat(B)[:, None, None]
                base[:, 1] = (base[:, 1].float() * ramp).half()
            qkv[..., :d] = base
            flat = qkv.view(B * N, 3 * heads * pad)
            P = heads * pad
            _lib.check(_lib.lib().tg_attn_set_variant(0), "tg_attn_set_variant")
            want = ops.attention_tc(flat[:, :P], flat[:, P:2 * P], flat[:, 2 * P:], B, heads, d)
            _lib.check(_lib.lib().tg_attn_set_variant(variant), "tg_attn_set_variant")
            got = ops.attention_tc(flat[:, :P], flat[:, P:2 * P], flat[:, 2 * P:], B, heads, d)
            q, k, v = [qkv[:, i, :, :d].float().reshape(B, N, heads, d).transpose(1, 2) for i in range(3)]
            ref = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B * N, heads * d)
            _close(got, ref, 3e-3, 1e-2, f"tcgen05 self-attn variant {variant} N={N} d={d}")
            assert (got.float() - want.float()).abs().max().item() < 2e-3
            assert torch.equal(got, ops.attention_tc(flat[:, :P], flat[:, P:2 * P], flat[:, 2 * P:], B, heads, d))
    finally:
        _lib.lib().tg_attn_set_variant(0)


def test_self_attention_cross_lengths():
    """Nq != Nk (the resampler: 16 queries over 273 keys) and sharper logits."""
    from theatergen_b200 import ops
    B, Nq, Nk, heads, d = 2, 16, 273, 12, 64
    q, k, v = _rand((B * Nq, heads * d), 1, 2.0), _rand((B * Nk, heads * d), 2, 2.0), _rand((B * Nk, heads * d), 3)
    out = ops.attention(q, k, v, B, heads)
    qq = q.float().reshape(B, Nq, heads, d).transpose(1, 2)
    kk = k.float().reshape(B, Nk, heads, d).transpose(1, 2)
    vv = v.float().reshape(B, Nk, heads, d).transpose(1, 2)
    ref = F.scaled_dot_product_attention(qq, kk, vv).transpose(1, 2).reshape(B * Nq, heads * d)
    _close(out, ref, 2e-3, 1e-2, "cross-length attn")


@pytest.fixture(params=[-1, 0, 1], ids=["auto", "mma_sync", "tcgen05"])
def cross_impl(request):
    """tg_attn_cross_set_impl: every cross-attention test runs on the default dispatch, on the mma.sync kernel alone and on
    the tcgen05 / TMEM kernel alone (all head dims)."""
    from theatergen_b200 import _lib
    _lib.check(_lib.lib().tg_attn_cross_set_impl(request.param), "tg_attn_cross_set_impl")
    yield request.param
    _lib.lib().tg_attn_cross_set_impl(-1)


@pytest.mark.parametrize("B,N,C,n_ip", [(2, 4096, 320, 4), (2, 1024, 640, 4), (4, 256, 1280, 4), (2, 64, 1280, 16),
                                        (2, 256, 1280, 0), (2, 200, 320, 16), (1, 64, 640, 4)])
def test_cross_attention_ip(B, N, C, n_ip, cross_impl):
    from theatergen_b200 import ops
    heads, n_text = 8, 77
    d = C // heads
    T = n_text + n_ip
    q = _rand((B * N, C), 1)
    kv = _rand((B * T, 4 * C), 2)
    scale = torch.tensor([0.0, 0.4, 1.0, 0.1][:B] if B <= 4 else [0.4] * B, dtype=torch.float32, device="cuda")
    probs = torch.zeros((B // 2, heads, N, n_text), dtype=torch.float16, device="cuda")
    out = ops.attention_cross_ip(q, kv[:, :C], kv[:, C:2 * C], kv[:, 2 * C:3 * C] if n_ip else None,
                                 kv[:, 3 * C:] if n_ip else None, B, heads, n_text, n_ip, ip_scale=scale, probs=probs,
                                 probs_b0=B // 2)
    qq = q.float().reshape(B, N, heads, d).transpose(1, 2)
    kvb = kv.float().reshape(B, T, 4, heads, d)
    kt, vt = kvb[:, :n_text, 0].transpose(1, 2), kvb[:, :n_text, 1].transpose(1, 2)
    p = torch.softmax(qq @ kt.transpose(-1, -2) * d ** -0.5, -1)
    ref = p @ vt
    if n_ip:
        ki, vi = kvb[:, n_text:, 2].transpose(1, 2), kvb[:, n_text:, 3].transpose(1, 2)
        pi = torch.softmax(qq @ ki.transpose(-1, -2) * d ** -0.5, -1)
        ref = ref + scale.view(B, 1, 1, 1) * (pi @ vi)
    ref = ref.transpose(1, 2).reshape(B * N, C)
    _close(out, ref, 3e-3, 1e-2, "cross-attn")
    _close(probs, p[B // 2:], 1e-3, 1e-2, "captured probs")
    # additive attention mask (attention_processor.py:202-208 baddbmm with beta = 1): masked text tokens get -inf-like bias
    mask = torch.zeros((B, n_text), dtype=torch.float16, device="cuda")
    mask[:, 50:] = -10000.0
    mask[0, 3] = -2.5
    out_m = ops.attention_cross_ip(q, kv[:, :C], kv[:, C:2 * C], kv[:, 2 * C:3 * C] if n_ip else None,
                                   kv[:, 3 * C:] if n_ip else None, B, heads, n_text, n_ip, ip_scale=scale, add_mask=mask)
    pm = torch.softmax(qq @ kt.transpose(-1, -2) * d ** -0.5 + mask.float().view(B, 1, 1, n_text), -1)
    refm = pm @ vt
    if n_ip:
        refm = refm + scale.view(B, 1, 1, 1) * (pi @ vi)
    _close(out_m, refm.transpose(1, 2).reshape(B * N, C), 3e-3, 1e-2, "cross-attn + mask")


@pytest.mark.parametrize("B,HW,C1,C2,silu", [(2, 4096, 320, 0, True), (2, 4096, 640, 320, True),
                                             (2, 256, 1280, 640, True), (2, 64, 1280, 1280, True),
                                             (2, 1024, 640, 0, False)])
def test_groupnorm(B, HW, C1, C2, silu):
    from theatergen_b200 import ops
    x1 = _rand((B, HW, C1), 1) + 0.5
    x2 = _rand((B, HW, C2), 2, 2.0) if C2 else None
    C = C1 + C2
    gamma, beta = _rand((C,), 3) * 0.1 + 1, _rand((C,), 4, 0.1)
    out = ops.groupnorm(x1, gamma, beta, 1e-5, x2=x2, silu=silu)
    x = x1 if x2 is None else torch.cat([x1, x2], -1)
    ref = F.group_norm(x.float().transpose(1, 2), 32, gamma.float(), beta.float(), 1e-5)
    if silu:
        ref = F.silu(ref)
    _close(out, ref.transpose(1, 2), 4e-3, 4e-3, "groupnorm")


@pytest.mark.parametrize("B,HW,C1,C2", [(10, 64, 1280, 1280), (10, 256, 1280, 0), (5, 1024, 640, 0), (3, 1024, 320, 0),
                                        (10, 1024, 1280, 640), (3, 4096, 640, 320)])
def test_groupnorm_paths_agree_and_are_batch_independent(B, HW, C1, C2):
    """single-kernel cluster path vs two-pass kernels (forced through tg_groupnorm_tune) on the same input, and an
    image's result must not depend on the batch it travels in (co-batched characters == solo)."""
    from theatergen_b200 import _lib, ops
    x1 = _rand((B, HW, C1), 1) + 0.5
    x2 = _rand((B, HW, C2), 2, 2.0) if C2 else None
    C = C1 + C2
    gamma, beta = _rand((C,), 3) * 0.1 + 1, _rand((C,), 4, 0.1)
    x = x1 if x2 is None else torch.cat([x1, x2], -1)
    ref = F.silu(F.group_norm(x.float().transpose(1, 2), 32, gamma.float(), beta.float(), 1e-5)).transpose(1, 2)
    try:
        outs = []
        for fused in (2, 0):
            _lib.lib().tg_groupnorm_tune(0, 0, fused)
            out = ops.groupnorm(x1, gamma, beta, 1e-5, x2=x2, silu=True)
            _close(out, ref, 4e-3, 4e-3, f"groupnorm fused={fused}")
            solo = ops.groupnorm(x1[1:2].contiguous(), gamma, beta, 1e-5,
                                 x2=None if x2 is None else x2[1:2].contiguous(), silu=True)
            assert torch.equal(solo[0], out[1]), f"fused={fused}: result depends on the batch"
            outs.append(out)
    finally:
        _lib.lib().tg_groupnorm_tune(0, 0, 1)


@pytest.mark.parametrize("M,C", [(8192, 320), (2048, 640), (512, 1280), (273, 768), (5, 768)])
def test_layernorm(M, C):
    from theatergen_b200 import ops
    x = _rand((M, C), 1, 3.0) + 1.0
    gamma, beta = _rand((C,), 3) * 0.1 + 1, _rand((C,), 4, 0.1)
    out = ops.layernorm(x, gamma, beta)
    _close(out, F.layer_norm(x.float(), (C,), gamma.float(), beta.float(), 1e-5), 3e-3, 3e-3, "layernorm")


def test_layout_kernels():
    from theatergen_b200 import ops
    x = _rand((2, 16, 16, 640), 1)
    up = ops.upsample2x(x)
    ref = F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest").permute(0, 2, 3, 1)
    assert torch.equal(up.float(), ref)
    a = ops.im2col_s2(x)
    unf = F.unfold(x.float().permute(0, 3, 1, 2), 3, padding=1, stride=2)  # [B, C*9, L] (c-major, tap-minor)
    unf = unf.reshape(2, 640, 9, 64).permute(0, 3, 2, 1).reshape(2 * 64, 9 * 640)
    assert torch.equal(a.float(), unf)
    n = _rand((2, 320, 8, 8), 2)
    assert torch.equal(ops.nchw_to_nhwc(n), n.permute(0, 2, 3, 1).contiguous())
    y = _rand((1000,), 3)
    y0 = y.clone()
    ops.add_inplace(y, y0)
    assert torch.equal(y, y0 + y0)


def test_downsample_conv_via_im2col():
    from theatergen_b200 import ops
    from theatergen_b200.weights import pack_conv3x3
    B, H, W, C = 2, 32, 32, 640
    x = _rand((B, H, W, C), 1)
    w = _rand((C, C, 3, 3), 2, (9 * C) ** -0.5)
    bias = _rand((C,), 3)
    out = ops.gemm(ops.im2col_s2(x), pack_conv3x3(w).reshape(C, 9 * C), bias)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), bias.float(), stride=2, padding=1)
    _close(out.reshape(B, H // 2, W // 2, C), ref.permute(0, 2, 3, 1), 3e-3, 3e-3, "downsample")


def test_conv_in_out():
    from theatergen_b200 import ops
    from theatergen_b200.weights import pack_conv3x3
    x = _rand((2, 4, 64, 64), 1)
    w, b = _rand((320, 4, 3, 3), 2, 1 / 6), _rand((320,), 3, 0.1)
    out = ops.conv_in(x, w, b)
    ref = F.conv2d(x.float(), w.float(), b.float(), padding=1).permute(0, 2, 3, 1)
    _close(out, ref, 2e-3, 2e-3, "conv_in")
    out_g = ops.conv_in_gemm(x, ops.pack_conv_in(w), b)  # the path TGUNet uses: im2col (K padded to 64) + tcgen05 GEMM
    _close(out_g, ref, 2e-3, 2e-3, "conv_in_gemm")
    x3 = _rand((1, 3, 128, 256), 7).abs()  # ControlNet hint embedder: 3 channels, SiLU, non-square
    w3, b3 = _rand((64, 3, 3, 3), 8, 0.3), _rand((64,), 9, 0.1)
    out3 = ops.conv_in_gemm(x3, ops.pack_conv_in(w3), b3, silu=True)
    _close(out3, F.silu(F.conv2d(x3.float(), w3.float(), b3.float(), padding=1)).permute(0, 2, 3, 1), 2e-3, 2e-3, "conv_in_gemm silu")
    h = _rand((2, 64, 64, 320), 4)
    w2, b2 = _rand((4, 320, 3, 3), 5, (2880) ** -0.5), _rand((4,), 6, 0.1)
    out2 = ops.conv_out(h, pack_conv3x3(w2), b2)
    ref2 = F.conv2d(h.float().permute(0, 3, 1, 2), w2.float(), b2.float(), padding=1)
    _close(out2, ref2, 2e-3, 2e-3, "conv_out")
    w64, b64 = ops.pack_conv_out_tc(pack_conv3x3(w2), b2)  # the path TGUNet uses: tensor-core conv, columns padded to 64
    _close(ops.conv_out_tc(h, w64, b64, 4), ref2, 2e-3, 2e-3, "conv_out_tc")
    h5 = _rand((5, 16, 8, 320), 10)  # pixel count not a multiple of the CTA's 64 pixels x batch
    out5 = ops.conv_out(h5, pack_conv3x3(w2), b2)
    _close(out5, F.conv2d(h5.float().permute(0, 3, 1, 2), w2.float(), b2.float(), padding=1), 2e-3, 2e-3, "conv_out ragged")


def test_timestep_embedding():
    from theatergen_b200 import ops
    t = torch.tensor([981.0, 981.0, 1.0, 501.0], device="cuda")
    out = ops.timestep_embedding(t)
    half = 160
    f = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32, device="cuda") / half)
    e = t[:, None] * f[None]
    ref = torch.cat([torch.cos(e), torch.sin(e)], -1)
    _close(out, ref, 2e-3, 0, "timestep embedding")


def test_cfg_ddim_composite():
    from theatergen_b200 import ops
    K = 3
    eps, lat = _rand((2 * K, 4, 64, 64), 1), _rand((K, 4, 64, 64), 2)
    frozen, mask = _rand((K, 4, 64, 64), 3), (torch.rand(64, 64) > 0.5).half().cuda()
    a_t, a_p, g = 0.3, 0.45, 7.5
    coef = torch.tensor([g, a_t ** 0.5, (1 - a_t) ** 0.5, a_p ** 0.5, (1 - a_p) ** 0.5], device="cuda")
    e = eps[:K].float() + g * (eps[K:].float() - eps[:K].float())
    x0 = (lat.float() - (1 - a_t) ** 0.5 * e) / a_t ** 0.5
    ref = a_p ** 0.5 * x0 + (1 - a_p) ** 0.5 * e
    l1, trace = lat.clone(), torch.zeros_like(lat)
    ops.cfg_ddim_composite(eps, l1, coef, latents_copy=trace)
    _close(l1, ref, 1.5e-2, 2e-3, "cfg+ddim")
    assert torch.equal(l1, trace)
    l2 = lat.clone()
    ops.cfg_ddim_composite(eps, l2, coef, frozen=frozen, mask=mask)
    ref2 = frozen.float() * mask.float() + ref * (1 - mask.float())
    _close(l2, ref2, 1.5e-2, 2e-3, "cfg+ddim+composite")
    assert torch.equal(l2[:, :, mask.bool()], frozen[:, :, mask.bool()])


@pytest.mark.parametrize("B,N,C", [(10, 4096, 320), (8, 4096, 320), (10, 1024, 640), (3, 1000, 320)])
def test_cross_attention_query_tile_walk_is_batch_independent(B, N, C, cross_impl):
    """A CTA of the cross-attention kernels (attn_cross_ip_kernel / attn_cross_tc_kernel) walks 1, 2, 4 or 8 query tiles depending on the launch size (K/V staged once per CTA).
    Every image of a large batch must come out bit-identical to the same image run alone (one tile per CTA), captured
    probabilities included; N = 1000 leaves a ragged last tile."""
    from theatergen_b200 import ops
    heads, n_text, n_ip = 8, 77, 4
    T = n_text + n_ip
    g = torch.Generator().manual_seed(B * 7 + N)
    q = torch.randn(B * N, C, generator=g).half().cuda()
    kv = torch.randn(B * T, 4 * C, generator=g).half().cuda()
    sc = torch.linspace(0.0, 0.9, B).cuda()
    probs = torch.empty((B, heads, N, n_text), dtype=torch.float16, device="cuda")
    out = ops.attention_cross_ip(q, kv[:, :C], kv[:, C:2 * C], kv[:, 2 * C:3 * C], kv[:, 3 * C:], B, heads, n_text, n_ip,
                                 ip_scale=sc, probs=probs)
    for b in (0, B // 2, B - 1):
        kvb = kv[b * T:(b + 1) * T]
        pb = torch.empty((1, heads, N, n_text), dtype=torch.float16, device="cuda")
        ob = ops.attention_cross_ip(q[b * N:(b + 1) * N], kvb[:, :C], kvb[:, C:2 * C], kvb[:, 2 * C:3 * C], kvb[:, 3 * C:], 1,
                                    heads, n_text, n_ip, ip_scale=sc[b:b + 1].contiguous(), probs=pb)
        assert torch.equal(out[b * N:(b + 1) * N], ob) and torch.equal(probs[b], pb[0])
