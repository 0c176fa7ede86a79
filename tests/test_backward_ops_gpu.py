"""Operator backwards of the guidance step (backward.cu; reference: `torch.autograd.grad(loss, [latents])`,
models/pipelines.py:102) against torch autograd in fp32 on the same fp16-rounded inputs."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rel(a, ref):
    e = (a.float() - ref.float())
    return (e.pow(2).mean().sqrt() / ref.float().pow(2).mean().sqrt().clamp_min(1e-20)).item()


def _r(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).half().cuda()


@pytest.mark.parametrize("M,C", [(64, 1280), (1000, 320), (4096, 640)])
def test_layernorm_bwd(M, C):
    from theatergen_b200 import ops
    x, gm, bt, dy = _r(M, C, seed=1, scale=2.0), 1 + 0.1 * _r(C, seed=2), _r(C, seed=3, scale=0.1), _r(M, C, seed=4)
    xr = x.float().requires_grad_(True)
    F.layer_norm(xr, (C,), gm.float(), bt.float(), 1e-5).backward(dy.float())
    dx = ops.layernorm_bwd(x, gm, dy)
    assert _rel(dx, xr.grad) < 2e-3


@pytest.mark.parametrize("B,HW,C,silu,eps", [(1, 4096, 320, True, 1e-5), (2, 256, 1280, True, 1e-5), (1, 64, 2560, True, 1e-5),
                                              (1, 1024, 640, False, 1e-6), (1, 1024, 960, True, 1e-5)])
def test_groupnorm_bwd(B, HW, C, silu, eps):
    from theatergen_b200 import ops
    x, gm, bt, dy = _r(B, HW, C, seed=1, scale=1.5), 1 + 0.1 * _r(C, seed=2), _r(C, seed=3, scale=0.1), _r(B, HW, C, seed=4)
    xr = x.float().requires_grad_(True)
    y = F.group_norm(xr.transpose(1, 2), 32, gm.float(), bt.float(), eps)
    y = F.silu(y) if silu else y
    y.backward(dy.float().transpose(1, 2))
    dx = ops.groupnorm_bwd(x, gm, bt, dy, eps, silu)
    assert _rel(dx, xr.grad) < 2e-3


@pytest.mark.parametrize("M,N", [(64, 64), (300, 1024), (128, 4096)])
def test_softmax_bwd_rows(M, N):
    from theatergen_b200 import ops
    s, dp = _r(M, N, seed=1, scale=3.0), _r(M, N, seed=2)
    scale = 0.37
    sr = s.float().requires_grad_(True)
    p32 = torch.softmax(sr * scale, -1)
    p32.backward(dp.float())
    p = ops.softmax_rows(s, scale)
    ds = ops.softmax_bwd_rows(p, dp.clone(), scale)
    assert _rel(ds, sr.grad) < 5e-3
    out = torch.empty_like(dp)
    assert torch.equal(ops.softmax_bwd_rows(p, dp, scale, out=out), ds)


def test_geglu_bwd_matches_fused_forward_layout():
    from theatergen_b200 import ops
    from theatergen_b200.weights import pack_geglu
    M, C = 256, 320
    x, w, b = _r(M, C, seed=1), _r(8 * C, C, seed=2, scale=C ** -0.5), _r(8 * C, seed=3, scale=0.1)
    dy = _r(M, 4 * C, seed=4)
    wp, bp = pack_geglu(w, b)
    pre = ops.gemm(x, wp, bp)                           # packed pre-activation
    y = ops.gemm(x, wp, bp, act=ops.TG_ACT_GEGLU)       # the fused forward
    pr = pre.float().requires_grad_(True)
    v = pr.view(M, -1, 64)
    yr = (v[..., :32] * F.gelu(v[..., 32:])).reshape(M, 4 * C)
    assert _rel(y, yr) < 2e-3
    yr.backward(dy.float())
    assert _rel(ops.geglu_bwd(pre, dy), pr.grad) < 2e-3


@pytest.mark.parametrize("N,C,n_ip,with_dprobs", [(256, 1280, 4, True), (64, 1280, 4, True), (1024, 640, 4, False),
                                                  (4096, 320, 4, False), (256, 1280, 0, True)])
def test_cross_ip_bwd_q(N, C, n_ip, with_dprobs):
    from theatergen_b200 import ops
    heads, n_text = 8, 77
    d = C // heads
    T = n_text + n_ip
    q = _r(N, C, seed=1)
    kv = _r(T, 4 * C, seed=2)
    dout = _r(N, C, seed=3)
    ip_scale = 0.4
    dprobs = _r(heads, N, n_text, seed=5) if with_dprobs else None
    kt, vt, ki, vi = kv[:, :C], kv[:, C:2 * C], kv[:, 2 * C:3 * C], kv[:, 3 * C:]
    # forward parity of the probabilities / output with the product kernel (B = 1)
    probs = torch.empty((1, heads, N, n_text), dtype=torch.float16, device="cuda")
    sc = torch.full((1,), ip_scale, dtype=torch.float32, device="cuda")
    o = ops.attention_cross_ip(q, kt, vt, ki if n_ip else None, vi if n_ip else None, 1, heads, n_text, n_ip,
                               ip_scale=sc if n_ip else None, probs=probs)
    qr = q.float().requires_grad_(True)

    def hd(t):
        return t.float().reshape(-1, heads, d).transpose(0, 1)
    pt = torch.softmax(hd(qr) @ hd(kt[:n_text]).transpose(1, 2) * d ** -0.5, -1)
    out = pt @ hd(vt[:n_text])
    if n_ip:
        pi = torch.softmax(hd(qr) @ hd(ki[n_text:]).transpose(1, 2) * d ** -0.5, -1)
        out = out + ip_scale * (pi @ hd(vi[n_text:]))
    out = out.transpose(0, 1).reshape(N, C)
    assert _rel(o, out) < 3e-3 and _rel(probs[0], pt) < 3e-3
    loss = (out * dout.float()).sum() + ((pt * dprobs.float()).sum() if with_dprobs else 0.0)
    loss.backward()
    dq = ops.attention_cross_ip_bwd_q(q, kt, vt, ki if n_ip else None, vi if n_ip else None, heads, n_text, n_ip, ip_scale,
                                      dout, dprobs)
    assert _rel(dq, qr.grad) < 3e-3


def test_upsample_and_col2im_bwd():
    from theatergen_b200 import ops
    B, H, W, C = 2, 16, 16, 320
    dy = _r(B, 2 * H, 2 * W, C, seed=1)
    xr = torch.zeros(B, C, H, W, device="cuda", requires_grad=True)
    F.interpolate(xr, scale_factor=2.0, mode="nearest").backward(dy.float().permute(0, 3, 1, 2))
    assert _rel(ops.upsample2x_bwd(dy).permute(0, 3, 1, 2), xr.grad) < 1e-3
    # stride-2 conv: im2col_s2 + GEMM forward, GEMM(W^T) + col2im backward
    x, w = _r(B, 2 * H, 2 * W, C, seed=2), _r(C, C, 3, 3, seed=3, scale=(9 * C) ** -0.5)
    dyo = _r(B * H * W, C, seed=4)
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    F.conv2d(xr, w.float(), None, stride=2, padding=1).backward(dyo.float().view(B, H, W, C).permute(0, 3, 1, 2))
    wp = w.permute(0, 2, 3, 1).reshape(C, 9 * C).contiguous()
    dcols = ops.gemm(dyo, wp.t().contiguous())
    dx = ops.col2im_s2(dcols, B, 2 * H, 2 * W, C)
    assert _rel(dx.permute(0, 3, 1, 2), xr.grad) < 3e-3


def test_conv3x3_dgrad_by_rotated_weights():
    from theatergen_b200 import ops
    from theatergen_b200.weights import pack_conv3x3, pack_conv3x3_dgrad
    B, H, W, Cin, Cout = 1, 16, 16, 640, 320
    x, w = _r(B, H, W, Cin, seed=1), _r(Cout, Cin, 3, 3, seed=2, scale=(9 * Cin) ** -0.5)
    dy = _r(B, H, W, Cout, seed=3)
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    F.conv2d(xr, w.float(), None, padding=1).backward(dy.float().permute(0, 3, 1, 2))
    dx = ops.conv3x3(dy, pack_conv3x3_dgrad(w))
    assert dx.shape == (B, H, W, Cin) and _rel(dx.permute(0, 3, 1, 2), xr.grad) < 3e-3


def test_guidance_update_and_overflow_flag():
    from theatergen_b200 import ops
    lat, g = _r(1, 4, 64, 64, seed=1), _r(1, 4, 64, 64, seed=2, scale=100.0)
    ref = (lat.float() - g.float() / 256.0 * 0.37).half()
    gout = torch.empty(lat.numel(), dtype=torch.float32, device="cuda")
    bad = torch.zeros(1, dtype=torch.int32, device="cuda")
    out = ops.guidance_update(lat.clone(), g, 1 / 256.0, 0.37, grad_out=gout, bad=bad)
    assert torch.equal(out, ref) and int(bad) == 0 and torch.allclose(gout.view_as(g), g.float() / 256.0)
    g2 = g.clone()
    g2.view(-1)[5] = float("inf")
    l2 = lat.clone()
    ops.guidance_update(l2, g2, 1 / 256.0, 0.37, bad=bad)
    assert int(bad) == 1 and l2.view(-1)[5] == lat.view(-1)[5]
