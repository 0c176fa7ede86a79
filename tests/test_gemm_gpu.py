"""GPU parity of the tcgen05 GEMM / implicit-GEMM conv against an fp32 torch reference of the same op."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rand(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).half().cuda()


def _close(got, ref, atol, rtol):
    got, ref = got.float(), ref.float()
    err = (got - ref).abs()
    tol = atol + rtol * ref.abs()
    assert bool((err <= tol).all()), f"max err {err.max().item():.4e} (tol violated by {(err - tol).max().item():.3e})"


@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (128, 128, 64), (256, 128, 320), (8192, 320, 320),
                                   (2048, 640, 2560), (512, 1280, 1280), (162, 1280, 768), (2, 1280, 320),
                                   (300, 192, 128), (512, 1280, 11520), (128, 1280, 2560), (2048, 960, 320),
                                   (1024, 2560, 640), (256, 49920, 768), (2560, 1280, 1280), (2560, 1280, 11520),
                                   (40960, 320, 1280), (10240, 640, 2560), (6400, 1280, 5120), (2432, 1280, 2560)])
def test_gemm_plain(M, N, K):
    from theatergen_b200 import ops
    a, w = _rand((M, K), 1), _rand((N, K), 2, K ** -0.5)
    out = ops.gemm(a, w)
    ref = a.float() @ w.float().t()
    _close(out, ref, 2e-3, 2e-3)


def test_gemm_epilogues():
    from theatergen_b200 import ops
    M, N, K = 512, 640, 320
    a, w = _rand((M, K), 1), _rand((N, K), 2, K ** -0.5)
    bias, res = _rand((N,), 3), _rand((M, N), 4)
    rowbias = _rand((4, N), 5)
    out = ops.gemm(a, w, bias, rowbias=rowbias, rows_per_image=128, residual=res, out_scale=0.5)
    ref = (a.float() @ w.float().t() + bias.float() + rowbias.float().repeat_interleave(128, 0) + res.float()) * 0.5
    _close(out, ref, 3e-3, 2e-3)
    out = ops.gemm(a, w, bias, act=ops.TG_ACT_SILU)
    _close(out, torch.nn.functional.silu(a.float() @ w.float().t() + bias.float()), 3e-3, 2e-3)


def test_gemm_splitk_with_epilogue_is_deterministic():
    from theatergen_b200 import ops
    M, N, K = 512, 1280, 5760
    a, w = _rand((M, K), 1), _rand((N, K), 2, K ** -0.5)
    bias, res, rowbias = _rand((N,), 3), _rand((M, N), 4), _rand((8, N), 5)
    out = ops.gemm(a, w, bias, rowbias=rowbias, rows_per_image=64, residual=res)
    ref = a.float() @ w.float().t() + bias.float() + rowbias.float().repeat_interleave(64, 0) + res.float()
    _close(out, ref, 3e-3, 2e-3)
    for _ in range(3):
        assert torch.equal(out, ops.gemm(a, w, bias, rowbias=rowbias, rows_per_image=64, residual=res))


def test_gemm_residual_is_deterministic_on_partial_waves():
    """Shapes whose tile count does not fill whole waves (50 / 320 pair tiles on 74 workers): bit-reproducible results."""
    from theatergen_b200 import ops
    for M, N, K in ((2560, 1280, 5760), (40960, 320, 640)):
        a, w = _rand((M, K), 1), _rand((N, K), 2, K ** -0.5)
        bias, res = _rand((N,), 3), _rand((M, N), 4)
        out = ops.gemm(a, w, bias, residual=res)
        ref = a.float() @ w.float().t() + bias.float() + res.float()
        _close(out, ref, 3e-3, 2e-3)
        for _ in range(3):
            assert torch.equal(out, ops.gemm(a, w, bias, residual=res))


def test_gemm_two_sources_and_strided_out():
    from theatergen_b200 import ops
    M, N, K1, K2 = 256, 128, 128, 192
    a1, a2, w = _rand((M, K1), 1), _rand((M, K2), 2), _rand((N, K1 + K2), 3, (K1 + K2) ** -0.5)
    big = torch.zeros((M, 3 * N), dtype=torch.float16, device="cuda")
    ops.gemm(a1, w, a2=a2, out=big[:, N:2 * N])
    ref = torch.cat([a1, a2], 1).float() @ w.float().t()
    _close(big[:, N:2 * N], ref, 2e-3, 2e-3)
    assert float(big[:, :N].abs().max()) == 0 and float(big[:, 2 * N:].abs().max()) == 0


def test_gemm_geglu():
    from theatergen_b200 import ops
    from theatergen_b200.weights import pack_geglu
    M, C = 256, 320
    a = _rand((M, C), 1)
    w, b = _rand((8 * C, C), 2, C ** -0.5), _rand((8 * C,), 3, 0.1)
    wp, bp = pack_geglu(w, b)
    out = ops.gemm(a, wp, bp, act=ops.TG_ACT_GEGLU)
    proj = (a.float() @ w.float().t() + b.float()).half().float()
    val, gate = proj.chunk(2, dim=-1)
    _close(out, val * torch.nn.functional.gelu(gate), 3e-3, 3e-3)


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(2, 64, 64, 64, 64), (2, 64, 64, 320, 320), (2, 32, 32, 640, 640),
                                            (2, 16, 16, 1280, 1280), (2, 8, 8, 1280, 1280), (4, 8, 8, 2560, 1280),
                                            (1, 32, 32, 960, 640), (10, 16, 16, 1280, 1280), (10, 64, 64, 320, 320)])
def test_conv3x3(B, H, W, Cin, Cout):
    from theatergen_b200 import ops
    x = _rand((B, H, W, Cin), 1)
    w = _rand((Cout, 3, 3, Cin), 2, (9 * Cin) ** -0.5)
    bias, temb = _rand((Cout,), 3), _rand((B, Cout), 4)
    out = ops.conv3x3(x, w, bias, rowbias=temb)
    ref = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), w.float().permute(0, 3, 1, 2), bias.float(),
                                     padding=1) + temb.float()[:, :, None, None]
    _close(out, ref.permute(0, 2, 3, 1), 3e-3, 3e-3)


def test_forced_and_tuned_tiles_agree_with_the_model_choice():
    """tg_gemm_force_tile / tg_gemm_set_tuned (tools/gemm_autotune.py): every valid configuration computes the same GEMM;
    invalid ones are refused; the table is consulted per shape and cleared again."""
    from theatergen_b200 import _lib, ops
    L = _lib.lib()
    g = torch.Generator().manual_seed(11)
    M, N, K = 1000, 640, 1280
    a = torch.randn(M, K, generator=g).half().cuda()
    w = (torch.randn(N, K, generator=g) * K ** -0.5).half().cuda()
    b = torch.randn(N, generator=g).half().cuda()
    ref = a.float() @ w.float().t() + b.float()
    base = ops.gemm(a, w, b)
    try:
        for bn, split, pair in [(64, 1, 0), (128, 1, 1), (160, 2, 0), (128, 4, 1), (64, 8, 0)]:
            assert L.tg_gemm_force_tile(bn, split, pair) == 0
            out = ops.gemm(a, w, b)
            assert (out.float() - ref).abs().max().item() < 2e-2, (bn, split, pair)
            assert (out.float() - base.float()).abs().max().item() < 8e-3
        L.tg_gemm_force_tile(256, 1, 0)  # 640 % 256 != 0
        with pytest.raises(RuntimeError):
            ops.gemm(a, w, b)
    finally:
        L.tg_gemm_force_tile(0, 1, 0)
    assert L.tg_gemm_set_tuned(M, N, K, 0, 0, 192, 1, 0) != 0      # invalid entry refused
    assert L.tg_gemm_set_tuned(M, N, K, 0, 0, 64, 2, 1) == 0
    tuned = ops.gemm(a, w, b)
    L.tg_gemm_force_tile(64, 2, 1)
    forced = ops.gemm(a, w, b)
    L.tg_gemm_force_tile(0, 1, 0)
    assert torch.equal(tuned, forced)                                # the table entry selected exactly that tile
    assert L.tg_gemm_clear_tuned() == 0
    assert torch.equal(ops.gemm(a, w, b), base)


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(2, 64, 64, 320, 640), (3, 32, 32, 640, 1280), (1, 16, 16, 1280, 640),
                                            (1, 64, 64, 128, 128)])
def test_conv3x3_patch_reuse_tiles(B, H, W, Cin, Cout):
    """The KY3 stages of the convolution (one (box_h + 2)-row input patch serves the three ky taps; pair tiles of 64 / 128 /
    160 columns, image rows <= 64 pixels) against torch, with and without split-K, beside the tap-per-k-block ring that
    256-column tiles and single-CTA tiles keep.  Borders exercise the TMA zero fill above / below / beside the image."""
    from theatergen_b200 import _lib, ops
    L = _lib.lib()
    x = _rand((B, H, W, Cin), 1)
    w = _rand((Cout, 3, 3, Cin), 2, (9 * Cin) ** -0.5)
    bias = _rand((Cout,), 3)
    res = _rand((B, H, W, Cout), 5)
    ref = (torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), w.float().permute(0, 3, 1, 2), bias.float(),
                                      padding=1).permute(0, 2, 3, 1) + res.float())
    outs = {}
    try:
        for bn, split, pair in [(64, 1, 1), (128, 1, 1), (160, 1, 1), (128, 3, 1), (64, 2, 1), (256, 1, 1), (128, 1, 0)]:
            if Cout % bn:
                continue
            assert L.tg_gemm_force_tile(bn, split, pair) == 0
            outs[(bn, split, pair)] = ops.conv3x3(x, w, bias, residual=res)
    finally:
        L.tg_gemm_force_tile(0, 1, 0)
    for k, o in outs.items():
        _close(o, ref, 4e-3, 4e-3)
    # same tile, same k order per accumulator -> the patch path is bit-identical to itself across launches
    L.tg_gemm_force_tile(128, 1, 1)
    try:
        assert torch.equal(ops.conv3x3(x, w, bias, residual=res), outs[(128, 1, 1)])
    finally:
        L.tg_gemm_force_tile(0, 1, 0)


@pytest.mark.parametrize("B,H,Cin,Cout,pad_lo", [(2, 64, 320, 320, 1), (10, 32, 640, 640, 1), (2, 16, 1280, 1280, 1),
                                                 (1, 512, 64, 64, 1), (1, 256, 128, 128, 0), (3, 32, 64, 192, 0)])
def test_conv3x3_stride2_implicit_gemm(B, H, Cin, Cout, pad_lo):
    """Downsample2D (ip_adapter/unet_2d_blocks.py:1035,1112,1161,1198): stride-2 3x3 convolution whose operand is gathered by
    the TMA with an element stride of 2 -- against F.conv2d and against the im2col + GEMM path of round 1 (bit exact: same
    products, same accumulation order per output)."""
    from theatergen_b200 import ops
    x = _rand((B, H, H, Cin), 1)
    w = _rand((Cout, Cin, 3, 3), 2, (9 * Cin) ** -0.5)
    bias = _rand((Cout,), 3, 0.1)
    wp = w.permute(0, 2, 3, 1).contiguous()
    out = ops.conv3x3_s2(x, wp, bias, pad_lo=pad_lo)
    xin = x.float().permute(0, 3, 1, 2)
    if pad_lo == 0:
        xin = F.pad(xin, (0, 1, 0, 1))
    ref = F.conv2d(xin, w.float(), bias.float(), stride=2, padding=1 if pad_lo else 0).permute(0, 2, 3, 1)
    assert out.shape == ref.shape
    _close(out, ref, 3e-3, 3e-3)
    old = ops.gemm(ops.im2col_s2(x, pad_lo=pad_lo), wp.reshape(Cout, 9 * Cin), bias).view(B, H // 2, H // 2, Cout)
    assert (out.float() - old.float()).abs().max().item() < 2e-3
    y = ops.conv3x3_s2(x, wp, bias, pad_lo=pad_lo, act=ops.TG_ACT_SILU)
    _close(y, F.silu(ref), 3e-3, 3e-3)


@pytest.mark.parametrize("B,H,Cin,Cout", [(2, 8, 1280, 1280), (10, 16, 1280, 1280), (2, 32, 640, 640), (1, 64, 512, 512),
                                          (1, 128, 256, 256), (1, 256, 64, 128)])
def test_upsample2x_conv3x3_four_phases(B, H, Cin, Cout):
    """Upsample2D (ip_adapter/unet_2d_blocks.py:2184,2251,2298,2332): nearest x2 + 3x3 convolution computed as four 2x2
    convolutions on the low-resolution grid (weights.pack_upsample_conv3x3), each phase stored through a 3D tensor map into
    its pixels of the [B,2H,2W,C] image -- against F.interpolate + F.conv2d in fp32 and against the materialising path."""
    from theatergen_b200 import ops
    from theatergen_b200.weights import pack_upsample_conv3x3
    x = _rand((B, H, H, Cin), 1)
    w = _rand((Cout, Cin, 3, 3), 2, (9 * Cin) ** -0.5)
    bias = _rand((Cout,), 3, 0.1)
    w4 = pack_upsample_conv3x3(w)
    assert w4.shape == (4, Cout, 2, 2, Cin)
    out = ops.upsample2x_conv3x3(x, w4, bias)
    up = F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2, mode="nearest")
    ref = F.conv2d(up, w.float(), bias.float(), padding=1).permute(0, 2, 3, 1)
    assert out.shape == ref.shape
    _close(out, ref, 3e-3, 3e-3)
    old = ops.conv3x3(ops.upsample2x(x), w.permute(0, 2, 3, 1).contiguous(), bias)
    assert (out.float() - old.float()).abs().max().item() < 4e-3


def _totals(out, images):
    """Reference GroupNorm totals of a stored fp16 tensor [M, N]: [images, N/2, 2] (sum, sum of squares) in float64."""
    v = out.double().view(images, -1, out.shape[-1] // 2, 2)
    return torch.stack([v.sum(dim=(1, 3)), (v * v).sum(dim=(1, 3))], dim=-1)


def _check_totals(stats, out2d, images):
    got = stats.double() / 2.0 ** 24
    ref = _totals(out2d, images)
    err = (got - ref).abs()
    tol = 1e-5 * ref.abs() + 1e-3  # fp32 partial sums over 64 values, 2^-24 fixed point
    assert bool((err <= tol).all()), (err.max().item(), ref.abs().max().item())


def test_groupnorm_totals_from_the_gemm_epilogue():
    """tg_gemm_set_stats: the epilogue of a linear / convolution / stride-2 convolution / upsample convolution (plain and
    split-K tiles, CTA pairs and single CTAs, residual and row-bias epilogues) accumulates per (image, channel pair) the sum
    and sum of squares of the fp16 values it stores; integer atomics -> two runs give identical bits."""
    from theatergen_b200 import ops
    from theatergen_b200.weights import pack_upsample_conv3x3
    # linear, 4 images x 1024 rows (pair tiles) and 2 images x 64 rows (one M tile spans both images; split-K: K = 5120)
    for (imgs, hw, N, K, seed) in [(4, 1024, 320, 320, 1), (2, 64, 1280, 5120, 2), (10, 256, 640, 640, 3)]:
        M = imgs * hw
        a, w = _rand((M, K), seed), _rand((N, K), seed + 10, K ** -0.5)
        res, bias = _rand((M, N), seed + 20), _rand((N,), seed + 30)
        st = torch.zeros((imgs, N // 2, 2), dtype=torch.int64, device="cuda")
        out = ops.gemm(a, w, bias, residual=res, stats=st)
        assert torch.equal(out, ops.gemm(a, w, bias, residual=res)), "the statistics epilogue must not change the output"
        _check_totals(st, out, imgs)
        st2 = torch.zeros_like(st)
        ops.gemm(a, w, bias, residual=res, stats=st2)
        assert torch.equal(st, st2), "totals must be bit-reproducible"
        st3 = torch.zeros_like(st)  # one-shot: the next launch is not armed
        ops.gemm(a, w, bias, residual=res)
        assert int(st3.abs().sum()) == 0
    # convolutions
    B, H, Cin, Cout = 2, 32, 128, 192
    x = _rand((B, H, H, Cin), 4)
    w = _rand((Cout, 3, 3, Cin), 5, (9 * Cin) ** -0.5)
    rb = _rand((B, Cout), 6)
    st = torch.zeros((B, Cout // 2, 2), dtype=torch.int64, device="cuda")
    out = ops.conv3x3(x, w, None, rowbias=rb, stats=st)
    _check_totals(st, out.view(-1, Cout), B)
    st = torch.zeros((B, Cout // 2, 2), dtype=torch.int64, device="cuda")
    out = ops.conv3x3_s2(x, w, None, stats=st)
    _check_totals(st, out.view(-1, Cout), B)
    st = torch.zeros((B, Cout // 2, 2), dtype=torch.int64, device="cuda")
    w4 = pack_upsample_conv3x3(_rand((Cout, Cin, 3, 3), 7, (9 * Cin) ** -0.5))
    out = ops.upsample2x_conv3x3(x, w4, None, stats=st)
    _check_totals(st, out.view(-1, Cout), B)
    # low resolution, long K (split-K last-arriver epilogue), 8x8 images: a 128-row tile holds two images
    B, H, Cin, Cout = 4, 8, 1280, 1280
    x = _rand((B, H, H, Cin), 8)
    w = _rand((Cout, 3, 3, Cin), 9, (9 * Cin) ** -0.5)
    st = torch.zeros((B, Cout // 2, 2), dtype=torch.int64, device="cuda")
    out = ops.conv3x3(x, w, None, stats=st)
    _check_totals(st, out.view(-1, Cout), B)


@pytest.mark.parametrize("B,HW,C1,C2,silu", [(10, 4096, 320, 0, True), (4, 1024, 640, 320, True), (2, 64, 1280, 1280, True),
                                             (3, 256, 1280, 640, False), (4, 4096, 320, 320, True)])
def test_groupnorm_apply_from_totals(B, HW, C1, C2, silu):
    """tg_groupnorm_apply (one pass, statistics from the producers' epilogues) == tg_groupnorm (self-measuring kernels) on
    the same tensors, for single and concatenated inputs (group boundaries straddling the two tensors: C = 960, 1920), and
    with the rows [b0, b1) of both inputs re-measured (ControlNet residual add)."""
    from theatergen_b200 import ops

    def produce(Bn, C, seed):
        a, w = _rand((Bn * HW, 64), seed, 2.0), _rand((C, 64), seed + 1, 0.3)
        st = torch.zeros((Bn, C // 2, 2), dtype=torch.int64, device="cuda")
        return ops.gemm(a, w, _rand((C,), seed + 2), stats=st).view(Bn, HW, C), st

    x1, s1 = produce(B, C1, 1)
    x2, s2 = produce(B, C2, 5) if C2 else (None, None)
    g, b = _rand((C1 + C2,), 9), _rand((C1 + C2,), 10)
    ref = ops.groupnorm(x1, g, b, 1e-5, x2=x2, silu=silu)
    got = ops.groupnorm(x1, g, b, 1e-5, x2=x2, silu=silu, st1=ops.GnStats(s1), st2=None if x2 is None else ops.GnStats(s2))
    assert (got.float() - ref.float()).abs().max().item() <= 4e-3
    xf = x1.float() if x2 is None else torch.cat([x1.float(), x2.float()], -1)
    r32 = F.group_norm(xf.permute(0, 2, 1), 32, g.float(), b.float(), 1e-5).permute(0, 2, 1)
    r32 = F.silu(r32) if silu else r32
    assert (got.float() - r32).abs().max().item() <= 6e-3
    if B >= 3:  # rows [1, 3) rewritten in place by a residual GEMM that re-measures them
        def rewrite(x, s, C, seed):
            a, w = _rand((2 * HW, 64), seed), _rand((C, 64), seed + 1, 0.5)
            dst = x[1:3].view(-1, C)
            alt = torch.zeros((2, C // 2, 2), dtype=torch.int64, device="cuda")
            ops.gemm(a, w, residual=dst, out=dst, stats=alt)
            return ops.GnStats(s, alt, 1, 3)
        t1 = rewrite(x1, s1, C1, 20)
        t2 = rewrite(x2, s2, C2, 30) if C2 else None
        ref = ops.groupnorm(x1, g, b, 1e-5, x2=x2, silu=silu)
        got = ops.groupnorm(x1, g, b, 1e-5, x2=x2, silu=silu, st1=t1, st2=t2)
        assert (got.float() - ref.float()).abs().max().item() <= 4e-3
