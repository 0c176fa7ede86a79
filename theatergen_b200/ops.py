"""Operator-level Python wrappers over the C-ABI (raw device pointers + the current CUDA stream).

Every function enqueues hand-written sm_100a kernels from libtheatergen.so; tensors are fp16, contiguous
along the last dimension, channels-last for images.  No torch compute ops are used on this path.
"""
from __future__ import annotations

import torch

from . import _lib

TG_ACT_NONE, TG_ACT_SILU, TG_ACT_GEGLU, TG_ACT_GELU, TG_ACT_QUICK_GELU = 0, 1, 2, 3, 4

# ---------------------------------------------------------------------------------------------------------
# optional per-launch timing (bench.py roofline leg): CUDA events on the launching stream around each op
# ---------------------------------------------------------------------------------------------------------
_PROF = None


class profile:
    """with ops.profile() as rec: ...  -> rec.rows = [(kernel, ms, flops, bytes)] after rec.finish()."""

    def __enter__(self):
        global _PROF
        self.events = []
        _PROF = self
        return self

    def __exit__(self, *a):
        global _PROF
        _PROF = None

    def finish(self):
        torch.cuda.synchronize()
        self.rows = [(k, s.elapsed_time(e), f, b) for k, s, e, f, b, _ in self.events]
        self.tagged = [(k, t, s.elapsed_time(e), f, b) for k, s, e, f, b, t in self.events]
        return self.rows

    def summary(self):
        agg = {}
        for k, ms, f, b in self.rows:
            a = agg.setdefault(k, [0, 0.0, 0.0, 0.0])
            a[0] += 1; a[1] += ms; a[2] += f; a[3] += b
        return {k: dict(launches=v[0], ms=v[1], tflops=(v[2] / (v[1] * 1e-3) / 1e12) if v[1] > 0 else 0.0,
                        gbs=(v[3] / (v[1] * 1e-3) / 1e9) if v[1] > 0 else 0.0, flops=v[2], bytes=v[3])
                for k, v in agg.items()}


def _timed(kernel, flops, nbytes, fn, tag=""):
    if _PROF is None:
        return fn()
    st = torch.cuda.current_stream()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(st)
    r = fn()
    e.record(st)
    _PROF.events.append((kernel, s, e, float(flops), float(nbytes), tag))
    return r


def _p(t):
    return 0 if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _chk16(*ts):
    for t in ts:
        if t is not None:
            assert t.is_cuda and t.dtype == torch.float16 and t.stride(-1) == 1, (t.dtype, t.stride(), t.device)


class GnStats:
    """GroupNorm totals of a tensor, accumulated by the epilogue of the GEMM that produced it (tg_gemm_set_stats):
    `main` int64 [images, C/2, 2]; images [b0, b1) read `alt` [b1-b0, C/2, 2] instead (rows re-measured by the in-place
    ControlNet residual GEMM)."""
    __slots__ = ("main", "alt", "b0", "b1")

    def __init__(self, main, alt=None, b0=0, b1=0):
        self.main, self.alt, self.b0, self.b1 = main, alt, b0, b1


def _arm_stats(stats, M):
    """Arms the next tg_gemm / tg_conv3x3* launch of this thread to add its output's GroupNorm totals to `stats`
    (zeroed int64 [images, N/2, 2])."""
    if stats is None:
        return
    assert stats.is_cuda and stats.dtype == torch.int64 and stats.is_contiguous() and stats.dim() == 3 and stats.shape[2] == 2
    assert M % stats.shape[0] == 0, (M, stats.shape)
    _lib.check(_lib.lib().tg_gemm_set_stats(stats.data_ptr(), M // stats.shape[0]), "tg_gemm_set_stats")


def gemm(a, w, bias=None, *, a2=None, rowbias=None, rows_per_image=1, residual=None, act=TG_ACT_NONE,
         out_scale=1.0, out=None, stats=None):
    """out[M,N] = act([a|a2] @ w.T + bias + rowbias[row // rows_per_image]) + residual ; w is [N,K].
    stats: see _arm_stats."""
    _chk16(a, w, bias, a2, rowbias, residual, out)
    assert w.is_contiguous(), "gemm: the weight operand must be a dense [N,K] matrix"
    M, K1 = a.shape
    N, K = w.shape
    if a2 is not None:
        assert a2.shape[0] == M and K1 + a2.shape[1] == K
    else:
        assert K1 == K, (a.shape, w.shape)
    n_out = N // 2 if act == TG_ACT_GEGLU else N
    if out is None:
        out = torch.empty((M, n_out), dtype=torch.float16, device=a.device)
    _arm_stats(stats, M)
    rc = _timed("gemm_tc_kernel", 2.0 * M * N * K, 2.0 * (M * K + N * K + M * n_out + (M * N if residual is not None else 0)),
                lambda: _lib.lib().tg_gemm(_p(a), a.stride(0), K1, _p(a2), 0 if a2 is None else a2.stride(0), _p(w), M, N, K,
                            _p(bias), _p(rowbias), rows_per_image, 0 if rowbias is None else rowbias.stride(0),
                            _p(residual), 0 if residual is None else residual.stride(0), act, float(out_scale),
                            _p(out), out.stride(0), _stream()), tag=f"lin M={M} N={N} K={K} act={act}")
    _lib.check(rc, "tg_gemm")
    return out


CROSS_SLOTS = 96  # key slots per head of the folded cross-attention: text keys at 0.., image tokens at 80.., rest zero


def cross_scores(x, kf, n_text, n_ip, ip_scale=None, out=None):
    """Folded cross-attention, first half (include/theatergen.h tg_cross_scores): x [B*HW, C], kf [B, heads*96, C] ->
    probabilities [B*HW, heads*96] (text softmax | ip_scale[b] * image-token softmax | zeros), softmax in the GEMM epilogue."""
    _chk16(x, kf, out)
    B, N, C = kf.shape
    M = x.shape[0]
    assert kf.is_contiguous() and x.shape[1] == C and N % (2 * CROSS_SLOTS) == 0 and M % B == 0
    heads, HW = N // CROSS_SLOTS, M // B
    if n_ip:
        assert ip_scale is not None and ip_scale.dtype == torch.float32 and ip_scale.numel() == B and ip_scale.is_cuda
    if out is None:
        out = torch.empty((M, N), dtype=torch.float16, device=x.device)
    rc = _timed("gemm_tc_kernel", 2.0 * M * N * C, 2.0 * (M * C + B * N * C + M * N),
                lambda: _lib.lib().tg_cross_scores(_p(x), x.stride(0), _p(kf), B, HW, C, heads, n_text, n_ip,
                                                   _p(ip_scale) if n_ip else 0, _p(out), out.stride(0), _stream()),
                tag=f"xscores M={M} N={N} K={C}")
    _lib.check(rc, "tg_cross_scores")
    return out


def cross_apply(p, vf, bias=None, residual=None, out=None):
    """Folded cross-attention, second half (tg_cross_apply): p [B*HW, heads*96], vf [B, C, heads*96] ->
    [B*HW, C] = p @ vf[b].T + bias + residual."""
    _chk16(p, vf, bias, residual, out)
    B, C, N = vf.shape
    M = p.shape[0]
    assert vf.is_contiguous() and p.shape[1] == N and N % CROSS_SLOTS == 0 and M % B == 0
    if out is None:
        out = torch.empty((M, C), dtype=torch.float16, device=p.device)
    rc = _timed("gemm_tc_kernel", 2.0 * M * N * C, 2.0 * (M * N + B * N * C + M * C + (M * C if residual is not None else 0)),
                lambda: _lib.lib().tg_cross_apply(_p(p), p.stride(0), _p(vf), B, M // B, C, N // CROSS_SLOTS, _p(bias),
                                                  _p(residual), 0 if residual is None else residual.stride(0), _p(out),
                                                  out.stride(0), _stream()),
                tag=f"xapply M={M} N={C} K={N}")
    _lib.check(rc, "tg_cross_apply")
    return out


def conv3x3(x, w, bias=None, *, rowbias=None, residual=None, act=TG_ACT_NONE, out_scale=1.0, out=None, stats=None):
    """x [B,H,W,Cin] NHWC, w [Cout,3,3,Cin] -> [B,H,W,Cout] (pad 1, stride 1)."""
    _chk16(x, w, bias, rowbias, residual, out)
    B, H, W, Cin = x.shape
    Cout = w.shape[0]
    assert x.is_contiguous() and w.is_contiguous() and w.shape[1:] == (3, 3, Cin)
    if out is None:
        out = torch.empty((B, H, W, Cout), dtype=torch.float16, device=x.device)
    res2 = None if residual is None else residual.reshape(B * H * W, -1)
    Mc = B * H * W
    _arm_stats(stats, Mc)
    rc = _timed("gemm_tc_kernel", 2.0 * Mc * Cout * 9 * Cin,
                2.0 * (Mc * Cin + Cout * 9 * Cin + Mc * Cout + (Mc * Cout if res2 is not None else 0)),
                lambda: _lib.lib().tg_conv3x3(_p(x), B, H, W, Cin, _p(w), Cout, _p(bias), _p(rowbias),
                               0 if rowbias is None else rowbias.stride(0), _p(res2),
                               0 if res2 is None else res2.stride(0), act, float(out_scale), _p(out), Cout, _stream()),
                tag=f"conv M={Mc} N={Cout} K={9 * Cin} B={B} H={H} W={W}")
    _lib.check(rc, "tg_conv3x3")
    return out


_GN_WS = {}


def _gn_workspace(device, B):
    key = (device, torch.cuda.current_stream().cuda_stream)
    ws = _GN_WS.get(key)
    if ws is None or ws.numel() < 16640 + B * 16384:
        # [256 arrival counters (start at zero, self-resetting)][256x32x2 mean/rstd][partials B x 256 x 32 x 2]
        ws = torch.zeros(16640 + max(B, 16) * 16384, dtype=torch.float32, device=device)
        _GN_WS[key] = ws
    return ws


GN_CHECK_LOG = None  # tests: a list here makes every statistics-fed groupnorm() also run the two-pass kernels and log the gap


def groupnorm(x1, gamma, beta, eps, *, x2=None, silu=False, out=None, workspace=None, st1=None, st2=None):
    """GroupNorm(32)(+SiLU) over channels-last [B,HW,C1] (| x2 [B,HW,C2]) -> [B,HW,C1+C2].
    st1 / st2 (GnStats): totals the producers of x1 / x2 accumulated in their epilogues -> one pass (tg_groupnorm_apply);
    without them the statistics are computed here (tg_groupnorm)."""
    _chk16(x1, x2, gamma, beta, out)
    B, HW, C1 = x1.shape
    C2 = 0 if x2 is None else x2.shape[-1]
    assert x1.is_contiguous() and (x2 is None or x2.is_contiguous())
    if out is None:
        out = torch.empty((B, HW, C1 + C2), dtype=torch.float16, device=x1.device)
    if st1 is not None and (x2 is None or st2 is not None):
        assert st1.main.shape == (B, C1 // 2, 2) and (st2 is None or st2.main.shape == (B, C2 // 2, 2))
        alts = [s for s in (st1, st2) if s is not None and s.alt is not None]
        b0, b1 = (alts[0].b0, alts[0].b1) if alts else (0, 0)
        assert all((s.b0, s.b1) == (b0, b1) for s in alts)
        rc = _timed("groupnorm", 0, 2.0 * 2 * B * HW * (C1 + C2),
                    lambda: _lib.lib().tg_groupnorm_apply(_p(x1), C1, _p(st1.main), _p(st1.alt), _p(x2), C2,
                                                          _p(None if st2 is None else st2.main),
                                                          _p(None if st2 is None else st2.alt), b0, b1, B, HW, _p(gamma),
                                                          _p(beta), float(eps), int(silu), _p(out), _stream()),
                    tag=f"gn-apply B={B} HW={HW} C={C1 + C2}")
        _lib.check(rc, "tg_groupnorm_apply")
        if GN_CHECK_LOG is not None:
            ref = groupnorm(x1, gamma, beta, eps, x2=x2, silu=silu)
            GN_CHECK_LOG.append((B, HW, C1, C2, (out.float() - ref.float()).abs().max().item()))
        return out
    ws = workspace if workspace is not None else _gn_workspace(x1.device, B)
    rc = _timed("groupnorm", 0, 2.0 * 3 * B * HW * (C1 + C2),
                lambda: _lib.lib().tg_groupnorm(_p(x1), C1, _p(x2), C2, B, HW, _p(gamma), _p(beta), float(eps), int(silu),
                                                _p(out), _p(ws), _stream()), tag=f"gn B={B} HW={HW} C={C1 + C2}")
    _lib.check(rc, "tg_groupnorm")
    return out


def layernorm(x, gamma, beta, eps=1e-5, out=None):
    _chk16(x, gamma, beta, out)
    M, Cc = x.shape
    if out is None:
        out = torch.empty((M, Cc), dtype=torch.float16, device=x.device)
    rc = _timed("layernorm", 0, 4.0 * M * Cc,
                lambda: _lib.lib().tg_layernorm(_p(x), x.stride(0), _p(gamma), _p(beta), float(eps), _p(out),
                                                out.stride(0), M, Cc, _stream()))
    _lib.check(rc, "tg_layernorm")
    return out


def attention(q, k, v, B, heads, scale=None, out=None, causal=False):
    """q [B*Nq, heads*d] (row pitch may exceed), k/v [B*Nk, heads*d] views -> out [B*Nq, heads*d]."""
    _chk16(q, k, v, out)
    d = q.shape[1] // heads
    Nq, Nk = q.shape[0] // B, k.shape[0] // B
    if scale is None:
        scale = d ** -0.5
    if out is None:
        out = torch.empty((B * Nq, heads * d), dtype=torch.float16, device=q.device)
    rc = _timed("attn_self_kernel", 4.0 * B * heads * Nq * Nk * d, 2.0 * B * heads * d * (2 * Nq + 2 * Nk),
                lambda: _lib.lib().tg_attention(_p(q), _p(k), _p(v), q.stride(0), k.stride(0), v.stride(0), _p(out),
                                                out.stride(0), B, heads, Nq, Nk, d, float(scale), int(causal), _stream()))
    _lib.check(rc, "tg_attention")
    return out


def attention_tc(q, k, v, B, heads, d, scale=None, out=None, packed=False):
    """tcgen05 flash attention: q/k/v are [B*N, heads*P] views, every head padded with zero columns to P = 64 (d <= 56)
    or P = 128 (d <= 120); for 128 <= d <= 176 (d % 16 == 0) -- and for every d with packed=True -- the heads are packed:
    [B*N, heads*d] views."""
    _chk16(q, k, v, out)
    N = q.shape[0] // B
    if scale is None:
        scale = d ** -0.5
    if out is None:
        out = torch.empty((B * N, heads * d), dtype=torch.float16, device=q.device)
    rc = _timed("attn_tc_kernel", 4.0 * B * heads * N * N * d, 2.0 * B * heads * d * 4 * N,
                lambda: (_lib.lib().tg_attention_tc_packed if packed else _lib.lib().tg_attention_tc)(
                    _p(q), _p(k), _p(v), q.stride(0), k.stride(0), v.stride(0), _p(out), out.stride(0), B, heads, N, d,
                    float(scale), _stream()))
    _lib.check(rc, "tg_attention_tc")
    return out


def attention_cross_ip(q, kt, vt, ki, vi, B, heads, n_text, n_ip, ip_scale=None, add_mask=None, probs=None,
                       probs_b0=0, out=None):
    """q [B*Nq, C]; kt/vt/ki/vi column views [B*(n_text+n_ip), C] of the K/V projection buffer."""
    _chk16(q, kt, vt, ki, vi, add_mask, probs, out)
    Cc = q.shape[1]
    d = Cc // heads
    Nq = q.shape[0] // B
    assert kt.stride(0) == vt.stride(0) and (ki is None or ki.stride(0) == kt.stride(0) == vi.stride(0))
    if ip_scale is not None:
        assert ip_scale.dtype == torch.float32 and ip_scale.numel() == B and ip_scale.is_cuda
    if out is None:
        out = torch.empty((B * Nq, Cc), dtype=torch.float16, device=q.device)
    rc = _timed("attn_cross_ip_kernel", 4.0 * B * heads * Nq * (n_text + n_ip) * d, 2.0 * 2 * B * Nq * Cc,
                lambda: _lib.lib().tg_attention_cross_ip(_p(q), q.stride(0), _p(kt), _p(vt), _p(ki), _p(vi), kt.stride(0),
                                                         n_text, n_ip, _p(ip_scale), _p(add_mask), _p(out),
                                                         out.stride(0), _p(probs), probs_b0, B, heads, Nq, d,
                                                         float(d ** -0.5), _stream()))
    _lib.check(rc, "tg_attention_cross_ip")
    return out


def conv3x3_s2(x, w, bias=None, *, pad_lo=1, act=TG_ACT_NONE, out=None, stats=None):
    """x [B,H,W,Cin] NHWC, w [Cout,3,3,Cin] -> [B,H/2,W/2,Cout]: 3x3 convolution with stride 2 (diffusers Downsample2D) as an
    implicit GEMM whose TMA gathers every other pixel -- no im2col buffer.  pad_lo=1: padding 1; pad_lo=0: the VAE encoder's
    (0, 1) padding."""
    _chk16(x, w, bias, out)
    B, H, W, Cin = x.shape
    Cout = w.shape[0]
    assert x.is_contiguous() and w.is_contiguous() and tuple(w.shape[1:]) == (3, 3, Cin)
    if out is None:
        out = torch.empty((B, H // 2, W // 2, Cout), dtype=torch.float16, device=x.device)
    Mc = B * (H // 2) * (W // 2)
    _arm_stats(stats, Mc)
    rc = _timed("gemm_tc_kernel", 2.0 * Mc * Cout * 9 * Cin, 2.0 * (B * H * W * Cin + Cout * 9 * Cin + Mc * Cout),
                lambda: _lib.lib().tg_conv3x3_s2(_p(x), B, H, W, Cin, _p(w), Cout, _p(bias), int(pad_lo), act, _p(out), Cout,
                                                 _stream()),
                tag=f"conv_s2 M={Mc} N={Cout} K={9 * Cin} B={B} H={H // 2} W={W // 2}")
    _lib.check(rc, "tg_conv3x3_s2")
    return out


def upsample2x_conv3x3(x, w4, bias=None, out=None, stats=None):
    """x [B,H,W,Cin] NHWC, w4 = weights.pack_upsample_conv3x3(w) [4,Cout,2,2,Cin] -> [B,2H,2W,Cout] = conv3x3(nearest_x2(x)):
    four 2x2 convolutions on the low-resolution grid (2.25x fewer MACs, no upsampled tensor)."""
    _chk16(x, w4, bias, out)
    B, H, W, Cin = x.shape
    Cout = w4.shape[1]
    assert x.is_contiguous() and w4.is_contiguous() and tuple(w4.shape) == (4, Cout, 2, 2, Cin)
    if out is None:
        out = torch.empty((B, 2 * H, 2 * W, Cout), dtype=torch.float16, device=x.device)
    Mc = B * H * W
    _arm_stats(stats, Mc)  # each of the four phase GEMMs covers H*W rows of every image
    rc = _timed("gemm_tc_kernel", 2.0 * 4 * Mc * Cout * 4 * Cin, 2.0 * (Mc * Cin + 16 * Cout * Cin + 4 * Mc * Cout),
                lambda: _lib.lib().tg_upsample2x_conv3x3(_p(x), B, H, W, Cin, _p(w4), Cout, _p(bias), _p(out), _stream()),
                tag=f"upconv M={4 * Mc} N={Cout} K={4 * Cin} B={B} H={2 * H} W={2 * W}")
    _lib.check(rc, "tg_upsample2x_conv3x3")
    return out


def upsample2x(x, out=None):
    _chk16(x, out)
    B, H, W, Cc = x.shape
    if out is None:
        out = torch.empty((B, 2 * H, 2 * W, Cc), dtype=torch.float16, device=x.device)
    _lib.check(_lib.lib().tg_upsample2x(_p(x), B, H, W, Cc, _p(out), _stream()), "tg_upsample2x")
    return out


def im2col_patch(x_nchw, P, Kpad):
    """[B,Cin,H,W] -> [B*(H/P)*(W/P), Kpad] rows of flattened P x P patches (k = c*P*P + ky*P + kx, zero padded)."""
    _chk16(x_nchw)
    B, Cin, H, W = x_nchw.shape
    assert x_nchw.is_contiguous()
    out = torch.empty((B * (H // P) * (W // P), Kpad), dtype=torch.float16, device=x_nchw.device)
    _lib.check(_lib.lib().tg_im2col_patch(_p(x_nchw), _p(out), B, Cin, H, W, P, Kpad, _stream()), "tg_im2col_patch")
    return out


def embed_tokens(ids, table, pos):
    """ids int64 [B,T] on the device; table [vocab,C]; pos [T,C] -> [B*T, C] = table[ids] + pos."""
    _chk16(table, pos)
    assert ids.dtype == torch.int64 and ids.is_cuda and ids.is_contiguous() and table.is_contiguous() and pos.is_contiguous()
    B, T = ids.shape
    C = table.shape[1]
    assert pos.shape[0] >= T and pos.shape[1] == C
    out = torch.empty((B * T, C), dtype=torch.float16, device=table.device)
    _lib.check(_lib.lib().tg_embed_tokens(_p(ids), _p(table), _p(pos), _p(out), B, T, C, table.shape[0], _stream()),
               "tg_embed_tokens")
    return out


def gaussian_sample(moments, noise, out_scale=1.0, out=None):
    """moments NCHW [B,2L,H,W] (mean | logvar), noise [B,L,H,W] or None -> (mean + std * noise) * out_scale."""
    _chk16(moments, noise, out)
    B, L2, H, W = moments.shape
    L = L2 // 2
    assert moments.is_contiguous() and (noise is None or (noise.is_contiguous() and noise.shape == (B, L, H, W)))
    if out is None:
        out = torch.empty((B, L, H, W), dtype=torch.float16, device=moments.device)
    _lib.check(_lib.lib().tg_gaussian_sample(_p(moments), _p(noise), _p(out), B, L, H * W, float(out_scale), _stream()),
               "tg_gaussian_sample")
    return out


def im2col_s2(x, out=None, pad_lo=1):
    _chk16(x, out)
    B, H, W, Cc = x.shape
    if out is None:
        out = torch.empty((B * (H // 2) * (W // 2), 9 * Cc), dtype=torch.float16, device=x.device)
    _lib.check(_lib.lib().tg_im2col_s2(_p(x), B, H, W, Cc, _p(out), int(pad_lo), _stream()), "tg_im2col_s2")
    return out


def conv_in(x_nchw, w, bias, out=None, silu=False):
    """x NCHW [B,Cin,H,W] (Cin 3 or 4), w [Cout,Cin,3,3] (torch layout) -> NHWC [B,H,W,Cout]."""
    _chk16(x_nchw, w, bias, out)
    B, Cin, H, W = x_nchw.shape
    assert Cin in (3, 4) and x_nchw.is_contiguous() and w.is_contiguous() and w.shape[1] == Cin
    Cout = w.shape[0]
    if out is None:
        out = torch.empty((B, H, W, Cout), dtype=torch.float16, device=x_nchw.device)
    _lib.check(_lib.lib().tg_conv_in(_p(x_nchw), _p(w), _p(bias), _p(out), B, Cin, H, W, Cout, int(silu), _stream()),
               "tg_conv_in")
    return out


def pack_conv_in(w):
    """torch [Cout,Cin,3,3] -> GEMM weight [Cout,64] (k = ci*9 + tap, zero padded) for conv_in_gemm."""
    co, cin = w.shape[:2]
    wp = torch.zeros((co, 64), dtype=torch.float16, device=w.device)
    wp[:, :cin * 9] = w.reshape(co, cin * 9)
    return wp


def conv_in_gemm(x_nchw, w_packed, bias, silu=False, out=None, stats=None):
    """3x3 / pad 1 convolution of a small-channel NCHW image as im2col (K padded to 64) + tensor-core GEMM
    -> NHWC [B,H,W,Cout]."""
    _chk16(x_nchw, w_packed, bias, out)
    B, Cin, H, W = x_nchw.shape
    assert x_nchw.is_contiguous() and w_packed.shape[1] == 64 and Cin * 9 <= 64
    cols = torch.empty((B * H * W, 64), dtype=torch.float16, device=x_nchw.device)
    _lib.check(_lib.lib().tg_im2col_small(_p(x_nchw), _p(cols), B, Cin, H, W, _stream()), "tg_im2col_small")
    Cout = w_packed.shape[0]
    y = gemm(cols, w_packed, bias, act=TG_ACT_SILU if silu else TG_ACT_NONE,
             out=None if out is None else out.view(B * H * W, Cout), stats=stats)
    return y.view(B, H, W, Cout)


def conv_out(x_nhwc, w, bias, out=None):
    """w packed [Cout,3,3,C] with Cout 3 or 4; returns NCHW [B,Cout,H,W]."""
    _chk16(x_nhwc, w, bias, out)
    B, H, W, Cc = x_nhwc.shape
    cout = w.shape[0]
    if out is None:
        out = torch.empty((B, cout, H, W), dtype=torch.float16, device=x_nhwc.device)
    _lib.check(_lib.lib().tg_conv_out(_p(x_nhwc), _p(w), _p(bias), _p(out), B, H, W, Cc, cout, _stream()), "tg_conv_out")
    return out


def pack_conv_out_tc(w_packed, bias):
    """conv_out weights [Cout,3,3,C] (Cout 3 or 4) -> ([64,3,3,C], [64]) zero padded: a 64-column tile of conv3x3."""
    cout, C = w_packed.shape[0], w_packed.shape[-1]
    wp = torch.zeros((64, 3, 3, C), dtype=torch.float16, device=w_packed.device)
    wp[:cout] = w_packed
    bp = torch.zeros((64,), dtype=torch.float16, device=w_packed.device)
    bp[:cout] = bias
    return wp, bp


def conv_out_tc(x_nhwc, w64, b64, cout, out=None):
    """conv_out on the tensor-core convolution: x [B,H,W,C] -> NCHW [B,cout,H,W] (the padded columns are dropped)."""
    _chk16(x_nhwc, w64, b64, out)
    B, H, W, _ = x_nhwc.shape
    y = conv3x3(x_nhwc, w64, b64)
    if out is None:
        out = torch.empty((B, cout, H, W), dtype=torch.float16, device=x_nhwc.device)
    _lib.check(_lib.lib().tg_nhwc_head_to_nchw(_p(y), _p(out), B, 64, cout, H * W, _stream()), "tg_nhwc_head_to_nchw")
    return out


def pointwise_nchw(x, w, bias, out=None):
    """1x1 convolution between tiny channel counts: x NCHW [B,Cin,H,W], w [Cout,Cin] -> NCHW [B,Cout,H,W]."""
    _chk16(x, w, bias, out)
    B, Cin, H, W = x.shape
    cout = w.shape[0]
    assert x.is_contiguous() and w.is_contiguous() and w.shape[1] == Cin
    if out is None:
        out = torch.empty((B, cout, H, W), dtype=torch.float16, device=x.device)
    _lib.check(_lib.lib().tg_pointwise_nchw(_p(x), _p(w), _p(bias), _p(out), B, Cin, cout, H * W, _stream()), "tg_pointwise_nchw")
    return out


def softmax_rows(x, scale, out=None):
    """out[r,:] = softmax(scale * x[r,:]) (fp32 math), x [M,N] fp16 with N <= 8192."""
    _chk16(x, out)
    M, N = x.shape
    if out is None:
        out = torch.empty((M, N), dtype=torch.float16, device=x.device)
    _lib.check(_lib.lib().tg_softmax_rows(_p(x), x.stride(0), _p(out), out.stride(0), M, N, float(scale), _stream()),
               "tg_softmax_rows")
    return out


def timestep_embedding(t_dev, dim=320, out=None):
    assert t_dev.dtype == torch.float32 and t_dev.is_cuda
    B = t_dev.numel()
    if out is None:
        out = torch.empty((B, dim), dtype=torch.float16, device=t_dev.device)
    _lib.check(_lib.lib().tg_timestep_embedding(_p(t_dev), _p(out), B, dim, _stream()), "tg_timestep_embedding")
    return out


def add_inplace(y, x):
    _chk16(y, x)
    assert y.is_contiguous() and x.is_contiguous() and y.numel() == x.numel()
    _lib.check(_lib.lib().tg_add_inplace(_p(y), _p(x), y.numel(), _stream()), "tg_add_inplace")
    return y


def nchw_to_nhwc(x, out=None):
    _chk16(x, out)
    B, Cc, H, W = x.shape
    if out is None:
        out = torch.empty((B, H, W, Cc), dtype=torch.float16, device=x.device)
    _lib.check(_lib.lib().tg_nchw_to_nhwc(_p(x), _p(out), B, Cc, H * W, _stream()), "tg_nchw_to_nhwc")
    return out


def duplicate_latents(lat, out):
    _chk16(lat, out)
    _lib.check(_lib.lib().tg_duplicate_latents(_p(lat), _p(out), lat.numel(), _stream()), "tg_duplicate_latents")
    return out


def cfg_ddim_composite(eps, latents, coef_dev, frozen=None, mask=None, latents_copy=None, k0=0, count=None):
    """eps [2K,C,H,W], latents [K,C,H,W] updated in place; coef_dev float32[5] on device.
    k0 / count restrict the update to trajectories [k0, k0+count) of the co-batch (frozen / latents_copy are then
    [count,C,H,W])."""
    _chk16(eps, latents, frozen, mask, latents_copy)
    K, Cc, H, W = latents.shape
    assert eps.shape[0] == 2 * K and coef_dev.dtype == torch.float32 and coef_dev.numel() >= 5
    count = K - k0 if count is None else count
    assert eps.is_contiguous() and latents.is_contiguous() and 0 <= k0 and k0 + count <= K
    off = k0 * Cc * H * W * 2
    rc = _lib.lib().tg_cfg_ddim_composite(_p(eps) + off, _p(latents) + off, _p(coef_dev), _p(frozen), _p(mask), count,
                                          Cc, H * W, _p(latents_copy), K, _stream())
    _lib.check(rc, "tg_cfg_ddim_composite")
    return latents


# ---------------------------------------------------------------------------------------------------------
# input-gradient operators of the guidance step (backward.cu)
# ---------------------------------------------------------------------------------------------------------
def layernorm_bwd(x, gamma, dy, eps=1e-5):
    _chk16(x, gamma, dy)
    M, Cc = x.shape
    dx = torch.empty((M, Cc), dtype=torch.float16, device=x.device)
    _lib.check(_lib.lib().tg_layernorm_bwd(_p(x), x.stride(0), _p(gamma), _p(dy), dy.stride(0), _p(dx), Cc, M, Cc,
                                           float(eps), _stream()), "tg_layernorm_bwd")
    return dx


def groupnorm_bwd(x, gamma, beta, dy, eps, silu):
    """x, dy [B,HW,C] contiguous -> dx [B,HW,C]"""
    _chk16(x, gamma, beta, dy)
    B, HW, Cc = x.shape
    assert x.is_contiguous() and dy.is_contiguous() and dy.shape == x.shape
    dx = torch.empty_like(x)
    _lib.check(_lib.lib().tg_groupnorm_bwd(_p(x), _p(gamma), _p(beta), _p(dy), _p(dx), B, HW, Cc, float(eps), int(silu),
                                           _stream()), "tg_groupnorm_bwd")
    return dx


def softmax_bwd_rows(p, dp, scale, out=None):
    """dS = scale * P o (dP - rowsum(P o dP)); writes into `out` (default: in place over dp)."""
    _chk16(p, dp, out)
    M, N = p.shape
    out = dp if out is None else out
    _lib.check(_lib.lib().tg_softmax_bwd_rows(_p(p), p.stride(0), _p(dp), dp.stride(0), _p(out), out.stride(0), M, N,
                                              float(scale), _stream()), "tg_softmax_bwd_rows")
    return out


def geglu_bwd(pre, dy):
    _chk16(pre, dy)
    M, F2 = pre.shape
    assert pre.is_contiguous() and dy.is_contiguous() and dy.shape == (M, F2 // 2)
    dpre = torch.empty_like(pre)
    _lib.check(_lib.lib().tg_geglu_bwd(_p(pre), _p(dy), _p(dpre), M, F2 // 2, _stream()), "tg_geglu_bwd")
    return dpre


def attention_cross_ip_bwd_q(q, kt, vt, ki, vi, heads, n_text, n_ip, ip_scale, dout, dprobs=None):
    """One image: q/dout [Nq,C]; kt/vt/ki/vi column views [n_text+n_ip, C] of the context projection; dprobs
    [heads,Nq,n_text] or None -> dq [Nq,C]."""
    _chk16(q, kt, vt, ki, vi, dout, dprobs)
    Nq, Cc = q.shape
    d = Cc // heads
    assert kt.stride(0) == vt.stride(0) and (ki is None or ki.stride(0) == kt.stride(0) == vi.stride(0))
    assert dprobs is None or (dprobs.is_contiguous() and dprobs.shape == (heads, Nq, n_text))
    dq = torch.empty((Nq, Cc), dtype=torch.float16, device=q.device)
    _lib.check(_lib.lib().tg_attention_cross_ip_bwd_q(_p(q), q.stride(0), _p(kt), _p(vt), _p(ki), _p(vi), kt.stride(0),
                                                      n_text, n_ip if ki is not None else 0, float(ip_scale), _p(dout),
                                                      dout.stride(0), _p(dprobs), _p(dq), Cc, Nq, heads, d,
                                                      float(d ** -0.5), _stream()), "tg_attention_cross_ip_bwd_q")
    return dq


def upsample2x_bwd(dy):
    _chk16(dy)
    B, H2, W2, Cc = dy.shape
    assert dy.is_contiguous()
    dx = torch.empty((B, H2 // 2, W2 // 2, Cc), dtype=torch.float16, device=dy.device)
    _lib.check(_lib.lib().tg_upsample2x_bwd(_p(dy), B, H2 // 2, W2 // 2, Cc, _p(dx), _stream()), "tg_upsample2x_bwd")
    return dx


def col2im_s2(dcols, B, H, W, Cc):
    _chk16(dcols)
    assert dcols.is_contiguous() and dcols.shape == (B * (H // 2) * (W // 2), 9 * Cc)
    dx = torch.empty((B, H, W, Cc), dtype=torch.float16, device=dcols.device)
    _lib.check(_lib.lib().tg_col2im_s2(_p(dcols), B, H, W, Cc, _p(dx), _stream()), "tg_col2im_s2")
    return dx


def guidance_update(latents, grad, inv_scale, step, grad_out=None, bad=None):
    _chk16(latents, grad)
    assert latents.is_contiguous() and grad.is_contiguous() and latents.numel() == grad.numel()
    assert grad_out is None or (grad_out.dtype == torch.float32 and grad_out.numel() == grad.numel())
    assert bad is None or bad.dtype == torch.int32
    _lib.check(_lib.lib().tg_guidance_update(_p(latents), _p(grad), float(inv_scale), float(step), _p(grad_out), _p(bad),
                                             latents.numel(), _stream()), "tg_guidance_update")
    return latents
