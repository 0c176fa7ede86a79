"""Operator-level Python wrappers over the C-ABI (raw device pointers + the current CUDA stream).

Every function enqueues hand-written sm_100a kernels from libtheatergen.so; tensors are fp16, contiguous
along the last dimension, channels-last for images.  No torch compute ops are used on this path.
"""
from __future__ import annotations

import torch

from . import _lib

TG_ACT_NONE, TG_ACT_SILU, TG_ACT_GEGLU = 0, 1, 2


def _p(t):
    return 0 if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _chk16(*ts):
    for t in ts:
        if t is not None:
            assert t.is_cuda and t.dtype == torch.float16 and t.stride(-1) == 1, (t.dtype, t.stride(), t.device)


def gemm(a, w, bias=None, *, a2=None, rowbias=None, rows_per_image=1, residual=None, act=TG_ACT_NONE,
         out_scale=1.0, out=None):
    """out[M,N] = act([a|a2] @ w.T + bias + rowbias[row // rows_per_image]) + residual ; w is [N,K]."""
    _chk16(a, w, bias, a2, rowbias, residual, out)
    M, K1 = a.shape
    N, K = w.shape
    if a2 is not None:
        assert a2.shape[0] == M and K1 + a2.shape[1] == K
    else:
        assert K1 == K, (a.shape, w.shape)
    n_out = N // 2 if act == TG_ACT_GEGLU else N
    if out is None:
        out = torch.empty((M, n_out), dtype=torch.float16, device=a.device)
    rc = _lib.lib().tg_gemm(_p(a), a.stride(0), K1, _p(a2), 0 if a2 is None else a2.stride(0), _p(w), M, N, K,
                            _p(bias), _p(rowbias), rows_per_image, 0 if rowbias is None else rowbias.stride(0),
                            _p(residual), 0 if residual is None else residual.stride(0), act, float(out_scale),
                            _p(out), out.stride(0), _stream())
    _lib.check(rc, "tg_gemm")
    return out


def conv3x3(x, w, bias=None, *, rowbias=None, residual=None, out_scale=1.0, out=None):
    """x [B,H,W,Cin] NHWC, w [Cout,3,3,Cin] -> [B,H,W,Cout] (pad 1, stride 1)."""
    _chk16(x, w, bias, rowbias, residual, out)
    B, H, W, Cin = x.shape
    Cout = w.shape[0]
    assert x.is_contiguous() and w.is_contiguous() and w.shape[1:] == (3, 3, Cin)
    if out is None:
        out = torch.empty((B, H, W, Cout), dtype=torch.float16, device=x.device)
    res2 = None if residual is None else residual.reshape(B * H * W, -1)
    rc = _lib.lib().tg_conv3x3(_p(x), B, H, W, Cin, _p(w), Cout, _p(bias), _p(rowbias),
                               0 if rowbias is None else rowbias.stride(0), _p(res2),
                               0 if res2 is None else res2.stride(0), float(out_scale), _p(out), Cout, _stream())
    _lib.check(rc, "tg_conv3x3")
    return out
