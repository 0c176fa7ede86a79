"""Weight repacking from diffusers / IP-Adapter state-dict layouts into the kernel layouts."""
from __future__ import annotations

import torch


def pack_geglu(w: torch.Tensor, b: torch.Tensor):
    """GEGLU proj weight [8C, C] (rows: 4C value | 4C gate, ip_adapter/attention.py:344-367) ->
    rows interleaved per 64: [32 value | 32 gate] so one GEMM tile holds matching value/gate columns."""
    n2 = w.shape[0]
    n = n2 // 2
    assert n % 32 == 0
    idx = torch.arange(n, device=w.device).view(-1, 32)
    perm = torch.cat([idx, idx + n], dim=1).reshape(-1)
    return w[perm].contiguous(), b[perm].contiguous()


def pack_conv3x3(w: torch.Tensor):
    """[Cout, Cin, 3, 3] (torch) -> [Cout, 3, 3, Cin] (K ordered tap-major, channel-minor)."""
    return w.permute(0, 2, 3, 1).contiguous()


def pack_conv3x3_dgrad(w: torch.Tensor):
    """[Cout, Cin, 3, 3] (torch) -> the packed weight of the INPUT-gradient convolution, [Cin, 3, 3, Cout] with the taps
    rotated by 180 degrees: dx = conv3x3(dy, this) for y = conv3x3(x, w) (stride 1, pad 1)."""
    return w.flip(2, 3).permute(1, 2, 3, 0).contiguous()
