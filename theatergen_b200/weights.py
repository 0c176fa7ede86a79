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


def pack_upsample_conv3x3(w: torch.Tensor):
    """[Cout, Cin, 3, 3] (torch) of the convolution behind a nearest-x2 upsample (diffusers Upsample2D) -> [4, Cout, 2, 2, Cin]:
    the four 2x2 convolutions on the LOW-resolution image that produce the output pixels of parity (py, px), phase = 2 py +
    px.  Output pixel (2i + py, 2j + px) reads upsampled rows 2i + py + ky - 1, i.e. input rows i - 1 (ky = 0) and i
    (ky = 1, 2) for py = 0, and rows i (ky = 0, 1) and i + 1 (ky = 2) for py = 1; taps that land on the same input pixel are
    summed (in fp32, rounded to fp16 once).  csrc/gemm.cu::upsample2x_conv3x3_nhwc."""
    groups = {0: ((0,), (1, 2)), 1: ((0, 1), (2,))}   # parity -> 3x3 taps summed into window position a = 0, 1
    wf = w.float()
    out = torch.zeros((4, w.shape[0], 2, 2, w.shape[1]), dtype=torch.float32, device=w.device)
    for py in range(2):
        for px in range(2):
            for a in range(2):
                for b in range(2):
                    acc = 0
                    for ky in groups[py][a]:
                        for kx in groups[px][b]:
                            acc = acc + wf[:, :, ky, kx]
                    out[py * 2 + px, :, a, b, :] = acc
    return out.to(w.dtype).contiguous()
