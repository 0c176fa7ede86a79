#!/usr/bin/env python
"""Which role of the warp-specialised GEMM waits for which?  Runs a few SD-1.5 shapes with the tg_gemm_set_prof hook and
prints, per shape (mean over CTAs): main-loop cycles, k-blocks and tiles per CTA, cycles per k-block, and the share of
the main loop that the TMA producer spent waiting for a free stage, that the MMA issuer spent waiting for data /
for a drained accumulator, and that epilogue warp 0 spent waiting for a finished accumulator."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from theatergen_b200 import _lib, ops


def main():
    B = 10
    r = lambda *s: torch.randn(*s, device="cuda").half()
    cases = []
    for name, M, N, K, geglu in (("ff1 1280 (geglu)", 2560, 10240, 1280, True), ("ff2 640", 10240, 640, 2560, False),
                                 ("qkv 320", 40960, 960, 320, False), ("square 8192", 8192, 8192, 8192, False),
                                 ("ff1 320 (geglu)", 40960, 2560, 320, True)):
        a, w = r(M, K), r(N, K)
        bias = r(N)
        if geglu:
            cases.append((f"lin {name}", M, N, K, lambda a=a, w=w, bias=bias: ops.gemm(a, w, bias, act=ops.TG_ACT_GEGLU)))
        else:
            cases.append((f"lin {name}", M, N, K, lambda a=a, w=w: ops.gemm(a, w)))
    for name, M, N, K in (("C->C 320 +res", 40960, 320, 320), ("C->C 640 +res", 10240, 640, 640),
                          ("C->C 1280 +res", 2560, 1280, 1280), ("ff2 320 +res", 40960, 320, 1280)):
        a, w, bias, res = r(M, K), r(N, K), r(N), r(M, N)
        cases.append((f"lin {name}", M, N, K, lambda a=a, w=w, bias=bias, res=res: ops.gemm(a, w, bias, residual=res)))
    for H, cin, cout in ((64, 320, 320), (32, 1280, 640), (16, 1280, 1280)):
        x, w = r(B, H, H, cin), r(cout, 3, 3, cin)
        cases.append((f"conv {H}x{H} {cin}->{cout}", B * H * H, cout, 9 * cin, lambda x=x, w=w: ops.conv3x3(x, w)))
    buf = torch.zeros(148 * 8, dtype=torch.int64, device="cuda")
    print(f"pair={os.environ.get('TG_GEMM_PAIR', '1')}")
    print(f"{'shape':28s} {'us':>8s} {'loop clk':>9s} {'kb/cta':>7s} {'tiles':>6s} {'clk/kb':>7s} {'prod wait':>10s} {'mma wait data':>14s} {'mma wait acc':>13s} {'epi wait':>9s}")
    for name, M, N, K, fn in cases:
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(10):
            fn()
        e.record()
        torch.cuda.synchronize()
        us = s.elapsed_time(e) / 10 * 1e3
        buf.zero_()
        _lib.lib().tg_gemm_set_prof(buf.data_ptr())
        fn()
        torch.cuda.synchronize()
        _lib.lib().tg_gemm_set_prof(None)
        c = buf.view(148, 8).double()
        c = c[c[:, 0] > 0]  # CTAs that issued MMAs (leaders in pair mode)
        tot, kb, tiles = c[:, 0].mean().item(), c[:, 5].mean().item(), c[:, 6].mean().item()
        pw = buf.view(148, 8).double()[:, 1]
        pw = pw[pw > 0].mean().item() if (pw > 0).any() else 0.0
        ew = buf.view(148, 8).double()[:, 4]
        ew = ew[ew > 0].mean().item() if (ew > 0).any() else 0.0
        print(f"{name:28s} {us:8.1f} {tot:9.0f} {kb:7.1f} {tiles:6.1f} {tot / max(kb, 1):7.0f} {100 * pw / tot:9.1f}% "
              f"{100 * c[:, 2].mean().item() / tot:13.1f}% {100 * c[:, 3].mean().item() / tot:12.1f}% {100 * ew / tot:8.1f}%")


if __name__ == "__main__":
    main()
