#!/usr/bin/env python
"""GroupNorm from epilogue totals vs the self-measuring kernels, and what the totals cost the producing GEMM.
Pairs (producer, GroupNorm) of the SD-1.5 levels at batch B; a CUDA graph of 10 (producer -> norm) pairs, so the norm reads
what the producer just left in L2, like in the step."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from theatergen_b200 import ops


def timed(fn, n=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5):
        g.replay()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / (5 * n) * 1e3


B = int(sys.argv[1]) if len(sys.argv) > 1 else 10
for H, C, C2 in ((64, 320, 0), (64, 320, 320), (32, 640, 0), (32, 640, 640), (16, 1280, 0), (16, 1280, 1280), (8, 1280, 0)):
    HW = H * H
    x = torch.randn(B, H, H, C, device="cuda").half()
    w = (torch.randn(C, 3, 3, C, device="cuda") * (9 * C) ** -0.5).half()
    wl = (torch.randn(C, C, device="cuda") * C ** -0.5).half()
    skip = torch.randn(B, HW, C2, device="cuda").half() if C2 else None
    st_skip = torch.zeros(B, C2 // 2, 2, dtype=torch.int64, device="cuda") if C2 else None
    if C2:
        ops.gemm(torch.randn(B * HW, 64, device="cuda").half(), torch.randn(C2, 64, device="cuda").half(), out=skip.view(-1, C2),
                 stats=st_skip)
    g_, b_ = torch.randn(C + C2, device="cuda").half(), torch.randn(C + C2, device="cuda").half()
    st = torch.zeros(B, C // 2, 2, dtype=torch.int64, device="cuda")
    out = torch.empty(B, H, H, C, device="cuda").half()
    res = torch.randn(B * HW, C, device="cuda").half()
    rows = {}
    for name, prod in (("conv", lambda s: ops.conv3x3(x, w, out=out, stats=s)),
                       ("lin+res", lambda s: ops.gemm(x.view(-1, C), wl, residual=res, out=out.view(-1, C), stats=s))):
        t_p = timed(lambda: prod(None))
        t_ps = timed(lambda: (st.zero_(), prod(st)))
        t_old = timed(lambda: (prod(None), ops.groupnorm(out.view(B, HW, C), g_, b_, 1e-5, x2=skip, silu=True)))
        t_new = timed(lambda: (st.zero_(), prod(st), ops.groupnorm(out.view(B, HW, C), g_, b_, 1e-5, x2=skip, silu=True,
                                                                    st1=ops.GnStats(st), st2=None if skip is None else ops.GnStats(st_skip))))
        rows[name] = (t_p, t_ps, t_old, t_new)
    for name, (t_p, t_ps, t_old, t_new) in rows.items():
        print(f"B={B} {H}x{H} C={C}+{C2} {name:8s}: producer {t_p:6.1f} us, +totals(+memset) {t_ps:6.1f} | producer+GN old {t_old:6.1f} "
              f"(GN {t_old - t_p:5.1f}) new {t_new:6.1f} (GN {t_new - t_ps:5.1f})")
