#!/usr/bin/env python
"""Fixed cost per kernel inside a CUDA graph: graphs of 64 back-to-back launches of tiny problems (dependent chain
through the stream), microseconds per launch."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from theatergen_b200 import ops


def graph_us(fn, n=64, reps=5):
    fn(); fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        g.replay()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / reps / n * 1e3


def main():
    r = lambda *s: torch.randn(*s, device="cuda").half()
    rows = []
    for M, N, K in ((128, 64, 64), (128, 1280, 1280), (512, 1280, 1280), (2560, 1280, 1280), (10240, 640, 640), (40960, 320, 320),
                    (128, 1280, 11520), (640, 1280, 11520)):
        a, w = r(M, K), r(N, K)
        out = torch.empty(M, N, device="cuda", dtype=torch.float16)
        rows.append((f"gemm M={M} N={N} K={K}", graph_us(lambda: ops.gemm(a, w, out=out))))
        if M >= 2560:
            res = r(M, N)
            rows.append((f"gemm M={M} N={N} K={K} +residual", graph_us(lambda: ops.gemm(a, w, residual=res, out=out))))
    for M, C in ((128, 1280), (40960, 320)):
        x, g_, b_ = r(M, C), r(C), r(C)
        o = torch.empty_like(x)
        rows.append((f"layernorm M={M} C={C}", graph_us(lambda: ops.layernorm(x, g_, b_, out=o))))
    for B, HW, C in ((2, 64, 1280), (10, 4096, 320)):
        x, g_, b_ = r(B, HW, C), r(C), r(C)
        o = torch.empty_like(x)
        rows.append((f"groupnorm B={B} HW={HW} C={C}", graph_us(lambda: ops.groupnorm(x, g_, b_, 1e-5, silu=True, out=o))))
    from theatergen_b200.weights import pack_conv3x3
    h, wo, bo = r(10, 64, 64, 320), pack_conv3x3(r(4, 320, 3, 3)), r(4)
    oo = torch.empty(10, 4, 64, 64, device="cuda", dtype=torch.float16)
    rows.append(("conv_out B=10 64x64x320 -> 4", graph_us(lambda: ops.conv_out(h, wo, bo, out=oo), n=16)))
    xi, wi, bi = r(10, 4, 64, 64), ops.pack_conv_in(r(320, 4, 3, 3)), r(320)
    rows.append(("conv_in (im2col + gemm) B=10 4 -> 320", graph_us(lambda: ops.conv_in_gemm(xi, wi, bi), n=16)))
    y, x2 = r(4096), r(4096)
    rows.append(("add_inplace 4096", graph_us(lambda: ops.add_inplace(y, x2))))
    for name, us in rows:
        print(f"{name:44s} {us:8.2f} us per launch (in-graph, back to back)")


if __name__ == "__main__":
    main()
