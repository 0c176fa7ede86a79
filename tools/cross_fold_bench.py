#!/usr/bin/env python
"""Folded cross-attention (tg_cross_scores + tg_cross_apply) against the three launches it replaces (to_q GEMM, cross-attention
kernel, to_out GEMM) per UNet level, B = 10: device time per launch from a CUDA graph of 20 launches (L2-warm)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from theatergen_b200 import ops
B, heads = int(sys.argv[1]) if len(sys.argv) > 1 else 10, 8


def gtime(f, n=20):
    for _ in range(3): f()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n): f()
    g.replay(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5): g.replay()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / (5 * n) * 1e3


for HW, C in ((4096, 320), (1024, 640), (256, 1280)):
    M, S = B * HW, ops.CROSS_SLOTS
    x, h = [torch.randn(M, C, device="cuda").half() for _ in range(2)]
    wq, wo = [(torch.randn(C, C, device="cuda") * C ** -0.5).half() for _ in range(2)]
    bo = torch.zeros(C, device="cuda").half()
    kv = torch.randn(B * 81, 4 * C, device="cuda").half()
    sc = torch.full((B,), 0.4, device="cuda")
    kf = (torch.randn(B, heads * S, C, device="cuda") * 0.05).half()
    vf = (torch.randn(B, C, heads * S, device="cuda") * 0.05).half()
    q = torch.empty(M, C, device="cuda").half(); a = torch.empty_like(q); o = torch.empty_like(q)
    pm = torch.empty(M, heads * S, device="cuda").half()
    t_q = gtime(lambda: ops.gemm(x, wq, out=q))
    t_a = gtime(lambda: ops.attention_cross_ip(q, kv[:, :C], kv[:, C:2 * C], kv[:, 2 * C:3 * C], kv[:, 3 * C:], B, heads, 77, 4,
                                               ip_scale=sc, out=a))
    t_o = gtime(lambda: ops.gemm(a, wo, bo, residual=h, out=o))
    t_s = gtime(lambda: ops.cross_scores(x, kf, 77, 4, sc, out=pm))
    t_p = gtime(lambda: ops.cross_apply(pm, vf, bo, residual=h, out=o))
    print(f"HW={HW} C={C}: to_q {t_q:.1f} + attention {t_a:.1f} + to_out {t_o:.1f} = {t_q + t_a + t_o:.1f} us | "
          f"scores {t_s:.1f} + apply {t_p:.1f} = {t_s + t_p:.1f} us")
