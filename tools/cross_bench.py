#!/usr/bin/env python
"""Cross-attention (77 text + 4 image tokens) timing per SD-1.5 level; TG_CROSS_TC=0 selects the mma.sync kernel."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from theatergen_b200 import ops

B, heads = int(sys.argv[1]) if len(sys.argv) > 1 else 10, 8
print(f"TG_CROSS_TC={os.environ.get('TG_CROSS_TC', '1')} B={B}")
for N, C in ((4096, 320), (1024, 640), (256, 1280), (64, 1280)):
    q = torch.randn(B * N, C, device="cuda").half()
    kv = torch.randn(B * 81, 4 * C, device="cuda").half()
    sc = torch.full((B,), 0.4, device="cuda")
    out = torch.empty_like(q)
    fn = lambda: ops.attention_cross_ip(q, kv[:, :C], kv[:, C:2 * C], kv[:, 2 * C:3 * C], kv[:, 3 * C:], B, heads, 77, 4,  # noqa: E731
                                        ip_scale=sc, out=out)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()   # a graph of 20 calls: device time, not the Python wrapper's
    with torch.cuda.graph(g):
        for _ in range(20):
            fn()
    g.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5):
        g.replay()
    e.record()
    torch.cuda.synchronize()
    us = s.elapsed_time(e) / 100 * 1e3
    print(f"  N={N:5d} C={C:5d}: {us:7.1f} us   {2 * B * N * C * 2 / us / 1e3:7.1f} GB/s (Q in + O out)")
