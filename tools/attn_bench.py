#!/usr/bin/env python
"""Time the tcgen05 self-attention kernel: N=4096 d=40 (64x64 level) and N=1024 d=80 (32x32 level), 8 heads, and the
mma.sync kernel on the d=80 shape for comparison."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from theatergen_b200 import ops
B, heads = int(sys.argv[1]) if len(sys.argv) > 1 else 8, 8


def timeit(f):
    for _ in range(3): f()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10): f()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / 10


for N, d in ((4096, 40), (1024, 80)):
    pad = 64 if d < 64 else 128
    qkv = torch.zeros((B * N, 3, heads, pad), dtype=torch.float16, device="cuda")
    qkv[..., :d] = torch.randn((B * N, 3, heads, d), device="cuda").half()
    flat = qkv.view(B * N, 3 * heads * pad)
    P = heads * pad
    out = torch.empty((B * N, heads * d), dtype=torch.float16, device="cuda")
    ms = timeit(lambda: ops.attention_tc(flat[:, :P], flat[:, P:2 * P], flat[:, 2 * P:], B, heads, d, out=out))
    fl = 4.0 * B * heads * N * N * d
    print(f"tcgen05 B={B} N={N} d={d}: {ms*1e3:.1f} us  {fl/ms/1e9:.1f} TFLOP/s  per-KV-tile/SM cycles ~ {ms*1e-3*1.9e9/ (B*heads*(N/128)*(N/64)/148):.0f}")
    if d == 80:
        q, k, v = [torch.randn(B * N, heads * d, device="cuda").half() for _ in range(3)]
        ms = timeit(lambda: ops.attention(q, k, v, B, heads, out=out))
        print(f"mma.sync B={B} N={N} d={d}: {ms*1e3:.1f} us  {fl/ms/1e9:.1f} TFLOP/s")

# d = 160 at 16x16 (N = 256): packed heads on the tcgen05 kernel vs the mma.sync kernel, device time from a CUDA graph
N, d = 256, 160
q, k, v = [torch.randn(B * N, heads * d, device="cuda").half() for _ in range(3)]
out = torch.empty((B * N, heads * d), dtype=torch.float16, device="cuda")
for name, f in (("tcgen05 packed", lambda: ops.attention_tc(q, k, v, B, heads, d, out=out)),
                ("mma.sync", lambda: ops.attention(q, k, v, B, heads, out=out))):
    for _ in range(3): f()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(20): f()
    g.replay(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5): g.replay()
    e.record(); torch.cuda.synchronize()
    print(f"{name} B={B} N={N} d={d}: {s.elapsed_time(e) / 100 * 1e3:.1f} us")
