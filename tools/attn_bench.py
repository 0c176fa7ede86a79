#!/usr/bin/env python
"""Time the tcgen05 self-attention kernel (N=4096, d=40, 8 heads). TG_ATTN_VARIANT selects ablations."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from theatergen_b200 import ops
B, N, d, heads = int(sys.argv[1]) if len(sys.argv) > 1 else 8, 4096, 40, 8
qkv = torch.zeros((B * N, 3, heads, 64), dtype=torch.float16, device="cuda")
qkv[..., :d] = torch.randn((B * N, 3, heads, d), device="cuda").half()
flat = qkv.view(B * N, 3 * heads * 64)
P = heads * 64
out = torch.empty((B * N, heads * d), dtype=torch.float16, device="cuda")
f = lambda: ops.attention_tc(flat[:, :P], flat[:, P:2 * P], flat[:, 2 * P:], B, heads, d, out=out)
for _ in range(3): f()
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(10): f()
e.record(); torch.cuda.synchronize()
ms = s.elapsed_time(e) / 10
fl = 4.0 * B * heads * N * N * d
print(f"variant={os.environ.get('TG_ATTN_VARIANT','0')} B={B}: {ms*1e3:.1f} us  {fl/ms/1e9:.1f} TFLOP/s  per-KV-tile/SM cycles ~ {ms*1e-3*1.9e9/ (B*heads*(N/128)*(N/64)/148):.0f}")
