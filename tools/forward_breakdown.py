#!/usr/bin/env python
"""Where does one co-batched forward go?  Eager pass of the bench's B=10 forward (+ B=2 ControlNet) with per-op CUDA
events, grouped by kernel and by GEMM / conv / GroupNorm shape (run on the GPU box)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from theatergen_b200 import ops, synth
from theatergen_b200.controlnet import TGControlNet
from theatergen_b200.unet import TGUNet


def main():
    dev = torch.device("cuda")
    net = TGUNet(synth.synth_state_dict(dev, 1234), dev)
    cn = TGControlNet(synth.synth_controlnet_state_dict(dev, 4321), dev)
    B = 10
    x = torch.randn(B, 4, 64, 64, device=dev).half()
    ctx = torch.randn(B, 81, 768, device=dev).half()
    t = torch.full((B,), 501.0, device=dev)
    sc = torch.full((B,), 0.4, device=dev)
    emb = cn.embed_condition(torch.rand(1, 3, 512, 512, device=dev).half())
    control = dict(rows=(8, 10), ctx=ctx[8:, :77].contiguous(), cond_emb=emb)
    for _ in range(2):
        net.forward_raw(x, t, ctx, ip_scale=sc, controlnet=cn, control=control)
    with ops.profile() as rec:
        for _ in range(3):
            net.forward_raw(x, t, ctx, ip_scale=sc, controlnet=cn, control=control)
    rec.finish()
    tot = sum(r[2] for r in rec.tagged) / 3
    print(f"timed ops: {tot:.3f} ms per forward (eager, CUDA events per op)")
    agg = {}
    for k, tag, ms, fl, by in rec.tagged:
        a = agg.setdefault((k, tag), [0, 0.0, 0.0])
        a[0] += 1; a[1] += ms / 3; a[2] += fl / 3
    print(f"{'kernel':22s} {'shape':40s} {'calls':>5s} {'ms':>8s} {'share':>6s} {'us/call':>8s} {'TFLOP/s':>8s}")
    for (k, tag), (n, ms, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
        print(f"{k:22s} {tag:40s} {n // 3:5d} {ms:8.3f} {100 * ms / tot:5.1f}% {1e3 * ms / (n / 3):8.1f} {fl / ms / 1e9 if ms else 0:8.1f}")


if __name__ == "__main__":
    main()
