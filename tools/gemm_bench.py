#!/usr/bin/env python
"""Per-shape throughput of the tcgen05 GEMM / conv kernels for the SD-1.5 layer shapes (run on the GPU box)."""
import json
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from theatergen_b200 import ops


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / reps


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    rows = []
    g = torch.Generator(device="cuda").manual_seed(0)
    r = lambda *s: torch.randn(*s, device="cuda", generator=g).half()
    # (name, M, N, K) linear shapes
    for C, hw in ((320, 4096), (640, 1024), (1280, 256), (1280, 64)):
        M = B * hw
        for name, N, K in (("proj/out", C, C), ("qkv", 3 * C, C), ("ff1(geglu)", 8 * C, C), ("ff2", C, 4 * C)):
            a, w = r(M, K), r(N, K)
            if "geglu" in name:
                bias = r(N)
                fn = lambda: ops.gemm(a, w, bias, act=ops.TG_ACT_GEGLU)
            else:
                fn = lambda: ops.gemm(a, w)
            ms = timeit(fn)
            rows.append((f"lin {name} C={C} M={M}", M, N, K, ms, 2.0 * M * N * K / ms / 1e9))
    for H, cin, cout in ((64, 320, 320), (64, 640, 320), (64, 960, 320), (32, 640, 640), (32, 1280, 640), (32, 1920, 640),
                         (16, 1280, 1280), (16, 2560, 1280), (8, 1280, 1280), (8, 2560, 1280)):
        x, w = r(B, H, H, cin), r(cout, 3, 3, cin)
        ms = timeit(lambda: ops.conv3x3(x, w))
        M = B * H * H
        rows.append((f"conv3x3 {H}x{H} {cin}->{cout}", M, cout, 9 * cin, ms, 2.0 * M * cout * 9 * cin / ms / 1e9))
    print(f"{'shape':36s} {'M':>6s} {'N':>6s} {'K':>6s} {'us':>9s} {'TFLOP/s':>9s}")
    for name, M, N, K, ms, tf in rows:
        print(f"{name:36s} {M:6d} {N:6d} {K:6d} {ms * 1e3:9.1f} {tf:9.1f}")


if __name__ == "__main__":
    main()
