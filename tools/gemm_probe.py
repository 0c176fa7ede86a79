#!/usr/bin/env python
"""Where does a GEMM k-block go?  Times the single-CTA tcgen05 GEMM / conv on a few SD-1.5 shapes with operand
streams or the MMAs switched off (TG_GEMM_DEBUG bits: 1 = no A loads, 2 = no B loads, 4 = no MMAs; results are garbage
in those modes), next to the CTA-pair kernel and to torch.matmul (cuBLAS) on the same GEMM shape as a calibration of what
the part delivers for that shape.  Run on the GPU box: TG_GEMM_PAIR=0 python tools/gemm_probe.py"""
import os
import sys

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
    return s.elapsed_time(e) / reps * 1e3


def main():
    B = 10
    r = lambda *s: torch.randn(*s, device="cuda").half()
    cases = []
    for name, M, N, K in (("ff1-like 1280", 2560, 10240, 1280), ("ff2 640", 10240, 640, 2560), ("qkv 320", 40960, 960, 320),
                          ("square 8192", 8192, 8192, 8192), ("conv-as-gemm 320", 40960, 320, 2880)):
        a, w = r(M, K), r(N, K)
        cases.append((f"lin {name}", M, N, K, (lambda a=a, w=w: ops.gemm(a, w)), (lambda a=a, w=w: torch.matmul(a, w.t()))))
    for H, cin, cout in ((64, 320, 320), (32, 1280, 640), (16, 1280, 1280)):
        x, w = r(B, H, H, cin), r(cout, 3, 3, cin)
        cases.append((f"conv {H}x{H} {cin}->{cout}", B * H * H, cout, 9 * cin, (lambda x=x, w=w: ops.conv3x3(x, w)), None))
    modes = [("full", "0"), ("noA", "1"), ("noB", "2"), ("noAB", "3"), ("noMMA", "4"), ("noAB+noMMA", "7")]
    print(f"{'shape':28s} {'M':>6s} {'N':>6s} {'K':>6s} " + " ".join(f"{m:>11s}" for m, _ in modes) + f" {'cuBLAS':>9s}   (us; TFLOP/s of 'full' / cuBLAS)")
    for name, M, N, K, fn, ref in cases:
        row = []
        for _, bits in modes:
            os.environ["TG_GEMM_DEBUG"] = bits
            row.append(timeit(fn))
        os.environ["TG_GEMM_DEBUG"] = "0"
        cb = timeit(ref) if ref is not None else float("nan")
        fl = 2.0 * M * N * K
        print(f"{name:28s} {M:6d} {N:6d} {K:6d} " + " ".join(f"{t:11.1f}" for t in row) + f" {cb:9.1f}   {fl / row[0] / 1e6:7.1f} / {fl / cb / 1e6:7.1f}")


if __name__ == "__main__":
    main()
