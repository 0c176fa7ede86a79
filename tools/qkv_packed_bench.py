import os, sys, torch
sys.path.insert(0, os.getcwd())
from theatergen_b200 import ops
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
B, heads = 10, 8
for N, d, C in ((4096, 40, 320), (1024, 80, 640)):
    M = B * N
    x = torch.randn(M, C, device="cuda").half()
    pad = 64 if d < 64 else 128
    wp = (torch.randn(3 * heads * pad, C, device="cuda") * 0.05).half()
    wu = (torch.randn(3 * C, C, device="cuda") * 0.05).half()
    qp = torch.empty(M, 3 * heads * pad, device="cuda").half(); qu = torch.empty(M, 3 * C, device="cuda").half()
    out = torch.empty(M, C, device="cuda").half()
    tg_p = gtime(lambda: ops.gemm(x, wp, out=qp)); tg_u = gtime(lambda: ops.gemm(x, wu, out=qu))
    P = heads * pad
    ta_p = gtime(lambda: ops.attention_tc(qp[:, :P], qp[:, P:2*P], qp[:, 2*P:], B, heads, d, out=out))
    ta_u = gtime(lambda: ops.attention_tc(qu[:, :C], qu[:, C:2*C], qu[:, 2*C:], B, heads, d, out=out, packed=True))
    print(f"N={N} d={d}: qkv GEMM padded {tg_p:.1f} us / packed {tg_u:.1f} us; attention padded {ta_p:.1f} / packed {ta_u:.1f} us")
