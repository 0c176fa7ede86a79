#!/usr/bin/env python
"""Where do the softmax warps of the tcgen05 self-attention kernel spend their cycles?  (tg_attn_set_prof hook)

    TG_ATTN_VARIANT=0|2|3 python tools/attn_prof.py [B]

Prints the launch time of the 64x64 (N=4096, d=40) and 32x32 (N=1024, d=80) layers and, from one instrumented launch, the
mean cycles per KV tile each softmax warp spent: waiting for S_j (mbarrier), waiting for tcgen05.ld, row max (+ lazy
rescale), exponentials + P store issue, store wait + arrive.  The instrumented kernel is slower (clock reads serialise);
read the SHARES."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from theatergen_b200 import _lib, ops

B, heads = int(sys.argv[1]) if len(sys.argv) > 1 else 10, 8


def timeit(f, n=10):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        f()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n


print(f"TG_ATTN_VARIANT={os.environ.get('TG_ATTN_VARIANT', '(default 2)')}  B={B}")
for N, d in ((4096, 40), (1024, 80)):
    pad = 64 if d < 64 else 128
    qkv = torch.zeros((B * N, 3, heads, pad), dtype=torch.float16, device="cuda")
    qkv[..., :d] = torch.randn((B * N, 3, heads, d), device="cuda").half()
    flat = qkv.view(B * N, 3 * heads * pad)
    P = heads * pad
    out = torch.empty((B * N, heads * d), dtype=torch.float16, device="cuda")
    run = lambda: ops.attention_tc(flat[:, :P], flat[:, P:2 * P], flat[:, 2 * P:], B, heads, d, out=out)  # noqa: E731
    ms = timeit(run)
    fl = 4.0 * B * heads * N * N * d
    tiles_per_sm = B * heads * (N / 128) * (N / 64) / 148
    print(f"N={N} d={d}: {ms * 1e3:.1f} us  {fl / ms / 1e9:.1f} TFLOP/s  ~{ms * 1e-3 * 1.9e9 / tiles_per_sm:.0f} cycles per KV tile per SM at 1.9 GHz")
    # parity against torch SDPA in fp32
    q, k, v = [qkv[:, i, :, :d].reshape(B, N, heads, d).permute(0, 2, 1, 3).float() for i in range(3)]
    ref = torch.nn.functional.scaled_dot_product_attention(q[:2], k[:2], v[:2]).permute(0, 2, 1, 3).reshape(2 * N, heads * d)
    print(f"   max|out - fp32 sdpa| = {(out[:2 * N].float() - ref).abs().max().item():.3e}")
    var = os.environ.get("TG_ATTN_VARIANT", "2")
    two = var == "3" and d < 64 and N % 256 == 0
    nw = 8 if two else 4
    ncta = (N // (256 if two else 128)) * heads * B
    T = N // 64
    buf = torch.zeros(ncta * nw * 8 + 5 * T, dtype=torch.int64, device="cuda")
    _lib.lib().tg_attn_set_prof(buf.data_ptr())
    msp = timeit(run, 3)
    _lib.lib().tg_attn_set_prof(None)
    torch.cuda.synchronize()
    r = buf[:ncta * nw * 8].view(ncta, nw, 8).double()
    tr = buf[ncta * nw * 8:].view(T, 5).cpu()
    tiles = r[..., 5].mean().item()
    names = ["wait S (mbarrier)", "wait tcgen05.ld", "max (+rescale)", "exp + P store issue", "st wait + arrive"]
    if two:
        names = ["wait S + tcgen05.ld", "max (+rescale) + scale", "wait for the MUFU turn", "exponentials", "P store + arrive"]
    tot = r[..., :5].sum(-1).mean().item()
    print(f"   instrumented launch {msp * 1e3:.1f} us; per softmax warp and KV tile ({tiles:.0f} tiles): total {tot / tiles:.0f} cycles")
    for i, n in enumerate(names):
        c = r[..., i].mean().item() / tiles
        print(f"     {n:22s} {c:7.0f} cycles  {100 * c * tiles / tot:5.1f} %")
    if not two:
        t0 = int(tr[8, 0])
        print("   trace of CTA 0 (cycles relative to tile 8's p_full arrive by softmax warp 2):")
        print("     tile | softmax arrives P_j | MMA past wait P_j | MMA committed S_j | MMA past kv wait (S_j) | softmax sees S_j")
        for j in range(8, 16):
            print(f"     {j:4d} | {int(tr[j, 0]) - t0:8d} | {int(tr[j, 1]) - t0:8d} | {int(tr[j, 2]) - t0:8d} | {int(tr[j, 4]) - t0:8d} | {int(tr[j, 3]) - t0:8d}")
