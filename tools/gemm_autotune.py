#!/usr/bin/env python
"""Measure, don't guess: time every valid (tile width, split-K, CTA pairing) of gemm_tc_kernel for each GEMM / conv
shape of the bench workload and write the winners to theatergen_b200/gemm_tuned.json (loaded by _lib at start-up; shapes
without an entry keep the built-in cost model).  Run on the GPU box:

    python tools/gemm_autotune.py [--report gpurun_out/gemm_autotune.md] [--min-gain 0.04]

Timing conditions mirror the product step: the weight is COLD (1.76 GB of weights stream through the 126 MB L2 between two
uses of a layer: L2 is flushed before every launch) and the activation is WARM (it was just written by the previous
kernel: re-written after the flush).  One CUDA-event pair per launch, median of the repetitions.  A candidate must also
reproduce the default configuration's result (guards the forced-tile path against an invalid combination).
"""
import argparse
import json
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from theatergen_b200 import _lib, ops, synth
from theatergen_b200.controlnet import TGControlNet
from theatergen_b200.unet import TGUNet

BNS = (256, 192, 160, 128, 64)
SPLITS = (1, 2, 3, 4, 6, 8)


def collect_shapes(dev):
    """{(kind, M, N, K, act2, B, H, W): calls per bench step} from eager passes of the bench's forwards."""
    os.environ["TG_CN_OVERLAP"] = "0"
    net = TGUNet(synth.synth_state_dict(dev, 1234), dev)
    cn = TGControlNet(synth.synth_controlnet_state_dict(dev, 4321), dev)
    emb = cn.embed_condition(torch.rand(1, 3, 512, 512, device=dev).half())
    shapes = {}

    def run(B, control_rows, weight):
        x = torch.randn(B, 4, 64, 64, device=dev).half()
        ctx = torch.randn(B, 81, 768, device=dev).half()
        t = torch.full((B,), 501.0, device=dev)
        sc = torch.full((B,), 0.4, device=dev)
        control = None
        if control_rows:
            r0, r1 = control_rows
            control = dict(rows=(r0, r1), ctx=ctx[r0:r1, :77].contiguous(), cond_emb=emb)
        with ops.profile() as rec:
            net.forward_raw(x, t, ctx, ip_scale=sc, controlnet=cn if control else None, control=control)
        rec.finish()
        for k, tag, ms, fl, by in rec.tagged:
            if k != "gemm_tc_kernel":
                continue
            m = re.match(r"(lin|conv) M=(\d+) N=(\d+) K=(\d+)(?: act=(\d+))?(?: B=(\d+) H=(\d+) W=(\d+))?", tag)
            kind, M, N, K = m.group(1), int(m.group(2)), int(m.group(3)), int(m.group(4))
            act2 = 1 if m.group(5) == "2" else 0
            geo = tuple(int(v) for v in m.group(6, 7, 8)) if m.group(6) else (0, 0, 0)
            key = (kind, M, N, K, act2) + geo
            shapes[key] = shapes.get(key, 0) + weight
    run(10, (8, 10), 1.0)   # the co-batched bench step (MixedStepGraph)
    run(8, None, 0.0)       # stage-1 alone / stage-2 alone: tuned too, weight 0 in the report's total
    run(2, (0, 2), 0.0)
    del net, cn
    torch.cuda.empty_cache()
    return shapes


def candidates(M, N, K, act2, sms=148):
    num_m, nk = (M + 127) // 128, K // 64
    out = []
    for pair in (1, 0):
        if pair and num_m < 2:
            continue
        workers = sms // 2 if pair else sms
        rows_m = (num_m + 1) // 2 if pair else num_m
        for bn in BNS:
            if N % bn or (act2 and bn % 128):
                continue
            tiles = rows_m * (N // bn)
            for s in SPLITS:
                if s > 1 and (act2 or nk < 4 * s or tiles * s > 2 * workers):
                    continue
                if s > 1 and (num_m * (N // bn) * s * 128 * bn * 4 > (64 << 20) or num_m * (N // bn) > 4096):
                    continue
                out.append((bn, s, pair))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=7)
    ap.add_argument("--min-gain", type=float, default=0.04, help="keep a measured choice only if it beats the model by this much")
    ap.add_argument("--report", default=None)
    ap.add_argument("--only", default=None, choices=["lin", "conv"], help="restrict the search to one kind of shape")
    ap.add_argument("--min-calls", type=float, default=0.0, help="skip shapes with fewer calls per bench step")
    ap.add_argument("--out", default=_lib.TUNED_PATH)
    args = ap.parse_args()
    dev = torch.device("cuda")
    L = _lib.lib()
    L.tg_gemm_clear_tuned()          # measure against the cost model, not against a previous table
    shapes = collect_shapes(dev)
    flush = torch.empty(192 << 20, dtype=torch.uint8, device=dev)
    g = torch.Generator(device=dev).manual_seed(0)

    def time_one(fn, a_dst, a_src):
        ts = []
        for _ in range(args.reps):
            flush.fill_(1)
            a_dst.copy_(a_src)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            fn()
            e.record()
            e.synchronize()
            ts.append(s.elapsed_time(e) * 1e3)
        ts.sort()
        return ts[len(ts) // 2]

    entries, rows = [], []
    saved = total_default = 0.0
    for key, calls in sorted(shapes.items(), key=lambda kv: -kv[1]):
        kind, M, N, K, act2, B, H, W = key
        if (args.only and kind != args.only) or calls < args.min_calls:
            continue
        wt = (torch.randn(N, K, device=dev, generator=g) * K ** -0.5).half()
        bias = (torch.randn(N, device=dev, generator=g) * 0.1).half()
        if kind == "conv":
            cin = K // 9
            a_src = torch.randn(B, H, W, cin, device=dev, generator=g).half()
            a = a_src.clone()
            w4 = wt.view(N, 3, 3, cin)
            out = torch.empty(B, H, W, N, device=dev, dtype=torch.float16)
            fn = lambda: ops.conv3x3(a, w4, bias, out=out)
        else:
            a_src = torch.randn(M, K, device=dev, generator=g).half()
            a = a_src.clone()
            out = torch.empty(M, N // 2 if act2 else N, device=dev, dtype=torch.float16)
            act = ops.TG_ACT_GEGLU if act2 else ops.TG_ACT_NONE
            fn = lambda: ops.gemm(a, wt, bias, act=act, out=out)
        L.tg_gemm_force_tile(0, 1, 0)
        fn()
        ref = out.float().clone()
        t_def = time_one(fn, a, a_src)
        best, t_best, results = None, t_def, []
        for bn, s, pair in candidates(M, N, K, act2):
            L.tg_gemm_force_tile(bn, s, pair)
            try:
                fn()
                err = (out.float() - ref).abs().max().item()
                if not err <= 2e-2 * max(1.0, ref.abs().max().item()):
                    results.append((bn, s, pair, None))
                    continue
                t = time_one(fn, a, a_src)
            except RuntimeError:
                continue
            finally:
                L.tg_gemm_force_tile(0, 1, 0)
            results.append((bn, s, pair, t))
            if t < t_best:
                best, t_best = (bn, s, pair), t
        gain = 1.0 - t_best / t_def
        keep = best is not None and gain >= args.min_gain
        if keep:
            entries.append([M, N, K, 2 if act2 else 0, 1 if kind == "conv" else 0, *best])
            saved += calls * (t_def - t_best)
        total_default += calls * t_def
        rows.append((kind, M, N, K, act2, calls, t_def, best, t_best, keep))
        print(f"{kind:4s} M={M:6d} N={N:5d} K={K:5d} act2={act2} calls={calls:4.0f} model {t_def:7.1f} us  best {best} "
              f"{t_best:7.1f} us  {'KEEP' if keep else '-'}", flush=True)
    json.dump({"source": "tools/gemm_autotune.py on 1x B200: cold weights, warm activations, median of %d" % args.reps,
               "fields": ["M", "N", "K", "act", "conv", "bn", "split", "pair"], "entries": entries},
              open(args.out, "w"), indent=0)
    msg = (f"shapes {len(shapes)}, entries kept {len(entries)}; GEMM time of one co-batched forward (cold-weight launches): "
           f"{total_default / 1e3:.2f} ms by the model, {(total_default - saved) / 1e3:.2f} ms with the table")
    print(msg)
    if args.report:
        with open(args.report, "w") as f:
            f.write("# gemm_autotune: measured tile choice per shape (cold weights, warm activations)\n\n" + msg + "\n\n")
            f.write("| kind | M | N | K | geglu | calls/step | model us | best (bn, split, pair) | best us | kept |\n|---|---|---|---|---|---|---|---|---|---|\n")
            for r in rows:
                f.write("| %s | %d | %d | %d | %d | %.0f | %.1f | %s | %.1f | %s |\n" % (r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7], r[8], "yes" if r[9] else ""))


if __name__ == "__main__":
    main()
