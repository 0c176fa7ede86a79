#!/usr/bin/env python
"""bench.py -- denoise-steps/sec of TheaterGen's hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    (N > 1: launched by torch.distributed.run, one rank per GPU)

Workload (BASELINE.json configs[2] / SURVEY.md 8d "config 3", the configuration the metric is quoted on):
SD-1.5 512x512, 4 characters with layout masks + attention-guidance energy, 50 DDIM steps, CFG 7.5, context =
77 text + 4 IP-Adapter tokens, seeded synthetic SD-1.5-shaped weights (no checkpoints exist offline).
One bench "step" is one timestep index of a turn:
    * the stage-1 denoise step of the 4 co-batched characters (one B=8 UNet forward, per-sample IP scales
      0.0/0.4/0.4/0.4, fused CFG+DDIM)                                          = 4 denoise steps
    * the stage-2 denoise step of the composed image (B=2 ControlNet-lineart forward feeding its 13 residuals into a
      B=2 UNet forward, IP scale 0.1, fused CFG+DDIM+masked composite, capture of the 4 guidance attention maps +
      guidance energy for index < 30)                                            = 1 denoise step
so value = 5 * K / seconds (x ranks; every rank owns independent dialogues -> weak scaling, no data-path
collective; the UNet/IP weights are NCCL-broadcast from rank 0 once at start-up).
By default the five trajectories share ONE UNet forward of batch 10 (per-sample IP scale; composite, capture and
energy apply to the stage-2 trajectory only): in steady state a rank always holds a turn in stage 1 and another
dialogue's turn in stage 2 (config 5: 8 dialogues per GPU), and co-batching them is a scheduling decision, not a
change of arithmetic.  `--separate-stage2` runs the two forwards as separate graphs (B=8 then B=2) instead.
The stage-2 ControlNet (reference models/pipelines.py:770-778) runs on our kernels inside the same
graph; its timestep-independent conditioning embedder runs once per trajectory, at timestep index 0.  The timed region
STARTS at index 0 (the warm-up steps are the last indices of the previous trajectory), so the embedder, the per-trajectory
host->device copies of the e2e leg and the energy-carrying indices (< 30) are inside every timed window.

Extra keys next to the contract's (all computed outside the timed region, rank 0, N = 1):
  parity               the 50-step trajectories of this very workload (the co-batched graphs that were timed) against the
                       oracle loops run in fp32 on the same GPU, with the oracle's fp16 run (the reference's numerics) as
                       the noise floor: max-abs latent drift, per step and final; guidance energy at step 0.
  gpu_eager_reference  the oracle -- the reference's op structure, materialised attention probabilities and all -- run by
                       PyTorch eager fp16 (cuBLAS / cuDNN) on this GPU: forward times and the reference's loop with its
                       per-step latents.cpu() ("the number to beat").
  sustained            the same device-resident step loop run for >= 3 s (250 steps = five 50-step trajectories).
The oracle is the checker in those legs, never the thing measured as ours.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOPS_PER_DENOISE_STEP = 1.607e12  # one B=2 UNet forward, 81 context tokens
FLOPS_CONTROLNET_STEP = 0.523e12   # ControlNet encoder + mid + zero-convs per stage-2 step (B=2)
KEYS = [("mid", 0, 0, 0), ("up", 1, 0, 0), ("up", 1, 1, 0), ("up", 1, 2, 0)]
XYWH = [[278, 251, 157, 185], [358, 328, 150, 180], [168, 4, 100, 80], [4, 331, 80, 70]]  # CMIGBench dialogue 4
WORKLOAD = ("sd15-512x512, 4 characters (stage-1 co-batched B=8, ip_scale 0/0.4/0.4/0.4) + stage-2 composite step "
            "(B=2 ControlNet-lineart + UNet, ip_scale 0.1, layout mask, 4-map guidance energy for index<30), "
            "50 DDIM steps, cfg 7.5, ctx 77+4 tokens")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1400.0), d.get("hbm_gbs", 6650.0), "measured"
    return 1400.0, 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw,enforced.power.limit")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        def _f(r, i):
            try:
                return float(r[i])
            except (IndexError, ValueError):
                return None
        pw = sorted(v for v in (_f(r, 6) for r in self.rows) if v is not None)
        lim = [v for v in (_f(r, 7) for r in self.rows) if v is not None]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "power_w": pw[len(pw) // 2] if pw else None,
                "power_limit_w": max(lim) if lim else None}


# =============================================================================================================
# reference arm / cpu baseline: the oracle (CPU restatement of the reference's PyTorch path), fp32, all host cores
# =============================================================================================================
def cpu_denoise_steps(n_steps, warmup):
    import torch
    from oracle import ddim as OD
    from oracle import unet as OU
    torch.set_num_threads(os.cpu_count() or 1)
    sd = OU.synth_state_dict(1234)
    g = torch.Generator().manual_seed(0)
    lat = torch.randn(1, 4, 64, 64, generator=g)
    ctx = torch.randn(2, 81, 768, generator=g)
    sched = OD.DDIM()
    sched.set_timesteps(50)
    times = []
    with torch.no_grad():
        for i in range(warmup + n_steps):
            t = int(sched.timesteps[i % 50])
            t0 = time.perf_counter()
            eps = OU.unet_forward(sd, torch.cat([lat] * 2), t, ctx, ip_scale=0.4)
            lat = sched.step(OD.cfg(eps, 7.5), t, lat)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    return sum(times), torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    total, cores = cpu_denoise_steps(args.steps, args.warmup)
    v = args.steps / total
    sample = f"{args.steps} stage-1 denoise steps of ONE character (B=2 UNet fwd + CFG + DDIM), fp32, oracle port on {cores} host threads"
    print(json.dumps({
        "impl": "reference", "metric": "denoise-steps/sec", "value": v, "unit": "steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "CPU arm times a bounded sample: " + sample},
        "cpu_baseline": {"value": v, "unit": "steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


# =============================================================================================================
# checker legs (outside every timed region): 50-step parity against the oracle, same-GPU eager reference timing
# =============================================================================================================
def parity_leg(torch, dev, gme, gm, embed_hint, d_hint, d_ts, d_coef, d_frozen, h_latm, sdu, sdc, ctx1, ctx2, lat1, lat2,
               ctrl_ctx, mask, bboxes, K, STEPS):
    """Runs the timed workload itself (co-batched B=10 graphs: 4 stage-1 characters + the stage-2 trajectory with
    ControlNet, composite, capture, energy) for one full 50-step trajectory and compares every step's latents with the
    oracle's loops (oracle/parity.py: fp32 = the yardstick, fp16 = the reference's own numerics)."""
    from oracle import parity as P
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    gme.latents.copy_(h_latm)
    trace = [gme.latents.clone()]
    energies = []
    for idx in range(STEPS):
        if idx == 0:
            embed_hint(d_hint)
        g = gme if idx < 30 else gm
        g.t.copy_(d_ts[idx].expand(2 * (K + 1))); g.coef.copy_(d_coef[idx]); g.frozen.copy_(d_frozen[idx + 1])
        g.replay()
        if idx == 29:
            gm.latents.copy_(gme.latents)
        trace.append(g.trace.clone())
        if idx < 30:
            energies.append(gme.loss.clone())
    torch.cuda.synchronize()
    ours = torch.stack(trace)                                   # [51, K+1, 4, 64, 64]
    energies = [e.item() for e in energies]
    scales = [0.0, 0.4, 0.4, 0.4]
    en = dict(bboxes=bboxes, positions=[[5, 6], [9], [12, 13, 14], [17]], upto=30,
              kwargs=dict(fg_top_p=0.2, bg_top_p=0.2, fg_weight=1.0, bg_weight=4.0))
    out = {}
    t0 = time.perf_counter()
    refs = {}
    for name, dt in (("fp32", torch.float32), ("fp16", torch.float16)):
        r1 = P.stage1_traces(sdu, lat1.to(dev), ctx1.to(dev), scales, STEPS, dt)
        r2, e2 = P.stage2_trace(sdu, sdc, lat2.to(dev), ctx2.to(dev), ctrl_ctx.to(dev), d_hint, d_frozen, mask.to(dev),
                                STEPS, STEPS, dt, energy=en)
        refs[name] = (torch.cat([r1, r2], 1), e2)
    s = P.summarize(ours, refs["fp32"][0], refs["fp16"][0])
    s1 = P.summarize(ours[:, :K], refs["fp32"][0][:, :K], refs["fp16"][0][:, :K])
    s2 = P.summarize(ours[:, K:], refs["fp32"][0][:, K:], refs["fp16"][0][:, K:])
    e32, e16 = refs["fp32"][1], refs["fp16"][1]
    rel = lambda a, b: max(abs(x - y) / max(abs(y), 1e-6) for x, y in zip(a, b))  # noqa: E731
    out.update({k: s[k] for k in ("steps", "trajectories", "max_abs_latent_vs_fp32_oracle", "fp16_oracle_gap",
                                  "first_step_max_abs", "first_step_fp16_oracle_gap", "final_max_abs",
                                  "final_fp16_oracle_gap", "rms_final", "rms_final_fp16_oracle")})
    out["stage1_4chars"] = {"max_abs": s1["max_abs_latent_vs_fp32_oracle"], "fp16_oracle_gap": s1["fp16_oracle_gap"]}
    out["stage2_controlnet_composite"] = {"max_abs": s2["max_abs_latent_vs_fp32_oracle"],
                                          "fp16_oracle_gap": s2["fp16_oracle_gap"]}
    out["per_step_max_abs_every_5"] = s["per_step_max_abs"][::5]
    out["per_step_fp16_oracle_gap_every_5"] = s["per_step_fp16_oracle_gap"][::5]
    out["guidance_energy"] = {"step0_ours": energies[0], "step0_fp32_oracle": e32[0],
                              "max_rel_diff_30_steps": rel(energies, e32), "fp16_oracle_max_rel_diff": rel(e16, e32)}
    out["north_star_tolerance"] = "1e-3 max-abs on latents"
    out["what"] = ("max over steps / trajectories / elements of |latents - fp32 oracle| for the timed co-batched workload "
                   "(fp16 storage, fp32 accumulate); fp16_oracle_gap = the same for the oracle run in fp16, i.e. the "
                   "reference's own arithmetic on this GPU")
    out["oracle_seconds"] = round(time.perf_counter() - t0, 1)
    return out


def eager_reference_leg(torch, dev, sdu, sdc, ctx1, ctx2, lat1, lat2, ctrl_ctx, d_hint, d_frozen, mask, K):
    """the oracle -- the reference's op sequence (baddbmm -> softmax -> bmm attention with materialised
    probabilities, unfused norms, NCHW convolutions) -- executed by PyTorch eager fp16 on this GPU, with the reference's
    per-step `latents.cpu()` (models/pipelines.py:449-453, 835).  One bench step of the reference = the 4 characters'
    stage-1 steps one after the other (B=2 each; the reference cannot co-batch: its IP scale is a module attribute) + one
    stage-2 step (ControlNet + UNet B=2 + composite)."""
    from oracle import controlnet as OC
    from oracle import ddim as OD
    from oracle import unet as OU
    u16 = {k: v.half() for k, v in sdu.items()}
    c16 = {k: v.half() for k, v in sdc.items()}
    sched = OD.DDIM()
    sched.set_timesteps(50)
    ts = [int(t) for t in sched.timesteps]
    scales = [0.0, 0.4, 0.4, 0.4]
    lat = [lat1[k:k + 1].to(dev) for k in range(K)] + [lat2.to(dev)]
    cx = [torch.stack([ctx1[k], ctx1[K + k]]).to(dev) for k in range(K)] + [ctx2.to(dev)]
    cc, cond = ctrl_ctx.to(dev), d_hint.repeat(2, 1, 1, 1)
    fm = mask.to(dev).half()
    host = []

    def fwd_ms(B, reps=5):
        g = torch.Generator().manual_seed(B)
        x = torch.randn(B, 4, 64, 64, generator=g).half().to(dev)
        c = torch.randn(B, 81, 768, generator=g).half().to(dev)
        for _ in range(2):
            OU.unet_forward(u16, x, 501, c, ip_scale=0.4)
        s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        s_.record()
        for _ in range(reps):
            OU.unet_forward(u16, x, 501, c, ip_scale=0.4)
        e_.record()
        torch.cuda.synchronize()
        return round(s_.elapsed_time(e_) / reps, 3)

    def bench_step(i):
        t = ts[i % 50]
        for k in range(K):
            eps = OU.unet_forward(u16, torch.cat([lat[k]] * 2), t, cx[k], ip_scale=scales[k])
            lat[k] = sched.step(OD.cfg(eps, 7.5), t, lat[k])
            host.append(lat[k].cpu())
        x = torch.cat([lat[K]] * 2)
        d, m = OC.controlnet_forward(c16, x, t, cc, cond)
        eps = OU.unet_forward(u16, x, t, cx[K], ip_scale=0.1, down_residuals=d, mid_residual=m)
        l2 = sched.step(OD.cfg(eps, 7.5), t, lat[K])
        lat[K] = d_frozen[(i % 50) + 1] * fm + l2 * (1.0 - fm)
        host.append(lat[K].cpu())

    with torch.no_grad():
        f2, f10 = fwd_ms(2), fwd_ms(10, 3)
        n, warm = 10, 2
        for i in range(warm):
            bench_step(i)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(warm, warm + n):
            bench_step(i)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    return {"value": round((K + 1) * n / dt, 2), "unit": "steps/s", "ms_per_bench_step": round(1e3 * dt / n, 3),
            "bench_steps_timed": n, "unet_forward_ms_B2": f2, "unet_forward_ms_B10": f10, "dtype": "f16",
            "what": "oracle (reference op structure) in PyTorch eager fp16 on this GPU: 4 sequential B=2 stage-1 steps + "
                    "1 stage-2 step (ControlNet + UNet + composite) per bench step, latents.cpu() after every denoise "
                    "step like the reference; no guidance energy (the reference never evaluates it)"}


# =============================================================================================================
# our arm
# =============================================================================================================
def build_inputs(torch, dev, K=4):
    g = torch.Generator().manual_seed(4321)
    ctx1 = torch.randn(2 * K, 81, 768, generator=g).half()          # rows [0,K) uncond, [K,2K) cond
    ctx2 = torch.randn(2, 81, 768, generator=g).half()
    lat1 = torch.randn(K, 4, 64, 64, generator=torch.Generator().manual_seed(7)).half()
    lat2 = torch.randn(1, 4, 64, 64, generator=torch.Generator().manual_seed(8)).half()
    x0 = torch.randn(1, 4, 64, 64, generator=torch.Generator().manual_seed(77))
    nz = torch.randn(1, 4, 64, 64, generator=torch.Generator().manual_seed(78))
    scale1 = torch.tensor([0.0, 0.4, 0.4, 0.4] * 2, dtype=torch.float32)
    return ctx1, ctx2, lat1, lat2, x0, nz, scale1


def run_ours(args):
    import torch
    import torch.distributed as dist
    from theatergen_b200 import _lib, ops, synth
    from theatergen_b200.controlnet import TGControlNet
    from theatergen_b200.engine import DenoiseEngine
    from theatergen_b200.unet import TGUNet
    from theatergen_b200.utils.guidance import EnergyPlan

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (ours) needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.lib()  # fail loudly if the extension is missing

    # ---- weights: rank 0 generates, NCCL broadcast to the others (the only collective on this path) ----
    def shared(sd):
        if world > 1:
            flat = torch.cat([v.reshape(-1) for v in sd.values()])
            dist.broadcast(flat, 0)
            off = 0
            for k, v in sd.items():
                sd[k] = flat[off:off + v.numel()].view_as(v)
                off += v.numel()
        return sd

    sdu = shared(synth.synth_state_dict(dev, 1234, fill=(rank == 0)))
    sdc = shared(synth.synth_controlnet_state_dict(dev, 4321, fill=(rank == 0)))
    net, cnet = TGUNet(sdu, dev), TGControlNet(sdc, dev)
    eng = DenoiseEngine(net, controlnet=cnet)
    K, STEPS = 4, 50
    ctx1, ctx2, lat1, lat2, x0, nz, scale1 = build_inputs(torch, dev, K)
    # synthetic line-art hint of the collage (stand-in for LineartDetector output): sparse strokes in [0,1]
    hint = (torch.rand(1, 1, 512, 512, generator=torch.Generator().manual_seed(99)) > 0.93).half().repeat(1, 3, 1, 1)
    h_hint = hint.contiguous().pin_memory()
    d_hint = hint.to(dev)
    ctrl_ctx = ctx2[:, :77].contiguous()  # the ControlNet sees the 77 text tokens only
    ts, coef = eng.tables(STEPS, 7.5)
    sched = eng.scheduler
    frozen_all = torch.cat([lat2.float()] + [sched.add_noise(x0, nz, torch.tensor([int(t)])) for t in sched.timesteps]).half()
    frozen_all = frozen_all.view(STEPS + 1, 1, 4, 64, 64)
    mask = torch.zeros(64, 64)
    for x, y, w, h in XYWH:
        mask[round(y / 8):round(y / 8) + round(h / 8), round(x / 8):round(x / 8) + round(w / 8)] = 1.0
    bboxes = [[x / 512, y / 512, (x + w) / 512, (y + h) / 512] for x, y, w, h in XYWH]
    plan = EnergyPlan(bboxes, [[5, 6], [9], [12, 13, 14], [17]], KEYS, fg_top_p=0.2, bg_top_p=0.2, fg_weight=1.0,
                      bg_weight=4.0)

    # pinned host staging (e2e leg) and device-resident copies (kernel leg)
    pin = lambda t: t.contiguous().pin_memory()
    h_ctx1, h_ctx2, h_lat1, h_lat2, h_frozen, h_mask = map(pin, (ctx1, ctx2, lat1, lat2, frozen_all, mask.half()))
    h_ts, h_coef = pin(ts.cpu()), pin(coef.cpu())
    h_ts8, h_ts2 = pin(ts.cpu()[:, None].repeat(1, 2 * K)), pin(ts.cpu()[:, None].repeat(1, 2))
    h_out1 = torch.empty((K, 4, 64, 64), dtype=torch.float16).pin_memory()
    h_out2 = torch.empty((1, 4, 64, 64), dtype=torch.float16).pin_memory()
    h_loss = torch.empty((1,), dtype=torch.float32).pin_memory()
    d_frozen, d_ts, d_coef = frozen_all.to(dev), ts, coef

    g1 = eng.step_graph(K, 81, False)
    loss = torch.zeros((1,), dtype=torch.float32, device=dev)
    ws = torch.empty((256,), dtype=torch.float32, device=dev)
    # stage 2: two graphs -- with and without the guidance energy (index < 30)
    from theatergen_b200.engine import StepGraph

    class Stage2(StepGraph):
        def __init__(self, energy):
            super().__init__(eng, 1, 81, True, KEYS, controlnet=True)
            self.energy = energy

        def _body(self):
            super()._body()
            if self.energy:
                plan.run([self.captured[k][0] for k in KEYS], loss, ws)

    g2e, g2 = Stage2(True).capture(), Stage2(False).capture()
    from theatergen_b200.engine import MixedStepGraph
    mixed = not args.separate_stage2
    gme = gm = None
    if mixed:
        gme = MixedStepGraph(eng, K, 1, 81, KEYS, plan, controlnet=True).capture()
        gm = MixedStepGraph(eng, K, 1, 81, None, None, controlnet=True).capture()
        for g in (gme, gm):  # batch rows: [stage-1 uncond | stage-1 cond | stage-2 uncond | stage-2 cond]
            g.set_context(torch.cat([ctx1, ctx2]), ctrl_ctx)   # context + its K / V projections: once per trajectory
            g.ip_scale.copy_(torch.tensor([0.0, 0.4, 0.4, 0.4] * 2 + [0.1, 0.1]))
            g.latents.copy_(torch.cat([lat1, lat2]))
            g.mask.copy_(h_mask)
        h_ctxm = pin(torch.cat([ctx1, ctx2]))
        h_ctrl = pin(ctrl_ctx)
        h_latm = pin(torch.cat([lat1, lat2]))
        h_tsm = pin(ts.cpu()[:, None].repeat(1, 2 * (K + 1)))
        h_outm = torch.empty((K + 1, 4, 64, 64), dtype=torch.float16).pin_memory()
    g1.set_context(h_ctx1); g1.ip_scale.copy_(scale1); g1.latents.copy_(h_lat1)
    for g in (g2e, g2):
        g.set_context(h_ctx2, ctrl_ctx); g.ip_scale.fill_(0.1); g.latents.copy_(h_lat2); g.mask.copy_(h_mask)
    stage2_graphs = (gme, gm) if mixed else (g2e, g2)

    def embed_hint(src):
        """per-trajectory ControlNet work: conditioning embedder on the hint image (timestep independent)"""
        emb = cnet.embed_condition(src)
        for g in stage2_graphs:
            g.cond_emb.copy_(emb)

    n0 = _lib.launch_count()
    embed_hint(d_hint)
    embed_launches = _lib.launch_count() - n0
    torch.cuda.synchronize()
    launches_per_step = {True: g1.launches_per_step + g2e.launches_per_step,
                         False: g1.launches_per_step + g2.launches_per_step}
    if mixed:
        launches_per_step = {True: gme.launches_per_step, False: gm.launches_per_step}

    def step_device_mixed(i):
        idx = i % STEPS
        if idx == 0:  # per-trajectory work: hint embedder + K / V projections of the (device-resident) context
            embed_hint(d_hint)
            gme.set_context(); gm.set_context()
        g = gme if idx < 30 else gm
        g.t.copy_(d_ts[idx].expand(2 * (K + 1))); g.coef.copy_(d_coef[idx]); g.frozen.copy_(d_frozen[idx + 1])
        g.replay()
        if idx == STEPS - 1:
            gme.latents.copy_(h_latm, non_blocking=True)
        elif idx == 29:
            gm.latents.copy_(gme.latents)

    def step_e2e_mixed(i):
        idx = i % STEPS
        g = gme if idx < 30 else gm
        if idx == 0:
            gme.set_context(h_ctxm, h_ctrl); gm.set_context(h_ctxm, h_ctrl)   # H2D + K / V projection, once per trajectory
            gme.latents.copy_(h_latm, non_blocking=True); gme.mask.copy_(h_mask, non_blocking=True)
            gm.mask.copy_(h_mask, non_blocking=True)
            d_hint.copy_(h_hint, non_blocking=True)
            embed_hint(d_hint)
        g.t.copy_(h_tsm[idx], non_blocking=True); g.coef.copy_(h_coef[idx], non_blocking=True)
        g.frozen.copy_(h_frozen[idx + 1], non_blocking=True)
        g.replay()
        if idx == 29:
            gm.latents.copy_(gme.latents)
        h_outm.copy_(g.trace, non_blocking=True)
        if idx < 30:
            h_loss.copy_(gme.loss, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def step_device(i):
        """inputs resident in HBM; only device-side work"""
        idx = i % STEPS
        if idx == 0:
            embed_hint(d_hint)
            g1.set_context(); g2e.set_context(); g2.set_context()
        g1.t.copy_(d_ts[idx].expand(2 * K)); g1.coef.copy_(d_coef[idx])
        g1.replay()
        g = g2e if idx < 30 else g2
        g.t.copy_(d_ts[idx].expand(2)); g.coef.copy_(d_coef[idx]); g.frozen.copy_(d_frozen[idx + 1])
        g.replay()
        if idx == STEPS - 1:  # start the next synthetic trajectory from fresh noise
            g1.latents.copy_(h_lat1, non_blocking=True)
            g2e.latents.copy_(h_lat2, non_blocking=True); g2.latents.copy_(h_lat2, non_blocking=True)
        elif idx == 29:
            g2.latents.copy_(g2e.latents)
        return idx < 30

    def step_e2e(i):
        """public-API view: that step's inputs come from pinned HOST memory, results are read back to the host
        and the host waits for them (the reference's per-step latents.cpu(), models/pipelines.py:449-453, 835)"""
        idx = i % STEPS
        g1.t.copy_(h_ts8[idx], non_blocking=True); g1.coef.copy_(h_coef[idx], non_blocking=True)
        if idx == 0:
            g1.set_context(h_ctx1); g1.latents.copy_(h_lat1, non_blocking=True)
            for g in (g2e, g2):
                g.set_context(h_ctx2, h_ctrl if mixed else ctrl_ctx); g.latents.copy_(h_lat2, non_blocking=True)
                g.mask.copy_(h_mask, non_blocking=True)
            d_hint.copy_(h_hint, non_blocking=True)
            embed_hint(d_hint)
        g1.replay()
        g = g2e if idx < 30 else g2
        g.t.copy_(h_ts2[idx], non_blocking=True); g.coef.copy_(h_coef[idx], non_blocking=True)
        g.frozen.copy_(h_frozen[idx + 1], non_blocking=True)
        g.replay()
        if idx == 29:
            g2.latents.copy_(g2e.latents)
        h_out1.copy_(g1.trace, non_blocking=True)
        h_out2.copy_(g.trace, non_blocking=True)
        if idx < 30:
            h_loss.copy_(loss, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n, warm, nvtx=None):
        for i in range(warm):
            fn(i)
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if nvtx:  # lets `ncu --nvtx --nvtx-include "<name>/"` list exactly the launches of the timed region
            torch.cuda.nvtx.range_push(nvtx)
        s.record()
        for i in range(warm, warm + n):
            fn(i)
        e.record()
        if nvtx:
            torch.cuda.nvtx.range_pop()
        barrier()
        ms = s.elapsed_time(e)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms

    Kb, Wb = args.steps, max(3, args.warmup)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # timestep index of loop iteration i: the timed region starts at index 0 of a trajectory
    dev_fn, e2e_fn = (step_device_mixed, step_e2e_mixed) if mixed else (step_device, step_e2e)
    ms_dev = timed(lambda i: dev_fn(i - Wb), Kb, Wb, nvtx="tg_timed")
    clocks = sampler.stop() if rank == 0 else None
    n_energy = sum(1 for i in range(Kb) if (i % STEPS) < 30)
    n_embed = sum(1 for i in range(Kb) if (i % STEPS) == 0)
    gpu_launches = n_energy * launches_per_step[True] + (Kb - n_energy) * launches_per_step[False] + n_embed * embed_launches
    ms_e2e = timed(lambda i: e2e_fn(i - Wb), Kb, Wb)

    # ---- sustained leg: >= 3 s of the same device-resident loop (DESIGN "law 5": dense work is power capped) ----
    sustained = None
    if not args.no_sustained:
        n_sus = 5 * STEPS
        s2 = ClockSampler(local)
        if rank == 0:
            s2.start()
        ms_sus = timed(lambda i: dev_fn(i - Wb), n_sus, Wb)
        c2 = s2.stop() if rank == 0 else None
        sustained = {"value": round((K + 1) * world * n_sus / (ms_sus * 1e-3), 2), "unit": "steps/s", "steps": n_sus,
                     "seconds": round(ms_sus * 1e-3, 3), "ms_per_step": round(ms_sus / n_sus, 4), "clocks": c2}

    # ---- replay time of each captured graph on its own (where the step time goes) ----
    def graph_ms(g, n=10):
        for _ in range(2):
            g.replay()
        s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        s_.record()
        for _ in range(n):
            g.replay()
        e_.record()
        torch.cuda.synchronize()
        return round(s_.elapsed_time(e_) / n, 4)
    graphs = None
    if rank == 0:
        graphs = {"stage1_B8": graph_ms(g1), "stage2_B2_capture_energy": graph_ms(g2e), "stage2_B2": graph_ms(g2)}
        if mixed:
            graphs.update({"cobatched_B10_capture_energy": graph_ms(gme), "cobatched_B10": graph_ms(gm)})

    # ---- roofline leg: every kernel family timed with CUDA events (eager, same shapes, same stream) ----
    roof, breakdown = None, None
    if rank == 0:
        net.ip_scale_per_sample = None
        gp = gme if mixed else g1
        os.environ["TG_CN_OVERLAP"] = "0"  # per-op timing: no second stream sharing the SMs while an op is timed
        for _ in range(2):
            gp._body()
        with ops.profile() as rec:
            for _ in range(3):
                gp._body()
        rec.finish()
        os.environ.pop("TG_CN_OVERLAP", None)
        summ = rec.summary()
        tot = sum(v["ms"] for v in summ.values())
        breakdown = {k: {"launches_per_fwd": v["launches"] // 3, "ms_per_fwd": round(v["ms"] / 3, 4),
                         "share": round(v["ms"] / tot, 4), "tflops": round(v["tflops"], 1), "gbs": round(v["gbs"], 1)}
                     for k, v in sorted(summ.items(), key=lambda kv: -kv[1]["ms"])}
        top = max(summ.items(), key=lambda kv: kv[1]["ms"])
        tf_peak, hbm_peak, src = peaks()
        # DRAM traffic per launch of the same kernel family over the same timed region, from the committed ncu pass
        # (tools/profile.sh -> profiles/gemm_traffic.json); null when the file is absent
        traffic, traffic_src = None, None
        tp = os.path.join(ROOT, "profiles", "gemm_traffic.json")
        if os.path.exists(tp):
            tj = json.load(open(tp))
            traffic, traffic_src = tj.get("mean_dram_bytes_per_launch"), tj.get("source")
        roof = {"kernel": top[0], "bound": "tensor", "achieved": round(top[1]["tflops"], 2), "peak": tf_peak,
                "unit": "TFLOP/s", "frac": round(top[1]["tflops"] / tf_peak, 4), "traffic": traffic,
                "traffic_source": traffic_src,
                "algorithmic_bytes_per_launch": round(top[1]["bytes"] / top[1]["launches"]),
                "peak_source": f"{src} (bf16_tflops_sustained: kernel timed inside a long step)",
                "avg_launch_us": round(1e3 * top[1]["ms"] / top[1]["launches"], 2),
                "algorithmic_flops_per_launch": round(top[1]["flops"] / top[1]["launches"]),
                "launches_timed": top[1]["launches"]}

    # one representative launch of the dominant kernel, timed alone (burst peak applies): the 3x3 conv 320->320 at
    # 64x64 for the co-batched batch of 10 (M = 40960, K = 2880); its DRAM traffic comes from the committed ncu capture
    if rank == 0:
        xb = torch.randn(10, 64, 64, 320, device=dev).half()
        wb = (torch.randn(320, 3, 3, 320, device=dev) * 0.02).half()
        ob = torch.empty(10, 64, 64, 320, device=dev, dtype=torch.float16)
        for _ in range(3):
            ops.conv3x3(xb, wb, out=ob)
        s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        s_.record()
        for _ in range(20):
            ops.conv3x3(xb, wb, out=ob)
        e_.record()
        torch.cuda.synchronize()
        us = s_.elapsed_time(e_) / 20 * 1e3
        fl = 2.0 * 40960 * 320 * 2880
        burst = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("bf16_tflops", 1590.0) \
            if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 1590.0
        roof["representative_launch"] = {
            "what": "conv3x3 320->320 @64x64, B=10 (gemm_tc_kernel<160,3,pair,KY3>: patch-reuse conv stages), timed alone, 20 launches",
            "us": round(us, 2), "achieved": round(fl / us / 1e6, 1), "peak": burst, "unit": "TFLOP/s",
            "frac": round(fl / us / 1e6 / burst, 4), "algorithmic_flops": fl,
            "algorithmic_bytes": 2.0 * (40960 * 320 * 2 + 320 * 2880),
            "traffic": "see profiles/ (dram__bytes_read+write of the captured launches)"}

    # second family by time: the 64x64 self-attention (N = 4096, d = 40, 8 heads, B = 10), timed alone.  Its bound is not the
    # tensor pipe: one exponential per score, 16 MUFU.EX2 per clock and SM (tools/probes/xu_probe.cu), against 4 d = 160 flop
    # per score -- at d = 40 the XU pipe saturates at 26 % tensor-pipe activity (profiles/r02_attention.md).
    if rank == 0:
        Ba, Na, da, ha = 10, 4096, 40, 8
        qkv_a = torch.randn(Ba * Na, 3 * ha * da, device=dev).half()
        oa = torch.empty(Ba * Na, ha * da, device=dev, dtype=torch.float16)
        Ca = ha * da

        def _attn():
            ops.attention_tc(qkv_a[:, :Ca], qkv_a[:, Ca:2 * Ca], qkv_a[:, 2 * Ca:], Ba, ha, da, out=oa, packed=True)
        for _ in range(3):
            _attn()
        s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        s_.record()
        for _ in range(10):
            _attn()
        e_.record()
        torch.cuda.synchronize()
        us = s_.elapsed_time(e_) / 10 * 1e3
        exps = float(Ba) * ha * Na * Na
        sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
        xu_peak = sm_count * 16 * 1.965e9 / 1e12  # Texp/s at the maximum SM clock
        roof["attention_64x64"] = {
            "what": "attn_tc_kernel<1>: self-attention N=4096 d=40 heads=8 B=10 (packed heads), timed alone, 10 launches",
            "us": round(us, 1), "bound": "xu (MUFU.EX2, 16 per clock and SM)", "achieved": round(exps / us / 1e6, 3),
            "peak": round(xu_peak, 3), "unit": "Texp/s", "frac": round(exps / us / 1e6 / xu_peak, 4),
            "tensor_tflops": round(4.0 * exps * da / us / 1e6, 1)}
        del qkv_a, oa

    # Reported beside the main number, never inside it: one iteration of latent_backward_guidance --
    # conditional forward to the last guidance key, energy, d(energy)/d(latents) through the UNet, latent update.  The
    # reference has no caller for it; this is the cost a caller would pay per iteration.
    grad_leg = None
    if rank == 0 and world == 1:
        try:
            from theatergen_b200.backward import GuidanceBackward
            from theatergen_b200.utils.guidance import EnergyPlan
            net.ip_scale_per_sample = None
            gb = GuidanceBackward(net)
            plan_g = EnergyPlan([[(0.1, 0.2, 0.5, 0.8)], [(0.55, 0.3, 0.95, 0.9)]], [[2, 3], [7]],
                                [("mid", 0, 0, 0), ("up", 1, 0, 0), ("up", 1, 1, 0), ("up", 1, 2, 0)])
            lat_g, ctx_g = lat2[:1].to(dev).contiguous(), ctx2[1:2].to(dev).contiguous()
            n0 = _lib.launch_count()
            for it in range(3):
                if it == 1:
                    torch.cuda.synchronize()
                    n0 = _lib.launch_count()
                    s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    s_.record()
                loss_g, grad_g = gb.energy_and_grad(lat_g, 801, ctx_g, plan_g, loss_scale=30.0)
                ops.guidance_update(lat_g.clone(), grad_g, 1.0 / gb.grad_scale, 0.5)
            e_.record()
            torch.cuda.synchronize()
            grad_leg = {"ms_per_iteration_eager": round(s_.elapsed_time(e_) / 2, 3),
                        "launches_per_iteration": (_lib.launch_count() - n0) // 2,
                        "finite": bool(torch.isfinite(grad_g).all()),
                        "what": "B=1 conditional forward to up_blocks.1 + energy + input-gradient backward + update"}
            from theatergen_b200.backward import GuidanceStepGraph
            sg = GuidanceStepGraph(gb, plan_g, ctx_g, 30.0).capture(lat_g, 801)
            sg.replay(lat_g, 801)
            torch.cuda.synchronize()
            s_.record()
            for it in range(5):
                loss_g, grad_g = sg.replay(lat_g, 801)
                ops.guidance_update(lat_g.clone(), grad_g, 1.0 / gb.grad_scale, 0.5)
            e_.record()
            torch.cuda.synchronize()
            grad_leg["ms_per_iteration"] = round(s_.elapsed_time(e_) / 5, 3)  # one CUDA graph per evaluation (the product path)
            del sg
        except Exception as ex:  # the leg is informational; the main line must still print
            grad_leg = {"error": f"{type(ex).__name__}: {ex}"}

    # ---- parity leg (checker): the co-batched graphs that were timed, 50 steps from index 0, vs the oracle loops ----
    parity = None
    if rank == 0 and world == 1 and mixed and not args.no_parity:
        try:
            parity = parity_leg(torch, dev, gme, gm, embed_hint, d_hint, d_ts, d_coef, d_frozen, h_latm, sdu, sdc,
                                ctx1, ctx2, lat1, lat2, ctrl_ctx, mask, bboxes, K, STEPS)
        except Exception as ex:  # informational: the main line must still print
            parity = {"error": f"{type(ex).__name__}: {ex}"}
    eager = None
    if rank == 0 and world == 1 and not args.no_eager_reference:
        try:
            eager = eager_reference_leg(torch, dev, sdu, sdc, ctx1, ctx2, lat1, lat2, ctrl_ctx, d_hint, d_frozen, mask, K)
        except Exception as ex:
            eager = {"error": f"{type(ex).__name__}: {ex}"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        total, cores = cpu_denoise_steps(2, 1)
        cpu = {"value": 2 / total, "unit": "steps/s", "cores": cores, "kind": "port",
               "sample": "2 stage-1 denoise steps of one character (B=2 UNet fwd + CFG + DDIM), fp32 oracle, after 1 warm-up"}

    if rank == 0:
        denoise_per_step = (K + 1) * world
        value = denoise_per_step * Kb / (ms_dev * 1e-3)
        e2e = denoise_per_step * Kb / (ms_e2e * 1e-3)
        h2d = 8 * 4 + 8 * 4 + 4 * 64 * 64 * 2 + (h_ctx1.numel() + h_ctx2.numel() + h_lat1.numel() + h_lat2.numel() + h_mask.numel()
                                                 + ctrl_ctx.numel() + h_hint.numel()) * 2 // STEPS
        d2h = (K + 1) * 4 * 64 * 64 * 2 + 4
        out = {
            "metric": "denoise-steps/sec", "value": round(value, 2), "unit": "steps/s", "n_gpus": world,
            "steps": Kb, "warmup": Wb, "ms_per_step": round(ms_dev / Kb, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f16 (fp32 accumulate)", "data": "synthetic",
            "config": {"workload": WORKLOAD, "denoise_steps_per_bench_step_per_gpu": K + 1,
                       "l2_policy": "inputs larger than L2 (1.76 GB of weights streamed per forward vs 126 MB L2)",
                       "cuda_graphs": True, "cobatch_stage1_stage2": mixed,
                       "controlnet": "computed on our kernels inside the stage-2 forward (conditioning embedder once per trajectory)",
                       "parallelism": f"dp{world} (independent dialogues per rank, weights NCCL-broadcast once)"},
            "e2e": {"value": round(e2e, 2), "unit": "steps/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": round(ms_e2e / Kb, 4)},
            "gpu_launches": int(gpu_launches), "clocks": clocks,
            "tensor_frac_of_step": round(Kb / (ms_dev * 1e-3) * (K * FLOPS_PER_DENOISE_STEP + FLOPS_PER_DENOISE_STEP
                                                                + FLOPS_CONTROLNET_STEP) / 1e12 / peaks()[0], 4),
            "roofline": roof, "cpu_baseline": cpu, "graph_replay_ms": graphs,
            "kernel_breakdown_per_forward": breakdown, "guidance_gradient": grad_leg,
            "parity": parity, "gpu_eager_reference": eager, "sustained": sustained,
        }
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


# =============================================================================================================
# --config dag4: BASELINE.json configs[3] -- the CMIGBench "dialogue 4" DAG (4 turns, 10 character generations, 4 ids)
# sharded over the ranks by theatergen_b200.runner.DagRunner (phases A / B / C, NCCL exchanges in between)
# =============================================================================================================
def run_dag4(args):
    import torch
    import torch.distributed as dist
    from theatergen_b200 import _lib, synth
    from theatergen_b200.controlnet import TGControlNet
    from theatergen_b200.engine import DenoiseEngine
    from theatergen_b200.ip_adapter import IPAdapter
    from theatergen_b200.ip_adapter.ip_adapter import _Pipe
    from theatergen_b200.runner import DagRunner
    from theatergen_b200.scheduler import DDIMScheduler
    from theatergen_b200.unet import TGUNet

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py --config dag4 needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.lib()

    def shared(sd):  # rank 0 generates, NCCL broadcast
        if world > 1:
            flat = torch.cat([v.reshape(-1) for v in sd.values()])
            dist.broadcast(flat, 0)
            off = 0
            for k, v in sd.items():
                sd[k] = flat[off:off + v.numel()].view_as(v)
                off += v.numel()
        return sd

    net = TGUNet(shared(synth.synth_state_dict(dev, 1234, fill=(rank == 0))), dev)
    cnet = TGControlNet(shared(synth.synth_controlnet_state_dict(dev, 4321, fill=(rank == 0))), dev)
    sched = DDIMScheduler()
    adapter = IPAdapter(_Pipe(net, sched), None, None, dev)
    g = torch.Generator().manual_seed(9)
    adapter.image_proj_model.load_state_dict(
        {"proj.weight": torch.randn(3072, 1024, generator=g) / 32.0, "proj.bias": torch.zeros(3072),
         "norm.weight": torch.ones(768), "norm.bias": torch.zeros(768)}, device=dev)
    eng = DenoiseEngine(net, sched, controlnet=cnet)
    dlg = json.load(open(os.path.join(ROOT, "tests", "golden", "cmigbench_dialogue4.json")))["dialogue 4"]
    STEPS = 50
    DagRunner(eng, adapter, cobatch=args.cobatch, num_steps=2).run_dialogue("dialogue 4", dlg, 3)  # graphs, allocator
    runner = DagRunner(eng, adapter, cobatch=args.cobatch, num_steps=STEPS)
    sampler = ClockSampler(local)
    reps = max(1, args.steps // 10)
    if rank == 0:
        sampler.start()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = _lib.launch_count() + eng.launches_replayed()
    s_.record()
    for _ in range(reps):
        s1, s2, stats = runner.run_dialogue("dialogue 4", dlg, 3)
    e_.record()
    torch.cuda.synchronize()
    launches = _lib.launch_count() + eng.launches_replayed() - n0   # this rank's kernels: graph replays + eager launches
    ms = s_.elapsed_time(e_)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    clocks = sampler.stop() if rank == 0 else None
    # per-phase idle fraction: 1 - (denoising time summed over ranks) / (ranks x phase wall time), last repetition
    phases = ["A", "B", "C"]
    mine = torch.tensor([[stats.phase_ms.get(p, 0.0), stats.busy_ms.get(p, 0.0), float(stats.denoise_steps.get(p, 0)),
                          float(stats.padded_steps.get(p, 0))] for p in phases] + [[stats.phase_ms.get("exchange", 0.0), 0, 0, 0]],
                        device=dev)
    allr = [torch.zeros_like(mine) for _ in range(world)]
    if world > 1:
        dist.all_gather(allr, mine)
    else:
        allr = [mine]
    if rank == 0:
        allr = torch.stack(allr).cpu()
        per_phase = {}
        for i, p in enumerate(phases):
            wall = float(allr[:, i, 0].max())
            per_phase[p] = {"wall_ms": round(wall, 2), "busy_ms_per_rank": [round(float(v), 2) for v in allr[:, i, 1]],
                            "idle_fraction": round(1.0 - float(allr[:, i, 1].sum()) / (world * wall), 4) if wall else None,
                            "denoise_steps": int(allr[:, i, 2].sum()), "padding_steps": int(allr[:, i, 3].sum())}
        useful = sum(v["denoise_steps"] for v in per_phase.values())
        value = useful * reps / (ms * 1e-3)
        print(json.dumps({
            "metric": "denoise-steps/sec", "value": round(value, 2), "unit": "steps/s", "n_gpus": world, "steps": reps,
            "warmup": 1, "ms_per_step": round(ms / reps, 3), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f16 (fp32 accumulate)", "data": "synthetic",
            "config": {"workload": "CMIGBench story 'dialogue 4': 4 turns, 10 character generations (4 first appearances at "
                                   "IP scale 0, 6 later appearances at 0.4), 4 stage-2 turns with ControlNet + masked composite, "
                                   "50 DDIM steps each = 700 denoise steps; one 'step' = the whole dialogue",
                       "partition": f"DAG phases A/B/C over {world} rank(s), item i of a phase on rank i % N, co-batches of "
                                    f"exactly {args.cobatch} stage-1 trajectories (padding counted as idle), NCCL all-gather of "
                                    "image tokens (A->B) and of stage-1 traces + masks (B->C)",
                       "stand_ins": "text / CLIP-image embeddings, SAM masks and the line-art hint are seeded stand-ins "
                                    "(theatergen_b200.runner.SyntheticModels)"},
            "e2e": {"value": round(value, 2), "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                    "note": "the runner is the public API of this path; its inputs (seeded latents, embeddings) are built on "
                            "the host and copied inside the timed region"},
            "gpu_launches": int(launches), "clocks": clocks,
            "dag": {"phases": per_phase, "exchange_ms": round(float(allr[:, 3, 0].max()), 2),
                    "useful_denoise_steps_per_dialogue": useful}}))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sustained", action="store_true", help="skip the >= 3 s sustained leg")
    ap.add_argument("--no-parity", action="store_true", help="skip the 50-step oracle parity leg (~40 s)")
    ap.add_argument("--no-eager-reference", action="store_true", help="skip the same-GPU PyTorch eager leg")
    ap.add_argument("--separate-stage2", action="store_true", help="stage-1 (B=8) and stage-2 (B=2) as separate forwards")
    ap.add_argument("--config", default="turn", choices=["turn", "dag4"],
                    help="turn (default): the co-batched 4-character turn step of the headline metric; dag4: the CMIGBench "
                         "dialogue-4 DAG sharded over the ranks (BASELINE.json configs[3])")
    ap.add_argument("--cobatch", type=int, default=2, help="dag4: stage-1 trajectories per forward (fixed group size)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.config == "dag4":
        run_dag4(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
