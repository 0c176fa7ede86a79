"""DenoiseEngine -- the per-timestep hot loop on one GPU: CFG-batched UNet forward -> fused CFG + DDIM
(+ masked composite) -> trace, with K co-batched characters (batch 2K, per-sample IP scale) and the whole step
replayed as a CUDA graph.

Mirrors the two loops of the reference: models/pipelines.py:406-453 (stage 1) and :742-835 (stage 2: ControlNet
(TGControlNet, fused into the UNet forward) -> UNet -> CFG -> DDIM -> masked composite).
The reference's per-step `latents.cpu()` (pipelines.py:449-453, 835) becomes an asynchronous device->pinned-host
copy that is only waited for at the end of the trajectory.
"""
from __future__ import annotations

import os

import torch

from . import ops
from .scheduler import DDIMScheduler
from .unet import TGUNet


class StepGraph:
    """One denoise step for a fixed (K, frozen?) shape, captured once and replayed with new t / coefficients."""

    def __init__(self, engine, K, T, composite, capture_keys=None, controlnet=False):
        dev = engine.device
        self.K = K
        self.use_controlnet = bool(controlnet)
        if self.use_controlnet:
            if engine.controlnet is None:
                raise RuntimeError("StepGraph(controlnet=True) needs DenoiseEngine.controlnet (a TGControlNet)")
            self.ctrl_ctx = torch.zeros((2 * K, 77, 768), dtype=torch.float16, device=dev)   # text-only embeddings
            self.cond_emb = torch.zeros((K, 64, 64, 320), dtype=torch.float16, device=dev)   # embed_condition() output
        self.latents = torch.zeros((K, 4, 64, 64), dtype=torch.float16, device=dev)
        self.x_in = torch.zeros((2 * K, 4, 64, 64), dtype=torch.float16, device=dev)
        self.ctx = torch.zeros((2 * K, T, 768), dtype=torch.float16, device=dev)
        self.ip_scale = torch.zeros((2 * K,), dtype=torch.float32, device=dev)
        self.t = torch.zeros((2 * K,), dtype=torch.float32, device=dev)
        self.coef = torch.zeros((8,), dtype=torch.float32, device=dev)
        self.trace = torch.zeros((K, 4, 64, 64), dtype=torch.float16, device=dev)
        self.frozen = torch.zeros((K, 4, 64, 64), dtype=torch.float16, device=dev) if composite else None
        self.mask = torch.zeros((64, 64), dtype=torch.float16, device=dev) if composite else None
        self.capture_keys = capture_keys
        self.captured = {} if capture_keys else None
        self.engine = engine
        # context K / V projections of every cross-attention layer: a function of `ctx` / `ctrl_ctx` only, refreshed by
        # set_context() once per trajectory and read by every replayed step
        self.ctxkv = torch.zeros((2 * K * T, engine.unet.w["ctxkv.w"].shape[0]), dtype=torch.float16, device=dev)
        self.ctrl_ctxkv = None
        if self.use_controlnet:
            self.ctrl_ctxkv = torch.zeros((2 * K * 77, engine.controlnet.w["ctxkv.w"].shape[0]), dtype=torch.float16, device=dev)
        self.graph = None
        self.launches_per_step = 0
        self.replays = 0
        # CFG prefix sharing (unet._BlockNet._encoder): batch row -> trajectory, trajectory -> one of its batch rows
        self.share_prefix = os.environ.get("TG_CFG_SHARE", "1") != "0"
        ar = torch.arange(K, device=dev)
        self.src, self.first = torch.cat([ar, ar]), ar.clone()
        self._ctrl_share = None

    def _control(self, r0, r1):
        if not self.use_controlnet:
            return None, None
        c = dict(rows=(r0, r1), ctx=self.ctrl_ctx, cond_emb=self.cond_emb, ctxkv=self.ctrl_ctxkv)
        if self.share_prefix:
            if self._ctrl_share is None:  # ControlNet rows [uncond k | cond k] -> its k unique rows (built before capture)
                ar = torch.arange((r1 - r0) // 2, device=self.latents.device)
                self._ctrl_share = (torch.cat([ar, ar]), ar)
            c["share"] = self._ctrl_share
        return self.engine.controlnet, c

    def set_context(self, ctx=None, ctrl_ctx=None):
        """Copy the trajectory's context into the graph's buffers (None: already there) and project it to the K / V of every
        cross-attention layer -- once per trajectory, outside the replayed step."""
        if ctx is not None:
            self.ctx.copy_(ctx, non_blocking=True)
        self.engine.unet.project_context(self.ctx, out=self.ctxkv)
        if self.use_controlnet:
            if ctrl_ctx is not None:
                self.ctrl_ctx.copy_(ctrl_ctx, non_blocking=True)
            self.engine.controlnet.project_context(self.ctrl_ctx, out=self.ctrl_ctxkv)

    def _body(self):
        net = self.engine.unet
        ops.duplicate_latents(self.latents, self.x_in)
        cn, control = self._control(0, 2 * self.K)
        eps = net.forward_raw(self.x_in, self.t, self.ctx, ip_scale=self.ip_scale, capture=self.captured,
                              save_keys=self.capture_keys, cond_only=True, controlnet=cn, control=control, ctxkv=self.ctxkv,
                              unique=(self.latents, self.src, self.first) if self.share_prefix else None)
        ops.cfg_ddim_composite(eps, self.latents, self.coef, frozen=self.frozen, mask=self.mask,
                               latents_copy=self.trace)

    def capture(self):
        from . import _lib
        self.set_context()  # the per-image operands of the folded cross-attention layers must exist before the capture
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):  # warm up: lazy kernel attribute setup, allocator pools
                self._body()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        n0 = _lib.launch_count()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._body()
        self.launches_per_step = _lib.launch_count() - n0
        return self

    def replay(self):
        self.graph.replay()
        self.replays += 1


class MixedStepGraph(StepGraph):
    """One denoise step for a co-batch of K1 stage-1 trajectories (plain CFG + DDIM) and K2 stage-2 trajectories
    (ControlNet residuals, masked composite, optional capture of the guidance maps + energy): independent trajectories
    of different turns / dialogues that happen to sit at the same timestep index share ONE UNet forward of batch
    2*(K1+K2) with per-sample IP scale.  Batch rows: [stage-1 uncond K1 | stage-1 cond K1 | stage-2 uncond K2 |
    stage-2 cond K2] -- `ctx`, `ip_scale`, `t` follow that order; the stage-2 rows are contiguous so the ControlNet
    (batch 2*K2) adds its residuals into a row range of the skip tensors.  `latents` / `trace` are [K1 + K2].
    `energy_plan` (utils.guidance.EnergyPlan) is evaluated on the first stage-2 trajectory."""

    def __init__(self, engine, K1, K2, T, capture_keys=None, energy_plan=None, controlnet=False):
        super().__init__(engine, K1 + K2, T, False, capture_keys, controlnet=False)
        dev = engine.device
        self.K1, self.K2 = K1, K2
        self.frozen = torch.zeros((K2, 4, 64, 64), dtype=torch.float16, device=dev)
        self.mask = torch.zeros((64, 64), dtype=torch.float16, device=dev)
        self.energy_plan = energy_plan
        self.loss = torch.zeros((1,), dtype=torch.float32, device=dev)
        self._ws = torch.empty((256,), dtype=torch.float32, device=dev)
        self.use_controlnet = bool(controlnet)
        if self.use_controlnet:
            if engine.controlnet is None:
                raise RuntimeError("MixedStepGraph(controlnet=True) needs DenoiseEngine.controlnet (a TGControlNet)")
            self.ctrl_ctx = torch.zeros((2 * K2, 77, 768), dtype=torch.float16, device=dev)
            self.cond_emb = torch.zeros((K2, 64, 64, 320), dtype=torch.float16, device=dev)
            self.ctrl_ctxkv = torch.zeros((2 * K2 * 77, engine.controlnet.w["ctxkv.w"].shape[0]), dtype=torch.float16, device=dev)
        # rows [stage-1 uncond K1 | stage-1 cond K1 | stage-2 uncond K2 | stage-2 cond K2] -> trajectory index into `latents`
        a1, a2 = torch.arange(K1, device=dev), K1 + torch.arange(K2, device=dev)
        self.src = torch.cat([a1, a1, a2, a2])
        self.first = torch.cat([a1, 2 * K1 + torch.arange(K2, device=dev)])

    def _body(self):
        net = self.engine.unet
        K1, K2 = self.K1, self.K2
        ops.duplicate_latents(self.latents[:K1], self.x_in[:2 * K1])
        ops.duplicate_latents(self.latents[K1:], self.x_in[2 * K1:])
        cn, control = self._control(2 * K1, 2 * (K1 + K2))
        eps = net.forward_raw(self.x_in, self.t, self.ctx, ip_scale=self.ip_scale, capture=self.captured,
                              save_keys=self.capture_keys, capture_b0=2 * K1 + K2, controlnet=cn, control=control,
                              ctxkv=self.ctxkv, unique=(self.latents, self.src, self.first) if self.share_prefix else None)
        if self.energy_plan is not None:
            self.energy_plan.run([self.captured[k][0] for k in self.capture_keys], self.loss, self._ws)
        ops.cfg_ddim_composite(eps[:2 * K1], self.latents[:K1], self.coef, latents_copy=self.trace[:K1])
        ops.cfg_ddim_composite(eps[2 * K1:], self.latents[K1:], self.coef, frozen=self.frozen, mask=self.mask,
                               latents_copy=self.trace[K1:])


class DenoiseEngine:
    def __init__(self, unet: TGUNet, scheduler: DDIMScheduler | None = None, use_graphs=True, controlnet=None):
        self.unet = unet
        self.controlnet = controlnet  # optional TGControlNet (stage 2)
        self.device = unet.device
        self.scheduler = scheduler or DDIMScheduler()
        self.use_graphs = use_graphs
        self._graphs = {}
        self._tables = {}

    def launches_replayed(self):
        """kernels launched through graph replays so far (tg_launch_count only sees eager launches and captures)"""
        return sum(g.launches_per_step * g.replays for g in self._graphs.values())

    def tables(self, num_steps, guidance_scale):
        key = (num_steps, float(guidance_scale))
        if key not in self._tables:
            self.scheduler.set_timesteps(num_steps)
            coef = self.scheduler.coef_table(guidance_scale, self.device)
            ts = self.scheduler.timesteps.to(self.device, torch.float32)
            self._tables[key] = (ts, coef)
        return self._tables[key]

    def step_graph(self, K, T, composite, capture_keys=None, controlnet=False):
        key = (K, T, composite, tuple(capture_keys) if capture_keys else None, bool(controlnet))
        if key not in self._graphs:
            sg = StepGraph(self, K, T, composite, capture_keys, controlnet=controlnet)
            if self.use_graphs:
                sg.capture()
            self._graphs[key] = sg
        return self._graphs[key]

    @torch.no_grad()
    def denoise(self, latents, ctx, ip_scale, num_steps=50, guidance_scale=7.5, frozen_all=None, frozen_mask=None,
                frozen_steps=0, trace_host=None, capture_keys=None, on_step=None, control_ctx=None, cond_emb=None):
        """latents [K,4,64,64], ctx [2K,T,768] (rows [0,K) unconditional), ip_scale float[2K] or float.
        frozen_all [steps+1,K,4,64,64] + frozen_mask [64,64] enable the stage-2 composite for index < frozen_steps.
        trace_host: optional pinned fp16 [steps+1,K,4,64,64] receiving every step's latents asynchronously.
        control_ctx [2K,77,768] (text-only embeddings) + cond_emb [K or 1,64,64,320] (TGControlNet.embed_condition
        of the hint image) switch on the ControlNet of stage 2 (needs self.controlnet).
        Returns the final latents (device)."""
        K, T = latents.shape[0], ctx.shape[1]
        ts, coef = self.tables(num_steps, guidance_scale)
        composite = frozen_all is not None and frozen_steps > 0
        use_cn = control_ctx is not None
        sg = self.step_graph(K, T, composite, capture_keys, controlnet=use_cn)
        sg.latents.copy_(latents, non_blocking=True)
        sg.set_context(ctx, control_ctx if use_cn else None)
        if use_cn:
            sg.cond_emb.copy_(cond_emb.expand_as(sg.cond_emb), non_blocking=True)
        if torch.is_tensor(ip_scale):
            sg.ip_scale.copy_(ip_scale.to(torch.float32), non_blocking=True)
        else:
            sg.ip_scale.fill_(float(ip_scale))
        if composite:
            sg.mask.copy_(frozen_mask.to(torch.float16), non_blocking=True)
        if trace_host is not None:
            trace_host[0].copy_(sg.latents, non_blocking=True)
        sg_plain = None
        for i in range(num_steps):
            cur = sg
            if composite and i >= frozen_steps:  # composite only for index < frozen_steps (pipelines.py:833)
                if sg_plain is None:
                    sg_plain = self.step_graph(K, T, False, capture_keys, controlnet=use_cn)
                    if use_cn:
                        sg_plain.cond_emb.copy_(sg.cond_emb)
                    sg_plain.set_context(sg.ctx, sg.ctrl_ctx if use_cn else None)
                    sg_plain.ip_scale.copy_(sg.ip_scale)
                    sg_plain.latents.copy_(sg.latents)
                cur = sg_plain
            cur.t.copy_(ts[i].expand(2 * K), non_blocking=True)
            cur.coef.copy_(coef[i], non_blocking=True)
            if cur.frozen is not None:
                cur.frozen.copy_(frozen_all[i + 1], non_blocking=True)
            if self.use_graphs:
                cur.replay()
            else:
                cur._body()
            if trace_host is not None:
                trace_host[i + 1].copy_(cur.trace, non_blocking=True)
            if on_step is not None:
                on_step(i, cur)
        # the step graph owns its `latents` buffer and the next call with the same shape overwrites it: hand out a copy
        return (sg_plain or sg).latents.clone()
