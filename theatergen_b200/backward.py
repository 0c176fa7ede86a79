"""d(energy)/d(latents) through the conditional UNet forward -- the gradient of `latent_backward_guidance`
(models/pipelines.py:62-128: forward with `save_attn_to_dict`, `compute_ca_lossv3 * loss_scale`,
`torch.autograd.grad(loss, [latents])`, `latents -= grad * sqrt(1 - alpha_prod_t)`).  SURVEY.md row a16.

What is differentiated: the energy reads the text cross-attention probabilities of the four guidance keys
(`("mid",0,0,0)`, `("up",1,{0,1,2},0)`, pipelines.py:21), so the forward is cut after the cross-attention of
`up_blocks.1.attentions.2` -- conv_in, the four down blocks, the mid block, up block 0 and up block 1 -- and the backward
walks the same operators in reverse.  The weights are frozen: every GEMM / convolution backward is the forward tensor-core
kernel on a transposed / 180-degree-rotated weight copy (built lazily, once), and the remaining operator backwards are the
kernels of csrc/backward.cu.  One image (the reference runs the conditional half alone, B = 1), fp16 tensors with fp32
arithmetic inside each kernel like the reference's fp16 autograd; gradients are carried at `grad_scale` times their value
against fp16 underflow and unscaled in fp32 by the update kernel.

Self-attention here is materialised per head (S = Q K^T GEMM -> row softmax -> P V GEMM, P kept for the backward; 32 MB per
head at 64x64): the flash kernels of the product path have no backward, and this path runs a handful of times per image
(the reference's loop: at most 5 iterations at the first 10 of 50 steps), not per denoise step.

The recording is a plain tape: each forward operator appends (output, function mapping d(output) -> [(input, d(input))]).
"""
from __future__ import annotations

import torch

from . import ops
from .unet import BLOCK_OUT, HEADS, IPAttnProcessorHandle
from .weights import pack_conv3x3_dgrad

GUIDANCE_KEYS = (("mid", 0, 0, 0), ("up", 1, 0, 0), ("up", 1, 1, 0), ("up", 1, 2, 0))


def _pad64(n):
    return (n + 63) // 64 * 64


class _Tape:
    def __init__(self):
        self.nodes = []  # (output tensor, backward fn, run-without-upstream-gradient)

    def add(self, out, fn, always=False):
        self.nodes.append((out, fn, always))
        return out

    def backward(self, seeds=None):
        """-> {data_ptr: gradient} of the tensors no node produced (the leaves)."""
        grads, owned = dict(seeds or {}), set()
        for out, fn, always in reversed(self.nodes):
            g = grads.pop(out.data_ptr(), None)
            if g is None and not always:
                continue
            for inp, gi in fn(None if g is None else g.view(out.shape)):
                k = inp if isinstance(inp, int) else inp.data_ptr()
                if k not in grads:
                    grads[k] = gi
                else:  # a second consumer: accumulate (never in place on a tensor another entry may alias)
                    acc = grads[k] if k in owned else grads[k].clone()
                    grads[k] = ops.add_inplace(acc.view(-1), gi.contiguous().view(-1)).view(gi.shape)
                    owned.add(k)
        return grads


class GuidanceBackward:
    def __init__(self, unet, grad_scale=256.0):
        self.unet = unet
        self.grad_scale = float(grad_scale)
        self._cache = {}

    # ---- derived weight copies (built once) --------------------------------------------------------------
    def _T(self, w):
        """[N,K] linear weight -> its transpose as a dense [K,N] GEMM weight (dx = dy @ W)."""
        k = ("T", w.data_ptr())
        if k not in self._cache:
            self._cache[k] = w.t().contiguous()
        return self._cache[k]

    def _R(self, w):
        """packed conv3x3 weight [Cout,3,3,Cin] -> packed input-gradient weight [Cin,3,3,Cout], taps rotated."""
        k = ("R", w.data_ptr())
        if k not in self._cache:
            self._cache[k] = pack_conv3x3_dgrad(w.permute(0, 3, 1, 2))
        return self._cache[k]

    def _attn_weights(self, t):
        """q/k/v rows and the out-projection columns of a self-attention, every head padded to a multiple of 64."""
        k = ("A", t["qkvw"].data_ptr())
        if k not in self._cache:
            C = t["C"]
            d = C // HEADS
            P = _pad64(d)
            src = t["qkvw"].view(3, HEADS, t["qkv_pad"] or d, C)[:, :, :d]
            wq = torch.zeros((3, HEADS, P, C), dtype=torch.float16, device=src.device)
            wq[:, :, :d] = src
            wo = torch.zeros((C, HEADS, P), dtype=torch.float16, device=src.device)
            wo[:, :, :d] = t["o1w"].view(C, HEADS, d)
            self._cache[k] = (wq.view(3 * HEADS * P, C), wo.view(C, HEADS * P), P)
        return self._cache[k]

    # ---- operators with their backward recorded ----------------------------------------------------------
    @staticmethod
    def _conv3x3(x, w, bias=None, rowbias=None, residual=None):
        """The implicit-GEMM conv tiles 128 pixels: an 8x8 image is run as a batch of two with a zero partner."""
        if x.shape[0] * x.shape[1] * x.shape[2] >= 128:
            return ops.conv3x3(x, w, bias, rowbias=rowbias, residual=residual)
        z = torch.zeros_like(x)
        y = ops.conv3x3(torch.cat([x, z]), w, bias, rowbias=None if rowbias is None else torch.cat([rowbias, rowbias]),
                        residual=None if residual is None else torch.cat([residual, torch.zeros_like(residual)]))
        return y[:1].contiguous()

    def _gn(self, tape, x, g, b, eps, silu, H, W):
        C = x.shape[-1]
        y = ops.groupnorm(x.view(1, H * W, C), g, b, eps, silu=silu).view(1, H, W, C)
        tape.add(y, lambda d: [(x, ops.groupnorm_bwd(x.view(1, H * W, C), g, b, d.contiguous().view(1, H * W, C), eps,
                                                    silu).view(x.shape))])
        return y

    def _ln(self, tape, x, g, b):
        y = ops.layernorm(x, g, b)
        tape.add(y, lambda d: [(x, ops.layernorm_bwd(x, g, d))])
        return y

    def _linear(self, tape, x, w, bias=None, residual=None):
        y = ops.gemm(x, w, bias, residual=residual)

        def bwd(d):
            out = [(x, ops.gemm(d, self._T(w)))]
            if residual is not None:
                out.append((residual, d))
            return out
        tape.add(y, bwd)
        return y

    def _conv(self, tape, x, w, bias, rowbias=None, residual=None):
        y = self._conv3x3(x, w, bias, rowbias, residual)

        def bwd(d):
            out = [(x, self._conv3x3(d.contiguous(), self._R(w)))]
            if residual is not None:
                out.append((residual, d))
            return out
        tape.add(y, bwd)
        return y

    def _resnet(self, tape, p, x, skip, temb_all, H, W):
        r = self.unet.resnets[p]
        cin, co = r["cin"], r["cout"]
        if skip is not None:
            c1 = x.shape[-1]
            xin = torch.cat([x, skip], dim=-1)
            tape.add(xin, lambda d: [(x, d[..., :c1].contiguous()), (skip, d[..., c1:].contiguous())])
        else:
            xin = x
        n1 = self._gn(tape, xin, r["n1g"], r["n1b"], 1e-5, True, H, W)
        h = self._conv(tape, n1, r["c1w"], r["c1b"], rowbias=temb_all[:, r["temb_off"]:r["temb_off"] + co])
        n2 = self._gn(tape, h, r["n2g"], r["n2b"], 1e-5, True, H, W)
        if "scw" in r:
            sc = self._linear(tape, xin.view(H * W, cin), r["scw"], r["scb"]).view(1, H, W, co)
        else:
            sc = xin
        return self._conv(tape, n2, r["c2w"], r["c2b"], residual=sc)

    def _self_attention(self, tape, t, ln1, M):
        C = t["C"]
        d = C // HEADS
        wq, wo, P = self._attn_weights(t)
        scale = d ** -0.5
        qkv = self._linear(tape, ln1, wq)                                   # [M, 3*8*P]
        hm = qkv.view(M, 3, HEADS, P).permute(1, 2, 0, 3).contiguous()      # [3, 8, M, P]   head-major
        hmT = hm.transpose(2, 3).contiguous()                               # [3, 8, P, M]
        a = torch.empty((M, HEADS * P), dtype=torch.float16, device=ln1.device)
        probs = []
        for h in range(HEADS):
            s = ops.gemm(hm[0, h], hm[1, h])                                # Q_h K_h^T
            pm = ops.softmax_rows(s, scale, out=s)
            ops.gemm(pm, hmT[2, h], out=a[:, h * P:(h + 1) * P])            # P V_h
            probs.append(pm)

        def bwd(da):                                                        # [M, 8*P]
            dhm = da.view(M, HEADS, P).permute(1, 0, 2).contiguous()        # [8, M, P]
            dhmT = dhm.transpose(1, 2).contiguous()                         # [8, P, M]
            dq = torch.empty((3, HEADS, M, P), dtype=torch.float16, device=da.device)
            for h in range(HEADS):
                pm = probs[h]
                ds = ops.softmax_bwd_rows(pm, ops.gemm(dhm[h], hm[2, h]), scale)       # dP = dO V^T -> dS
                ops.gemm(ds, hmT[1, h], out=dq[0, h])                                   # dQ = dS K
                ops.gemm(ops.nchw_to_nhwc(ds.view(1, M, M, 1)).view(M, M), hmT[0, h], out=dq[1, h])   # dK = dS^T Q
                ops.gemm(ops.nchw_to_nhwc(pm.view(1, M, M, 1)).view(M, M), dhmT[h], out=dq[2, h])     # dV = P^T dO
            return [(qkv, dq.permute(2, 0, 1, 3).reshape(M, 3 * HEADS * P))]
        tape.add(a, bwd)
        return a, wo

    def _transformer(self, tape, p, x, ctxkv, H, W, n_text, n_ip, ip_scale, ip_scale_dev, maps, dmaps, last=False):
        t = self.unet.transformers[p]
        C, M = t["C"], H * W
        n = self._gn(tape, x, t["ng"], t["nb"], 1e-6, False, H, W)
        h0 = self._linear(tape, n.view(M, C), t["piw"], t["pib"])
        ln1 = self._ln(tape, h0, t["norm1g"], t["norm1b"])
        a, wo = self._self_attention(tape, t, ln1, M)
        h1 = self._linear(tape, a, wo, t["o1b"], residual=h0)
        ln2 = self._ln(tape, h1, t["norm2g"], t["norm2b"])
        q = self._linear(tape, ln2, t["q2w"])
        o = t["kv_off"]
        kt, vt = ctxkv[:, o:o + C], ctxkv[:, o + C:o + 2 * C]
        ki, vi = (ctxkv[:, o + 2 * C:o + 3 * C], ctxkv[:, o + 3 * C:o + 4 * C]) if n_ip else (None, None)
        key = t["key"]
        probs = None
        if key in dmaps:
            probs = torch.empty((1, HEADS, M, n_text), dtype=torch.float16, device=x.device)
        a2 = ops.attention_cross_ip(q, kt, vt, ki, vi, 1, HEADS, n_text, n_ip, ip_scale=ip_scale_dev if n_ip else None,
                                    probs=probs)
        if probs is not None:
            maps[key] = probs[0]

        def cross_bwd(d):
            if d is None:
                d = torch.zeros_like(q)
            return [(q, ops.attention_cross_ip_bwd_q(q, kt, vt, ki, vi, HEADS, n_text, n_ip, ip_scale, d, dmaps.get(key)))]
        tape.add(a2, cross_bwd, always=probs is not None)
        if last:
            return None
        h2 = self._linear(tape, a2, t["o2w"], t["o2b"], residual=h1)
        ln3 = self._ln(tape, h2, t["norm3g"], t["norm3b"])
        ff = ops.gemm(ln3, t["ff1w"], t["ff1b"], act=ops.TG_ACT_GEGLU)

        def ff_bwd(d):  # the packed pre-activation is recomputed instead of kept
            pre = ops.gemm(ln3, t["ff1w"], t["ff1b"])
            return [(ln3, ops.gemm(ops.geglu_bwd(pre, d.contiguous()), self._T(t["ff1w"])))]
        tape.add(ff, ff_bwd)
        h3 = self._linear(tape, ff, t["ff2w"], t["ff2b"], residual=h2)
        return self._linear(tape, h3, t["pow"], t["pob"], residual=x.view(M, C)).view(1, H, W, C)

    # ---- forward to the last guidance key, energy, backward ----------------------------------------------
    def _ip_scale(self):
        u = self.unet
        if not u.has_ip:
            return 0.0
        return float(next(p.scale for p in u.attn_processors.values() if isinstance(p, IPAttnProcessorHandle)))

    def forward(self, sample, t_dev, ctx, dmaps):
        """sample [1,4,64,64], t_dev float32 [1], ctx [1,T,768] -> (tape, {key: probabilities [8,HW,n_text]}).
        `dmaps` {key: gradient buffer} names the captured keys; its tensors are read when the tape runs backward."""
        u, w = self.unet, self.unet.w
        assert sample.shape[0] == 1 and sample.is_contiguous() and ctx.shape[0] == 1
        _, _, H, W = sample.shape
        n_ip = u.num_ip_tokens if u.has_ip else 0
        n_text = ctx.shape[1] - n_ip
        ip_scale = self._ip_scale()
        ip_dev = torch.full((1,), ip_scale, dtype=torch.float32, device=sample.device)
        temb_all, ctxkv = u._time_and_context(t_dev, ctx, 1)
        tape, maps = _Tape(), {}

        def tf(p, x, h_, w_, last=False):
            return self._transformer(tape, p, x, ctxkv, h_, w_, n_text, n_ip, ip_scale, ip_dev, maps, dmaps, last)

        x = ops.conv_in_gemm(sample, w["conv_in.w"], w["conv_in.b"])
        w_in = w["conv_in.w"][:, :36].reshape(320, 4, 3, 3)
        zero4 = torch.zeros((4,), dtype=torch.float16, device=sample.device)
        tape.add(x, lambda d: [(sample, ops.conv_out(d.contiguous(), pack_conv3x3_dgrad(w_in), zero4))])
        skips = [x]
        h_, w_ = H, W
        for i in range(4):
            for j in range(2):
                x = self._resnet(tape, f"down_blocks.{i}.resnets.{j}", x, None, temb_all, h_, w_)
                if i < 3:
                    x = tf(f"down_blocks.{i}.attentions.{j}", x, h_, w_)
                skips.append(x)
            if i < 3:
                c = BLOCK_OUT[i]
                src, hh, ww = x, h_, w_
                y = ops.gemm(ops.im2col_s2(src), w[f"down{i}.w"], w[f"down{i}.b"]).view(1, hh // 2, ww // 2, c)
                tape.add(y, lambda d, src=src, hh=hh, ww=ww, c=c, i=i: [
                    (src, ops.col2im_s2(ops.gemm(d.contiguous().view(-1, c), self._T(w[f"down{i}.w"])), 1, hh, ww, c))])
                x = y
                h_, w_ = h_ // 2, w_ // 2
                skips.append(x)
        x = self._resnet(tape, "mid_block.resnets.0", x, None, temb_all, h_, w_)
        x = tf("mid_block.attentions.0", x, h_, w_)
        x = self._resnet(tape, "mid_block.resnets.1", x, None, temb_all, h_, w_)
        for j in range(3):
            x = self._resnet(tape, f"up_blocks.0.resnets.{j}", x, skips.pop(), temb_all, h_, w_)
        src = x
        up = ops.upsample2x(src)
        tape.add(up, lambda d: [(src, ops.upsample2x_bwd(d.contiguous()))])
        h_, w_ = 2 * h_, 2 * w_
        x = self._conv(tape, up, w["up0.w"], w["up0.b"])
        for j in range(3):
            x = self._resnet(tape, f"up_blocks.1.resnets.{j}", x, skips.pop(), temb_all, h_, w_)
            x = tf(f"up_blocks.1.attentions.{j}", x, h_, w_, last=(j == 2))
        return tape, maps

    def energy_and_grad(self, latents, t, ctx, plan, loss_scale=1.0, ref_maps=None):
        """-> (energy * loss_scale as a float32 device scalar, d(that)/d(latents) * grad_scale as fp16 [1,4,H,W]).
        `plan` is a utils.guidance.EnergyPlan over keys drawn from GUIDANCE_KEYS."""
        dev = latents.device
        sample = latents.to(torch.float16).contiguous()
        t_dev = t if (torch.is_tensor(t) and t.is_cuda and t.dtype == torch.float32) else \
            torch.full((1,), float(t), dtype=torch.float32, device=dev)
        for k in plan.keys:
            if k not in GUIDANCE_KEYS:
                raise NotImplementedError(f"guidance key {k} lies beyond up block 1; the recorded forward stops there")
        hw = {("mid", 0, 0, 0): (sample.shape[2] // 8) * (sample.shape[3] // 8)}
        n_ip = self.unet.num_ip_tokens if self.unet.has_ip else 0
        n_text = ctx.shape[1] - n_ip
        dmaps = {k: torch.zeros((HEADS, hw.get(k, (sample.shape[2] // 4) * (sample.shape[3] // 4)), n_text),
                                dtype=torch.float16, device=dev) for k in plan.keys}
        tape, maps = self.forward(sample, t_dev, ctx.to(torch.float16).contiguous(), dmaps)
        loss = torch.zeros((1,), dtype=torch.float32, device=dev)
        ws = torch.empty((256,), dtype=torch.float32, device=dev)
        plan.run([maps[k] for k in plan.keys], loss, ws, dmaps=[dmaps[k] for k in plan.keys], ref_maps=ref_maps,
                 grad_scale=loss_scale * self.grad_scale)
        grads = tape.backward()
        return loss[0] * loss_scale, grads[sample.data_ptr()]


class GuidanceStepGraph:
    """One `energy_and_grad` evaluation as a CUDA graph (about 1300 kernel launches of a B = 1 forward + backward are
    launch-bound when enqueued from Python).  The plan's descriptor arrays are baked in by value; latents, timestep and
    context are static buffers the caller fills before `replay()`."""

    def __init__(self, gb: GuidanceBackward, plan, ctx, loss_scale=1.0, latent_shape=(1, 4, 64, 64)):
        dev = ctx.device
        self.gb, self.plan, self.loss_scale = gb, plan, float(loss_scale)
        self.latents = torch.zeros(latent_shape, dtype=torch.float16, device=dev)
        self.t = torch.zeros((1,), dtype=torch.float32, device=dev)
        self.ctx = ctx.to(torch.float16).contiguous().clone()
        self.graph = None

    def capture(self, latents, t):
        self.latents.copy_(latents)
        self.t.fill_(float(t))
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):  # warm-up: builds the transposed / rotated weight copies outside the capture
            self.gb.energy_and_grad(self.latents, self.t, self.ctx, self.plan, self.loss_scale)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss, self.grad = self.gb.energy_and_grad(self.latents, self.t, self.ctx, self.plan, self.loss_scale)
        return self

    def replay(self, latents, t):
        self.latents.copy_(latents)
        self.t.fill_(float(t))
        self.graph.replay()
        return self.loss, self.grad
