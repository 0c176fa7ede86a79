"""TGUNet -- the SD-1.5 UNet forward executed entirely by libtheatergen.so kernels (module-level drop-in for
`adapter.pipe.unet`, called exactly like the reference calls the diffusers module:
models/pipelines.py:428 (stage 1), :802-816 (stage 2, with ControlNet residuals), :90 (guidance, with
save_attn_to_dict / save_keys in cross_attention_kwargs)).

Data layout in HBM: activations are channels-last fp16 -- images [B,H,W,C], tokens [B*H*W, C] (the same bytes,
so the reference's NCHW<->NHWC permutes around every Transformer2DModel disappear); weights are repacked once:
  conv3x3 [Cout,3,3,Cin]; q/k/v of self-attention fused [3C,C]; to_k/to_v/to_k_ip/to_v_ip of all 16
  cross-attentions fused into one [sum 4C, 768] matrix (one GEMM per forward projects the context for every
  layer); the 22 time_emb_proj fused into one [20160,1280] matrix; GEGLU rows interleaved per 64.
The wiring follows ip_adapter/unet_2d_condition.py:799-1125 and the block files cited in oracle/unet.py.
"""
from __future__ import annotations

import os

import torch

from . import _lib, ops
from .weights import pack_conv3x3, pack_geglu, pack_upsample_conv3x3

BLOCK_OUT = (320, 640, 1280, 1280)
HEADS = 8
# self-attention Q/K/V with heads packed at stride d for d = 40 / 80 too (TG_QKV_PACKED=0: zero-padded to 64 / 128 columns)
QKV_PACKED = os.environ.get("TG_QKV_PACKED", "1") != "0"
FOLD_FF2_PROJ_OUT = os.environ.get("TG_FOLD_FF2", "1") != "0"
# channel widths whose cross-attention layers run folded (TG_CROSS_FOLD="320,640,1280": every level, "" = none); the mid block
# (8x8: an image has fewer rows than a GEMM tile) always keeps the attention kernel.  At C = 320 (64x64) the folded form has
# MORE flops than the head-wise one (768 vs 401 per row and channel) and the power-limited sustained loop is 0.3 % faster with
# the tcgen05 cross-attention kernel there (390.5 vs 389.0 steps/s), so the default folds the 32x32 and 16x16 levels only.
OVERLAP_PDL_DEFAULT = "00"
CROSS_FOLD_C = tuple(int(v) for v in os.environ.get("TG_CROSS_FOLD", "640,1280").split(",") if v)


class UNetOutput:
    """Stand-in for diffusers' UNet2DConditionOutput: `.sample`, and `[0]` for return_dict=False call sites."""

    def __init__(self, sample):
        self.sample = sample

    def __getitem__(self, i):
        return (self.sample,)[i]


class _Config(dict):
    __getattr__ = dict.__getitem__


class IPAttnProcessorHandle:
    """What `unet.attn_processors` holds for attn2 layers: carries the reference's mutable `scale` attribute
    (ip_adapter/attention_processor.py:415; mutated by IPAdapter.set_scale, ip_adapter/ip_adapter.py:155-158)."""

    def __init__(self, hidden_size, cross_attention_dim, scale=1.0, num_tokens=4):
        self.hidden_size, self.cross_attention_dim = hidden_size, cross_attention_dim
        self.scale, self.num_tokens = scale, num_tokens


class AttnProcessorHandle:
    pass


def _resnet_specs():
    out, ch, skips = [], BLOCK_OUT[0], [BLOCK_OUT[0]]
    for i, co in enumerate(BLOCK_OUT):
        for j in range(2):
            out.append((f"down_blocks.{i}.resnets.{j}", ch, 0, co))
            ch = co
            skips.append(ch)
        if i < 3:
            skips.append(ch)
    out += [("mid_block.resnets.0", ch, 0, ch), ("mid_block.resnets.1", ch, 0, ch)]
    for i, co in enumerate(reversed(BLOCK_OUT)):
        for j in range(3):
            s = skips.pop()
            out.append((f"up_blocks.{i}.resnets.{j}", ch, s, co))
            ch = co
    return out


def _transformer_specs():
    out = []
    for i in range(3):
        for j in range(2):
            out.append((f"down_blocks.{i}.attentions.{j}", BLOCK_OUT[i], ("down", i, j, 0)))
    rev = list(reversed(BLOCK_OUT))
    for i in range(1, 4):
        for j in range(3):
            out.append((f"up_blocks.{i}.attentions.{j}", rev[i], ("up", i, j, 0)))
    out.append(("mid_block.attentions.0", 1280, ("mid", 0, 0, 0)))
    return out


def _getter(sd, dev):
    def g(name):
        t = sd[name].to(device=dev, dtype=torch.float16).contiguous()
        # TMA descriptors / 16-byte vector loads need 16-byte aligned bases (views into a flat broadcast buffer are not)
        return t.clone() if t.data_ptr() % 16 else t
    return g


def pack_resnet(g, sd, p, cin, co, temb_off):
    """diffusers ResnetBlock2D parameters -> kernel layouts (conv3x3 [Cout,3,3,Cin], 1x1 shortcut [Cout,Cin])."""
    r = dict(cin=cin, cout=co, temb_off=temb_off)
    r["n1g"], r["n1b"] = g(p + ".norm1.weight"), g(p + ".norm1.bias")
    r["c1w"], r["c1b"] = pack_conv3x3(g(p + ".conv1.weight")), g(p + ".conv1.bias")
    r["n2g"], r["n2b"] = g(p + ".norm2.weight"), g(p + ".norm2.bias")
    r["c2w"], r["c2b"] = pack_conv3x3(g(p + ".conv2.weight")), g(p + ".conv2.bias")
    if (p + ".conv_shortcut.weight") in sd:
        r["scw"] = g(p + ".conv_shortcut.weight").reshape(co, cin).contiguous()
        r["scb"] = g(p + ".conv_shortcut.bias")
    return r, g(p + ".time_emb_proj.weight"), g(p + ".time_emb_proj.bias")


def pack_transformer(g, dev, p, c, key, kv_off, has_ip, has_key=lambda k: True):
    """Transformer2DModel + BasicTransformerBlock parameters -> kernel layouts; returns (dict, context K/V rows)."""
    tb_ = p + ".transformer_blocks.0"
    t = dict(C=c, key=key, kv_off=kv_off)
    t["ng"], t["nb"] = g(p + ".norm.weight"), g(p + ".norm.bias")
    t["piw"], t["pib"] = g(p + ".proj_in.weight").reshape(c, c).contiguous(), g(p + ".proj_in.bias")
    t["pow"], t["pob"] = g(p + ".proj_out.weight").reshape(c, c).contiguous(), g(p + ".proj_out.bias")
    for n in ("norm1", "norm2", "norm3"):
        t[n + "g"], t[n + "b"] = g(f"{tb_}.{n}.weight"), g(f"{tb_}.{n}.bias")
    qkv = [g(f"{tb_}.attn1.to_{x}.weight") for x in "qkv"]
    # tcgen05 attention path: every head padded (zero weight rows) to 64 columns for d = 40 and to 128 (two 64-column
    # chunks) for d = 80; d = 160 stays unpadded on the mma.sync kernel
    d_ = c // HEADS
    t["qkv_pad"] = 0 if QKV_PACKED else (64 if d_ + 1 <= 64 else (128 if d_ + 1 <= 128 else 0))
    if t["qkv_pad"]:
        pad = t["qkv_pad"]
        padded = torch.zeros((3, HEADS, pad, c), dtype=torch.float16, device=dev)
        for i_, m_ in enumerate(qkv):
            padded[i_, :, :d_] = m_.view(HEADS, d_, c)
        t["qkvw"] = padded.view(3 * HEADS * pad, c).contiguous()
    else:
        t["qkvw"] = torch.cat(qkv, 0).contiguous()
    t["o1w"], t["o1b"] = g(f"{tb_}.attn1.to_out.0.weight"), g(f"{tb_}.attn1.to_out.0.bias")
    t["q2w"] = g(f"{tb_}.attn2.to_q.weight")
    t["o2w"], t["o2b"] = g(f"{tb_}.attn2.to_out.0.weight"), g(f"{tb_}.attn2.to_out.0.bias")
    rows = [g(f"{tb_}.attn2.to_k.weight"), g(f"{tb_}.attn2.to_v.weight")]
    if has_ip:
        # the IP-Adapter rows always exist when the UNet is built for image tokens: a stock diffusers state dict has no
        # to_k_ip / to_v_ip keys -- they arrive later through IPAdapter.load_ip_adapter -> TGUNet.load_ip_weights
        # (ip_adapter/ip_adapter.py:127-140); zero rows until then (an unloaded adapter contributes nothing)
        for n_ in ("to_k_ip", "to_v_ip"):
            k_ = f"{tb_}.attn2.processor.{n_}.weight"
            rows.append(g(k_) if has_key(k_) else torch.zeros_like(rows[0]))
    t["kv_rows"] = len(rows)
    if c in CROSS_FOLD_C and not p.startswith("mid_block"):
        fold_cross_weights(t, rows)
    t["ff1w"], t["ff1b"] = pack_geglu(g(f"{tb_}.ff.net.0.proj.weight"), g(f"{tb_}.ff.net.0.proj.bias"))
    t["ff2w"], t["ff2b"] = g(f"{tb_}.ff.net.2.weight"), g(f"{tb_}.ff.net.2.bias")
    if FOLD_FF2_PROJ_OUT:
        # ff.net.2 (+ residual h) followed by proj_out (+ residual x) is linear in (h, ff): one GEMM over the concatenation
        #   out = [h | ff] . [Wpo | Wpo W2]^T + (Wpo b2 + bpo) + x        (same flops, one launch and one [M, C] round trip less)
        wpo, w2 = t["pow"].float(), t["ff2w"].float()
        t["fpw"] = torch.cat([wpo, wpo @ w2], 1).half().contiguous()
        t["fpb"] = (wpo @ t["ff2b"].float() + t["pob"].float()).half()
    return t, rows


def _fold_pair(t, wk, wv):
    """[heads*C | heads*C, 768]: the key rows  scale*log2(e) * Wq_h^T Wk_h  and the value rows  Wo_h Wv_h  of every head."""
    c, d_ = t["C"], t["C"] // HEADS
    wq = t["q2w"].float().view(HEADS, d_, c)
    wo = t["o2w"].float().view(c, HEADS, d_)
    fk = torch.einsum("hdc,hdk->hck", wq, wk.float().view(HEADS, d_, -1)) * (d_ ** -0.5 * 1.4426950408889634)
    fv = torch.einsum("chd,hdk->hck", wo, wv.float().view(HEADS, d_, -1))
    return torch.cat([fk.reshape(HEADS * c, -1), fv.reshape(HEADS * c, -1)], 0).half().contiguous()


def fold_cross_weights(t, rows):
    """Folded cross-attention (include/theatergen.h tg_cross_scores): attn2.to_q is folded into the key projection and
    attn2.to_out into the value projection of every head, so the per-trajectory context projection yields, per image, a
    [heads*96, C] score operand and a [C, heads*96] output operand and a denoise step runs two GEMMs instead of
    to_q + attention + to_out.  Same algebra as attention_processor.py:461-553, summed in a different order."""
    t["fkt"] = _fold_pair(t, rows[0], rows[1])
    if len(rows) == 4:
        t["fki"] = _fold_pair(t, rows[2], rows[3])


def downsample_conv(x, w9, bias, cout, pad_lo=1, act=ops.TG_ACT_NONE, stats=None):
    """diffusers Downsample2D: 3x3 convolution, stride 2.  w9 = packed weight [Cout, 9*Cin] (the [Cout,3,3,Cin] tensor flattened).
    Default: implicit GEMM with a stride-2 TMA gather; TG_CONV_S2=0 selects the im2col buffer + GEMM of round 1 (A/B)."""
    B, H, W, cin = x.shape
    if os.environ.get("TG_CONV_S2", "1") != "0":
        return _set(ops.conv3x3_s2(x, w9.view(cout, 3, 3, cin), bias, pad_lo=pad_lo, act=act, stats=stats), stats)
    return _set(ops.gemm(ops.im2col_s2(x, pad_lo=pad_lo), w9, bias, act=act, stats=stats).view(B, H // 2, W // 2, cout), stats)


def upsample_conv(x, w, name, stats=None):
    """diffusers Upsample2D: nearest x2 + 3x3 convolution.  Default: four 2x2 convolutions on the low-resolution grid
    (w[name + ".w4"], weights.pack_upsample_conv3x3); TG_UPCONV=0 selects upsample kernel + 3x3 convolution (A/B)."""
    if os.environ.get("TG_UPCONV", "1") != "0":
        return _set(ops.upsample2x_conv3x3(x, w[name + ".w4"], w[name + ".b"], stats=stats), stats)
    return _set(ops.conv3x3(ops.upsample2x(x), w[name + ".w"], w[name + ".b"], stats=stats), stats)


def _st(x):
    """GroupNorm totals travelling with a tensor (set by its producer through _set), or None."""
    return None if x is None else getattr(x, "_gn", None)


def _set(x, stats):
    if stats is not None:
        x._gn = ops.GnStats(stats)
    return x


class _BlockNet:
    """Execution of the blocks the UNet and the ControlNet share (ResnetBlock2D, Transformer2DModel, the time
    embedding and the encoder half), on packed weights held in self.w / self.resnets / self.transformers.

    GroupNorm statistics: every GroupNorm input of these networks is the output of a tensor-core GEMM / convolution, so
    the producer's epilogue accumulates the per-image (sum, sum of squares) totals (ops.gemm(..., stats=)) and the
    normalisation is ONE pass over the activation (ops.groupnorm(..., st1=, st2=)).  The totals of one forward live in an
    arena (`_stats_begin` zeroes it, `_take` hands out blocks) and travel with the tensors as the attribute `_gn`; a tensor
    without it (foreign input, TG_GN_STATS=0) takes the self-contained GroupNorm kernels."""
    dtype = torch.float16
    _STATS_CHANNELS = 112 * 1024  # >= channels of all GroupNorm-feeding tensors of one forward (UNet: ~96 K)

    def _stats_begin(self, B):
        self._arena = None
        if os.environ.get("TG_GN_STATS", "1") == "0":
            return
        arenas = self.__dict__.setdefault("_arenas", {})  # one per batch size, never freed: captured graphs point into them
        a = arenas.get(B)
        if a is None:
            a = arenas[B] = [torch.zeros(B * self._STATS_CHANNELS, dtype=torch.int64, device=self.device), None]
        else:
            (a[0] if a[1] is None else a[0][:a[1]]).zero_()
        self._arena, self._arena_off = a, 0

    def _stats_end(self):
        if getattr(self, "_arena", None) is not None:
            self._arena[1] = max(self._arena[1] or 0, self._arena_off)
        self._arena = None

    _STATS_MIN_HW = 1024  # measured (tools/gn_stats_bench.py, B = 10): the one-pass norm wins at 64x64 and 32x32 (-7..-16 us per
    #                       site); at 16x16 / 8x8 the cluster kernel (slab on chip) is as fast and the producers' epilogue
    #                       work would be pure cost

    def _take(self, images, C, hw):
        """A zeroed [images, C/2, 2] block of this forward's arena for a tensor of `hw` pixels per image (None when the
        statistics epilogues are off or the tensor's GroupNorm is better left to the self-measuring kernel)."""
        a = getattr(self, "_arena", None)
        if a is None or hw < self._STATS_MIN_HW:
            return None
        n = images * C
        assert self._arena_off + n <= a[0].numel(), "GroupNorm statistics arena exhausted"
        v = a[0][self._arena_off:self._arena_off + n].view(images, C // 2, 2)
        self._arena_off += n
        return v

    def project_context(self, ctx, out=None):
        """K / V projections of the context tokens for every cross-attention layer in one GEMM: [B, T, 768] ->
        [B*T, sum of 2C (text) + 2C (image tokens) per layer].  They depend on the prompt / image tokens only -- not on the
        timestep or the latents -- so a denoise loop computes them ONCE per trajectory (engine.StepGraph.set_context)
        instead of once per step like the reference's processors do (attention_processor.py:475-476, 497-498): same values."""
        kv = ops.gemm(ctx.reshape(ctx.shape[0] * ctx.shape[1], -1), self.w["ctxkv.w"], out=out)
        self._fold_context(ctx, kv)
        return kv

    def _fold_context(self, ctx, kv):
        """Per-image operands of the folded cross-attention layers (fold_cross_weights), kept as `kv.fold[layer] = (kf
        [B, heads, 96, C], vf [B, C, heads, 96])`: allocated on first use, refreshed in place afterwards so a captured
        graph keeps reading the same buffers."""
        fold = getattr(kv, "fold", None)
        if fold is None:
            fold = kv.fold = {}
        B, T = ctx.shape[:2]
        n_ip = self.num_ip_tokens if getattr(self, "has_ip", False) else 0
        n_text = T - n_ip
        ct = ci = None
        for p, t in self.transformers.items():
            if "fkt" not in t:
                continue
            C = t["C"]
            if ct is None:
                ct = ctx[:, :n_text].reshape(B * n_text, -1)
                ci = ctx[:, n_text:].reshape(B * n_ip, -1) if n_ip else None
            if p not in fold or fold[p][0].shape[0] != B:
                fold[p] = (torch.zeros((B, HEADS, ops.CROSS_SLOTS, C), dtype=torch.float16, device=ctx.device),
                           torch.zeros((B, C, HEADS, ops.CROSS_SLOTS), dtype=torch.float16, device=ctx.device))
            kf, vf = fold[p]
            pr = ops.gemm(ct, t["fkt"]).view(B, n_text, 2, HEADS, C)
            kf[:, :, :n_text].copy_(pr[:, :, 0].permute(0, 2, 1, 3))
            vf[..., :n_text].copy_(pr[:, :, 1].permute(0, 3, 2, 1))
            if n_ip:
                pi = ops.gemm(ci, t["fki"]).view(B, n_ip, 2, HEADS, C)
                kf[:, :, 80:80 + n_ip].copy_(pi[:, :, 0].permute(0, 2, 1, 3))
                vf[..., 80:80 + n_ip].copy_(pi[:, :, 1].permute(0, 3, 2, 1))

    def _time_and_context(self, t_dev, ctx, B, ctxkv=None):
        w = self.w
        emb = ops.timestep_embedding(t_dev, 320)
        emb = ops.gemm(emb, w["t1.w"], w["t1.b"], act=ops.TG_ACT_SILU)
        emb = ops.gemm(emb, w["t2.w"], w["t2.b"], act=ops.TG_ACT_SILU)  # = silu(time_embedding(t)), its only use
        temb_all = ops.gemm(emb, w["temb.w"], w["temb.b"])
        if ctxkv is None:
            ctxkv = self.project_context(ctx)
        return temb_all, ctxkv

    def _encoder(self, x, temb_all, tf, B, H, W, share=None):
        """conv_in output -> (last down output, 12 skip tensors, h, w): CrossAttnDownBlock2D x3 + DownBlock2D
        (ip_adapter/unet_2d_blocks.py:1045-1116, 1171-1202).

        share = (src, first): CFG prefix sharing.  The unconditional and the conditional row of a trajectory enter the network
        with the SAME latents and timestep (`torch.cat([latents] * 2)`, models/pipelines.py:409, 747) and differ only in the
        context, which the network first looks at in the cross-attention of down_blocks.0.attentions.0.  Everything before it
        -- conv_in, the first ResnetBlock2D, GroupNorm, proj_in and the whole 64x64 self-attention of the first transformer --
        is therefore identical for the two rows and is computed once per trajectory: `x` holds the U unique rows,
        `src[b]` is the unique row of batch row b (device int64 [B]) and `first[u]` a batch row that carries unique row u's
        timestep (device int64 [U]).  Same values as the duplicated computation, half the work for that prefix."""
        w = self.w
        h_, w_ = H, W
        if share is not None:
            src, first = share
            U = x.shape[0]
            skips = [x.index_select(0, src)]
            if _st(x) is not None:
                _set(skips[0], _st(x).main.index_select(0, src))
            xu = self._resnet("down_blocks.0.resnets.0", x, None, temb_all.index_select(0, first), U, h_, w_)
            x = tf("down_blocks.0.attentions.0", xu, h_, w_, share=src)
            skips.append(x)
        else:
            skips = [x]
        for i in range(4):
            for j in range(2):
                if share is not None and i == 0 and j == 0:
                    continue
                x = self._resnet(f"down_blocks.{i}.resnets.{j}", x, None, temb_all, B, h_, w_)
                if i < 3:
                    x = tf(f"down_blocks.{i}.attentions.{j}", x, h_, w_)
                skips.append(x)
            if i < 3:
                c = BLOCK_OUT[i]
                x = downsample_conv(x, w[f"down{i}.w"], w[f"down{i}.b"], c, stats=self._take(B, c, (h_ // 2) * (w_ // 2)))
                h_, w_ = h_ // 2, w_ // 2
                skips.append(x)
        return x, skips, h_, w_

    def _mid(self, x, temb_all, tf, B, h_, w_):
        x = self._resnet("mid_block.resnets.0", x, None, temb_all, B, h_, w_)
        x = tf("mid_block.attentions.0", x, h_, w_)
        return self._resnet("mid_block.resnets.1", x, None, temb_all, B, h_, w_)


    def _resnet(self, p, x1, x2, temb_all, B, H, W):
        r = self.resnets[p]
        cin, co = r["cin"], r["cout"]
        HW = H * W
        sc, side = None, None
        if "scw" in r and getattr(self, "fork_shortcut", False):
            # the 1x1 shortcut depends on the block input only: a parallel branch (its own stream / graph branch and
            # split-K slot) beside GroupNorm -> conv1 -> GroupNorm, joined in front of conv2, whose epilogue adds it
            main, side = torch.cuda.current_stream(), self._shortcut_stream()
            side.wait_stream(main)
            _lib.check(_lib.lib().tg_gemm_set_scratch_slot(2), "tg_gemm_set_scratch_slot")
            try:
                with torch.cuda.stream(side):
                    sc = ops.gemm(x1.view(B * HW, -1), r["scw"], r["scb"], a2=None if x2 is None else x2.view(B * HW, -1))
            finally:
                _lib.check(_lib.lib().tg_gemm_set_scratch_slot(0), "tg_gemm_set_scratch_slot")
        n1 = ops.groupnorm(x1.view(B, HW, -1), r["n1g"], r["n1b"], 1e-5, x2=None if x2 is None else x2.view(B, HW, -1),
                           silu=True, st1=_st(x1), st2=_st(x2))
        sh, so = self._take(B, co, HW), self._take(B, co, HW)
        h = ops.conv3x3(n1.view(B, H, W, cin), r["c1w"], r["c1b"],
                        rowbias=temb_all[:, r["temb_off"]:r["temb_off"] + co], stats=sh)
        n2 = ops.groupnorm(h.view(B, HW, co), r["n2g"], r["n2b"], 1e-5, silu=True, st1=_st(_set(h, sh)))
        if side is not None:
            torch.cuda.current_stream().wait_stream(side)
            sc = sc.view(B, H, W, co)
        elif "scw" in r:
            sc = ops.gemm(x1.view(B * HW, -1), r["scw"], r["scb"], a2=None if x2 is None else x2.view(B * HW, -1))
            sc = sc.view(B, H, W, co)
        else:
            assert x2 is None
            sc = x1
        return _set(ops.conv3x3(n2.view(B, H, W, co), r["c2w"], r["c2b"], residual=sc, stats=so), so)

    def _transformer(self, p, x, ctxkv, B, H, W, n_text, n_ip, ip_scale, capture, save_keys, cond_only, token_idx,
                     capture_b0=None, share=None):
        """share (device int64 [B], CFG prefix sharing, see _encoder): `x` holds U unique rows; GroupNorm, proj_in and the
        self-attention run on them once and are expanded to the B batch rows in front of the cross-attention."""
        t = self.transformers[p]
        Bu = x.shape[0] if share is not None else B   # rows the context-independent part runs on
        C, M, Mu = t["C"], B * H * W, Bu * H * W
        n = ops.groupnorm(x.view(Bu, H * W, C), t["ng"], t["nb"], 1e-6, silu=False, st1=_st(x))
        h = ops.gemm(n.view(Mu, C), t["piw"], t["pib"])
        # self attention
        ln = ops.layernorm(h, t["norm1g"], t["norm1b"])
        qkv = ops.gemm(ln, t["qkvw"])
        if t["qkv_pad"] and (H * W) % 128 == 0:
            P = HEADS * t["qkv_pad"]
            a = ops.attention_tc(qkv[:, :P], qkv[:, P:2 * P], qkv[:, 2 * P:], Bu, HEADS, C // HEADS)
        elif t["qkv_pad"]:
            raise RuntimeError("padded-QKV layers need a token count that is a multiple of 128")
        elif (H * W) % 128 == 0 and (C // HEADS) % 8 == 0 and C // HEADS <= 176 and os.environ.get("TG_ATTN_TC160", "1") != "0":
            # packed heads on the tcgen05 kernel (d = 160 at 16x16: three 64-column chunks per head; d = 40 / 80 with
            # QKV_PACKED: the QKV projection has 3C output columns instead of 3 * 8 * 64 / 3 * 8 * 128)
            a = ops.attention_tc(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], Bu, HEADS, C // HEADS, packed=True)
        else:
            a = ops.attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], Bu, HEADS)
        h = ops.gemm(a, t["o1w"], t["o1b"], residual=h)
        if share is not None:  # from here on the two rows of a trajectory see different context tokens
            h = h.view(Bu, H * W * C).index_select(0, share).view(M, C)
            x = x.view(Bu, H * W * C).index_select(0, share)
        # cross attention (+ image tokens)
        ln = ops.layernorm(h, t["norm2g"], t["norm2b"])
        probs, b0 = None, 0
        want_probs = capture is not None and (save_keys is None or t["key"] in save_keys)
        if want_probs:
            b0 = capture_b0 if capture_b0 is not None else (B // 2 if cond_only else 0)
        fold = getattr(ctxkv, "fold", {}).get(p) if (H * W) % 256 == 0 else None
        if fold is not None:
            # folded layer: probabilities = softmax epilogue of ln . Kf[b]^T, output = P . Vf[b]^T + bias + residual
            NS = HEADS * ops.CROSS_SLOTS
            pm = ops.cross_scores(ln, fold[0].view(B, NS, C), n_text, n_ip, ip_scale if n_ip else None)
            if want_probs:
                probs = pm.view(B, H * W, HEADS, ops.CROSS_SLOTS)[b0:, :, :, :n_text].permute(0, 2, 1, 3).contiguous()
            h = ops.cross_apply(pm, fold[1].view(B, C, NS), t["o2b"], residual=h)
        else:
            q = ops.gemm(ln, t["q2w"])
            o = t["kv_off"]
            if want_probs:
                probs = torch.empty((B - b0, HEADS, H * W, n_text), dtype=torch.float16, device=x.device)
            a = ops.attention_cross_ip(q, ctxkv[:, o:o + C], ctxkv[:, o + C:o + 2 * C],
                                       ctxkv[:, o + 2 * C:o + 3 * C] if n_ip else None,
                                       ctxkv[:, o + 3 * C:o + 4 * C] if n_ip else None, B, HEADS, n_text, n_ip,
                                       ip_scale=ip_scale if n_ip else None, probs=probs, probs_b0=b0)
            h = ops.gemm(a, t["o2w"], t["o2b"], residual=h)
        if probs is not None:
            if token_idx is not None:
                probs = probs[..., token_idx:token_idx + 1] if isinstance(token_idx, int) else probs[..., token_idx]
            capture[t["key"]] = probs
        # feed forward (GEGLU fused in the GEMM epilogue)
        ln = ops.layernorm(h, t["norm3g"], t["norm3b"])
        ff = ops.gemm(ln, t["ff1w"], t["ff1b"], act=ops.TG_ACT_GEGLU)
        so = self._take(B, C, H * W)
        if "fpw" in t:  # ff.net.2 and proj_out as one GEMM over [h | ff] (pack_transformer)
            return _set(ops.gemm(h, t["fpw"], t["fpb"], a2=ff, residual=x.view(M, C), stats=so).view(B, H, W, C), so)
        h = ops.gemm(ff, t["ff2w"], t["ff2b"], residual=h)
        return _set(ops.gemm(h, t["pow"], t["pob"], residual=x.view(M, C), stats=so).view(B, H, W, C), so)


class TGUNet(_BlockNet):
    def __init__(self, state_dict, device="cuda", num_ip_tokens=4, cross_attention_dim=768):
        self.device = torch.device(device)
        self.num_ip_tokens = num_ip_tokens
        self.config = _Config(in_channels=4, cross_attention_dim=cross_attention_dim, block_out_channels=BLOCK_OUT,
                              sample_size=64)
        self._pack(state_dict)
        self.attn_processors = {}
        for p, c, _ in _transformer_specs():
            self.attn_processors[f"{p}.transformer_blocks.0.attn1.processor"] = AttnProcessorHandle()
            self.attn_processors[f"{p}.transformer_blocks.0.attn2.processor"] = IPAttnProcessorHandle(
                c, cross_attention_dim, 1.0, num_ip_tokens)
        self.ip_scale_per_sample = None  # optional float32 [B] device tensor (co-batched characters)
        self.fork_shortcut = os.environ.get("TG_SC_FORK", "1") != "0"  # _resnet: shortcut GEMM as a parallel branch
        self._scale_buf = {}

    # -------------------------------------------------------------------------------------------------
    def _pack(self, sd):
        dev = self.device
        g = _getter(sd, dev)
        w = {}
        w["conv_in.w"], w["conv_in.b"] = ops.pack_conv_in(g("conv_in.weight")), g("conv_in.bias")
        w["t1.w"], w["t1.b"] = g("time_embedding.linear_1.weight"), g("time_embedding.linear_1.bias")
        w["t2.w"], w["t2.b"] = g("time_embedding.linear_2.weight"), g("time_embedding.linear_2.bias")
        self.resnets = {}
        tw, tb, off = [], [], 0
        for p, c1, c2, co in _resnet_specs():
            self.resnets[p], tw_, tb_ = pack_resnet(g, sd, p, c1 + c2, co, off)
            tw.append(tw_)
            tb.append(tb_)
            off += co
        w["temb.w"], w["temb.b"] = torch.cat(tw, 0).contiguous(), torch.cat(tb, 0).contiguous()
        self.transformers = {}
        kvw, off = [], 0
        has_ip = self.num_ip_tokens > 0
        self.has_ip = has_ip
        for p, c, key in _transformer_specs():
            self.transformers[p], rows = pack_transformer(g, dev, p, c, key, off, has_ip, sd.__contains__)
            kvw += rows
            off += c * len(rows)
        w["ctxkv.w"] = torch.cat(kvw, 0).contiguous()
        for i in range(3):
            c = BLOCK_OUT[i]
            w[f"down{i}.w"] = pack_conv3x3(g(f"down_blocks.{i}.downsamplers.0.conv.weight")).reshape(c, 9 * c)
            w[f"down{i}.b"] = g(f"down_blocks.{i}.downsamplers.0.conv.bias")
        for i in range(3):
            w[f"up{i}.w"] = pack_conv3x3(g(f"up_blocks.{i}.upsamplers.0.conv.weight"))   # (guidance-gradient path)
            w[f"up{i}.w4"] = pack_upsample_conv3x3(g(f"up_blocks.{i}.upsamplers.0.conv.weight"))
            w[f"up{i}.b"] = g(f"up_blocks.{i}.upsamplers.0.conv.bias")
        w["no.g"], w["no.b"] = g("conv_norm_out.weight"), g("conv_norm_out.bias")
        w["conv_out.w"], w["conv_out.b"] = pack_conv3x3(g("conv_out.weight")), g("conv_out.bias")
        # conv_out (320 -> 4) on the tensor-core convolution, output columns padded to one 64-wide tile: ~20 us instead of
        # ~100 us at B = 10 for the CUDA-core kernel (which stays for the VAE's 512x512 conv_out)
        w["conv_out.w64"], w["conv_out.b64"] = ops.pack_conv_out_tc(w["conv_out.w"], w["conv_out.b"])
        self.w = w

    def load_ip_weights(self, ip_state_dict):
        """`ip_adapter.{1,3,..,31}.to_{k,v}_ip.weight` in processor order (down, up, mid -- the order of
        unet.attn_processors the reference loads through a ModuleList, ip_adapter/ip_adapter.py:139-140)."""
        if not self.has_ip:
            raise RuntimeError("TGUNet was built with num_ip_tokens=0: it has no IP-Adapter rows to load into")
        for n, (p, c, _) in enumerate(_transformer_specs()):
            idx = 2 * n + 1
            o = self.transformers[p]["kv_off"]
            for j, name in enumerate(("to_k_ip", "to_v_ip")):
                wt = ip_state_dict[f"{idx}.{name}.weight"].to(device=self.device, dtype=torch.float16)
                self.w["ctxkv.w"][o + (2 + j) * c:o + (3 + j) * c].copy_(wt)
            t = self.transformers[p]
            if "fki" in t:  # folded layer: the image-token operands follow the new weights
                t["fki"] = _fold_pair(t, self.w["ctxkv.w"][o + 2 * c:o + 3 * c], self.w["ctxkv.w"][o + 3 * c:o + 4 * c])

    def weight_bytes(self):
        n = sum(v.numel() for v in self.w.values())
        for d in list(self.resnets.values()) + list(self.transformers.values()):
            n += sum(v.numel() for v in d.values() if torch.is_tensor(v))
        return 2 * n

    # -------------------------------------------------------------------------------------------------
    def _shortcut_stream(self):
        if getattr(self, "_sc_side", None) is None:
            self._sc_side = torch.cuda.Stream(device=self.device)
        return self._sc_side

    def _side_stream(self):
        if getattr(self, "_side", None) is None:
            self._side = torch.cuda.Stream(device=self.device)
        return self._side

    def _ip_scale_tensor(self, B):
        if self.ip_scale_per_sample is not None:
            assert self.ip_scale_per_sample.numel() == B
            return self.ip_scale_per_sample
        s = float(next(p.scale for p in self.attn_processors.values() if isinstance(p, IPAttnProcessorHandle)))
        key = (B, s)
        if key not in self._scale_buf:
            self._scale_buf[key] = torch.full((B,), s, dtype=torch.float32, device=self.device)
        return self._scale_buf[key]

    # -------------------------------------------------------------------------------------------------
    def forward_raw(self, sample, t_dev, ctx, ip_scale=None, down_residuals=None, mid_residual=None, capture=None,
                    save_keys=None, cond_only=False, token_idx=None, capture_b0=None, controlnet=None, control=None,
                    ctxkv=None, unique=None):
        """sample [B,4,H,W] fp16 (B even), t_dev float32 [B] on device, ctx [B,T,768] fp16 -> eps [B,4,H,W].

        ControlNet residuals arrive either as NCHW tensors (down_residuals / mid_residual: the reference call
        surface, ip_adapter/unet_2d_condition.py:1044-1053, 1077-1078) or fused: `controlnet` (a TGControlNet) with
        `control = dict(rows=(r0, r1), ctx=[r1-r0,77,768], cond_emb=...)` runs the ControlNet encoder on batch rows
        [r0, r1) of this very forward and its zero-convolutions add straight into the skip tensors (GEMM epilogue).
        capture_b0: first batch row whose text-attention probabilities are captured (default B/2 when cond_only).
        ctxkv: `project_context(ctx)` computed by the caller (once per trajectory); `control["ctxkv"]` likewise for the
        ControlNet's text context.
        unique = (sample_u [U,4,H,W], src int64 [B], first int64 [U]): CFG prefix sharing (see _encoder) -- `sample` must be
        sample_u[src]; `control["share"] = (src, first)` relative to the ControlNet's rows likewise."""
        w = self.w
        B, _, H, W = sample.shape
        assert B % 2 == 0, "co-batched CFG pairs: batch must be even"
        T = ctx.shape[1]
        n_ip = self.num_ip_tokens if self.has_ip else 0
        n_text = T - n_ip
        if n_ip and ip_scale is None:
            ip_scale = self._ip_scale_tensor(B)
        temb_all, ctxkv = self._time_and_context(t_dev, ctx, B, ctxkv)
        self._stats_begin(B)
        try:
            return self._forward_blocks(sample, t_dev, temb_all, ctxkv, ip_scale, down_residuals, mid_residual, capture,
                                        save_keys, cond_only, token_idx, capture_b0, controlnet, control, unique, n_text, n_ip)
        finally:
            self._stats_end()
            _lib.lib().tg_set_pdl(-1)

    def _forward_blocks(self, sample, t_dev, temb_all, ctxkv, ip_scale, down_residuals, mid_residual, capture, save_keys,
                        cond_only, token_idx, capture_b0, controlnet, control, unique, n_text, n_ip):
        w = self.w
        B, _, H, W = sample.shape

        def tf(p, x, h_, w_, share=None):
            return self._transformer(p, x, ctxkv, B, h_, w_, n_text, n_ip, ip_scale, capture, save_keys, cond_only,
                                     token_idx, capture_b0, share=share)

        ctrl = side = None
        if controlnet is not None:
            # The ControlNet encoder (batch r1-r0, small grids) runs on a side stream beside the UNet encoder: fork here,
            # join before its residuals are added.  Under CUDA-graph capture the two become parallel graph branches.
            r0, r1 = control["rows"]
            if os.environ.get("TG_CN_OVERLAP", "1") == "0":  # A/B measurements: ControlNet in line on the main stream
                ctrl = controlnet.encode(sample[r0:r1], t_dev[r0:r1], control["ctx"], control["cond_emb"],
                                         ctxkv=control.get("ctxkv"), share=control.get("share"))
            else:
                main = torch.cuda.current_stream()
                side = self._side_stream()
                side.wait_stream(main)
                _lib.check(_lib.lib().tg_gemm_set_scratch_slot(1), "tg_gemm_set_scratch_slot")  # own split-K workspace
                # programmatic dependent launch while the two encoders share the GPU (include/theatergen.h tg_set_pdl): a
                # pre-launched UNet kernel would sit on the SMs its predecessor's tail frees instead of leaving them to the
                # ControlNet branch.  TG_OVERLAP_PDL = "<UNet><ControlNet>" (A/B), default "00" (measured 12.13 ms per co-batched step vs 12.22 for "11", 12.50 for "01", 12.30 for "10")
                mode = os.environ.get("TG_OVERLAP_PDL", OVERLAP_PDL_DEFAULT)
                try:
                    _lib.lib().tg_set_pdl(int(mode[1]))
                    with torch.cuda.stream(side):
                        ctrl = controlnet.encode(sample[r0:r1], t_dev[r0:r1], control["ctx"], control["cond_emb"],
                                                 ctxkv=control.get("ctxkv"), share=control.get("share"))
                finally:
                    _lib.check(_lib.lib().tg_gemm_set_scratch_slot(0), "tg_gemm_set_scratch_slot")
                    _lib.lib().tg_set_pdl(int(mode[0]))
        if unique is not None:
            s0 = self._take(unique[0].shape[0], 320, H * W)
            x = _set(ops.conv_in_gemm(unique[0], w["conv_in.w"], w["conv_in.b"], stats=s0), s0)
            x, skips, h_, w_ = self._encoder(x, temb_all, tf, B, H, W, share=(unique[1], unique[2]))
        else:
            s0 = self._take(B, 320, H * W)
            x = _set(ops.conv_in_gemm(sample, w["conv_in.w"], w["conv_in.b"], stats=s0), s0)
            x, skips, h_, w_ = self._encoder(x, temb_all, tf, B, H, W)
        if down_residuals is not None:  # ControlNet: new tensors, the mid-block input stays untouched
            skips = [ops.add_inplace(s.clone(), ops.nchw_to_nhwc(r)) for s, r in zip(skips, down_residuals)]
        x = self._mid(x, temb_all, tf, B, h_, w_)
        if mid_residual is not None:
            x = ops.add_inplace(x, ops.nchw_to_nhwc(mid_residual))
            x._gn = None  # rewritten in place: the producer's totals no longer describe it
        if ctrl is not None:  # after the mid block: the last skip tensor is also the mid block's input
            if side is not None:
                _lib.lib().tg_set_pdl(-1)
                torch.cuda.current_stream().wait_stream(side)
            controlnet.add_residuals_inplace(ctrl, skips, x, r0, r1, take=self._take)
        for i in range(4):
            for j in range(3):
                x = self._resnet(f"up_blocks.{i}.resnets.{j}", x, skips.pop(), temb_all, B, h_, w_)
                if i > 0:
                    x = tf(f"up_blocks.{i}.attentions.{j}", x, h_, w_)
            if i < 3:
                x = upsample_conv(x, w, f"up{i}", stats=self._take(B, x.shape[-1], 4 * h_ * w_))
                h_, w_ = 2 * h_, 2 * w_
        n = ops.groupnorm(x.view(B, h_ * w_, 320), w["no.g"], w["no.b"], 1e-5, silu=True, st1=_st(x))
        return ops.conv_out_tc(n.view(B, h_, w_, 320), w["conv_out.w64"], w["conv_out.b64"], 4)

    # the reference call surface (ip_adapter/unet_2d_condition.py:799-813)
    def __call__(self, sample, timestep, encoder_hidden_states=None, cross_attention_kwargs=None,
                 down_block_additional_residuals=None, mid_block_additional_residual=None, return_dict=True,
                 **unused):
        if not (sample.is_cuda and encoder_hidden_states.is_cuda):
            raise RuntimeError("TGUNet runs on a CUDA device only (no CPU fallback)")
        sample = sample.to(torch.float16).contiguous()
        ctx = encoder_hidden_states.to(torch.float16).contiguous()
        B = sample.shape[0]
        pad = B % 2
        if pad:  # guidance path runs the conditional half alone (models/pipelines.py:90)
            sample, ctx = torch.cat([sample, sample[-1:]]), torch.cat([ctx, ctx[-1:]])
        if torch.is_tensor(timestep) and timestep.is_cuda and timestep.dtype == torch.float32 and timestep.numel() == B + pad:
            t_dev = timestep
        else:
            t_dev = torch.full((B + pad,), float(timestep), dtype=torch.float32, device=sample.device)
        kw = cross_attention_kwargs or {}
        capture = kw.get("save_attn_to_dict")
        save_keys = kw.get("save_keys")
        if save_keys is not None:
            save_keys = [tuple(k) for k in save_keys]
        cap_local = {} if capture is not None else None
        eps = self.forward_raw(sample, t_dev, ctx, down_residuals=down_block_additional_residuals,
                               mid_residual=mid_block_additional_residual, capture=cap_local, save_keys=save_keys,
                               cond_only=bool(kw.get("return_cond_ca_only", False)) and not pad,
                               token_idx=kw.get("return_token_ca_only"))
        if capture is not None:
            for k, v in cap_local.items():
                v = v[:B] if pad else v
                capture[k] = v.cpu() if kw.get("offload_cross_attn_to_cpu", False) else v
        if pad:
            eps = eps[:B]
        return UNetOutput(eps) if return_dict else (eps,)

    forward = __call__
