"""TGControlNet -- the SD-1.5 ControlNet (line-art) forward of stage 2 on libtheatergen.so kernels.

The reference builds the stock diffusers `ControlNetModel` (generate.py:91) and calls it once per stage-2 step
(models/pipelines.py:770-778); its 12 down residuals + 1 mid residual enter the UNet at
ip_adapter/unet_2d_condition.py:1044-1053 / 1077-1078.  SURVEY.md section 8(f) row 1.

B200-first structure:
  * the conditioning embedder (8 convolutions on the 512x512 hint image) does not depend on the timestep: it runs
    ONCE per trajectory (`embed_condition`) instead of once per step, on the tcgen05 implicit-GEMM conv with channels
    zero-padded to multiples of 64 (3->16->16->32->32->96->96->256->320 becomes 3->64->64->64->64->128->128->256->320);
  * the encoder half reuses the UNet's ResBlock / transformer execution (`_BlockNet`), text-only cross-attention
    (the 77 text tokens: SURVEY.md A.3 #15), K/V of its 7 cross-attentions projected by one fused GEMM;
  * the 13 zero-convolutions are GEMMs whose epilogue ADDS the result straight into the UNet's skip tensors / mid
    output (`add_residuals_inplace`): no residual tensor, no NCHW round trip, no separate add kernels.
`__call__` keeps the diffusers call surface (NCHW residual tuple) for callers that want the tensors.
"""
from __future__ import annotations

import torch

from . import ops
from .unet import BLOCK_OUT, _BlockNet, _getter, downsample_conv, pack_resnet, pack_transformer
from .weights import pack_conv3x3

COND_CHANNELS = (16, 32, 96, 256)
SKIP_CHANNELS = (320, 320, 320, 320, 640, 640, 640, 1280, 1280, 1280, 1280, 1280)


def _pad64(c):
    return (c + 63) // 64 * 64


def _resnet_specs():
    out, ch = [], BLOCK_OUT[0]
    for i, co in enumerate(BLOCK_OUT):
        for j in range(2):
            out.append((f"down_blocks.{i}.resnets.{j}", ch, co))
            ch = co
    return out + [("mid_block.resnets.0", ch, ch), ("mid_block.resnets.1", ch, ch)]


def _transformer_specs():
    out = [(f"down_blocks.{i}.attentions.{j}", BLOCK_OUT[i], ("down", i, j, 0)) for i in range(3) for j in range(2)]
    return out + [("mid_block.attentions.0", 1280, ("mid", 0, 0, 0))]


class TGControlNet(_BlockNet):
    def __init__(self, state_dict, device="cuda"):
        self.device = torch.device(device)
        self._pack(state_dict)

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
        for p, cin, co in _resnet_specs():
            self.resnets[p], tw_, tb_ = pack_resnet(g, sd, p, cin, co, off)
            tw.append(tw_)
            tb.append(tb_)
            off += co
        w["temb.w"], w["temb.b"] = torch.cat(tw, 0).contiguous(), torch.cat(tb, 0).contiguous()
        self.transformers = {}
        kvw, off = [], 0
        for p, c, key in _transformer_specs():
            self.transformers[p], rows = pack_transformer(g, dev, p, c, key, off, False)
            kvw += rows
            off += c * len(rows)
        w["ctxkv.w"] = torch.cat(kvw, 0).contiguous()
        for i in range(3):
            c = BLOCK_OUT[i]
            w[f"down{i}.w"] = pack_conv3x3(g(f"down_blocks.{i}.downsamplers.0.conv.weight")).reshape(c, 9 * c)
            w[f"down{i}.b"] = g(f"down_blocks.{i}.downsamplers.0.conv.bias")
        for i, c in enumerate(SKIP_CHANNELS):
            w[f"zc{i}.w"] = g(f"controlnet_down_blocks.{i}.weight").reshape(c, c).contiguous()
            w[f"zc{i}.b"] = g(f"controlnet_down_blocks.{i}.bias")
        w["zcm.w"] = g("controlnet_mid_block.weight").reshape(1280, 1280).contiguous()
        w["zcm.b"] = g("controlnet_mid_block.bias")

        # conditioning embedder, channels zero-padded to multiples of 64 (padded outputs are silu(0) = 0)
        def padded(name, cout_p, cin_p):
            wt, bs = g(name + ".weight"), g(name + ".bias")
            co, ci = wt.shape[:2]
            wp = torch.zeros((cout_p, cin_p, 3, 3), dtype=torch.float16, device=dev)
            wp[:co, :ci] = wt
            bp = torch.zeros((cout_p,), dtype=torch.float16, device=dev)
            bp[:co] = bs
            return wp, bp

        pe = "controlnet_cond_embedding"
        c0 = _pad64(COND_CHANNELS[0])
        wp, bp = padded(pe + ".conv_in", c0, 3)
        w["ce.in.w"], w["ce.in.b"] = ops.pack_conv_in(wp.contiguous()), bp
        self.cond_layers = []  # (weight, bias, stride, cout_padded)
        for i in range(3):
            ci, co = _pad64(COND_CHANNELS[i]), _pad64(COND_CHANNELS[i + 1])
            wp, bp = padded(f"{pe}.blocks.{2 * i}", ci, ci)
            self.cond_layers.append((pack_conv3x3(wp), bp, 1, ci))
            wp, bp = padded(f"{pe}.blocks.{2 * i + 1}", co, ci)
            self.cond_layers.append((pack_conv3x3(wp).reshape(co, 9 * ci).contiguous(), bp, 2, co))
        wp, bp = padded(pe + ".conv_out", 320, _pad64(COND_CHANNELS[3]))
        w["ce.out.w"], w["ce.out.b"] = pack_conv3x3(wp), bp
        self.w = w

    def weight_bytes(self):
        n = sum(v.numel() for v in self.w.values()) + sum(l[0].numel() + l[1].numel() for l in self.cond_layers)
        for d in list(self.resnets.values()) + list(self.transformers.values()):
            n += sum(v.numel() for v in d.values() if torch.is_tensor(v))
        return 2 * n

    # -------------------------------------------------------------------------------------------------
    def embed_condition(self, cond):
        """ControlNetConditioningEmbedding: hint image NCHW [n,3,512,512] (values in [0,1]) -> NHWC [n,64,64,320].
        Timestep independent: callers compute it once per trajectory.  The CFG duplicate of the image
        (`prepare_image(..., do_classifier_free_guidance=True)`, models/pipelines.py:712-723) is one image: pass n = 1."""
        w = self.w
        cond = cond.to(device=self.device, dtype=torch.float16).contiguous()
        n, _, H, W = cond.shape
        h = ops.conv_in_gemm(cond, w["ce.in.w"], w["ce.in.b"], silu=True)
        for wt, bs, stride, cout in self.cond_layers:
            if stride == 1:
                h = ops.conv3x3(h, wt, bs, act=ops.TG_ACT_SILU)
            else:
                H, W = H // 2, W // 2
                h = downsample_conv(h, wt, bs, cout, act=ops.TG_ACT_SILU)
        return ops.conv3x3(h, w["ce.out.w"], w["ce.out.b"])

    def encode(self, sample, t_dev, ctx, cond_emb, ctxkv=None, share=None):
        """sample NCHW [B,4,64,64], t_dev float32 [B], ctx [B,77,768] (text only), cond_emb NHWC [n,64,64,320] (row b uses image b % n)
        -> (12 skip tensors, mid output), NHWC, BEFORE the zero-convolutions."""
        w = self.w
        B, _, H, W = sample.shape
        T = ctx.shape[1]
        temb_all, ctxkv = self._time_and_context(t_dev, ctx.contiguous(), B, ctxkv)  # ctxkv: project_context(ctx), optional
        self._stats_begin(B)  # GroupNorm totals from the producers' epilogues (unet._BlockNet); zeroed on this stream
        try:
            return self._encode_blocks(sample, temb_all, ctxkv, cond_emb, share, B, H, W, T)
        finally:
            self._stats_end()

    def _encode_blocks(self, sample, temb_all, ctxkv, cond_emb, share, B, H, W, T):
        w = self.w

        def tf(p, x, h_, w_, share=None):
            return self._transformer(p, x, ctxkv, B, h_, w_, T, 0, None, None, None, False, None, share=share)

        per = H * W * 320
        n = cond_emb.shape[0]  # 1 (one image), K (rows [uncond K | cond K] share the K images) or B
        if share is not None:
            # CFG prefix sharing (unet._BlockNet._encoder): rows [uncond K | cond K] carry the same latents, timestep and
            # hint image, so conv_in + hint, the first ResnetBlock2D and the first self-attention run on the K unique rows
            src, first = share
            U = first.numel()
            assert B == 2 * U and n in (1, U), "shared prefix expects rows [uncond K | cond K] and one hint per trajectory"
            # (conv_in's output is rewritten in place by the hint add: no epilogue totals, its GroupNorm measures itself)
            x = ops.conv_in_gemm(sample[:U].contiguous(), w["conv_in.w"], w["conv_in.b"])
            for b in range(U):
                ops.add_inplace(x.view(-1)[b * per:(b + 1) * per], cond_emb[b % n].reshape(-1))
            x, skips, h_, w_ = self._encoder(x, temb_all, tf, B, H, W, share=share)
            x = self._mid(x, temb_all, tf, B, h_, w_)
            return skips, x
        x = ops.conv_in_gemm(sample, w["conv_in.w"], w["conv_in.b"])
        for b in range(B):  # + conditioning embedding
            ops.add_inplace(x.view(-1)[b * per:(b + 1) * per], cond_emb[b % n].reshape(-1))
        x, skips, h_, w_ = self._encoder(x, temb_all, tf, B, H, W)
        x = self._mid(x, temb_all, tf, B, h_, w_)
        return skips, x

    def add_residuals_inplace(self, ctrl, unet_skips, unet_mid, r0, r1, take=None):
        """unet_skips[i][r0:r1] += zero_conv_i(skip_i);  unet_mid[r0:r1] += zero_conv_mid(mid)   (conditioning scale
        1.0, the only value the reference uses: SURVEY.md A.3 #6).  One GEMM each, the addition is its epilogue.
        take: the UNet's statistics allocator -- a tensor that carries GroupNorm totals gets fresh ones for its rewritten
        rows from the same epilogue (ops.GnStats.alt)."""
        w = self.w
        skips, mid = ctrl

        def add(src, u, wk, bk):
            C = src.shape[-1]
            dst = u[r0:r1].view(-1, C)
            st = getattr(u, "_gn", None)
            alt = take(r1 - r0, C, u.shape[1] * u.shape[2]) if (st is not None and take is not None) else None
            ops.gemm(src.view(-1, C), w[wk], w[bk], residual=dst, out=dst, stats=alt)
            u._gn = None if alt is None else ops.GnStats(st.main, alt, r0, r1)

        for i, (s, u) in enumerate(zip(skips, unet_skips)):
            add(s, u, f"zc{i}.w", f"zc{i}.b")
        add(mid, unet_mid, "zcm.w", "zcm.b")

    # the diffusers call surface (models/pipelines.py:770-778)
    def __call__(self, sample, timestep, encoder_hidden_states=None, controlnet_cond=None, conditioning_scale=1.0,
                 guess_mode=False, return_dict=True, cond_emb=None, **unused):
        if guess_mode:
            raise NotImplementedError("guess_mode is never used by the reference (models/pipelines.py:776)")
        if not sample.is_cuda:
            raise RuntimeError("TGControlNet runs on a CUDA device only (no CPU fallback)")
        sample = sample.to(torch.float16).contiguous()
        ctx = encoder_hidden_states.to(torch.float16).contiguous()
        B = sample.shape[0]
        if torch.is_tensor(timestep) and timestep.is_cuda and timestep.dtype == torch.float32 and timestep.numel() == B:
            t_dev = timestep
        else:
            t_dev = torch.full((B,), float(timestep), dtype=torch.float32, device=sample.device)
        if cond_emb is None:
            cond_emb = self.embed_condition(controlnet_cond)
        skips, mid = self.encode(sample, t_dev, ctx, cond_emb)
        w = self.w
        sc = float(conditioning_scale)
        down = []
        for i, s in enumerate(skips):
            Bc, H, W, C = s.shape
            r = ops.gemm(s.view(-1, C), w[f"zc{i}.w"], w[f"zc{i}.b"], out_scale=sc)
            down.append(ops.nchw_to_nhwc(r.view(Bc, H * W, C, 1)).view(Bc, C, H, W))  # [B,HW,C] -> [B,C,HW]
        m = ops.gemm(mid.view(-1, 1280), w["zcm.w"], w["zcm.b"], out_scale=sc)
        Bc, H, W, C = mid.shape
        mid_r = ops.nchw_to_nhwc(m.view(Bc, H * W, C, 1)).view(Bc, C, H, W)
        return (down, mid_r)

    forward = __call__
