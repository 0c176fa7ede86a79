"""the IPAdapter control plane (ip_adapter/ip_adapter.py:67-158, 289-317) on a TGUNet.

* a UNet packed from a STOCK state dict (no to_k_ip / to_v_ip keys -- what `unet.state_dict()` of a freshly loaded
  diffusers UNet holds) + `IPAdapter(pipe, ..., ip_ckpt)` -> `load_ip_adapter` -> `TGUNet.load_ip_weights` gives,
  bit for bit, the forward of a UNet packed with the IP rows merged in.  The checkpoint uses the reference's layout
  `{"image_proj": {...}, "ip_adapter": {"1.to_k_ip.weight", "1.to_v_ip.weight", "3...", ..., "31..."}}`: index = position
  in `unet.attn_processors` (down blocks, up blocks, mid block; attn1 even, attn2 odd -- ip_adapter.py:139-140 loads it
  through `ModuleList(unet.attn_processors.values())`).
* `.bin` (torch.load) and `.safetensors` checkpoints, `set_scale`, `get_image_embeds`.
* IPAdapterPlus: 16 image tokens (context 77 + 16 = 93) through the Resampler and the UNet against the oracle.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _attn_processor_names():
    """The order of `unet.attn_processors` of the SD-1.5 UNet, written out independently of the product code:
    named_children order down_blocks, up_blocks, mid_block (ip_adapter/unet_2d_condition.py:655-676)."""
    names = []
    for i in range(3):
        for j in range(2):
            names.append(f"down_blocks.{i}.attentions.{j}.transformer_blocks.0")
    for i in (1, 2, 3):
        for j in range(3):
            names.append(f"up_blocks.{i}.attentions.{j}.transformer_blocks.0")
    names.append("mid_block.attentions.0.transformer_blocks.0")
    out = []
    for n in names:
        out += [n + ".attn1.processor", n + ".attn2.processor"]
    return out


def _split_ip(sd):
    """merged synthetic state dict -> (stock UNet dict, reference-layout `ip_adapter` dict)"""
    stock = {k: v for k, v in sd.items() if "_ip." not in k}
    ip = {}
    for idx, name in enumerate(_attn_processor_names()):
        if name.endswith("attn2.processor"):
            for w in ("to_k_ip", "to_v_ip"):
                ip[f"{idx}.{w}.weight"] = sd[f"{name}.{w}.weight"]
    assert sorted(int(k.split(".")[0]) for k in ip if "to_k" in k) == list(range(1, 32, 2))
    return stock, ip


def _inputs(T, seed=0, B=2):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(B, 4, 64, 64, generator=g).half().cuda(), torch.randn(B, T, 768, generator=g).half().cuda())


def test_load_ip_adapter_into_stock_unet_equals_merged_packing(tmp_path):
    from oracle import synth
    from oracle import unet as OU
    from theatergen_b200.ip_adapter import IPAdapter
    from theatergen_b200.ip_adapter.ip_adapter import _Pipe
    from theatergen_b200.scheduler import DDIMScheduler
    from theatergen_b200.unet import TGUNet
    sd = OU.synth_state_dict(1234)
    stock, ip = _split_ip(sd)
    assert list(TGUNet(stock).attn_processors.keys()) == _attn_processor_names()
    x, ctx = _inputs(81)
    merged = TGUNet(sd)
    for p in merged.attn_processors.values():
        if hasattr(p, "scale"):
            p.scale = 0.4
    want = merged(x, 481, encoder_hidden_states=ctx).sample

    proj = synth.fill(synth.image_proj_shapes(), 9)
    ckpt = {"image_proj": proj, "ip_adapter": ip}
    torch.save(ckpt, tmp_path / "ip-adapter_sd15.bin")
    from safetensors.torch import save_file
    flat = {f"image_proj.{k}": v.contiguous() for k, v in proj.items()}
    flat.update({f"ip_adapter.{k}": v.contiguous() for k, v in ip.items()})
    save_file(flat, str(tmp_path / "ip-adapter_sd15.safetensors"))

    for path in ("ip-adapter_sd15.bin", "ip-adapter_sd15.safetensors"):
        net = TGUNet(stock)
        # before the checkpoint is loaded the image tokens contribute nothing (zero to_k_ip / to_v_ip rows): the result
        # equals the merged UNet at scale 0 up to the sign of zero
        adapter = IPAdapter(_Pipe(net, DDIMScheduler()), None, None, "cuda")
        adapter.set_scale(0.4)
        before = net(x, 481, encoder_hidden_states=ctx).sample
        assert not torch.equal(before, want)
        adapter.ip_ckpt = str(tmp_path / path)
        adapter.load_ip_adapter()
        got = net(x, 481, encoder_hidden_states=ctx).sample
        assert torch.equal(got, want), (path, (got.float() - want.float()).abs().max().item())
        # set_scale reaches every IP processor handle (ip_adapter.py:155-158)
        adapter.set_scale(0.0)
        assert all(p.scale == 0.0 for p in net.attn_processors.values() if hasattr(p, "scale"))
        for p in merged.attn_processors.values():
            if hasattr(p, "scale"):
                p.scale = 0.0
        assert torch.equal(net(x, 481, encoder_hidden_states=ctx).sample, merged(x, 481, encoder_hidden_states=ctx).sample)
        for p in merged.attn_processors.values():
            if hasattr(p, "scale"):
                p.scale = 0.4
        # image tokens: cond from the embedding, uncond from zeros (ip_adapter.py:142-153)
        e = torch.randn(1, 1024, generator=torch.Generator().manual_seed(10)).half().cuda()
        c, u = adapter.get_image_embeds(clip_image_embeds=e)
        from oracle import resampler as ORS
        p32 = {k: v.half().float().cuda() for k, v in proj.items()}
        assert c.shape == u.shape == (1, 4, 768)
        assert torch.allclose(c.float(), ORS.image_proj(p32, e.float()), atol=1e-2, rtol=1e-2)
        assert torch.allclose(u.float(), ORS.image_proj(p32, torch.zeros_like(e).float()), atol=1e-2, rtol=1e-2)

    # the constructor path of the reference: IPAdapter(pipe, image_encoder_path, ip_ckpt, device) loads at once
    net = TGUNet(stock)
    IPAdapter(_Pipe(net, DDIMScheduler()), None, str(tmp_path / "ip-adapter_sd15.bin"), "cuda").set_scale(0.4)
    assert torch.equal(net(x, 481, encoder_hidden_states=ctx).sample, want)
    # a wrong token count is an error, not a silent re-split of the context
    with pytest.raises(RuntimeError):
        IPAdapter(_Pipe(TGUNet(stock), DDIMScheduler()), None, None, "cuda", num_tokens=16)
    with pytest.raises(RuntimeError):
        TGUNet(stock, num_ip_tokens=0).load_ip_weights(ip)


def test_ip_adapter_plus_16_tokens_vs_oracle():
    from oracle import resampler as ORS
    from oracle import synth
    from oracle import unet as OU
    from theatergen_b200.ip_adapter import IPAdapterPlus
    from theatergen_b200.ip_adapter.ip_adapter import _Pipe
    from theatergen_b200.scheduler import DDIMScheduler
    from theatergen_b200.unet import TGUNet
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    sd = OU.synth_state_dict(1234)
    stock, ip = _split_ip(sd)
    net = TGUNet(stock, num_ip_tokens=16)
    adapter = IPAdapterPlus(_Pipe(net, DDIMScheduler()), None, None, "cuda")
    rs_sd = synth.fill(synth.resampler_shapes(), 21)
    adapter.load_ip_adapter({"image_proj": rs_sd, "ip_adapter": ip})
    adapter.set_scale(0.4)
    g = torch.Generator().manual_seed(3)
    feats = torch.randn(1, 257, 1280, generator=g).half().cuda()         # hidden_states[-2] of CLIP ViT-H
    feats0 = torch.randn(1, 257, 1280, generator=g).half().cuda()        # ... of the all-zero image
    cond, uncond = adapter.get_image_embeds(clip_image_embeds=feats, uncond_clip_image_embeds=feats0)
    assert cond.shape == uncond.shape == (1, 16, 768)
    r32 = {k: v.half().float().cuda() for k, v in rs_sd.items()}
    with torch.no_grad():
        want_c, want_u = ORS.resampler(r32, feats.float()), ORS.resampler(r32, feats0.float())
    for a, b in ((cond, want_c), (uncond, want_u)):
        err = (a.float() - b).abs().max().item()
        print("resampler max|ours - fp32 oracle| =", err)
        assert err < 1e-2
    x, text = _inputs(77, seed=5)
    ctx = torch.cat([text, torch.cat([uncond, cond])], 1).contiguous()    # [2, 93, 768]
    eps = net(x, 601, encoder_hidden_states=ctx).sample
    sd32 = {k: v.half().float().cuda() for k, v in sd.items()}
    sd16 = {k: v.half().cuda() for k, v in sd.items()}
    with torch.no_grad():
        ref32 = OU.unet_forward(sd32, x.float(), 601, ctx.float(), ip_scale=0.4, num_ip_tokens=16)
        ref16 = OU.unet_forward(sd16, x, 601, ctx, ip_scale=0.4, num_ip_tokens=16)
        ref4 = OU.unet_forward(sd32, x.float(), 601, ctx.float()[:, :81], ip_scale=0.4, num_ip_tokens=4)
    e, e16 = (eps.float() - ref32).abs().max().item(), (ref16.float() - ref32).abs().max().item()
    rr = ((eps.float() - ref32).pow(2).mean().sqrt() / ref32.pow(2).mean().sqrt()).item()
    print(f"IPAdapterPlus UNet forward (T=93): max_abs {e:.3e} rel_rms {rr:.3e} | fp16 oracle {e16:.3e}")
    assert e <= max(1.5 * e16, 4e-3) and rr <= 2e-3
    assert (ref32 - ref4).abs().max().item() > 10 * e   # the 16 image tokens matter: not a 4-token result
