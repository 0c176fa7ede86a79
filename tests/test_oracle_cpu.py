"""CPU suite: the oracle restatement against the golden vectors produced by the reference's own code
(tests/golden/make_golden.py), plus closed-form pins for the restated third-party scheduler."""
import math
import os

import pytest
import torch

from oracle import attention as OA
from oracle import ddim as OD
from oracle import guidance as OG
from oracle import latents as OL
from oracle import resampler as OR
from oracle import synth
from oracle import unet as OU

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    return torch.load(os.path.join(GOLD, name), weights_only=False)


def gen(seed):
    return torch.Generator().manual_seed(seed)


@pytest.mark.parametrize("case", ["s64", "s256", "s1024", "s4096"])
def test_attention_processors_match_reference(case):
    g = load("attention.pt")[case]
    N, C, rs, cs = g["N"], g["C"], g["row_stride"], g["col_stride"]
    sd = synth.fill(synth.attention_shapes(C), g["seeds"]["params"])
    x = torch.randn(2, N, C, generator=gen(g["seeds"]["x"]))
    ctx = torch.randn(2, 81, 768, generator=gen(g["seeds"]["ctx"]))
    y1 = OA.attn_self(sd, "attn1", x)
    y2, probs = OA.attn_cross_ip(sd, "attn2", x, ctx, 0.4, return_probs=True)
    y0 = OA.attn_cross_ip(sd, "attn2", x, ctx, 0.0)
    assert torch.allclose(y1[:, ::rs, ::cs], g["y_self"], atol=2e-5, rtol=1e-4)
    assert torch.allclose(y2[:, ::rs, ::cs], g["y_cross"], atol=2e-5, rtol=1e-4)
    assert torch.allclose(y0[:, ::rs, ::cs], g["y_cross_scale0"], atol=2e-5, rtol=1e-4)
    assert torch.allclose(probs[1:, :, ::rs], g["probs_cond"], atol=1e-6, rtol=1e-4)  # return_cond_ca_only
    # per-sample scale tensor == python-float scale
    y2b = OA.attn_cross_ip(sd, "attn2", x, ctx, torch.tensor([0.4, 0.4]))
    assert torch.allclose(y2b, y2, atol=1e-6)


def test_resampler_and_image_proj_match_reference():
    g = load("resampler.pt")
    sd = synth.fill(synth.resampler_shapes(), g["seeds"]["params"])
    assert sum(v.numel() for v in sd.values()) == g["n_params"]
    x = torch.randn(2, 257, 1280, generator=gen(g["seeds"]["x"]))
    y = OR.resampler(sd, x)
    assert y.shape == (2, 16, 768)  # the reference's own (only) assertion is this shape: test_resampler.py:40
    assert torch.allclose(y, g["y"], atol=3e-5, rtol=1e-4)
    g = load("image_proj.pt")
    sd = synth.fill(synth.image_proj_shapes(), g["seeds"]["params"])
    e = torch.randn(3, 1024, generator=gen(g["seeds"]["e"]))
    assert torch.allclose(OR.image_proj(sd, e), g["tokens"], atol=2e-5, rtol=1e-4)
    assert torch.allclose(OR.image_proj(sd, torch.zeros_like(e)), g["tokens_uncond"], atol=2e-5, rtol=1e-4)


def test_guidance_energy_matches_reference():
    g = load("guidance.pt")
    saved = {k: v.float() for k, v in g["saved"].items()}
    l1 = OG.compute_ca_lossv3(saved, g["bboxes"], g["positions"], g["keys"], **g["kwargs"])
    assert abs(l1.item() - g["loss_plain"].item()) < 2e-5 * max(1, abs(g["loss_plain"].item()))
    l2 = OG.compute_ca_lossv3(saved, g["bboxes"], g["positions"], g["keys"], ref_ca_saved_attns=g["ref_attns"], index=1,
                              ref_ca_loss_weight=2.0, **g["kwargs"])
    assert abs(l2.item() - g["loss_ref"].item()) < 2e-5 * max(1, abs(g["loss_ref"].item()))
    for hw, masks in g["masks"].items():
        for b, m in zip(g["bboxes"], masks):
            assert torch.equal(OL.proportion_to_mask(b, hw, hw), m.cpu())
    assert OG.compute_ca_lossv3(saved, [], [], g["keys"]).item() == 0.0  # empty object list (guidance.py:255-256)


def test_latent_utils_match_reference():
    g = load("utils.pt")
    for (xo, yo), want, mwant in zip(g["shift_cases"], g["shifts"], g["mask_shifts"]):
        assert torch.equal(OL.shift_tensor(g["t"], xo, yo, offset_normalized=True), want)
        assert torch.equal(OL.shift_tensor(g["mask"], xo, yo, offset_normalized=True), mwant)
    assert torch.equal(OL.binary_mask_to_box_mask(g["mask"]), g["box_mask"])
    assert [int(v) for v in OL.binary_mask_to_box(g["mask"])] == [int(v) for v in g["box"]]
    cx, cy = OL.binary_mask_to_center(g["mask"], normalize=True)
    assert abs(cx - g["center"][0]) < 1e-6 and abs(cy - g["center"][1]) < 1e-6
    for (b, H, W), want in zip(g["scale_cases"], g["scaled"]):
        assert OG.scale_proportion(b, H, W) == tuple(want)
    ts = torch.arange(981, 0, -20)
    for (a, r), want in zip(((15, 2), (49, 2), (5, 3)), g["fast_schedule"]):
        assert torch.equal(OL.get_fast_schedule(ts, a, r), want)


def test_ddim_closed_form():
    s = OD.DDIM()
    ts = s.set_timesteps(50)
    assert ts[0].item() == 981 and ts[-1].item() == 1 and len(ts) == 50 and (ts[:-1] - ts[1:] == 20).all()
    assert s.set_timesteps(20)[0].item() == 951
    s.set_timesteps(50)
    betas = torch.linspace(0.00085 ** 0.5, 0.012 ** 0.5, 1000) ** 2
    assert torch.allclose(s.alphas_cumprod, torch.cumprod(1 - betas, 0))
    assert abs(s.alphas_cumprod[0].item() - (1 - 0.00085)) < 1e-7
    # one step with the true noise lands exactly on add_noise at the previous timestep
    x0, n = torch.randn(1, 4, 8, 8, generator=gen(1), dtype=torch.float64), torch.randn(1, 4, 8, 8, generator=gen(2),
                                                                                       dtype=torch.float64)
    xt = s.add_noise(x0, n, torch.tensor([981]))
    xp = s.step(n, 981, xt)
    assert torch.allclose(xp, s.add_noise(x0, n, torch.tensor([961])), atol=1e-6)
    # last step uses final_alpha_cumprod = alphas_cumprod[0] (set_alpha_to_one=False)
    a_t, a_prev = s._alphas(1)
    assert a_prev.item() == s.alphas_cumprod[0].item()


def test_blend_and_compose():
    lat, bg = OL.get_input_latents_list(7, 7 + 123456789, 0.01, 512, 512, [[0.25, 0.25, 0.75, 0.75]])
    assert lat[0].dtype == torch.float16 and lat[0].shape == (1, 4, 64, 64)
    m = OL.proportion_to_mask([0.25, 0.25, 0.75, 0.75], 64, 64).bool()
    assert torch.equal(lat[0][..., ~m], bg[..., ~m]) and not torch.equal(lat[0][..., m], bg[..., m])
    la = [torch.randn(51, 1, 4, 64, 64, generator=gen(i)).half() for i in range(2)]
    masks = [torch.zeros(64, 64, dtype=torch.bool) for _ in range(2)]
    masks[0][10:30, 10:30] = True
    masks[1][20:50, 25:60] = True
    comp, idx = OL.compose_latents(la, masks, bg, 50)
    assert comp.shape == (51, 1, 4, 64, 64)
    assert torch.equal(comp[5][..., masks[0]], la[0][5][..., masks[0]])          # smaller mask composed last
    only1 = masks[1] & ~masks[0]
    assert torch.equal(comp[5][..., only1], la[1][5][..., only1])
    assert int(idx[15, 15]) == 1 and int(idx[40, 40]) == 2 and int(idx[0, 0]) == 0
    assert float(comp[5][..., ~(masks[0] | masks[1])].abs().max()) == 0.0           # steps > 0 hold masked latents only


def test_unet_structure():
    shapes_only = {}
    sd = OU.synth_state_dict(0)
    n_ip = sum(v.numel() for k, v in sd.items() if ".processor." in k)
    assert OU.count_params(sd) - n_ip == 859_520_964        # SD-1.5 UNet parameter count
    assert n_ip == 19_169_280                                # 16 x (to_k_ip + to_v_ip)
    assert len(OU.resnet_specs()) == 22 and len(OU.transformer_specs()) == 16
    # processor order = IP-Adapter checkpoint index order 1,3,..,31 (ip_adapter/ip_adapter.py:139-140)
    keys = [k for _, _, k in OU.transformer_specs()]
    assert keys[0] == ("down", 0, 0, 0) and keys[6] == ("up", 1, 0, 0) and keys[-1] == ("mid", 0, 0, 0)


def test_controlnet_oracle_structure_and_determinism():
    """oracle/controlnet.py: diffusers ControlNetModel key set with the published SD-1.5 parameter count
    (361,279,120), residual shapes that line up with the UNet's 12 skips + mid, zero 'zero-convs' -> zero residuals."""
    from oracle import controlnet as OC
    from oracle import unet as OU
    sd = OC.synth_state_dict(4321)
    assert sum(v.numel() for v in sd.values()) == 361_279_120
    assert not any(k.startswith("up_blocks") or k.startswith("conv_out") for k in sd)
    g = torch.Generator().manual_seed(0)
    x, ctx = torch.randn(2, 4, 64, 64, generator=g), torch.randn(2, 77, 768, generator=g)
    cond = torch.rand(2, 3, 512, 512, generator=g)
    for k in list(sd):
        if k.startswith(("controlnet_down_blocks", "controlnet_mid_block")):
            sd[k] = torch.zeros_like(sd[k])
    torch.set_num_threads(8)
    with torch.no_grad():
        emb = OC.cond_embedding(sd, cond)
        down, mid = OC.controlnet_forward(sd, x, 981, ctx, cond)
    assert emb.shape == (2, 320, 64, 64)
    want = [(320, 64)] * 3 + [(320, 32)] + [(640, 32)] * 2 + [(640, 16)] + [(1280, 16)] * 2 + [(1280, 8)] * 3
    assert [(t.shape[1], t.shape[2]) for t in down] == want and mid.shape == (2, 1280, 8, 8)
    assert all(float(t.abs().max()) == 0 for t in down) and float(mid.abs().max()) == 0
    # zero residuals leave the UNet unchanged (unet_2d_condition.py:1044-1053)
    sdu = OU.synth_state_dict(1234)
    ctx81 = torch.cat([ctx, torch.randn(2, 4, 768, generator=g)], 1)
    with torch.no_grad():
        a = OU.unet_forward(sdu, x, 981, ctx81, ip_scale=0.1)
        b = OU.unet_forward(sdu, x, 981, ctx81, ip_scale=0.1, down_residuals=down, mid_residual=mid)
    assert torch.equal(a, b)


def test_vae_decoder_oracle_structure():
    """oracle/vae.py: diffusers AutoencoderKL decoder key set with the published parameter count of the SD VAE decoder
    (49,490,179 + 20 for post_quant_conv), 8x upsampling."""
    from oracle import vae as OV
    sd = OV.synth_state_dict(777)
    assert sum(v.numel() for v in sd.values()) == 49_490_199
    assert "decoder.mid_block.attentions.0.to_q.weight" in sd and "decoder.up_blocks.3.resnets.2.conv2.weight" in sd
    assert sd["decoder.up_blocks.2.resnets.0.conv_shortcut.weight"].shape == (256, 512, 1, 1)
    g = torch.Generator().manual_seed(0)
    z = torch.randn(1, 4, 16, 16, generator=g)  # small latent: the decoder is fully convolutional
    with torch.no_grad():
        img = OV.decode(sd, z)
    assert img.shape == (1, 3, 128, 128) and torch.isfinite(img).all()


def test_vae_encoder_oracle_structure():
    """encoder half: published parameter count of the SD VAE encoder (34,163,592 + 72 for quant_conv; 83,653,863 for the
    whole AutoencoderKL), 8x downsampling with one-sided padding, posterior sampling."""
    from oracle import vae as OV
    sd = OV.synth_encoder_state_dict(778)
    assert sum(v.numel() for v in sd.values()) == 34_163_664
    assert sum(v.numel() for v in sd.values()) + sum(v.numel() for v in OV.synth_state_dict(777).values()) == 83_653_863
    g = torch.Generator().manual_seed(0)
    x = torch.rand(1, 3, 64, 64, generator=g) * 2 - 1
    with torch.no_grad():
        m = OV.encode_moments(sd, x)
    assert m.shape == (1, 8, 8, 8)
    z = OV.sample_posterior(m, torch.zeros(1, 4, 8, 8))
    assert torch.equal(z, m[:, :4])
