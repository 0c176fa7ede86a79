"""VAE decode on the kernel path (reference call sites models/pipelines.py:468, 162-173) vs the
oracle restatement of diffusers' AutoencoderKL decoder (oracle/vae.py, unpinned third-party arithmetic), plus the small
kernels it adds."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rand(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).half().cuda()


def test_softmax_rows_pointwise_conv_out3():
    from theatergen_b200 import ops
    from theatergen_b200.weights import pack_conv3x3
    x = _rand((300, 4096), 1, 8.0)
    out = ops.softmax_rows(x, 0.0442)
    ref = (x.float() * 0.0442).softmax(-1)
    assert (out.float() - ref).abs().max().item() < 2e-4 and abs(out.float().sum(-1) - 1).max().item() < 2e-3
    y = x.clone()
    ops.softmax_rows(y, 0.0442, out=y)  # in place
    assert torch.equal(y, out)
    z = _rand((2, 4, 64, 64), 2)
    w, b = _rand((4, 4), 3, 0.7), _rand((4,), 4, 0.1)
    pw = ops.pointwise_nchw(z, w, b)
    refp = F.conv2d(z.float(), w.float()[:, :, None, None], b.float())
    assert (pw.float() - refp).abs().max().item() < 3e-3
    h = _rand((1, 64, 128, 128), 5)
    w3, b3 = _rand((3, 128, 3, 3), 6, 1152 ** -0.5), _rand((3,), 7, 0.1)
    o3 = ops.conv_out(h, pack_conv3x3(w3), b3)
    ref3 = F.conv2d(h.float().permute(0, 3, 1, 2), w3.float(), b3.float(), padding=1)
    assert o3.shape == (1, 3, 64, 128) and (o3.float() - ref3).abs().max().item() < 4e-3


def test_groupnorm_large_image():
    """HW = 65536 (256x256) at batch 1: the statistics kernel takes more slabs per image."""
    from theatergen_b200 import ops
    x = _rand((1, 65536, 256), 1) * 1.5 + 0.3
    gamma, beta = _rand((256,), 2) * 0.1 + 1, _rand((256,), 3, 0.1)
    out = ops.groupnorm(x, gamma, beta, 1e-6, silu=True)
    ref = F.silu(F.group_norm(x.float().transpose(1, 2), 32, gamma.float(), beta.float(), 1e-6)).transpose(1, 2)
    assert (out.float() - ref).abs().max().item() < 6e-3


def test_vae_decode_parity():
    from oracle import vae as OV
    from theatergen_b200.vae import TGAutoencoderKL
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    sd = OV.synth_state_dict(777)
    vae = TGAutoencoderKL(sd)
    z = _rand((1, 4, 64, 64), 11)
    img = vae.decode(z, return_dict=False)[0]
    assert img.shape == (1, 3, 512, 512) and vae.decode(z).sample.shape == img.shape
    sd32 = {k: v.half().float().cuda() for k, v in sd.items()}
    sd16 = {k: v.half().cuda() for k, v in sd.items()}
    with torch.no_grad():
        ref32 = OV.decode(sd32, z.float())
        ref16 = OV.decode(sd16, z)
    e = (img.float() - ref32).abs()
    e16 = (ref16.float() - ref32).abs()
    rms = lambda t: t.pow(2).mean().sqrt().item()
    print(f"vae decode: max_abs {e.max().item():.3e} rel_rms {rms(e) / rms(ref32):.3e} | fp16 oracle max_abs "
          f"{e16.max().item():.3e} rel_rms {rms(e16) / rms(ref32):.3e}")
    assert e.max().item() <= max(1.5 * e16.max().item(), 1e-2)
    assert rms(e) / rms(ref32) <= max(1.5 * rms(e16) / rms(ref32), 3e-3)
    assert torch.equal(img, vae.decode(z, return_dict=False)[0])  # deterministic


def test_vae_decode_timing_context():
    """Context numbers (not a parity check): one 512x512 decode on our kernels vs the oracle run by PyTorch eager fp16 on
    the same GPU."""
    import time
    from oracle import vae as OV
    from theatergen_b200.vae import TGAutoencoderKL
    sd = OV.synth_state_dict(777)
    vae = TGAutoencoderKL(sd)
    sd16 = {k: v.half().cuda() for k, v in sd.items()}
    z = _rand((1, 4, 64, 64), 11)
    out = {}
    for name, fn in (("ours", lambda: vae.decode(z, return_dict=False)[0]), ("torch_eager_fp16", lambda: OV.decode(sd16, z))):
        with torch.no_grad():
            fn(); fn()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(5):
                fn()
            torch.cuda.synchronize()
        out[name] = (time.perf_counter() - t0) / 5 * 1e3
    print("VAE decode 512x512, ms:", {k: round(v, 2) for k, v in out.items()})
    assert out["ours"] < 2.0 * out["torch_eager_fp16"]


def test_vae_encode_parity():
    """Encoder moments (mean | logvar) and the posterior sample vs the oracle; one-sided padding of the stride-2 convs."""
    from oracle import vae as OV
    from theatergen_b200 import ops
    from theatergen_b200.vae import TGAutoencoderKL
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    sd = {**OV.synth_state_dict(777), **OV.synth_encoder_state_dict(778)}
    vae = TGAutoencoderKL(sd)
    g = torch.Generator().manual_seed(21)
    img = (torch.rand(1, 3, 512, 512, generator=g) * 2 - 1).half().cuda()
    dist = vae.encode(img).latent_dist
    m = dist.moments
    assert m.shape == (1, 8, 64, 64)
    sd32 = {k: v.half().float().cuda() for k, v in sd.items()}
    sd16 = {k: v.half().cuda() for k, v in sd.items()}
    with torch.no_grad():
        ref32 = OV.encode_moments(sd32, img.float())
        ref16 = OV.encode_moments(sd16, img)
    e, e16 = (m.float() - ref32).abs(), (ref16.float() - ref32).abs()
    rms = lambda t: t.pow(2).mean().sqrt().item()
    print(f"vae encode moments: max_abs {e.max().item():.3e} rel_rms {rms(e) / rms(ref32):.3e} | fp16 oracle max_abs "
          f"{e16.max().item():.3e} rel_rms {rms(e16) / rms(ref32):.3e}")
    assert e.max().item() <= max(1.5 * e16.max().item(), 1e-2)
    assert rms(e) / rms(ref32) <= max(1.5 * rms(e16) / rms(ref32), 3e-3)
    noise = torch.randn(1, 4, 64, 64, generator=g).half().cuda()
    z = dist.sample(noise=noise)
    zref = OV.sample_posterior(m.float(), noise.float())
    assert (z.float() - zref).abs().max().item() < 4e-3 * max(1.0, zref.abs().max().item())
    assert torch.equal(dist.mode(), ops.gaussian_sample(m, None)) and (dist.mode().float() - m[:, :4].float()).abs().max().item() == 0
    # stride-2 im2col with the VAE's (0,1,0,1) padding
    x = _rand((1, 16, 16, 64), 5)
    a = ops.im2col_s2(x, pad_lo=0)
    unf = F.unfold(F.pad(x.float().permute(0, 3, 1, 2), (0, 1, 0, 1)), 3, padding=0, stride=2)
    unf = unf.reshape(1, 64, 9, 64).permute(0, 3, 2, 1).reshape(64, 9 * 64)
    assert torch.equal(a.float(), unf)
