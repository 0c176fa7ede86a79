"""CLIP ViT-H/14 image encoder on the kernel path (reference ip_adapter/ip_adapter.py:78-81, 142-153,
301-317) against the transformers implementation ITSELF (CLIPVisionModelWithProjection, random-initialised from the
ViT-H/14 config with a fixed seed -- no checkpoint exists offline), fp32 on the same GPU, with transformers' own fp16 run
as the noise floor."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _hf(layers=32, seed=0):
    from transformers import CLIPVisionConfig, CLIPVisionModelWithProjection
    cfg = CLIPVisionConfig(hidden_size=1280, intermediate_size=5120, num_hidden_layers=layers, num_attention_heads=16,
                           patch_size=14, image_size=224, projection_dim=1024, hidden_act="gelu")
    torch.manual_seed(seed)
    m = CLIPVisionModelWithProjection(cfg).eval()
    # transformers initialises with std 0.02-ish weights and unit LayerNorms: perturb norms / biases so they matter
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if n.endswith("bias"):
                p.add_(0.02 * torch.randn(p.shape, generator=g))
            elif "norm" in n and n.endswith("weight"):
                p.add_(0.1 * torch.randn(p.shape, generator=g))
    return m


def _rel(a, ref):
    e = (a.float() - ref).abs()
    return e.max().item(), (e.pow(2).mean().sqrt() / ref.pow(2).mean().sqrt().clamp_min(1e-12)).item()


def test_patch_im2col_and_embedding():
    from theatergen_b200 import ops
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 3, 224, 224, generator=g).half().cuda()
    cols = ops.im2col_patch(x, 14, 640)
    ref = torch.nn.functional.unfold(x.float(), 14, stride=14).transpose(1, 2).reshape(2 * 256, 588)
    assert cols.shape == (512, 640) and torch.equal(cols[:, :588].float(), ref) and float(cols[:, 588:].abs().max()) == 0


@pytest.mark.parametrize("layers,B", [(2, 2), (32, 1)])
def test_clip_vision_parity_vs_transformers(layers, B):
    from theatergen_b200.clip_vision import TGCLIPVision
    torch.backends.cuda.matmul.allow_tf32 = False
    m = _hf(layers)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    ours = TGCLIPVision(sd)
    g = torch.Generator().manual_seed(7)
    px = torch.randn(B, 3, 224, 224, generator=g).half().cuda()  # CLIPImageProcessor output is roughly unit-normal
    out = ours(px, output_hidden_states=True)
    assert out.image_embeds.shape == (B, 1024) and len(out.hidden_states) == layers + 1
    m32 = m.cuda().float()
    for p in m32.parameters():  # both arms see the fp16-rounded weights the kernel path consumes
        p.data = p.data.half().float()
    with torch.no_grad():
        r32 = m32(px.float(), output_hidden_states=True)
        r16 = m32.half()(px, output_hidden_states=True)
    for name, a, r, h in (("image_embeds", out.image_embeds, r32.image_embeds, r16.image_embeds),
                          ("hidden_states[-2]", out.hidden_states[-2], r32.hidden_states[-2], r16.hidden_states[-2])):
        mx, rr = _rel(a, r)
        mx16, rr16 = _rel(h, r)
        print(f"CLIP-H {layers} layers {name}: max_abs {mx:.3e} rel_rms {rr:.3e} | transformers fp16 {mx16:.3e} {rr16:.3e}")
        assert rr <= max(1.5 * rr16, 3e-3) and mx <= max(2.0 * mx16, 2e-2 * r.abs().max().item())
