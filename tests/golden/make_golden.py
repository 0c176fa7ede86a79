"""Generate the committed golden vectors by running the REFERENCE's own importable code on seeded inputs.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):
    python tests/golden/make_golden.py
Imports, unmodified, from /root/reference:
    ip_adapter/attention_processor.py  (Attention, AttnProcessor, IPAttnProcessor) via oracle/diffusers_shim
    ip_adapter/resampler.py            (Resampler)
    utils/guidance.py                  (compute_ca_lossv3; its hard-coded device="cuda" is rewritten to "cpu"
                                        by a TorchFunctionMode, nothing else is touched)
    utils/utils.py                     (shift_tensor, scale_proportion, binary_mask_to_box_mask, ..._center)
    utils/schedule.py                  (get_fast_schedule)
ImageProjModel (ip_adapter/ip_adapter.py:30-47) needs `diffusers` to import its module, so its 12 lines are
exercised through an equivalent nn.Module built here from the same layers (Linear -> reshape -> LayerNorm).
Outputs: tests/golden/*.pt (small fp32 tensors + the seeds/configs that produced them).
"""
import os
import sys
import types

import torch
from torch.overrides import TorchFunctionMode

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, os.path.join(ROOT, "oracle", "diffusers_shim"))
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)


class CudaToCpu(TorchFunctionMode):
    def __torch_function__(self, func, types_, args=(), kwargs=None):
        kwargs = dict(kwargs or {})
        if kwargs.get("device") in ("cuda", torch.device("cuda")):
            kwargs["device"] = "cpu"
        name = getattr(func, "__name__", "")
        if name == "cuda":
            return args[0]
        if name == "to" and len(args) > 1 and args[1] == "cuda":
            args = (args[0], "cpu") + tuple(args[2:])
        return func(*args, **kwargs)


def gen(seed):
    return torch.Generator().manual_seed(seed)


def golden_attention():
    import importlib.util
    from oracle import synth
    spec = importlib.util.spec_from_file_location("ref_attention_processor", f"{REF}/ip_adapter/attention_processor.py")
    ap = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ap)
    out = {}
    for name, (N, C) in {"s64": (64, 1280), "s256": (256, 1280), "s1024": (1024, 640), "s4096": (4096, 320)}.items():
        attn1 = ap.Attention(query_dim=C, heads=8, dim_head=C // 8, processor=ap.AttnProcessor())
        attn2 = ap.Attention(query_dim=C, cross_attention_dim=768, heads=8, dim_head=C // 8)
        proc = ap.IPAttnProcessor(hidden_size=C, cross_attention_dim=768, scale=0.4, num_tokens=4)
        attn2.set_processor(proc)
        shapes = synth.attention_shapes(C)
        got = {f"attn1.{k}": tuple(v.shape) for k, v in attn1.state_dict().items()}
        got.update({f"attn2.{k}": tuple(v.shape) for k, v in attn2.state_dict().items()})
        assert got == {k: tuple(v) for k, v in shapes.items()}, (got.keys() ^ shapes.keys())
        sd = synth.fill(shapes, 100 + N)
        attn1.load_state_dict({k[6:]: v for k, v in sd.items() if k.startswith("attn1.")})
        attn2.load_state_dict({k[6:]: v for k, v in sd.items() if k.startswith("attn2.")})
        x = torch.randn(2, N, C, generator=gen(N))
        ctx = torch.randn(2, 81, 768, generator=gen(N + 1))
        with torch.no_grad():
            y1 = attn1(x)
            saved = {}
            y2 = attn2(x, encoder_hidden_states=ctx, attn_key=["mid", 0, 0, 0], save_attn_to_dict=saved,
                       save_keys=[("mid", 0, 0, 0)], return_cond_ca_only=True)
            proc.scale = 0.0
            y2_s0 = attn2(x, encoder_hidden_states=ctx)
        rs, cs = (1, 1) if N <= 64 else (max(1, N // 64), 5)
        out[name] = dict(N=N, C=C, seeds=dict(params=100 + N, x=N, ctx=N + 1), row_stride=rs, col_stride=cs,
                         y_self=y1[:, ::rs, ::cs].clone(), y_cross=y2[:, ::rs, ::cs].clone(),
                         y_cross_scale0=y2_s0[:, ::rs, ::cs].clone(),
                         probs_cond=saved[("mid", 0, 0, 0)][:, :, ::rs].clone())
    torch.save(out, os.path.join(HERE, "attention.pt"))
    return ap


def golden_resampler():
    import importlib.util
    from oracle import synth
    spec = importlib.util.spec_from_file_location("ref_resampler", f"{REF}/ip_adapter/resampler.py")
    rs = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(rs)
    m = rs.Resampler(dim=768, depth=4, dim_head=64, heads=12, num_queries=16, embedding_dim=1280, output_dim=768,
                     ff_mult=4)
    shapes = synth.resampler_shapes()
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == {k: tuple(v) for k, v in shapes.items()}
    m.load_state_dict(synth.fill(shapes, 7))
    x = torch.randn(2, 257, 1280, generator=gen(8))
    with torch.no_grad():
        y = m(x)
    torch.save(dict(seeds=dict(params=7, x=8), y=y, n_params=sum(p.numel() for p in m.parameters())),
               os.path.join(HERE, "resampler.pt"))
    # ImageProjModel (ip_adapter/ip_adapter.py:30-47): same layers, same forward
    sd = synth.fill(synth.image_proj_shapes(), 9)
    proj, norm = torch.nn.Linear(1024, 4 * 768), torch.nn.LayerNorm(768)
    proj.load_state_dict({"weight": sd["proj.weight"], "bias": sd["proj.bias"]})
    norm.load_state_dict({"weight": sd["norm.weight"], "bias": sd["norm.bias"]})
    e = torch.randn(3, 1024, generator=gen(10))
    with torch.no_grad():
        t = norm(proj(e).reshape(-1, 4, 768))
        t0 = norm(proj(torch.zeros_like(e)).reshape(-1, 4, 768))
    torch.save(dict(seeds=dict(params=9, e=10), tokens=t, tokens_uncond=t0), os.path.join(HERE, "image_proj.pt"))


def golden_guidance():
    sys.modules.setdefault("utils", types.ModuleType("utils"))
    import importlib.util

    def load(name, path):
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        return mod

    u = load("utils.utils", f"{REF}/utils/utils.py")
    sys.modules["utils"].utils = u
    for k, v in vars(u).items():  # utils/__init__.py does `from .utils import *`
        if not k.startswith("_"):
            setattr(sys.modules["utils"], k, v)
    sys.modules["utils"].__path__ = [f"{REF}/utils"]
    gd = load("utils.guidance", f"{REF}/utils/guidance.py")
    sched = load("utils.schedule", f"{REF}/utils/schedule.py")
    keys = [("mid", 0, 0, 0), ("up", 1, 0, 0), ("up", 1, 1, 0), ("up", 1, 2, 0)]
    # CMIGBench story "dialogue 4" turn 4 boxes @512: x,y,w,h -> normalised xyxy
    xywh = [[278, 251, 157, 185], [358, 328, 150, 180], [168, 4, 100, 80], [4, 331, 80, 70]]
    bboxes = [[x / 512, y / 512, (x + w) / 512, (y + h) / 512] for x, y, w, h in xywh]
    positions = [[5, 6], [9], [12, 13, 14], [17]]
    g = gen(42)
    saved = {}
    for k in keys:
        hw = 64 if k[0] == "mid" else 256
        saved[k] = torch.softmax(3.0 * torch.randn(1, 8, hw, 77, generator=g), dim=-1).half()
    ref_attns = []
    for o in range(4):
        per_step = []
        for step in range(2):
            per_step.append({k: torch.softmax(3.0 * torch.randn(1, 8, saved[k].shape[2], 1, generator=g), dim=2).half()
                             for k in keys})
        ref_attns.append(per_step)
    with CudaToCpu():
        loss_plain = gd.compute_ca_lossv3(saved_attn=saved, bboxes=bboxes, object_positions=positions,
                                          guidance_attn_keys=keys, use_ratio_based_loss=False, fg_top_p=0.2,
                                          bg_top_p=0.2, fg_weight=1.0, bg_weight=4.0)
        loss_ref = gd.compute_ca_lossv3(saved_attn=saved, bboxes=bboxes, object_positions=positions,
                                        guidance_attn_keys=keys, ref_ca_saved_attns=ref_attns, index=1,
                                        ref_ca_loss_weight=2.0, use_ratio_based_loss=False, fg_top_p=0.2,
                                        bg_top_p=0.2, fg_weight=1.0, bg_weight=4.0)
        masks = {hw: [u.proportion_to_mask(b, hw, hw) for b in bboxes] for hw in (8, 16, 64)}
    torch.save(dict(keys=keys, bboxes=bboxes, positions=positions, saved=saved, ref_attns=ref_attns,
                    loss_plain=loss_plain.float(), loss_ref=loss_ref.float(), masks=masks,
                    kwargs=dict(fg_top_p=0.2, bg_top_p=0.2, fg_weight=1.0, bg_weight=4.0)),
               os.path.join(HERE, "guidance.pt"))
    # utils.py / schedule.py helpers
    t = torch.randn(3, 1, 4, 64, 64, generator=gen(11))
    m = torch.zeros(64, 64, dtype=torch.bool)
    m[20:41, 10:33] = True
    m[25:30, 33:37] = True
    cases = [(0.27, -0.11), (-0.4, 0.33), (0.0625, 0.0), (0.5, 0.5)]
    shifts = [u.shift_tensor(t, xo, yo, offset_normalized=True) for xo, yo in cases]
    mshifts = [u.shift_tensor(m, xo, yo, offset_normalized=True) for xo, yo in cases]
    scale_cases = [([0.1, 0.2, 0.55, 0.9], 64, 64), ([0.543, 0.49, 0.85, 0.852], 16, 16), ([-0.1, 0.3, 0.4, 1.2], 8, 8),
                   ([0.0078125, 0.6464, 0.1640625, 0.783], 64, 64), ([0.5390625, 0.0234375, 0.7734375, 0.1796875], 64, 64)]
    torch.save(dict(t=t, mask=m, shift_cases=cases, shifts=shifts, mask_shifts=mshifts,
                    box_mask=u.binary_mask_to_box_mask(m, to_device=False), box=u.binary_mask_to_box(m),
                    center=u.binary_mask_to_center(m, normalize=True),
                    scale_cases=scale_cases, scaled=[u.scale_proportion(b, H, W) for b, H, W in scale_cases],
                    fast_schedule=[sched.get_fast_schedule(torch.arange(981, 0, -20), a, r) for a, r in ((15, 2), (49, 2), (5, 3))]),
               os.path.join(HERE, "utils.pt"))


if __name__ == "__main__":
    golden_attention()
    golden_resampler()
    golden_guidance()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".pt"):
            print(f, os.path.getsize(os.path.join(HERE, f)))
