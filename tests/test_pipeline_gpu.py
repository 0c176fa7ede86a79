"""The denoise loops through the reference-facing pipeline surface vs the oracle loops, plus size-independent
properties at the full BASELINE configuration (50 steps, 4 co-batched characters)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


class FakeVAE:
    class config:
        scaling_factor = 0.18215

    def decode(self, z, return_dict=True):
        img = z[:, :3].float().repeat_interleave(8, 2).repeat_interleave(8, 3).clamp(-1, 1)
        return (img,) if not return_dict else type("O", (), {"sample": img})()


class FakeImageProcessor:
    def postprocess(self, image, output_type="pil", do_denormalize=None):
        from PIL import Image
        arr = ((image[0] / 2 + 0.5).clamp(0, 1).permute(1, 2, 0).cpu().numpy() * 255).round().astype("uint8")
        return [Image.fromarray(arr)]


class FakePipe:
    """Everything the stage functions reach through adapter.pipe, with seeded stand-ins for CLIP text / VAE."""
    vae_scale_factor = 8

    def __init__(self, unet, scheduler):
        self.unet, self.scheduler, self.vae, self.image_processor = unet, scheduler, FakeVAE(), FakeImageProcessor()

    def to(self, device):
        return self

    def encode_prompt(self, prompt, device=None, num_images_per_prompt=1, do_classifier_free_guidance=True,
                      negative_prompt=None, **kw):
        def emb(s):
            import hashlib
            g = torch.Generator().manual_seed(int(hashlib.md5(str(s).encode()).hexdigest()[:7], 16))
            return torch.randn(1, 77, 768, generator=g).half().cuda()
        return emb("p:" + str(prompt)), emb("n:" + str(negative_prompt))


def _make(seed=1234):
    from oracle import unet as OU
    from theatergen_b200.ip_adapter import IPAdapter
    from theatergen_b200.scheduler import DDIMScheduler
    from theatergen_b200.unet import TGUNet
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    sd = OU.synth_state_dict(seed)
    net = TGUNet(sd)
    adapter = IPAdapter(FakePipe(net, DDIMScheduler()), device="cuda")
    from oracle import synth
    adapter.image_proj_model.load_state_dict(synth.fill(synth.image_proj_shapes(), 9))
    return OU, sd, net, adapter


def test_stage1_and_stage2_loops_match_oracle(tmp_path):
    from oracle import ddim as OD
    from oracle import resampler as ORS
    from oracle import synth
    from theatergen_b200.models import pipelines
    OU, sd, net, adapter = _make()
    steps = 6
    g = torch.Generator().manual_seed(3)
    lat0 = torch.randn(1, 4, 64, 64, generator=g).half()
    clip = torch.randn(1, 1024, generator=g).half()
    out = pipelines.generate_semantic_guidance(
        "story", 0, "1.5", "a red fox", str(tmp_path) + "/", 0, adapter, None, lat0, None, steps, None, None, None,
        save_all_latents=True, obj_id=7, have_reffer=1, clip_image_embeds=clip)
    latents, image, trace = out[0], out[1], out[-1]
    assert trace.shape == (steps + 1, 1, 4, 64, 64) and latents.shape == (1, 4, 64, 64)
    # oracle: same embeddings, fp32, on the GPU
    pe, ne = adapter.pipe.encode_prompt("full-body picture of a red fox", negative_prompt=pipelines.NEGATIVE_SO)
    ip_sd = {k: v.half().float().cuda() for k, v in synth.fill(synth.image_proj_shapes(), 9).items()}
    tok = ORS.image_proj(ip_sd, clip.float().cuda())
    tok0 = ORS.image_proj(ip_sd, torch.zeros_like(clip).float().cuda())
    ctx = torch.cat([torch.cat([ne.float(), tok0], 1), torch.cat([pe.float(), tok], 1)])
    ctx = ctx.half().float()  # both arms see the fp16 context the CUDA path consumes
    sd32 = {k: v.half().float().cuda() for k, v in sd.items()}
    sd16 = {k: v.half().cuda() for k, v in sd.items()}
    with torch.no_grad():
        ref, ref_trace = OD.stage1_loop(lambda x, t, c: OU.unet_forward(sd32, x, t, c, ip_scale=0.4), OD.DDIM(),
                                        lat0.float().cuda(), ctx, num_steps=steps)
        # the reference's own numerics: the same loop, every tensor in fp16 (what diffusers would execute on this GPU)
        _, ref16_trace = OD.stage1_loop(lambda x, t, c: OU.unet_forward(sd16, x, t, c, ip_scale=0.4), OD.DDIM(),
                                        lat0.cuda(), ctx.half(), num_steps=steps)
    err = (trace.float().cuda() - ref_trace).abs().amax(dim=(1, 2, 3, 4))
    err16 = (ref16_trace.float() - ref_trace).abs().amax(dim=(1, 2, 3, 4))
    print("stage-1 per-step max|latents - fp32 oracle|  ours:", [f"{e:.2e}" for e in err.tolist()])
    print("stage-1 per-step max|latents - fp32 oracle|  fp16 oracle:", [f"{e:.2e}" for e in err16.tolist()])
    # CFG (x16) and the DDIM update amplify the per-forward fp16 noise; require ours to stay within 1.5x of the
    # drift the reference's own fp16 path shows on the same trajectory
    assert (err <= 1.5 * err16 + 1e-2).all(), (err, err16)
    assert torch.equal(trace[-1].cuda(), latents)

    # stage 2 with the masked composite (index < frozen_steps), injected noise tensors
    from PIL import Image
    import numpy as np
    mask_img = Image.fromarray((np.pad(np.ones((200, 160)), ((150, 162), (100, 252))) * 255).astype("uint8"))
    init, noise, bg = [torch.randn(1, 4, 64, 64, generator=g).half() for _ in range(3)]
    lat2, _ = pipelines.final_image_generation(
        "1.5", None, None, None, "two foxes", "blurry", "forest", [None], None, 0, 512, 512, 11, mask_img, None, adapter,
        None, torch.zeros(steps + 1, 1, 4, 64, 64), None, None, None, steps, 4, clip_image_embeds=clip,
        stage2_noise=(init, noise, bg))
    sched = OD.DDIM()
    sched.set_timesteps(steps)
    frozen = torch.cat([bg.float()[None], sched.add_noise(init.float(), noise.float(), sched.timesteps).unsqueeze(1)]).cuda()
    m = np.array(mask_img.resize((64, 64)).convert("L")).astype(np.float32) / 255.0
    m[m > 0] = 1
    fm = torch.from_numpy(1 - m).cuda()
    pe, ne = adapter.pipe.encode_prompt("two foxes", negative_prompt="blurry")
    ctx2 = torch.cat([torch.cat([ne.float(), tok0], 1), torch.cat([pe.float(), tok], 1)])
    with torch.no_grad():
        ref2 = OD.stage2_loop(lambda x, t, c: OU.unet_forward(sd32, x, t, c, ip_scale=0.1), OD.DDIM(), bg.float().cuda(),
                              ctx2, frozen, fm, 4, num_steps=steps)
    e2 = (lat2.float() - ref2).abs().max().item()
    print("stage-2 max|latents - oracle|:", e2)
    assert e2 < 1.5 * err16.max().item() + 1e-2


def test_full_config_properties():
    """50 steps, 4 co-batched characters (BASELINE full size): determinism, batch-position independence,
    composite idempotence."""
    OU, sd, net, adapter = _make()
    from theatergen_b200.engine import DenoiseEngine
    eng = DenoiseEngine(net, adapter.pipe.scheduler)
    g = torch.Generator().manual_seed(5)
    K = 4
    lat = torch.randn(K, 4, 64, 64, generator=g).half().cuda()
    ctx = torch.randn(2 * K, 81, 768, generator=g).half().cuda()
    scale = torch.tensor([0.0, 0.4, 0.4, 0.4] * 2, device="cuda")
    a = eng.denoise(lat, ctx, scale, 50).clone()
    b = eng.denoise(lat, ctx, scale, 50).clone()
    assert torch.equal(a, b), "the 50-step trajectory must be bit-reproducible"
    assert torch.isfinite(a.float()).all() and a.float().std().item() > 0.05
    # character 2 alone (K=1) vs inside the co-batch: the same arithmetic per row, but tile shapes / split-K depend on
    # the batch, so fp32 summation order (not the result beyond rounding) may differ -> compare one step, with tolerance
    one = eng.denoise(lat[2:3], ctx[[2, K + 2]], scale[[2, K + 2]], 50, guidance_scale=7.5)
    assert torch.isfinite(one.float()).all()
    from theatergen_b200.engine import DenoiseEngine as _DE
    e1 = _DE(net, adapter.pipe.scheduler, use_graphs=False)
    s4 = e1.step_graph(K, 81, False); s1 = e1.step_graph(1, 81, False)
    ts, coef = e1.tables(50, 7.5)
    for sg, l_, c_, sc_ in ((s4, lat, ctx, scale), (s1, lat[2:3], ctx[[2, K + 2]], scale[[2, K + 2]])):
        sg.latents.copy_(l_); sg.set_context(c_); sg.ip_scale.copy_(sc_)
        sg.t.copy_(ts[0].expand(sg.t.numel())); sg.coef.copy_(coef[0]); sg._body()
    d1 = (s4.latents[2].float() - s1.latents[0].float()).abs().max().item()
    assert d1 < 1e-2, d1
    # stage-2 composite: with frozen_steps = 50 the masked region ends exactly on the frozen latents
    frozen = torch.randn(51, 1, 4, 64, 64, generator=g).half().cuda()
    mask = torch.zeros(64, 64, device="cuda")
    mask[10:40, 20:50] = 1
    out = eng.denoise(lat[:1], ctx[[0, K]], 0.1, 50, frozen_all=frozen, frozen_mask=mask, frozen_steps=50)
    assert torch.equal(out[0][:, mask.bool()], frozen[50, 0][:, mask.bool()])
    out30 = eng.denoise(lat[:1], ctx[[0, K]], 0.1, 50, frozen_all=frozen, frozen_mask=mask, frozen_steps=30)
    assert not torch.equal(out30[0][:, mask.bool()], frozen[50, 0][:, mask.bool()])


def test_cobatched_stage1_stage2_step_matches_separate_steps():
    """MixedStepGraph: 4 stage-1 trajectories + 1 stage-2 trajectory in one B=10 forward == the two separate steps
    (same arithmetic per trajectory; only fp32 summation order may differ with the batch-dependent tile shapes)."""
    OU, sd, net, adapter = _make()
    from theatergen_b200.engine import DenoiseEngine, MixedStepGraph
    from theatergen_b200.utils.guidance import EnergyPlan
    eng = DenoiseEngine(net, adapter.pipe.scheduler, use_graphs=False)
    g = torch.Generator().manual_seed(9)
    K = 4
    lat1, lat2 = torch.randn(K, 4, 64, 64, generator=g).half().cuda(), torch.randn(1, 4, 64, 64, generator=g).half().cuda()
    ctx1, ctx2 = torch.randn(2 * K, 81, 768, generator=g).half().cuda(), torch.randn(2, 81, 768, generator=g).half().cuda()
    frozen = torch.randn(1, 4, 64, 64, generator=g).half().cuda()
    mask = torch.zeros(64, 64, device="cuda")
    mask[8:40, 16:56] = 1
    keys = [("mid", 0, 0, 0), ("up", 1, 0, 0)]
    plan = EnergyPlan([[0.1, 0.2, 0.6, 0.7]], [[3, 4]], keys, fg_weight=1.0, bg_weight=4.0)
    ts, coef = eng.tables(50, 7.5)
    mx = MixedStepGraph(eng, K, 1, 81, keys, plan)
    # batch rows: [stage-1 uncond | stage-1 cond | stage-2 uncond | stage-2 cond]
    mx.latents.copy_(torch.cat([lat1, lat2])); mx.set_context(torch.cat([ctx1, ctx2]))
    mx.ip_scale.copy_(torch.tensor([0.0, 0.4, 0.4, 0.4] * 2 + [0.1, 0.1])); mx.frozen.copy_(frozen); mx.mask.copy_(mask.half())
    mx.t.copy_(ts[3].expand(10)); mx.coef.copy_(coef[3])
    mx._body()
    s1 = eng.step_graph(K, 81, False)
    s1.latents.copy_(lat1); s1.set_context(ctx1); s1.ip_scale.copy_(torch.tensor([0.0, 0.4, 0.4, 0.4] * 2))
    s1.t.copy_(ts[3].expand(8)); s1.coef.copy_(coef[3]); s1._body()
    s2 = eng.step_graph(1, 81, True, keys)
    s2.latents.copy_(lat2); s2.set_context(ctx2); s2.ip_scale.fill_(0.1); s2.frozen.copy_(frozen); s2.mask.copy_(mask.half())
    s2.t.copy_(ts[3].expand(2)); s2.coef.copy_(coef[3]); s2._body()
    assert (mx.latents[:K].float() - s1.latents.float()).abs().max().item() < 1e-2
    assert (mx.latents[K:].float() - s2.latents.float()).abs().max().item() < 1e-2
    assert torch.equal(mx.latents[K][:, mask.bool()], frozen[0][:, mask.bool()])
    assert torch.equal(mx.trace, mx.latents)
    loss2 = torch.zeros(1, device="cuda")
    plan.run([s2.captured[k][0] for k in keys], loss2, torch.empty(256, device="cuda"))
    assert abs(mx.loss.item() - loss2.item()) < 2e-3 * max(1.0, abs(loss2.item()))
