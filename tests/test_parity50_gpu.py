"""50-step parity at BASELINE.json configs 2 and 3 (models/pipelines.py:406-453, 742-835): the product loops
(DenoiseEngine, CUDA-graph steps, co-batched characters, ControlNet fused into the stage-2 forward, masked composite,
guidance-energy evaluation) against the oracle loops run in fp32 on the same GPU, with the oracle's own fp16 run -- the
reference's numerics -- as the noise floor.  The measured drift is printed; bench.py reports the same numbers in its
`parity` key.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

STEPS = 50
XYWH = [[278, 251, 157, 185], [358, 328, 150, 180], [168, 4, 100, 80], [4, 331, 80, 70]]  # CMIGBench dialogue 4, turn 4


@pytest.fixture(scope="module")
def world():
    from oracle import controlnet as OC
    from oracle import parity as P
    from oracle import unet as OU
    from theatergen_b200.controlnet import TGControlNet
    from theatergen_b200.engine import DenoiseEngine
    from theatergen_b200.unet import TGUNet
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    sdu, sdc = OU.synth_state_dict(1234), OC.synth_state_dict(4321)
    net, cn = TGUNet(sdu), TGControlNet(sdc)
    eng = DenoiseEngine(net, controlnet=cn)
    g = torch.Generator().manual_seed(2024)
    K = 4
    w = dict(P=P, eng=eng, cn=cn, K=K,
             sdu={k: v.cuda() for k, v in sdu.items()}, sdc={k: v.cuda() for k, v in sdc.items()},
             lat=torch.randn(K, 4, 64, 64, generator=g).half().cuda(),
             ctx=torch.randn(2 * K, 81, 768, generator=g).half().cuda(),
             scales=[0.0, 0.4, 0.4, 0.4])
    # the oracle's stage-1 trajectories of the four characters, fp32 and fp16, once for both configs
    w["ref32"] = P.stage1_traces(w["sdu"], w["lat"], w["ctx"], w["scales"], STEPS, torch.float32)
    w["ref16"] = P.stage1_traces(w["sdu"], w["lat"], w["ctx"], w["scales"], STEPS, torch.float16)
    return w


def _ours_stage1(w, K):
    eng = w["eng"]
    idx = list(range(K)) + [w["K"] + k for k in range(K)]
    trace = torch.empty((STEPS + 1, K, 4, 64, 64), dtype=torch.float16).pin_memory()
    scale = torch.tensor(w["scales"][:K] * 2, dtype=torch.float32, device="cuda")
    eng.denoise(w["lat"][:K], w["ctx"][idx].contiguous(), scale, num_steps=STEPS, trace_host=trace)
    torch.cuda.synchronize()
    return trace.cuda()


def _check(name, s):
    print(f"\n{name}: max|latents - fp32 oracle| over {s['steps']} steps = {s['max_abs_latent_vs_fp32_oracle']:.4e} "
          f"(fp16 oracle: {s['fp16_oracle_gap']:.4e}); after step 1: {s['first_step_max_abs']:.3e} "
          f"({s['first_step_fp16_oracle_gap']:.3e}); final: {s['final_max_abs']:.3e} ({s['final_fp16_oracle_gap']:.3e}); "
          f"final rms {s['rms_final']:.3e} ({s['rms_final_fp16_oracle']:.3e})")
    print("  per-step ours      :", s["per_step_max_abs"][::5])
    print("  per-step fp16 oracle:", s["per_step_fp16_oracle_gap"][::5])
    ours, ref = torch.tensor(s["per_step_max_abs"]), torch.tensor(s["per_step_fp16_oracle_gap"])
    # per-step drift no worse than 1.5x what the reference's own fp16 arithmetic shows on the same trajectories
    assert (ours <= 1.5 * ref + 5e-3).all(), (ours, ref)
    assert s["rms_final"] <= 1.5 * s["rms_final_fp16_oracle"] + 1e-3


def test_config2_two_characters_50_steps(world):
    w = world
    ours = _ours_stage1(w, 2)
    s = w["P"].summarize(ours, w["ref32"][:, :2], w["ref16"][:, :2])
    _check("config 2 (2 characters, IP scales 0.0 / 0.4, 50 steps)", s)
    w["ours2"] = ours


def test_config3_four_characters_stage2_composite_energy_50_steps(world):
    from theatergen_b200.utils.guidance import EnergyPlan
    w, P = world, world["P"]
    ours4 = _ours_stage1(w, 4)
    s = P.summarize(ours4, w["ref32"], w["ref16"])
    _check("config 3 stage 1 (4 co-batched characters, 50 steps)", s)
    if "ours2" in w:
        # Batch POSITION never changes a result (tests/test_pipeline_gpu.py: co-batch == solo at equal batch size), but the
        # batch SIZE selects tile widths / split-K per GEMM shape, i.e. a different fp32 summation order -- like cuBLAS
        # heuristics do for the reference.  The two runs of the same characters must stay inside the fp16 noise floor.
        d = (ours4[:, :2].float() - w["ours2"].float()).abs().amax(dim=(1, 2, 3, 4))
        e16 = (w["ref16"][:, :2].float() - w["ref32"][:, :2]).abs().amax(dim=(1, 2, 3, 4))
        print("  K=4 vs K=2 co-batch, same characters, per-step max diff:", [round(float(v), 5) for v in d[::5]])
        assert (d <= e16 + 5e-3).all()

    # stage 2: ControlNet + composite on every step + capture of the guidance maps + energy for index < 30
    g = torch.Generator().manual_seed(77)
    lat2 = torch.randn(1, 4, 64, 64, generator=g).half().cuda()
    ctx2 = torch.randn(2, 81, 768, generator=g).half().cuda()
    x0, nz = torch.randn(1, 4, 64, 64, generator=g), torch.randn(1, 4, 64, 64, generator=g)
    hint = (torch.rand(1, 1, 512, 512, generator=g) > 0.93).half().repeat(1, 3, 1, 1).cuda()
    ctrl_ctx = ctx2[:, :77].contiguous()
    sched = w["eng"].scheduler
    sched.set_timesteps(STEPS)
    frozen = torch.cat([lat2.float().cpu()] + [sched.add_noise(x0, nz, torch.tensor([int(t)])) for t in sched.timesteps])
    frozen = frozen.half().view(STEPS + 1, 1, 4, 64, 64).cuda()
    mask = torch.zeros(64, 64)
    for x, y, ww, hh in XYWH:
        mask[round(y / 8):round(y / 8) + round(hh / 8), round(x / 8):round(x / 8) + round(ww / 8)] = 1.0
    mask = mask.cuda()
    bboxes = [[x / 512, y / 512, (x + ww) / 512, (y + hh) / 512] for x, y, ww, hh in XYWH]
    positions = [[5, 6], [9], [12, 13, 14], [17]]
    kw = dict(fg_top_p=0.2, bg_top_p=0.2, fg_weight=1.0, bg_weight=4.0)
    plan = EnergyPlan(bboxes, positions, P.KEYS, **kw)
    loss, ws = torch.zeros(1, device="cuda"), torch.empty(256, device="cuda")
    energies = []

    def on_step(i, sg):
        if i < 30:
            plan.run([sg.captured[k][0] for k in P.KEYS], loss, ws)
            energies.append(loss.item())

    trace = torch.empty((STEPS + 1, 1, 4, 64, 64), dtype=torch.float16).pin_memory()
    emb = w["cn"].embed_condition(hint)
    w["eng"].denoise(lat2, ctx2, 0.1, num_steps=STEPS, frozen_all=frozen, frozen_mask=mask, frozen_steps=STEPS,
                     trace_host=trace, capture_keys=P.KEYS, on_step=on_step, control_ctx=ctrl_ctx, cond_emb=emb)
    torch.cuda.synchronize()
    en = dict(bboxes=bboxes, positions=positions, kwargs=kw, upto=30)
    r32, e32 = P.stage2_trace(w["sdu"], w["sdc"], lat2, ctx2, ctrl_ctx, hint, frozen, mask, STEPS, STEPS, torch.float32,
                              energy=en)
    r16, e16 = P.stage2_trace(w["sdu"], w["sdc"], lat2, ctx2, ctrl_ctx, hint, frozen, mask, STEPS, STEPS, torch.float16,
                              energy=en)
    s2 = P.summarize(trace.cuda(), r32, r16)
    _check("config 3 stage 2 (ControlNet + composite + capture, 50 steps)", s2)
    # inside the layout mask the composite pins the latents to the frozen trajectory: exact
    mb = mask.bool()
    assert torch.equal(trace[-1].cuda()[0][:, mb], frozen[STEPS][0][:, mb])
    de = [abs(a - b) / max(abs(b), 1e-6) for a, b in zip(energies, e32)]
    de16 = [abs(a - b) / max(abs(b), 1e-6) for a, b in zip(e16, e32)]
    print(f"  guidance energy, 30 evaluations: max rel diff vs fp32 oracle {max(de):.3e} (fp16 oracle {max(de16):.3e}); "
          f"energy[0] ours {energies[0]:.6f} oracle {e32[0]:.6f}")
    assert len(energies) == 30 and abs(energies[0] - e32[0]) <= 2e-3 * abs(e32[0])
    assert max(de) <= 1.5 * max(de16) + 5e-3
