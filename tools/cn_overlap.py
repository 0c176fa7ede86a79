import os, sys, torch
sys.path.insert(0, os.getcwd())
from theatergen_b200 import synth
from theatergen_b200.controlnet import TGControlNet
from theatergen_b200.engine import DenoiseEngine, MixedStepGraph
from theatergen_b200.unet import TGUNet
dev = torch.device("cuda")
net = TGUNet(synth.synth_state_dict(dev, 1234), dev)
cn = TGControlNet(synth.synth_controlnet_state_dict(dev, 4321), dev)
eng = DenoiseEngine(net, controlnet=cn)
def t(g, n=20):
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): g.replay()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n
res = {}
for name, ov, use_cn in (("B10 no ControlNet", "1", False), ("B10 + ControlNet side stream", "1", True), ("B10 + ControlNet in line", "0", True)):
    os.environ["TG_CN_OVERLAP"] = ov
    g = MixedStepGraph(eng, 4, 1, 81, None, None, controlnet=use_cn).capture()
    g.ctx.normal_(); g.latents.normal_(); g.ip_scale.fill_(0.4)
    if use_cn: g.ctrl_ctx.normal_(); g.cond_emb.normal_()
    g.t.fill_(501.0); g.coef.copy_(torch.tensor([7.5, .7, .7, .75, .66, 0, 0, 0]))
    res[name] = t(g)
    print(f"{name:32s} {res[name]:.3f} ms per step")
