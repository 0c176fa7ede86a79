import os, sys, torch
sys.path.insert(0, os.getcwd())
from theatergen_b200 import ops
M, N, K = 40960, 2560, 320
a = torch.randn(M, K, device="cuda").half(); w = (torch.randn(N, K, device="cuda") * 0.05).half(); b = torch.randn(N, device="cuda").half()
out = torch.empty(M, N // 2, device="cuda").half()
for _ in range(4):
    ops.gemm(a, w, b, act=ops.TG_ACT_GEGLU, out=out)
torch.cuda.synchronize()
