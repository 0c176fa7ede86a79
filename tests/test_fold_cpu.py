"""Host-side algebra of the late round-2 folds (no GPU): the folded cross-attention operands (unet._fold_pair) and the
ff.net.2 + proj_out weight concatenation reproduce the reference's op order (ip_adapter/attention_processor.py:461-553,
ip_adapter/attention.py:339-367, transformer_2d.py:330-342) in plain torch."""
import torch

from theatergen_b200 import unet as U

H = U.HEADS


def test_fold_pair_reproduces_decoupled_cross_attention():
    torch.manual_seed(0)
    C, T, n_ip, B, HW, D = 80, 9, 2, 2, 16, 32
    d, n_text, S = C // H, T - n_ip, 96
    wq, wo = torch.randn(C, C) * 0.2, torch.randn(C, C) * 0.2
    wk, wv, wki, wvi = [torch.randn(C, D) * 0.2 for _ in range(4)]
    ctx, x, ips = torch.randn(B, T, D), torch.randn(B, HW, C), torch.tensor([0.4, 0.7])

    def heads(t):
        return t.view(B, -1, H, d).transpose(1, 2)
    q = heads(x @ wq.T)
    ct, ci = ctx[:, :n_text], ctx[:, n_text:]
    a = torch.softmax(q @ heads(ct @ wk.T).transpose(-1, -2) * d ** -0.5, -1) @ heads(ct @ wv.T)
    a = a + ips.view(B, 1, 1, 1) * (torch.softmax(q @ heads(ci @ wki.T).transpose(-1, -2) * d ** -0.5, -1) @ heads(ci @ wvi.T))
    ref = a.transpose(1, 2).reshape(B, HW, C) @ wo.T

    t = dict(C=C, q2w=wq, o2w=wo)
    fkt, fki = U._fold_pair(t, wk, wv).float(), U._fold_pair(t, wki, wvi).float()
    assert fkt.shape == (2 * H * C, D)
    kf, vf = torch.zeros(B, H, S, C), torch.zeros(B, C, H, S)
    pr = (ct.reshape(B * n_text, D) @ fkt.T).view(B, n_text, 2, H, C)
    kf[:, :, :n_text] = pr[:, :, 0].permute(0, 2, 1, 3)
    vf[..., :n_text] = pr[:, :, 1].permute(0, 3, 2, 1)
    pi = (ci.reshape(B * n_ip, D) @ fki.T).view(B, n_ip, 2, H, C)
    kf[:, :, 80:80 + n_ip] = pi[:, :, 0].permute(0, 2, 1, 3)
    vf[..., 80:80 + n_ip] = pi[:, :, 1].permute(0, 3, 2, 1)
    sc = torch.einsum("bmc,bnc->bmn", x, kf.view(B, H * S, C)).view(B, HW, H, S)  # scores * log2(e), as the epilogue sees them
    P = torch.zeros_like(sc)
    P[..., :n_text] = torch.softmax(sc[..., :n_text] * 0.6931471805599453, -1)
    P[..., 80:80 + n_ip] = torch.softmax(sc[..., 80:80 + n_ip] * 0.6931471805599453, -1) * ips.view(B, 1, 1, 1)
    out = torch.einsum("bmn,bcn->bmc", P.view(B, HW, H * S), vf.view(B, C, H * S))
    # the folded operands are stored in fp16: 2^-11 relative per operand
    assert (out - ref).abs().max().item() < 2e-2 * ref.abs().max().item()


def test_ff2_proj_out_concatenation():
    torch.manual_seed(1)
    C, M = 64, 10
    w2, b2, wpo, bpo = torch.randn(C, 4 * C) * 0.1, torch.randn(C) * 0.1, torch.randn(C, C) * 0.1, torch.randn(C) * 0.1
    h, ff, x = torch.randn(M, C), torch.randn(M, 4 * C), torch.randn(M, C)
    ref = (h + ff @ w2.T + b2) @ wpo.T + bpo + x
    fpw = torch.cat([wpo, wpo @ w2], 1)
    fpb = wpo @ b2 + bpo
    out = torch.cat([h, ff], 1) @ fpw.T + fpb + x
    assert torch.allclose(out, ref, atol=1e-4, rtol=1e-4)
