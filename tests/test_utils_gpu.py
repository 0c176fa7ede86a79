"""the PRODUCT box / mask / latent helpers (theatergen_b200/utils/{utils,latents,schedule}.py)
against the golden vectors produced by the reference's own utils/utils.py and utils/schedule.py
(tests/golden/utils.pt, tests/golden/make_golden.py) and, for utils/latents.py (not importable: matplotlib), against
the oracle restatement -- bit exact: this is index arithmetic and 0/1-mask blending.

The device-agnostic helpers are checked on the CPU in the `not gpu` run and again on the GPU; the helpers that pin
their result to "cuda" (like the reference, utils/utils.py:6,50) only in the `gpu` run.
"""
import os

import pytest
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "utils.pt")


def _gold():
    return torch.load(GOLD, map_location="cpu", weights_only=False)


def _index_helpers(dev):
    from theatergen_b200.utils import schedule as SC
    from theatergen_b200.utils import utils as U
    g = _gold()
    t, m = g["t"].to(dev), g["mask"].to(dev)
    for (xo, yo), want, mwant in zip(g["shift_cases"], g["shifts"], g["mask_shifts"]):
        got = U.shift_tensor(t, xo, yo, offset_normalized=True)
        assert got.device == t.device and torch.equal(got.cpu(), want)
        assert torch.equal(U.shift_tensor(m, xo, yo, offset_normalized=True).cpu(), mwant)
    # ignore_last_dim (channels-last tensors): same shift on the H/W axes of a permuted copy
    tl = t.permute(0, 1, 3, 4, 2).contiguous()
    for (xo, yo), want in zip(g["shift_cases"], g["shifts"]):
        got = U.shift_tensor(tl, xo, yo, offset_normalized=True, ignore_last_dim=True)
        assert torch.equal(got.permute(0, 1, 4, 2, 3).cpu(), want)
    assert [int(v) for v in U.binary_mask_to_box(m)] == [int(v) for v in g["box"]]
    assert torch.equal(U.binary_mask_to_box_mask(m, to_device=False), g["box_mask"])
    cx, cy = U.binary_mask_to_center(m, normalize=True)
    assert abs(cx - g["center"][0]) < 1e-6 and abs(cy - g["center"][1]) < 1e-6
    for (b, H, W), want in zip(g["scale_cases"], g["scaled"]):
        assert U.scale_proportion(b, H, W) == tuple(want)
    ts = torch.arange(981, 0, -20, device=dev)
    for (a, r), want in zip(((15, 2), (49, 2), (5, 3)), g["fast_schedule"]):
        assert torch.equal(SC.get_fast_schedule(ts, a, r).cpu(), want)
    with pytest.raises(ValueError):
        U.binary_mask_to_box(torch.zeros(8, 8, dtype=torch.bool, device=dev))      # utils/utils.py:75-76
    with pytest.raises(AssertionError):
        U.shift_tensor(torch.zeros(1, 4, 60, 60, device=dev), 0.1, 0.1, offset_normalized=True)  # not a multiple of 8


def test_index_helpers_match_reference_golden_cpu():
    _index_helpers("cpu")


@pytest.mark.gpu
def test_index_helpers_match_reference_golden_gpu():
    _index_helpers("cuda")


@pytest.mark.gpu
def test_device_pinned_helpers_match_reference_golden():
    from theatergen_b200.utils import utils as U
    g = _gold()
    bm = U.binary_mask_to_box_mask(g["mask"].cuda())
    assert bm.is_cuda and torch.equal(bm.cpu(), g["box_mask"])
    gg = torch.load(os.path.join(os.path.dirname(GOLD), "guidance.pt"), map_location="cpu", weights_only=False)
    for hw, masks in gg["masks"].items():
        for b, m in zip(gg["bboxes"], masks):
            got = U.proportion_to_mask(b, hw, hw)
            assert got.is_cuda and torch.equal(got.cpu(), m.cpu())


class _Adapter:
    """what utils/latents.py reaches through `adapter` (latents.py:258-262): unet.config.in_channels, unet.dtype,
    scheduler.init_noise_sigma"""

    class pipe:
        class unet:
            dtype = torch.float16

            class config:
                in_channels = 4

        class scheduler:
            init_noise_sigma = 1.0


@pytest.mark.gpu
def test_latent_helpers_match_oracle_bit_exact():
    from oracle import latents as OL
    from theatergen_b200.utils import latents as L
    boxes = [[0.3, 0.2, 0.7, 0.8], [0.543, 0.49, 0.85, 0.852], [0.0078125, 0.6464, 0.1640625, 0.783],
             [-0.1, 0.3, 0.4, 1.2]]
    bg_seed, fg_seed = 3, 3 + 123456789
    got, got_bg, seeds = L.get_input_latents_list(None, bg_seed, fg_seed, 0.01, 512, 512, _Adapter, so_boxes=boxes)
    want, want_bg = OL.get_input_latents_list(bg_seed, fg_seed, 0.01, 512, 512, boxes)
    assert seeds == [fg_seed] * 4 and got_bg.is_cuda and got_bg.dtype == torch.float16
    assert torch.equal(got_bg.cpu(), want_bg)
    for a, b in zip(got, want):
        assert a.dtype == torch.float16 and torch.equal(a.cpu(), b)
    assert torch.equal(L.get_input_latents_lne(2, _Adapter, None, bg_seed, fg_seed, 0.01, 512, 512, so_boxes=boxes).cpu(),
                       want[2])
    # blend on its own, incl. the reference's assert on identical inputs (latents.py:157)
    fg = OL.get_unscaled_latents(1, 4, 512, 512, torch.manual_seed(9), torch.float16)
    m = OL.proportion_to_mask(boxes[0], 64, 64)
    assert torch.equal(L.blend_latents(want_bg.cuda(), fg.cuda(), m.cuda(), 0.05).cpu(),
                       OL.blend_latents(want_bg, fg, m, 0.05))
    with pytest.raises(AssertionError):
        L.blend_latents(fg.cuda(), fg.cuda(), m.cuda())

    # align + compose on [51,1,4,64,64] stacks with ragged masks (one of them shifted partly out of the frame)
    g = torch.Generator().manual_seed(5)
    stacks = [torch.randn(51, 1, 4, 64, 64, generator=g).half() for _ in range(4)]
    masks = []
    for i, (y0, y1, x0, x1) in enumerate([(20, 41, 10, 33), (5, 15, 40, 60), (30, 62, 28, 40), (0, 9, 0, 12)]):
        mk = torch.zeros(64, 64, dtype=torch.bool)
        mk[y0:y1, x0:x1] = True
        mk[y0 + 2:y0 + 5, x1:min(x1 + 3, 64)] = True
        masks.append(mk)
    for hso in (False, True):
        a_l, a_m, a_o = L.align_with_bboxes([s.cuda() for s in stacks], [m.cuda() for m in masks], boxes,
                                            horizontal_shift_only=hso)
        o_l, o_m, o_o = OL.align_with_bboxes(stacks, masks, boxes, horizontal_shift_only=hso)
        for a, b in zip(a_l + a_m, o_l + o_m):
            assert torch.equal(a.cpu(), b)
        assert all(abs(p[0] - q[0]) < 1e-6 and abs(p[1] - q[1]) < 1e-6 for p, q in zip(a_o, o_o))
    bg = want_bg
    for box_to_bg in (True, False):
        comp, idx = L.compose_latents(_Adapter, None, [t.cuda() for t in o_l], [m.cuda() for m in o_m], 50, 1, 512, 512,
                                      latents_bg=bg.cuda(), compose_box_to_bg=box_to_bg)
        comp_o, idx_o = OL.compose_latents(o_l, o_m, bg, 50, compose_box_to_bg=box_to_bg)
        assert comp.is_cuda and comp.shape == (51, 1, 4, 64, 64) and torch.equal(comp.cpu(), comp_o)
        assert torch.equal(idx.cpu(), idx_o)
    # latents_bg=None: seeded background (latents.py:175-177)
    comp2, _ = L.compose_latents(_Adapter, None, [t.cuda() for t in o_l], [m.cuda() for m in o_m], 50, 1, 512, 512,
                                 bg_seed=bg_seed)
    assert torch.equal(comp2.cpu(), OL.compose_latents(o_l, o_m, want_bg, 50)[0])
