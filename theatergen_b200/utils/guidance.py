"""Mirror of the reference's utils/guidance.py energy functions on top of `tg_guidance_energy`.

compute_ca_lossv3 keeps the reference signature (utils/guidance.py:244) and returns a 0-d float32 CUDA tensor.
The host code only flattens the python description (boxes / token positions / keys) into the item list the
kernel takes by value; all arithmetic runs in guidance.cu.  `get_phrase_indices` / `get_token_map` are host
string code and stay with the caller's tokenizer.
"""
from __future__ import annotations

import ctypes as C
from collections.abc import Iterable

import torch

from .. import _lib

MAX_KEYS, MAX_BOXES, MAX_ITEMS, MAX_REF = 8, 16, 128, 32


class EnergyPlan:
    """Pre-flattened description of one energy evaluation (reusable across timesteps / graph capture)."""

    def __init__(self, bboxes, object_positions, guidance_attn_keys, fg_top_p=0.2, bg_top_p=0.2, fg_weight=1.0,
                 bg_weight=1.0, ref_positions=None, ref_ca_loss_weight=1.0, use_ratio_based_loss=False, **unused):
        # Deviation from the reference's DEFAULT, on purpose: add_ca_loss_per_attn_map_to_loss defaults to
        # use_ratio_based_loss=True (utils/guidance.py:91), but every caller in the reference passes False explicitly
        # (theatergen.py:367, 430) and the ratio branch is marked deprecated there; here False is the default and True raises.
        if use_ratio_based_loss:
            raise NotImplementedError("ratio-based loss is deprecated in the reference (utils/guidance.py:122) and "
                                      "never selected by theatergen.py; only the max-based energy is implemented")
        self.keys = [tuple(k) for k in guidance_attn_keys]
        n_obj, n_keys = len(bboxes), len(self.keys)
        self.n_obj = n_obj
        boxes, box_obj, self.obj_boxes = [], [], []
        for o, ob in enumerate(bboxes):
            if not isinstance(ob[0], Iterable):
                ob = [ob]
            ids = []
            for b in ob:
                ids.append(len(boxes))
                boxes.append([float(v) for v in b])
                box_obj.append(o)
            self.obj_boxes.append(ids)
        assert n_keys <= MAX_KEYS and len(boxes) <= MAX_BOXES
        self.boxes = (C.c_double * (4 * max(1, len(boxes))))(*[v for b in boxes for v in b])
        self.box_obj = (C.c_int * max(1, len(boxes)))(*box_obj)
        self.n_box = len(boxes)
        items = []
        for k in range(n_keys):
            for o in range(n_obj):
                pos = list(object_positions[o])
                for p in pos:
                    items.append((k, o, 0, 0, int(p), 1.0 / (len(pos) * n_obj * n_keys)))
        assert len(items) <= MAX_ITEMS, f"{len(items)} energy items > {MAX_ITEMS}"
        self.items = (_lib.GuidanceItem * max(1, len(items)))(*[_lib.GuidanceItem(*it) for it in items])
        self.n_items = len(items)
        # reference-attention transfer items: (key, box, token, ref-map slot)
        self.ref_slots = []  # (obj, box_local_idx, key) per ref map slot
        ref_items = []
        if ref_positions is not None:
            for o in range(n_obj):
                pos = list(ref_positions[o])
                nb = len(self.obj_boxes[o])
                for bl, bi in enumerate(self.obj_boxes[o]):
                    for k in range(n_keys):
                        slot = len(self.ref_slots)
                        self.ref_slots.append((o, bl, self.keys[k]))
                        for p in pos:
                            w = ref_ca_loss_weight / (nb * len(pos)) / (n_obj * n_keys)
                            ref_items.append((k, o, bi, slot, int(p), w))
        assert len(ref_items) <= MAX_REF, f"{len(ref_items)} reference items > {MAX_REF}"
        self.ref_items = (_lib.GuidanceItem * max(1, len(ref_items)))(*[_lib.GuidanceItem(*it) for it in ref_items])
        self.n_ref = len(ref_items)
        self.fg_top_p, self.bg_top_p, self.fg_weight, self.bg_weight = fg_top_p, bg_top_p, fg_weight, bg_weight

    def run(self, maps, loss_out, workspace, dmaps=None, ref_maps=None, grad_scale=1.0, ref_eps=1e-5):
        """maps: list (one per key) of contiguous fp16 CUDA tensors [heads, HW, n_tok]."""
        n_keys = len(self.keys)
        heads, _, n_tok = maps[0].shape
        for m in maps:
            assert m.is_cuda and m.dtype == torch.float16 and m.is_contiguous() and m.shape[0] == heads and m.shape[2] == n_tok
        mp = (C.c_void_p * n_keys)(*[m.data_ptr() for m in maps])
        dp = None if dmaps is None else (C.c_void_p * n_keys)(*[d.data_ptr() for d in dmaps])
        hw = (C.c_int * n_keys)(*[m.shape[1] for m in maps])
        rp = None
        if self.n_ref:
            assert ref_maps is not None and len(ref_maps) == len(self.ref_slots)
            rp = (C.c_void_p * len(ref_maps))(*[r.data_ptr() for r in ref_maps])
        rc = _lib.lib().tg_guidance_energy(mp, dp, hw, n_keys, heads, n_tok, self.boxes, self.box_obj, self.n_box,
                                           self.items, self.n_items, self.ref_items, rp, self.n_ref,
                                           self.fg_top_p, self.bg_top_p, self.fg_weight, self.bg_weight,
                                           float(grad_scale), float(ref_eps), loss_out.data_ptr(),
                                           workspace.data_ptr(), torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "tg_guidance_energy")
        return loss_out


def _cond_map(t):
    """saved_attn entries are [1, heads, HW, n_tok] (cond half); the kernel wants contiguous fp16 [heads,HW,n_tok]."""
    if not t.is_cuda:
        t = t.cuda()
    t = t.squeeze(0) if t.dim() == 4 else t
    return t.to(torch.float16).contiguous()


def build_plan(bboxes, object_positions, guidance_attn_keys, ref_ca_saved_attns=None, ref_ca_last_token_only=True,
               ref_ca_word_token_only=False, word_token_indices=None, index=None, ref_ca_loss_weight=1.0, **kwargs):
    """The python description of one energy evaluation (utils/guidance.py:244-286) -> (EnergyPlan, reference maps)."""
    keys = [tuple(k) for k in guidance_attn_keys]
    ref_positions = None
    if ref_ca_saved_attns is not None and ref_ca_loss_weight != 0.0:
        if ref_ca_word_token_only:
            ref_positions = [[word_token_indices[o]] for o in range(len(bboxes))]
        elif ref_ca_last_token_only:
            ref_positions = [[object_positions[o][-1]] for o in range(len(bboxes))]
        else:
            ref_positions = [list(object_positions[o]) for o in range(len(bboxes))]
    kwargs.pop("verbose", None)
    plan = EnergyPlan(bboxes, object_positions, keys, ref_positions=ref_positions,
                      ref_ca_loss_weight=ref_ca_loss_weight, **kwargs)
    ref_maps = None
    if plan.n_ref:
        ref_maps = []
        for (o, bl, k) in plan.ref_slots:
            per_box = ref_ca_saved_attns[o]
            if not isinstance(bboxes[o][0], Iterable):
                per_box = [per_box]
            r = per_box[bl][index][k]
            r = r if r.is_cuda else r.cuda()
            ref_maps.append(r[0, :, :, 0].to(torch.float16).contiguous())
    return plan, ref_maps


def compute_ca_lossv3(saved_attn, bboxes, object_positions, guidance_attn_keys, ref_ca_saved_attns=None,
                      ref_ca_last_token_only=True, ref_ca_word_token_only=False, word_token_indices=None, index=None,
                      ref_ca_loss_weight=1.0, verbose=False, **kwargs):
    """Same contract as utils/guidance.py:244-286; returns a 0-d float32 CUDA tensor."""
    dev = torch.device("cuda")
    if len(bboxes) == 0:
        return torch.zeros((), dtype=torch.float32, device=dev)
    plan, ref_maps = build_plan(bboxes, object_positions, guidance_attn_keys, ref_ca_saved_attns, ref_ca_last_token_only,
                                ref_ca_word_token_only, word_token_indices, index, ref_ca_loss_weight, **kwargs)
    maps = [_cond_map(saved_attn[k]) for k in plan.keys]
    loss = torch.zeros((1,), dtype=torch.float32, device=dev)
    ws = torch.empty((MAX_ITEMS + MAX_REF,), dtype=torch.float32, device=dev)
    plan.run(maps, loss, ws, ref_maps=ref_maps)
    return loss[0]
