"""Cross-attention guidance energy restated, device agnostic (TEST INFRASTRUCTURE, see oracle/__init__.py).

Follows /root/reference/utils/guidance.py (which hard-codes device="cuda"):
  add_ca_loss_per_attn_map_to_loss        :91-148   (max-based branch, use_ratio_based_loss=False)
  add_ref_ca_loss_per_attn_map_to_lossv2  :150-242
  compute_ca_lossv3                       :244-286
and utils/utils.py:55-68 (scale_proportion).  Pinned against the reference functions themselves (imported with a
device-rewriting TorchFunctionMode) by tests/golden/make_golden.py.
"""
from __future__ import annotations

import math
from collections.abc import Iterable

import torch


def scale_proportion(obj_box, H, W):
    x_min, y_min = round(obj_box[0] * W), round(obj_box[1] * H)
    box_w, box_h = round((obj_box[2] - obj_box[0]) * W), round((obj_box[3] - obj_box[1]) * H)
    x_max, y_max = x_min + box_w, y_min + box_h
    x_min, y_min = max(x_min, 0), max(y_min, 0)
    x_max, y_max = min(x_max, W), min(y_max, H)
    return x_min, y_min, x_max, y_max


def box_mask(boxes, H, W, device):
    if not isinstance(boxes[0], Iterable):
        boxes = [boxes]
    mask = torch.zeros((H, W), device=device)
    for b in boxes:
        x0, y0, x1, y1 = scale_proportion(b, H, W)
        mask[y0:y1, x0:x1] = 1
    return mask


def ca_loss_per_map(loss, attn_map, bboxes, object_positions, fg_top_p=0.2, bg_top_p=0.2, fg_weight=1.0, bg_weight=1.0):
    b, i, j = attn_map.shape
    H = W = int(math.sqrt(i))
    for obj_idx in range(len(bboxes)):
        obj_loss = 0
        mask = box_mask(bboxes[obj_idx], H, W, attn_map.device)
        k_fg = (mask.sum() * fg_top_p).long().clamp_(min=1)
        k_bg = ((1 - mask).sum() * bg_top_p).long().clamp_(min=1)
        mask_1d = mask.view(1, -1)
        for pos in object_positions[obj_idx]:
            ca = attn_map[:, :, pos]
            obj_loss += (1 - (ca * mask_1d).topk(k=k_fg).values.mean(dim=1)).sum(dim=0) * fg_weight
            obj_loss += ((ca * (1 - mask_1d)).topk(k=k_bg).values.mean(dim=1)).sum(dim=0) * bg_weight
        loss = loss + obj_loss / len(object_positions[obj_idx])
    return loss


def ref_ca_loss(saved_attn, bboxes, object_positions, keys, ref_ca_saved_attns, index, loss_weight=1.0,
                ref_ca_last_token_only=True, ref_ca_word_token_only=False, word_token_indices=None, eps=1e-5):
    device = next(iter(saved_attn.values())).device
    loss = torch.tensor(0.0, device=device)
    if loss_weight == 0.0:
        return loss
    for obj_idx in range(len(bboxes)):
        obj_loss = 0
        obj_boxes, obj_refs = bboxes[obj_idx], ref_ca_saved_attns[obj_idx]
        if not isinstance(obj_boxes[0], Iterable):
            obj_boxes, obj_refs = [obj_boxes], [obj_refs]
        for box, ref in zip(obj_boxes, obj_refs):
            ref = ref[index]
            for key in keys:
                attn_map = saved_attn[key].squeeze(dim=0)
                b, i, j = attn_map.shape
                H = W = int(math.sqrt(i))
                ref_map = ref[key].to(device)[0, :, :, 0]
                m = box_mask(box, H, W, device).reshape(1, -1)
                if ref_ca_word_token_only:
                    positions = [word_token_indices[obj_idx]]
                elif ref_ca_last_token_only:
                    positions = [object_positions[obj_idx][-1]]
                else:
                    positions = object_positions[obj_idx]
                for pos in positions:
                    a = attn_map[:, :, pos] * m
                    a = a / (a.sum(dim=-1, keepdim=True) + eps)
                    r = ref_map * m
                    r = r / (r.sum(dim=-1, keepdim=True) + eps)
                    obj_loss += torch.mean((a - r).abs().sum(dim=-1), dim=0)
        loss = loss + loss_weight * obj_loss / (len(obj_boxes) * len(positions))
    return loss


def compute_ca_lossv3(saved_attn, bboxes, object_positions, guidance_attn_keys, ref_ca_saved_attns=None,
                      ref_ca_last_token_only=True, ref_ca_word_token_only=False, word_token_indices=None, index=None,
                      ref_ca_loss_weight=1.0, **kwargs):
    device = next(iter(saved_attn.values())).device
    loss = torch.tensor(0.0, device=device)
    n_obj = len(bboxes)
    if n_obj == 0:
        return loss
    for key in guidance_attn_keys:
        loss = ca_loss_per_map(loss, saved_attn[key].squeeze(dim=0), bboxes, object_positions, **kwargs)
    n_attn = len(guidance_attn_keys)
    if n_attn > 0:
        loss = loss / (n_obj * n_attn)
    if ref_ca_saved_attns is not None:
        ref = ref_ca_loss(saved_attn, bboxes, object_positions, guidance_attn_keys, ref_ca_saved_attns, index,
                          ref_ca_loss_weight, ref_ca_last_token_only, ref_ca_word_token_only, word_token_indices)
        loss = loss + ref / (n_obj * n_attn)
    return loss
