"""Latent / mask utilities restated, device agnostic (TEST INFRASTRUCTURE, see oracle/__init__.py).

Follows /root/reference/utils/latents.py (cannot be imported: needs matplotlib) and utils/utils.py:
  get_unscaled_latents :138-149, blend_latents :156-166, compose_latents :168-218, align_with_bboxes :220-240,
  get_input_latents_list :257-295; utils.py: proportion_to_mask :45-53, binary_mask_to_box(_mask) :70-98,
  binary_mask_to_center :100-119, shift_tensor :143-178.
utils/utils.py imports here as-is, so shift_tensor / scale_proportion / binary_mask_* are pinned against it.
"""
from __future__ import annotations

import numpy as np
import torch

from .guidance import scale_proportion


def proportion_to_mask(box, H, W, device="cpu"):
    x0, y0, x1, y1 = scale_proportion(box, H, W)
    mask = torch.zeros(H, W, device=device)
    mask[y0:y1, x0:x1] = 1.0
    return mask


def get_unscaled_latents(batch_size, in_channels, height, width, generator, dtype):
    return torch.randn((batch_size, in_channels, height // 8, width // 8), generator=generator, dtype=dtype)


def blend_latents(latents_bg, latents_fg, fg_mask, fg_blending_ratio=0.01):
    dtype = latents_bg.dtype
    latents = latents_bg * (1.0 - fg_mask) + (
        latents_bg * np.sqrt(1.0 - fg_blending_ratio) + latents_fg * np.sqrt(fg_blending_ratio)) * fg_mask
    return latents.to(dtype=dtype)


def get_input_latents_list(bg_seed, fg_seed_start, fg_blending_ratio, height, width, so_boxes, dtype=torch.float16,
                           in_channels=4, init_noise_sigma=1.0):
    latents_bg = get_unscaled_latents(1, in_channels, height, width, torch.manual_seed(bg_seed), dtype)
    out = []
    for box in so_boxes:
        fg_mask = proportion_to_mask(box, height // 8, width // 8)
        latents_fg = get_unscaled_latents(1, in_channels, height, width, torch.manual_seed(fg_seed_start), dtype)
        out.append(blend_latents(latents_bg, latents_fg, fg_mask, fg_blending_ratio) * init_noise_sigma)
    return out, latents_bg * init_noise_sigma


def binary_mask_to_box(mask):
    ys, xs = torch.where(mask)
    h, w = mask.shape
    ymin, ymax = max(int(ys.min()) - 1, 0), min(int(ys.max()) + 1, h)
    xmin, xmax = max(int(xs.min()) - 1, 0), min(int(xs.max()) + 1, w)
    return [xmin, ymin, xmax, ymax]


def binary_mask_to_box_mask(mask):
    x0, y0, x1, y1 = binary_mask_to_box(mask)
    out = torch.zeros(mask.shape)
    out[y0:y1 + 1, x0:x1 + 1] = 1.0
    return out


def binary_mask_to_center(mask, normalize=False):
    h, w = mask.shape
    total = mask.sum()
    x = ((mask.sum(dim=0) @ torch.arange(w)) / total).item()
    y = ((mask.sum(dim=1) @ torch.arange(h)) / total).item()
    if normalize:
        x, y = x / w, y / h
    return x, y


def shift_tensor(tensor, x_offset, y_offset, base_w=8, base_h=8, offset_normalized=False):
    th, tw = tensor.shape[-2:]
    if offset_normalized:
        x_offset, y_offset = round(x_offset * base_w) * (tw // base_w), round(y_offset * base_h) * (th // base_h)
    new = torch.zeros_like(tensor)
    ow, oh = tw - abs(x_offset), th - abs(y_offset)
    ys, yd = (0, y_offset) if y_offset >= 0 else (-y_offset, 0)
    xs, xd = (0, x_offset) if x_offset >= 0 else (-x_offset, 0)
    new[..., yd:yd + oh, xd:xd + ow] = tensor[..., ys:ys + oh, xs:xs + ow]
    return new


def align_with_bboxes(latents_all_list, mask_tensor_list, bboxes, horizontal_shift_only=False):
    new_l, new_m, offs = [], [], []
    for latents_all, mask, bbox in zip(latents_all_list, mask_tensor_list, bboxes):
        xs, ys = binary_mask_to_center(mask, normalize=True)
        xd, yd = (bbox[0] + bbox[2]) / 2, (bbox[1] + bbox[3]) / 2
        xo, yo = xd - xs, (0.0 if horizontal_shift_only else yd - ys)
        new_l.append(shift_tensor(latents_all, xo, yo, offset_normalized=True))
        new_m.append(shift_tensor(mask, xo, yo, offset_normalized=True))
        offs.append((xo, yo))
    return new_l, new_m, offs


def compose_latents(latents_all_list, mask_tensor_list, latents_bg, num_inference_steps, compose_box_to_bg=True):
    """utils/latents.py:168-218 (use_fast_schedule=False).  masks: bool [64,64]; latents_all: [T+1,1,4,64,64]."""
    dtype = latents_bg.dtype
    composed = torch.zeros((num_inference_steps + 1, *latents_bg.shape), dtype=dtype)
    composed[0] = latents_bg
    fg_idx = torch.zeros(latents_bg.shape[-2:], dtype=torch.long)
    sizes = np.array([m.sum().item() for m in mask_tensor_list])
    order = np.argsort(-sizes)
    if compose_box_to_bg:
        for i in order:
            bm = binary_mask_to_box_mask(mask_tensor_list[i])[None, None, None, ...].to(dtype)
            composed[0] = composed[0] * (1.0 - bm) + latents_all_list[i][0] * bm
    for i in order:
        m = mask_tensor_list[i]
        fg_idx = fg_idx * (~m) + (i + 1) * m
        me = m[None, None, None, ...].to(dtype)
        composed = composed * (1.0 - me) + latents_all_list[i] * me
    return composed, fg_idx


def get_fast_schedule(origial_timesteps, fast_after_steps, fast_rate):
    """utils/schedule.py:4-8"""
    if fast_after_steps >= len(origial_timesteps) - 1:
        return origial_timesteps
    new = torch.cat((origial_timesteps[:fast_after_steps], origial_timesteps[fast_after_steps + 1::fast_rate]), dim=0)
    return new
