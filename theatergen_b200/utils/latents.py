"""Mirror of the reference's utils/latents.py latent math (the PIL / matplotlib collage helpers stay out of scope).

Seeded initial latents and the fg/bg blend follow the reference exactly (CPU default generator, fp16 randn:
latents.py:138-149, 257-295) because those values define "identical seeds".  compose_latents / align_with_bboxes
keep the reference signatures and semantics (latents.py:168-240) but run on the device the tensors live on instead
of the reference's CPU round trip.
"""
from __future__ import annotations

import numpy as np
import torch

from . import utils

torch_device = "cuda"


def get_unscaled_latents(batch_size, in_channels, height, width, generator, dtype):
    return torch.randn((batch_size, in_channels, height // 8, width // 8), generator=generator,
                       dtype=dtype).to(torch_device, dtype=dtype)


def get_scaled_latents(batch_size, in_channels, height, width, generator, dtype, scheduler):
    return get_unscaled_latents(batch_size, in_channels, height, width, generator, dtype) * scheduler.init_noise_sigma


def blend_latents(latents_bg, latents_fg, fg_mask, fg_blending_ratio=0.01):
    assert not torch.allclose(latents_bg, latents_fg), "latents_bg should be independent with latents_fg"
    dtype = latents_bg.dtype
    latents = latents_bg * (1. - fg_mask) + (latents_bg * np.sqrt(1. - fg_blending_ratio)
                                             + latents_fg * np.sqrt(fg_blending_ratio)) * fg_mask
    return latents.to(dtype=dtype)


def get_input_latents_list(model_dict, bg_seed, fg_seed_start, fg_blending_ratio, height, width, adapter,
                           so_prompt_phrase_box_list=None, so_boxes=None, verbose=False):
    unet, scheduler, dtype = adapter.pipe.unet, adapter.pipe.scheduler, adapter.pipe.unet.dtype
    latents_bg = get_unscaled_latents(1, unet.config.in_channels, height, width, torch.manual_seed(bg_seed), dtype)
    if so_boxes is None:
        so_boxes = [item[-1] for item in so_prompt_phrase_box_list]
    input_latents_list, fg_seed_list = [], []
    for obj_box in so_boxes:
        fg_mask = utils.proportion_to_mask(obj_box, height // 8, width // 8, device=latents_bg.device)
        fg_seed_list.append(fg_seed_start)  # every foreground shares one seed (latents.py:282)
        latents_fg = get_unscaled_latents(1, unet.config.in_channels, height, width, torch.manual_seed(fg_seed_start),
                                          dtype)
        input_latents_list.append(blend_latents(latents_bg, latents_fg, fg_mask, fg_blending_ratio)
                                  * scheduler.init_noise_sigma)
    return input_latents_list, latents_bg * scheduler.init_noise_sigma, fg_seed_list


def get_input_latents_lne(idx, adapter, model_dict, bg_seed, fg_seed_start, fg_blending_ratio, height, width,
                          so_prompt_phrase_box_list=None, so_boxes=None, verbose=False):
    lst, _, _ = get_input_latents_list(model_dict, bg_seed, fg_seed_start, fg_blending_ratio, height, width, adapter,
                                       so_prompt_phrase_box_list, so_boxes)
    return lst[idx]


def align_with_bboxes(latents_all_list, mask_tensor_list, bboxes, horizontal_shift_only=False):
    new_latents, new_masks, offsets = [], [], []
    for latents_all, mask_tensor, bbox in zip(latents_all_list, mask_tensor_list, bboxes):
        x_src, y_src = utils.binary_mask_to_center(mask_tensor, normalize=True)
        x_dst, y_dst = (bbox[0] + bbox[2]) / 2, (bbox[1] + bbox[3]) / 2
        x_off, y_off = x_dst - x_src, (0. if horizontal_shift_only else y_dst - y_src)
        new_latents.append(utils.shift_tensor(latents_all, x_off, y_off, offset_normalized=True))
        new_masks.append(utils.shift_tensor(mask_tensor, x_off, y_off, offset_normalized=True))
        offsets.append((x_off, y_off))
    return new_latents, new_masks, offsets


@torch.no_grad()
def compose_latents(adapter, model_dict, latents_all_list, mask_tensor_list, num_inference_steps, overall_batch_size,
                    height, width, latents_bg=None, bg_seed=None, compose_box_to_bg=True, use_fast_schedule=False,
                    fast_after_steps=None):
    unet, scheduler, dtype = adapter.pipe.unet, adapter.pipe.scheduler, adapter.pipe.unet.dtype
    if latents_bg is None:
        latents_bg = get_scaled_latents(overall_batch_size, unet.config.in_channels, height, width,
                                        torch.manual_seed(bg_seed), dtype, scheduler)
    dev = latents_bg.device
    n = (fast_after_steps if use_fast_schedule else num_inference_steps) + 1
    composed = torch.zeros((n, *latents_bg.shape), dtype=dtype, device=dev)
    composed[0] = latents_bg
    fg_idx = torch.zeros(latents_bg.shape[-2:], dtype=torch.long, device=dev)
    sizes = np.array([m.sum().item() for m in mask_tensor_list])
    order = np.argsort(-sizes)  # compose the largest mask first
    if compose_box_to_bg:
        for i in order:
            bm = utils.binary_mask_to_box_mask(mask_tensor_list[i], to_device=False).to(dev)[None, None, None].to(dtype)
            composed[0] = composed[0] * (1. - bm) + latents_all_list[i][0].to(dev) * bm
    for i in order:
        m = mask_tensor_list[i].to(dev)
        fg_idx = fg_idx * (~m) + (i + 1) * m
        me = m[None, None, None].to(dtype)
        composed = composed * (1. - me) + latents_all_list[i][:n].to(dev) * me
    return composed, fg_idx


@torch.no_grad()
def compose_latents_with_alignment(basever, adapter, repeat_ind, mask_tensor_list_512, single_obj_img_list, model_dict,
                                   latents_all_list, mask_tensor_list, num_inference_steps, overall_batch_size, height,
                                   width, align_with_overall_bboxes=True, overall_bboxes=None,
                                   horizontal_shift_only=False, prepare_mid_image=None, **kwargs):
    """utils/latents.py:242-255.  The PIL collage + inpainting mask of `prepare_mid_image` (latents.py:48-135) is image
    munging outside the hot path: pass the reference's function as `prepare_mid_image` to get (inp_mask, inp_img);
    without it both are None."""
    inp_mask = inp_img = None
    if align_with_overall_bboxes and len(latents_all_list):
        expanded = utils.expand_overall_bboxes(overall_bboxes)
        latents_all_list, mask_tensor_list, _ = align_with_bboxes(latents_all_list, mask_tensor_list, bboxes=expanded,
                                                                  horizontal_shift_only=horizontal_shift_only)
        if prepare_mid_image is not None:
            inp_mask, inp_img = prepare_mid_image(basever, repeat_ind, mask_tensor_list_512, single_obj_img_list,
                                                  bboxes=expanded)
    composed, fg_idx = compose_latents(adapter, model_dict, latents_all_list, mask_tensor_list, num_inference_steps,
                                       overall_batch_size, height, width, **kwargs)
    return composed, fg_idx, inp_mask, inp_img
