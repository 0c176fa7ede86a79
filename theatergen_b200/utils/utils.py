"""Host-side box / mask helpers with the reference's names and semantics (utils/utils.py:18-178).
Index arithmetic on tiny tensors: stays in Python, runs on whatever device the tensors live on."""
from __future__ import annotations

import numpy as np
import torch

torch_device = "cuda"


def get_centered_box(box, horizontal_center_only=True, vertical_placement='centered', vertical_center=0.5,
                     floor_padding=None):
    x_min, y_min, x_max, y_max = box
    w = x_max - x_min
    x_min_new, x_max_new = 0.5 - w / 2, 0.5 + w / 2
    if horizontal_center_only:
        return [x_min_new, y_min, x_max_new, y_max]
    h = y_max - y_min
    if vertical_placement == 'centered':
        assert floor_padding is None, "Set vertical_placement to floor_padding to use floor padding"
        y_min_new, y_max_new = vertical_center - h / 2, vertical_center + h / 2
    elif vertical_placement == 'floor_padding':
        y_max_new = 1 - floor_padding
        y_min_new = y_max_new - h
    else:
        raise ValueError(f"Unknown vertical placement: {vertical_placement}")
    return [x_min_new, y_min_new, x_max_new, y_max_new]


def scale_proportion(obj_box, H, W, use_legacy=False):
    if use_legacy:
        return int(obj_box[0] * W), int(obj_box[1] * H), int(obj_box[2] * W), int(obj_box[3] * H)
    x_min, y_min = round(obj_box[0] * W), round(obj_box[1] * H)
    box_w, box_h = round((obj_box[2] - obj_box[0]) * W), round((obj_box[3] - obj_box[1]) * H)
    x_max, y_max = x_min + box_w, y_min + box_h
    x_min, y_min = max(x_min, 0), max(y_min, 0)
    x_max, y_max = min(x_max, W), min(y_max, H)
    return x_min, y_min, x_max, y_max


def proportion_to_mask(obj_box, H, W, use_legacy=False, return_np=False, device=None):
    x_min, y_min, x_max, y_max = scale_proportion(obj_box, H, W, use_legacy)
    if return_np:
        mask = np.zeros((H, W))
    else:
        mask = torch.zeros(H, W, device=device if device is not None else torch_device)
    mask[y_min:y_max, x_min:x_max] = 1.
    return mask


def binary_mask_to_box(mask, enlarge_box_by_one=True, w_scale=1, h_scale=1):
    mask_loc = torch.where(mask) if isinstance(mask, torch.Tensor) else np.where(mask)
    height, width = mask.shape
    if len(mask_loc) == 0 or len(mask_loc[0]) == 0:
        raise ValueError('The mask is empty')
    if enlarge_box_by_one:
        ymin, ymax = max(int(min(mask_loc[0])) - 1, 0), min(int(max(mask_loc[0])) + 1, height)
        xmin, xmax = max(int(min(mask_loc[1])) - 1, 0), min(int(max(mask_loc[1])) + 1, width)
    else:
        ymin, ymax = int(min(mask_loc[0])), int(max(mask_loc[0]))
        xmin, xmax = int(min(mask_loc[1])), int(max(mask_loc[1]))
    return [xmin * w_scale, ymin * h_scale, xmax * w_scale, ymax * h_scale]


def binary_mask_to_box_mask(mask, to_device=True):
    x_min, y_min, x_max, y_max = binary_mask_to_box(mask)
    H, W = mask.shape
    out = torch.zeros(H, W)
    if to_device:
        out = out.to(torch_device)
    out[y_min:y_max + 1, x_min:x_max + 1] = 1.
    return out


def binary_mask_to_center(mask, normalize=False):
    h, w = mask.shape
    total = mask.sum()
    if isinstance(mask, torch.Tensor):
        # the reference's integer dot product (`mask.sum(dim=0) @ arange`, utils/utils.py:107-108) written as mul + sum:
        # same exact int64 value, and defined for integer tensors on a CUDA device as well
        x_coord = (((mask.sum(dim=0) * torch.arange(w, device=mask.device)).sum()) / total).item()
        y_coord = (((mask.sum(dim=1) * torch.arange(h, device=mask.device)).sum()) / total).item()
    else:
        x_coord = (mask.sum(axis=0) @ np.arange(w)) / total
        y_coord = (mask.sum(axis=1) @ np.arange(h)) / total
    if normalize:
        x_coord, y_coord = x_coord / w, y_coord / h
    return x_coord, y_coord


def iou(mask, masks, eps=1e-6):
    mask = mask[None].astype(bool)
    masks = masks.astype(bool)
    i = (mask & masks).sum(axis=(1, 2))
    u = (mask | masks).sum(axis=(1, 2))
    return i / (u + eps)


def expand_overall_bboxes(overall_bboxes):
    return sum(overall_bboxes, start=[])


def shift_tensor(tensor, x_offset, y_offset, base_w=8, base_h=8, offset_normalized=False, ignore_last_dim=False):
    """utils/utils.py:143-178: zero-filled translation, offsets snapped to the 8x8 grid of the coarsest map."""
    tensor_h, tensor_w = tensor.shape[-3:-1] if ignore_last_dim else tensor.shape[-2:]
    if offset_normalized:
        assert tensor_h % base_h == 0 and tensor_w % base_w == 0, f"{tensor_h, tensor_w} is not a multiple of {base_h, base_w}"
        x_offset = round(x_offset * base_w) * (tensor_w // base_w)
        y_offset = round(y_offset * base_h) * (tensor_h // base_h)
    new_tensor = torch.zeros_like(tensor)
    overlap_w, overlap_h = tensor_w - abs(x_offset), tensor_h - abs(y_offset)
    y_src, y_dst = (0, y_offset) if y_offset >= 0 else (-y_offset, 0)
    x_src, x_dst = (0, x_offset) if x_offset >= 0 else (-x_offset, 0)
    if ignore_last_dim:
        new_tensor[..., y_dst:y_dst + overlap_h, x_dst:x_dst + overlap_w, :] = \
            tensor[..., y_src:y_src + overlap_h, x_src:x_src + overlap_w, :]
    else:
        new_tensor[..., y_dst:y_dst + overlap_h, x_dst:x_dst + overlap_w] = \
            tensor[..., y_src:y_src + overlap_h, x_src:x_src + overlap_w]
    return new_tensor
