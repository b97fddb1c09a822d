"""Mirror of the box epilogue of reference core/utils/image_process.py: letterbox inverse
(`reverse_letter_box_numpy` :69-97, `reverse_letter_box` :100-129, `yolo_correct_boxes` :161-181).

Host arithmetic kept exactly as the reference performs it: the scale and padding are Python doubles,
the box arrays stay float32 and every in-place op rounds once in float32."""
from __future__ import annotations

import numpy as np
import torch


def _letterbox_params(image_h, image_w, input_h, input_w):
    scale = max(image_h / input_h, image_w / input_w)
    top = (input_h - image_h / scale) // 2
    left = (input_w - image_w / scale) // 2
    return scale, top, left


def reverse_letter_box_numpy(image_shape, input_shape, boxes, xywh=True):
    """boxes normalised to the network input -> pixels of the original image (numpy)."""
    if xywh:
        out = np.concatenate((boxes[..., 0:2] - boxes[..., 2:4] / 2, boxes[..., 0:2] + boxes[..., 2:4] / 2), axis=-1)
    else:
        out = boxes.copy()
    out[..., 0::2] *= input_shape[1]
    out[..., 1::2] *= input_shape[0]
    scale, top, left = _letterbox_params(image_shape[0], image_shape[1], input_shape[0], input_shape[1])
    out[..., 0] -= left
    out[..., 2] -= left
    out[..., 1] -= top
    out[..., 3] -= top
    out *= scale
    return out


def reverse_letter_box(h, w, input_size, boxes, xywh=True):
    """torch twin of reverse_letter_box_numpy.  CUDA tensors go through cvpp_letterbox_reverse (one kernel);
    CPU tensors (the reference also calls this on host tensors) keep the reference's eager arithmetic."""
    if boxes.is_cuda and boxes.dtype == torch.float32 and boxes.shape[-1] == 4 and boxes.numel() > 0:
        from ... import ops
        return ops.letterbox_reverse(boxes, h, w, input_size, xywh)
    if xywh:
        out = torch.cat((boxes[..., 0:2] - boxes[..., 2:4] / 2, boxes[..., 0:2] + boxes[..., 2:4] / 2), dim=-1)
    else:
        out = boxes.clone()
    out[..., 0::2] *= input_size[1]
    out[..., 1::2] *= input_size[0]
    scale, top, left = _letterbox_params(h, w, input_size[0], input_size[1])
    out[..., 0] -= left
    out[..., 2] -= left
    out[..., 1] -= top
    out[..., 3] -= top
    out *= scale
    return out


def yolo_correct_boxes(box_xy, box_wh, input_shape, image_shape, letterbox_image):
    """Normalised (centre, size) boxes -> (xmin, ymin, xmax, ymax) pixels of the original image."""
    xywh = np.concatenate([box_xy, box_wh], axis=-1)
    if letterbox_image:
        return reverse_letter_box_numpy(image_shape, input_shape, xywh, xywh=True)
    half = xywh[..., 2:4] / 2
    out = np.concatenate((xywh[..., 0:2] - half, xywh[..., 0:2] + half), axis=-1)
    out[:, 0::2] *= image_shape[1]
    out[:, 1::2] *= image_shape[0]
    return out
