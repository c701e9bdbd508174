"""Train-time augmentation of the 5-channel input on the device.

The reference augments on CPU data-loader workers (configs/depthformer/depthformer_v.py:13-33): ``KBCrop`` ->
``Resize(ratio_range=(0.5, 2.0))`` -> ``Padding`` -> ``RandomRotate(prob=0.5, degree=2.5)`` -> ``RandomFlip(prob=0.5)`` ->
``RandomCrop((352, 704))`` -> ``ColorAug(prob=0.5)`` -> ``Normalize`` (depth/datasets/pipelines/transforms.py), i.e. about 30 MB
of float32 resampling per frame in cv2 before a single byte reaches the GPU.  Here the un-augmented frame is uploaded once
(uint8 image, the two ground-plane maps, sparse depth, slope labels) and two kernels (csrc/augment.cu) produce the
network's input - with OpenCV's own arithmetic, so the tensors are bit-identical to the reference's for the same drawn
parameters.  ``draw_params`` draws those parameters from numpy's global RandomState and python's ``random`` in the
reference's order: seeding both the way the reference's workers are seeded reproduces its augmentation stream.
"""
from __future__ import annotations

import ctypes as C
import math
import random as _pyrandom
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import kernels as K
from .inputs import KITTI_MEAN, KITTI_STD

KB_H, KB_W = 352, 1216          # KBCrop / Padding target (transforms.py:159,66)
CROP_HW = (352, 704)            # RandomCrop (depthformer_v.py:23)


def draw_params(h: int = KB_H, w: int = KB_W, crop=CROP_HW, ratio_range=(0.5, 2.0), rotate_prob=0.5, degree=2.5,
                flip_prob=0.5, color_prob=0.5, gamma_range=(0.9, 1.1), brightness_range=(0.9, 1.1),
                color_range=(0.9, 1.1), scale=None) -> Dict:
    """One frame's augmentation parameters, drawn exactly as the transforms draw them (same generators, same order)."""
    p: Dict = {}
    if scale is None:
        ratio = np.random.random_sample() * (ratio_range[1] - ratio_range[0]) + ratio_range[0]      # Resize :611
        scale = (int(w * ratio), int(h * ratio))
    sf = min(max(scale) / max(h, w), min(scale) / min(h, w))                                        # mmcv.rescale_size
    nw, nh = int(w * float(sf) + 0.5), int(h * float(sf) + 0.5)
    p["new_w"], p["new_h"] = nw, nh
    if nh < h or nw < w:                                                                            # Padding :82-87
        p["pad_y"] = _pyrandom.randint(0, h - nh)
        p["pad_x"] = _pyrandom.randint(0, w - nw)
        ch, cw = h, w
    else:
        p["pad_y"] = p["pad_x"] = 0
        ch, cw = nh, nw
    p["canvas_h"], p["canvas_w"] = ch, cw
    p["rotate"] = bool(np.random.rand() < rotate_prob)                                              # RandomRotate :262-263
    p["degree"] = float(np.random.uniform(-degree, degree))
    p["flip"] = bool(np.random.rand() < flip_prob)                                                  # RandomFlip :333
    p["crop_y"] = int(np.random.randint(0, max(ch - crop[0], 0) + 1))                               # RandomCrop :371-374
    p["crop_x"] = int(np.random.randint(0, max(cw - crop[1], 0) + 1))
    p["color"] = bool(np.random.rand() < color_prob)                                                # ColorAug :447
    if p["color"]:
        p["gamma"] = float(np.random.uniform(min(*gamma_range), max(*gamma_range)))
        p["brightness"] = float(np.random.uniform(min(*brightness_range), max(*brightness_range)))
        p["colors"] = [float(c) for c in np.random.uniform(min(*color_range), max(*color_range), size=3)]
    else:
        p["gamma"], p["brightness"], p["colors"] = 1.0, 1.0, [1.0, 1.0, 1.0]
    return p


def _inverse_rotation(w: int, h: int, degree: float) -> List[float]:
    """mmcv.imrotate: cv2.getRotationMatrix2D(((w-1)/2, (h-1)/2), -degree, 1.0), then the inversion cv2.warpAffine applies
    (imgwarp.cpp), all in double."""
    ang = -degree * (math.pi / 180.0)                   # cv2: angle *= CV_PI / 180
    al, be = float(np.cos(ang)), float(np.sin(ang))
    cx, cy = (w - 1) * 0.5, (h - 1) * 0.5
    m = [al, be, (1 - al) * cx - be * cy, -be, al, be * cx + (1 - al) * cy]
    D = m[0] * m[4] - m[1] * m[3]
    D = 1.0 / D if D != 0 else 0.0
    a11, a22 = m[4] * D, m[0] * D
    m[0] = a11; m[1] *= -D; m[3] *= -D; m[4] = a22
    b1 = -m[0] * m[2] - m[1] * m[5]
    b2 = -m[3] * m[2] - m[4] * m[5]
    m[2] = b1; m[5] = b2
    return m


class TrainAugmenter:
    """Device-side KITTI train pipeline.  Holds the canvas workspace (sized for the largest resize: 2 x the KB window)."""

    def __init__(self, device, crop_hw=CROP_HW, kb_hw=(KB_H, KB_W), max_ratio: float = 2.0, mean=KITTI_MEAN, std=KITTI_STD,
                 depth_scale: float = 200.0):
        self.device = torch.device(device)
        self.crop_hw, self.kb_hw = tuple(crop_hw), tuple(kb_hw)
        self.depth_scale = float(depth_scale)
        mh, mw = int(kb_hw[0] * max_ratio) + 2, int(kb_hw[1] * max_ratio) + 2
        self._slot = 7 * mh * mw                                                            # 5 image planes, depth, labels
        self.canvas = torch.empty(self._slot, dtype=torch.float32, device=self.device)      # grows to one slot per frame
        self._mean = (C.c_float * 3)(*mean)
        self._std = (C.c_float * 3)(*std)
        self._frame_bytes = int(K.load().ged_aug_frame_bytes())
        self._desc_dev = torch.empty(0, dtype=torch.uint8, device=self.device)

    def frame_planes(self, bgr_u8: torch.Tensor, pe_clamped: torch.Tensor, pe_raw: torch.Tensor) -> torch.Tensor:
        """(5, H0, W0) float32 frame as the loader stacks it (loading.py:524-527): BGR as float, the clamped plane map
        (load_pe) and the raw one (load_pe_comput)."""
        H0, W0 = int(bgr_u8.shape[0]), int(bgr_u8.shape[1])
        assert bgr_u8.is_cuda and bgr_u8.dtype == torch.uint8 and bgr_u8.is_contiguous()
        out = torch.empty(5, H0, W0, dtype=torch.float32, device=self.device)
        K._call("ged_aug_u8_to_planes", K._p(bgr_u8), K._p(out), H0, W0, K._stream())
        out[3].copy_(pe_clamped)
        out[4].copy_(pe_raw)
        return out

    def __call__(self, frames: Sequence[torch.Tensor], depth_gt: Sequence[torch.Tensor], pe_k_gt: Sequence[torch.Tensor],
                 params: Sequence[Dict]):
        """frames[i] (5, H0, W0) float32, depth_gt[i] / pe_k_gt[i] (H0, W0) float32, params[i] from ``draw_params``.
        Returns img (B, 5, h, w), depth_gt (B, 1, h, w), pe_k_gt (B, h, w)."""
        B = len(frames)
        oh, ow = self.crop_hw
        img = torch.empty(B, 5, oh, ow, dtype=torch.float32, device=self.device)
        dep = torch.empty(B, 1, oh, ow, dtype=torch.float32, device=self.device)
        lab = torch.empty(B, oh, ow, dtype=torch.float32, device=self.device)
        sh, sw = self.kb_hw
        if self.canvas.numel() < B * self._slot:
            self.canvas = torch.empty(B * self._slot, dtype=torch.float32, device=self.device)
        if self._desc_dev.numel() < B * self._frame_bytes:
            self._desc_dev = torch.empty(B * self._frame_bytes, dtype=torch.uint8, device=self.device)
        desc = C.create_string_buffer(B * self._frame_bytes)
        for i in range(B):
            f, d, l, p = frames[i], depth_gt[i], pe_k_gt[i], params[i]
            H0, W0 = int(f.shape[1]), int(f.shape[2])
            top, left = int(H0 - sh), int((W0 - sw) / 2)                                             # KBCrop :177-178
            cw, ch = int(p["canvas_w"]), int(p["canvas_h"])
            if 7 * cw * ch > self._slot:
                raise ValueError(f"canvas {cw}x{ch} exceeds the workspace")
            minv = (C.c_double * 6)(*(_inverse_rotation(cw, ch, p["degree"]) if p["rotate"] else [0.0] * 6))
            colors = (C.c_double * 3)(*[float(c) for c in p["colors"]])
            rc = K.load().ged_aug_pack_frame(
                C.cast(desc, C.c_void_p), i, K._p(f), K._p(d), K._p(l), K._p(self.canvas[i * self._slot:]), K._p(img[i]), K._p(dep[i]),
                K._p(lab[i]), H0, W0, top, left, sh, sw, int(p["new_w"]), int(p["new_h"]), int(p["pad_x"]), int(p["pad_y"]), cw, ch,
                C.cast(minv, C.c_void_p), int(p["rotate"]), int(p["flip"]), int(p["crop_x"]), int(p["crop_y"]), ow, oh,
                int(p["color"]), float(np.float32(p["gamma"])), float(np.float32(p["brightness"])), C.cast(colors, C.c_void_p),
                C.cast(self._mean, C.c_void_p), C.cast(self._std, C.c_void_p), self.depth_scale)
            if rc != 0:
                raise RuntimeError(f"ged_aug_pack_frame failed for frame {i}: {rc}")
        # cudaMemcpyAsync from pageable memory returns once `desc` has been staged, so the buffer may go out of scope here
        K._call("ged_aug_train_batch", C.cast(desc, C.c_void_p), B, K._p(self._desc_dev), K._stream())
        return img, dep, lab
