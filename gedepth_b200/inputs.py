"""Test-time input construction on the device (part of SURVEY.md §8(f) row 3).

The reference's test pipeline (configs/depthformer/depthformer_v.py:33-53) runs on CPU workers:
``LoadImageFromFile(USEPE)`` reads the image and the pre-computed ``pe_165.npy`` and stacks them to H x W x 5
(depth/datasets/pipelines/loading.py:490-527), ``KBCrop`` takes the 352 x 1216 window (transforms.py:176-197),
``RandomFlip`` mirrors the flip-TTA view, ``Normalize`` applies mmcv.imnormalize to RGB and divides the clamped
ground depth by ``depth_scale`` (transforms.py:40-48).  Here the uint8 image goes to the GPU once and both views
are produced there: RGB by one crop + flip + normalise kernel (bit-exact with cv2's arithmetic), channels 3 / 4 by
evaluating the ground plane analytically on the cropped (and mirrored) pixel grid - no ``.npy`` file, no CPU pass.
"""
from __future__ import annotations

from typing import Sequence

import torch

from . import ops
from .metrics import kb_crop_window

KITTI_MEAN = (123.675, 116.28, 103.53)
KITTI_STD = (58.395, 57.12, 57.375)


def build_test_views(bgr_u8: torch.Tensor, plane_coef: Sequence[float], crop_hw=(352, 1216), mean=KITTI_MEAN,
                     std=KITTI_STD, to_rgb: bool = True, depth_scale: float = 200.0, clamp_max: float = 200.0,
                     flip_view: bool = True):
    """bgr_u8: (H0, W0, 3) uint8 CUDA tensor (as cv2.imread gives it).  plane_coef: (num, c_u, c_v, c_1) of
    pe[v,u] = num / (c_u u + c_v v + c_1) (preprocess_data_kitti.py:47-53).
    Returns (imgs, img_metas, pe_ori_point): imgs = [(1,5,H,W) plain view, (1,5,H,W) mirrored view] ready for
    ``model(img=imgs, img_metas=img_metas, return_loss=False, pe_ori_point=[p, p])``."""
    ops.require_cuda(bgr_u8)
    k = ops._k()
    H0, W0 = int(bgr_u8.shape[0]), int(bgr_u8.shape[1])
    H, W = crop_hw
    top, left = kb_crop_window(H0, W0) if (H0, W0) != (H, W) else (0, 0)
    num, cu, cv, c1 = [float(v) for v in plane_coef]
    pe_ori_point = num / (cu * (W0 - 1) + cv * (H0 - 1) + c1)        # pe[-1,-1] of the un-cropped map (loading.py:524)
    imgs, metas = [], []
    for flip in ([False, True] if flip_view else [False]):
        img = torch.empty(1, 5, H, W, dtype=torch.float32, device=bgr_u8.device)
        k.rgb_crop_normalize_into(img[0], bgr_u8, top, left, flip, mean, std, to_rgb)
        # mirrored view: u = left + (W-1-x)
        k.ground_plane_into(img, plane_coef, u0=(left + W - 1) if flip else left, v0=top, depth_scale=depth_scale,
                            clamp_max=clamp_max, su=-1.0 if flip else 1.0, sv=1.0)
        imgs.append(img)
        metas.append([dict(ori_shape=(H, W, 3), img_shape=(H, W, 3), pad_shape=(H, W, 3), flip=flip,
                           flip_direction="horizontal" if flip else None,
                           img_norm_cfg=dict(mean=mean, std=std, to_rgb=to_rgb))])
    p = torch.tensor([pe_ori_point], dtype=torch.float32)
    return imgs, metas, [p] * len(imgs)
