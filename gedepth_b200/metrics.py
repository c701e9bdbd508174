"""Depth metrics on the device.

Mirrors depth/core/evaluation/metrics.py:8-45 (``calculate`` / ``metrics``), ``pre_eval_to_metrics`` (:77-100,
nan-mean over images) and the evaluation mask of depth/datasets/kitti.py:355-385 (``eval_kb_crop``, ``eval_mask``:
Garg / Eigen crop AND min/max depth).  The reference moves every prediction to the host (``.cpu().numpy()``,
depth/apis/test.py:209-218) and runs numpy per image; here one reduction kernel per batch leaves 10 fp64 sums per
image on the device and the host only sees the final numbers.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Optional

import torch

from . import ops

NAMES = ("a1", "a2", "a3", "abs_rel", "rmse", "log_10", "rmse_log", "silog", "sq_rel")


def kb_crop_window(height: int, width: int):
    """(top, left) of the 352x1216 KB crop the reference applies to the raw ground truth (kitti.py:355-366)."""
    return int(height - 352), int((width - 1216) / 2)


def crop_rect(height: int, width: int, garg_crop: bool = True, eigen_crop: bool = False):
    """(y0, y1, x0, x1) of the evaluation rectangle (kitti.py:373-383); None = whole image."""
    if garg_crop:
        return (int(0.40810811 * height), int(0.99189189 * height), int(0.03594771 * width), int(0.96405229 * width))
    if eigen_crop:
        return (int(0.3324324 * height), int(0.91351351 * height), int(0.0359477 * width), int(0.96405229 * width))
    return None


def finalize(sums: torch.Tensor) -> torch.Tensor:
    """(B,10) sums -> (B,9) metrics in the reference's order; images with an empty mask are NaN (metrics.py:9-10)."""
    s = sums.double()
    n = s[:, 0]
    mean = s[:, 1:] / n.unsqueeze(1)                      # n == 0 -> nan/inf rows, replaced below
    a1, a2, a3, abs_rel, sq_rel, mse, msle, merr, log10 = mean.unbind(1)
    silog = torch.sqrt(msle - merr * merr) * 100.0
    silog = torch.where(torch.isnan(silog), torch.zeros_like(silog), silog)     # metrics.py:29-31
    out = torch.stack([a1, a2, a3, abs_rel, torch.sqrt(mse), log10, torch.sqrt(msle), silog, sq_rel], 1)
    return torch.where((n > 0).unsqueeze(1), out, torch.full_like(out, float("nan")))


class DepthMetrics:
    """Accumulates per-image metrics over an evaluation run; ``compute()`` = ``pre_eval_to_metrics``."""

    def __init__(self, min_depth: float = 1e-3, max_depth: float = 80.0, garg_crop: bool = True,
                 eigen_crop: bool = False):
        self.min_depth, self.max_depth = float(min_depth), float(max_depth)
        self.garg_crop, self.eigen_crop = garg_crop, eigen_crop
        self._rows = []

    def update(self, pred: torch.Tensor, gt: torch.Tensor) -> torch.Tensor:
        """pred, gt: (B,H,W) or (B,1,H,W) CUDA tensors of the same size (gt already KB-cropped).  Returns (B,9)."""
        ops.require_cuda(pred, gt)
        H, W = gt.shape[-2], gt.shape[-1]
        sums = ops.depth_metric_sums(pred, gt, crop_rect(H, W, self.garg_crop, self.eigen_crop), self.min_depth,
                                     self.max_depth)
        per_image = finalize(sums)
        self._rows.append(per_image)
        return per_image

    def compute(self) -> "OrderedDict[str, float]":
        allm = torch.cat(self._rows, 0)
        mean = torch.nanmean(allm, 0).tolist()            # ONE device->host copy for the whole run
        return OrderedDict(zip(NAMES, mean))


def flip_tta_average(pred: torch.Tensor, pred_of_flipped: torch.Tensor) -> torch.Tensor:
    """encoder_decoder.py:226-233,262-270: un-flip the prediction of the mirrored image and average."""
    ops.require_cuda(pred, pred_of_flipped)
    return ops.tta_merge(pred, pred_of_flipped)
