"""Deterministic synthetic weights and KITTI/DDAD-shape inputs.

There is no network for checkpoints or datasets, and parity tests must never depend on matching
initialisers, so every tensor of a ``state_dict`` is generated from its *name*:
the reference model (oracle/ref_harness.py), the oracle and the CUDA build all load the same bytes.
Pure host-side numpy/torch-CPU code; no oracle import.
"""
from __future__ import annotations

import zlib
from typing import Dict

import numpy as np
import torch

KITTI_MEAN = [123.675, 116.28, 103.53]
KITTI_STD = [58.395, 57.12, 57.375]


def synth_state_dict(template: Dict[str, torch.Tensor], seed: int = 0) -> Dict[str, torch.Tensor]:
    """Name-keyed deterministic fill of a state_dict template (shapes/dtypes taken from it).

    weights (dim>1): N(0, 1/fan_in) - keeps activations O(1) through 12-24 blocks; biases N(0,0.02);
    norm weights 1+0.1 N; BN running_var 1+0.1|N|; integer buffers are kept."""
    out = {}
    for name in sorted(template.keys()):
        t = template[name]
        if not t.dtype.is_floating_point:
            out[name] = t.clone()
            continue
        g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(name.encode())) % (2 ** 31))
        r = torch.randn(t.shape, generator=g, dtype=torch.float32)
        leaf = name.rsplit(".", 1)[-1]
        is_norm = any(s in name for s in (".norm", ".bn", "norm.", ".ln")) or name.startswith("backbone.norm")
        if leaf == "running_var":
            v = 1.0 + 0.1 * r.abs()
        elif leaf == "running_mean":
            v = 0.05 * r
        elif leaf == "relative_position_bias_table":
            v = 0.2 * r
        elif leaf == "level_embed":
            v = r
        elif is_norm and leaf == "weight":
            v = 1.0 + 0.1 * r
        elif leaf == "bias":
            v = 0.02 * r
        elif t.dim() > 1:
            fan_in = int(np.prod(t.shape[1:]))
            v = r / np.sqrt(fan_in)
        else:
            v = 0.02 * r
        if "sampling_offsets.bias" in name:
            v = 1.5 * r          # spread the deformable sampling points over a few pixels
        if "conv_depth.bias" in name:
            v = v + 10.0         # metric depth head: keep relu(conv) alive and O(10 m)
        out[name] = v.to(t.dtype)
    return out


def kitti_plane_coef(height: float = 1.65):
    """(num, c_u, c_v, c_1) of the public KITTI 2011_09_26 calibration - fp64 host math, the
    eight lines of tools/preprocess_data_kitti.py:29-53 (P_rect_02: depth/datasets/kitti.py:182-184)."""
    P2 = np.array([[721.5377, 0.0, 609.5593, 44.85728], [0.0, 721.5377, 172.854, 0.2163791],
                   [0.0, 0.0, 1.0, 0.002745884]])
    R0 = np.eye(4)
    R0[:3, :3] = [[0.9999239, 0.00983776, -0.007445048], [-0.009869795, 0.9999421, -0.004278459],
                  [0.007402527, 0.004351614, 0.9999631]]
    Tr = np.eye(4)
    Tr[:3, :3] = [[0.007533745, -0.9999714, -0.000616602], [0.01480249, 0.0007280733, -0.9998902],
                  [0.9998621, 0.00752379, 0.01480755]]
    Tr[:3, 3] = [-0.004069766, -0.07631618, -0.2717806]
    A = P2 @ R0 @ Tr
    Rinv = np.linalg.inv(A[:3, :3])
    RT = Rinv @ A[:3, 3]
    return float(RT[2] - height), float(Rinv[2, 0]), float(Rinv[2, 1]), float(Rinv[2, 2])


def synth_batch(B: int, H: int, W: int, seed: int = 1234, depth_scale: float = 200.0,
                max_depth: float = 80.0, adaptive: bool = False, u0: int = 61, v0: int = 23,
                sparsity: float = 0.05):
    """Host (numpy) batch of the reference's input contract: img (B,5,H,W) fp32 with ch0-2 =
    normalised RGB, ch3 = clamp(pe,0,S)/S, ch4 = raw pe (loading.py:388-403,524-527,
    transforms.py:40-48); depth_gt (B,1,H,W) LiDAR-sparse; pe_k_gt (B,H,W) in {0..10}|255."""
    rng = np.random.default_rng(seed)
    num, cu, cv, c1 = kitti_plane_coef()
    # full-resolution crop window scaled so any (H,W) sees the horizon near 40-50 % of the height
    sy, sx = 352.0 / H, 1120.0 / W
    u = (np.arange(W, dtype=np.float64) * sx + u0)[None, :]
    v = (np.arange(H, dtype=np.float64) * sy + v0)[:, None]
    with np.errstate(divide="ignore"):
        pe = (num / (cu * u + cv * v + c1)).astype(np.float32)
    ch3 = pe.copy()
    ch3[ch3 > depth_scale] = 0
    ch3[ch3 < 0] = 0
    ch3[ch3 > 0] /= np.float32(depth_scale)
    rgb = rng.integers(0, 256, (B, H, W, 3), dtype=np.uint8).astype(np.float32)
    rgb = (rgb - np.array(KITTI_MEAN, np.float32)) / np.array(KITTI_STD, np.float32)
    img = np.empty((B, 5, H, W), np.float32)
    img[:, 0:3] = rgb.transpose(0, 3, 1, 2)
    img[:, 3] = ch3
    img[:, 4] = pe
    ground = (pe > 0) & (pe <= max_depth)
    gt = np.where(ground[None], pe[None] * (1 + 0.05 * rng.standard_normal((B, H, W))),
                  rng.uniform(5, max_depth, (B, H, W))).astype(np.float32)
    gt = np.clip(gt, 0.5, max_depth)
    keep = rng.random((B, H, W)) < sparsity
    gt = np.where(keep, gt, 0).astype(np.float32)[:, None]
    out = dict(img=img, depth_gt=gt)
    if adaptive:
        k = rng.integers(0, 11, (B, H, W)).astype(np.float32)
        out["pe_k_gt"] = np.where(gt[:, 0] > 0, k, 255).astype(np.float32)
    return out


def synth_raw_frame(seed: int, h: int = 375, w: int = 1242):
    """Un-augmented KITTI-shaped frame as the train loader stacks it (loading.py:362,388-403,524-527): (h, w, 5) float32 =
    BGR image as float, the clamped ground-plane map, the raw one; sparse depth (h, w); slope labels (h, w) in {0..10} with
    255 where there is no LiDAR return.  Input of the device augmentation (augment.py) in bench.py and the tests."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    base = np.stack([127 + 90 * np.sin(xx / (37.0 + 5 * c) + c) * np.cos(yy / (23.0 + 3 * c)) for c in range(3)], -1)
    bgr = np.clip(base + rng.normal(0, 12, (h, w, 3)), 0, 255).astype(np.uint8)
    num, cu, cv, c1 = kitti_plane_coef()
    with np.errstate(divide="ignore"):
        pe = (num / (cu * np.arange(w, dtype=np.float64)[None, :] + cv * np.arange(h, dtype=np.float64)[:, None] + c1)).astype(np.float32)
    pe_c = pe.copy()
    pe_c[pe_c > 200] = 0
    pe_c[pe_c < 0] = 0
    img5 = np.concatenate([bgr.astype(np.float32), pe_c[:, :, None], pe[:, :, None]], -1).astype(np.float32)
    keep = rng.random((h, w)) < 0.06
    depth = np.where(keep, rng.uniform(2, 80, (h, w)), 0).astype(np.float32)
    lab = np.where(keep, rng.integers(0, 11, (h, w)), 255).astype(np.float32)
    return img5, depth, lab
