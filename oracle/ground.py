"""ORACLE (test infrastructure, not product code) - numpy restatement of the reference's CPU
ground-embedding path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this package.

Pinned against: the reference's own files executed under mmcv stubs in the build container
(oracle/ref_harness.py -> tests/golden/*.npz).  The reference ships no tests or golden vectors
(SURVEY.md §4), so for the offline numpy script below parity is pinned only by re-reading:
`tools/preprocess_data_kitti.py` has module-level disk I/O and cannot be executed; its eight
arithmetic lines are restated one for one.

Reference lines followed
  ground_plane_kitti   tools/preprocess_data_kitti.py:29-53
  ground_plane_ddad    tools/preprocess_data_ddad.py:30-41
  load_pe / load_pe_comput / concat   depth/datasets/pipelines/loading.py:366-403,524-527 (DDAD :851-887)
  normalize_pe         depth/datasets/pipelines/transforms.py:40-48
  find_k_kitti         tools/preprocess_data_kitti.py:59-63,86-89
  find_k_ddad          tools/preprocess_data_ddad.py:47-51,77-82
"""
from __future__ import annotations

import numpy as np

# KITTI 2011_09_26 calibration.  P_rect_02 is in the reference (depth/datasets/kitti.py:182-184);
# R_rect_00 / Tr_velo_to_cam are read from dataset files by the reference and are the public
# 2011_09_26 values [external].
KITTI_P2 = np.array([[721.5377, 0.0, 609.5593, 44.85728],
                     [0.0, 721.5377, 172.854, 0.2163791],
                     [0.0, 0.0, 1.0, 0.002745884]], dtype=np.float64)
KITTI_R0 = np.array([[0.9999239, 0.00983776, -0.007445048],
                     [-0.009869795, 0.9999421, -0.004278459],
                     [0.007402527, 0.004351614, 0.9999631]], dtype=np.float64)
KITTI_VELO_R = np.array([[0.007533745, -0.9999714, -0.000616602],
                         [0.01480249, 0.0007280733, -0.9998902],
                         [0.9998621, 0.00752379, 0.01480755]], dtype=np.float64)
KITTI_VELO_T = np.array([-0.004069766, -0.07631618, -0.2717806], dtype=np.float64)
KITTI_CAM_HEIGHT = 1.65
DDAD_CAM_HEIGHTS = (1.56, 1.57, 1.53, 1.53)  # tools/preprocess_data_ddad.py:68-75


def kitti_projection(P2=KITTI_P2, R0=KITTI_R0, velo_R=KITTI_VELO_R, velo_T=KITTI_VELO_T):
    """A = P2 * R0_rect(4x4) * Tr_velo_to_cam(4x4)  (preprocess_data_kitti.py:29-47)."""
    R0_4 = np.eye(4)
    R0_4[:3, :3] = R0
    Tr = np.eye(4)
    Tr[:3, :3] = velo_R
    Tr[:3, 3] = velo_T
    return P2 @ R0_4 @ Tr  # 3x4


def plane_coefficients(A: np.ndarray, height: float):
    """(numerator, c_u, c_v, c_1) with pe[v,u] = numerator / (c_u*u + c_v*v + c_1).

    Rinv = inv(A[:3,:3]); RT = Rinv @ A[:3,3]; numerator = RT[2] - height
    (preprocess_data_kitti.py:49-53; DDAD passes height=0, preprocess_data_ddad.py:36-41)."""
    A = np.asarray(A, dtype=np.float64)
    Rinv = np.linalg.inv(A[:3, :3])
    RT = Rinv @ A[:3, 3]
    return float(RT[2] - height), float(Rinv[2, 0]), float(Rinv[2, 1]), float(Rinv[2, 2])


def pixel_grid(H: int, W: int, u0: int = 0, v0: int = 0):
    """int64 (u, v) grid, 'xy' indexing (preprocess_data_kitti.py:52).  u0/v0 express a crop
    window of the full-resolution grid (KBCrop, transforms.py:179-180)."""
    u, v = np.meshgrid(np.arange(u0, u0 + W, dtype=np.int64), np.arange(v0, v0 + H, dtype=np.int64),
                       indexing="xy")
    return u, v


def ground_plane(coef, H: int, W: int, u0: int = 0, v0: int = 0) -> np.ndarray:
    """float64 H x W ground-plane depth (preprocess_data_kitti.py:53)."""
    num, cu, cv, c1 = coef
    u, v = pixel_grid(H, W, u0, v0)
    with np.errstate(divide="ignore", invalid="ignore"):
        return num / (cu * u + cv * v + c1)


def ground_plane_kitti(H: int, W: int, u0: int = 0, v0: int = 0, height: float = KITTI_CAM_HEIGHT):
    return ground_plane(plane_coefficients(kitti_projection(), height), H, W, u0, v0)


def ddad_projection(fx, fy, cx, cy, pitch_rad: float, cam_height: float):
    """Synthetic DDAD-style A = K @ inv(cam_pose) @ lidar_pose with a pitch-only camera mounted
    cam_height above the plane z=0 of the lidar-pose frame (the reference
    reads real poses from the TRI dgp SDK, preprocess_data_ddad.py:25-35, which is not installed)."""
    K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1.0]])
    # camera axes in world (x fwd, y left, z up): cam z = forward tilted by pitch, cam y = down
    c, s = np.cos(pitch_rad), np.sin(pitch_rad)
    R_wc = np.array([[0.0, -s, c], [-1.0, 0.0, 0.0], [0.0, -c, -s]])  # columns: cam x,y,z in world
    pose = np.eye(4)
    pose[:3, :3] = R_wc
    pose[:3, 3] = [0.0, 0.0, cam_height]
    A4 = np.eye(4)
    A4[:3, :3] = K
    return (A4 @ np.linalg.inv(pose))[:3, :]


def load_channels(pe: np.ndarray, clamp_max: float = 200.0):
    """(ch3, ch4) before Normalize: ch3 = pe with >clamp_max -> 0 and <0 -> 0, ch4 = raw pe, both
    cast to float32 first (loading.py:375,397-401; DDAD clamps at 250, :863-864)."""
    pe32 = pe.astype(np.float32)
    ch3 = pe32.copy()
    ch3[ch3 > clamp_max] = 0
    ch3[ch3 < 0] = 0
    return ch3, pe32.copy()


def normalize_pe(ch3: np.ndarray, depth_scale: float = 200.0) -> np.ndarray:
    """ch3[ch3>0] /= depth_scale (transforms.py:43-44)."""
    out = ch3.copy()
    m = out > 0
    out[m] = out[m] / np.float32(depth_scale)
    return out


def normalize_rgb(img_u8: np.ndarray, mean, std) -> np.ndarray:
    """mmcv.imnormalize(img, mean, std, to_rgb=True) [external]: BGR->RGB, (x-mean)/std in fp32."""
    img = img_u8[..., ::-1].astype(np.float32)
    mean = np.asarray(mean, dtype=np.float64).reshape(1, -1)
    stdinv = 1.0 / np.asarray(std, dtype=np.float64).reshape(1, -1)
    return ((img - mean.astype(np.float32)) * stdinv.astype(np.float32)).astype(np.float32)


def find_k_kitti(gt: np.ndarray, pe: np.ndarray, h: float = KITTI_CAM_HEIGHT) -> np.ndarray:
    """Slope labels, degrees rounded half-to-even, clipped to +-5, 255 where gt==0
    (preprocess_data_kitti.py:59-63,80-89).  pe is float32 (cast at :81), gt float64."""
    with np.errstate(divide="ignore", invalid="ignore"):
        k = (h / gt) + ((-h) / pe.astype(np.float32))
        k = np.around(np.rad2deg(np.arctan(k)))
    k[k > 5] = 5
    k[k < -5] = -5
    k[gt == 0] = 255
    return k


def find_k_ddad(gt: np.ndarray, pe: np.ndarray, h: float) -> np.ndarray:
    """DDAD variant: int64 truncation instead of rounding (preprocess_data_ddad.py:47-51,77-82)."""
    with np.errstate(divide="ignore", invalid="ignore"):
        k = (h / gt) + ((-h) / pe)
        k = np.rad2deg(np.arctan(k))
    k = np.where(np.isfinite(k), k, 0).astype(np.int64)
    k[k > 5] = 5
    k[k < -5] = -5
    k[gt == 0] = 255
    return k


def abs_rel(gt: np.ndarray, pred: np.ndarray, min_depth=1e-3, max_depth=80.0) -> float:
    """depth/core/evaluation/metrics.py:17 with the (min,max) mask of :35-45."""
    m = np.logical_and(gt > min_depth, gt < max_depth)
    g, p = gt[m], pred[m]
    if g.shape[0] == 0:
        return float("nan")
    return float(np.mean(np.abs(g - p) / g))


def depth_metrics(gt: np.ndarray, pred: np.ndarray, min_depth=1e-3, max_depth=80.0):
    """All nine numbers of metrics.py:8-33 in the reference's order."""
    m = np.logical_and(gt > min_depth, gt < max_depth)
    gt, pred = gt[m], pred[m]
    if gt.shape[0] == 0:
        return (np.nan,) * 9
    thresh = np.maximum(gt / pred, pred / gt)
    a1, a2, a3 = (thresh < 1.25).mean(), (thresh < 1.25 ** 2).mean(), (thresh < 1.25 ** 3).mean()
    ar = np.mean(np.abs(gt - pred) / gt)
    sq = np.mean(((gt - pred) ** 2) / gt)
    rmse = np.sqrt(((gt - pred) ** 2).mean())
    rmse_log = np.sqrt(((np.log(gt) - np.log(pred)) ** 2).mean())
    err = np.log(pred) - np.log(gt)
    silog = np.sqrt(np.mean(err ** 2) - np.mean(err) ** 2) * 100
    if np.isnan(silog):
        silog = 0
    log10 = np.abs(np.log10(gt) - np.log10(pred)).mean()
    return a1, a2, a3, ar, rmse, log10, rmse_log, silog, sq
