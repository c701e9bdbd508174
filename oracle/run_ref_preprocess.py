"""ORACLE tooling: execute the reference's own offline script tools/preprocess_data_kitti.py VERBATIM (as a subprocess,
from /root/reference, build container only) on a small synthetic data/kitti tree and store what it writes as the golden
vectors of the ground-plane back-projection (a1, `pe_165.npy`, :15-56) and of the slope labels (f1, `find_k`, :59-89).

    python -m oracle.run_ref_preprocess        # rewrites tests/golden/ref_preprocess_kitti.npz

The tree: one date (2011_09_26) with calib_cam_to_cam.txt / calib_velo_to_cam.txt carrying the public calibration the
oracle hard-codes (oracle/ground.py), one 375 x 1242 image, one LiDAR-like 16-bit ground-truth PNG and a one-line split
file.  Tests regenerate the same ground truth from `synth_gt` and compare the oracle (CPU) and the CUDA kernels with the
script's outputs bit for bit."""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF_SCRIPT = "/root/reference/tools/preprocess_data_kitti.py"
OUT = os.path.join(ROOT, "tests", "golden", "ref_preprocess_kitti.npz")
H, W = 375, 1242
DATE, DRIVE = "2011_09_26", "2011_09_26_drive_0001_sync"


def synth_gt(seed: int = 7) -> np.ndarray:
    """uint16 KITTI-style ground truth (depth * 256, 0 = no return): ~6 % of the pixels, depths around the ground plane
    with +-30 % slope-inducing noise below the horizon, 5..80 m above it."""
    from oracle import ground as og
    rng = np.random.default_rng(seed)
    pe = og.ground_plane(og.plane_coefficients(og.kitti_projection(), og.KITTI_CAM_HEIGHT), H, W)
    ground = (pe > 0) & (pe < 80)
    d = np.where(ground, pe * (1 + 0.3 * rng.standard_normal((H, W))), rng.uniform(5, 80, (H, W)))
    d = np.clip(d, 0.5, 85.0)
    d = np.where(rng.random((H, W)) < 0.06, d, 0.0)
    return np.round(d * 256).astype(np.uint16)


def _fmt(vals):
    return " ".join(f"{v:.9e}" for v in np.asarray(vals, dtype=np.float64).ravel())


def build_tree(root: str):
    import cv2
    from oracle import ground as og
    d = os.path.join(root, "data", "kitti", "input", DATE)
    os.makedirs(os.path.join(d, DRIVE, "image_02", "data"))
    cam = [f"filler_{i}: 0" for i in range(34)]                    # KITTI's file has 34 lines; the script reads [8] and [25]
    cam[8] = "R_rect_00: " + _fmt(og.KITTI_R0)
    cam[25] = "P_rect_02: " + _fmt(og.KITTI_P2)
    open(os.path.join(d, "calib_cam_to_cam.txt"), "w").write("\n".join(cam) + "\n")
    velo = ["calib_time: 15-Mar-2012 11:37:16", "R: " + _fmt(og.KITTI_VELO_R), "T: " + _fmt(og.KITTI_VELO_T)]
    open(os.path.join(d, "calib_velo_to_cam.txt"), "w").write("\n".join(velo) + "\n")
    img = np.random.default_rng(0).integers(0, 256, (H, W, 3), dtype=np.uint8)
    cv2.imwrite(os.path.join(d, DRIVE, "image_02", "data", "0000000000.png"), img)
    gt_rel = f"{DRIVE}/proj_depth/groundtruth/image_02/0000000005.png"
    gt_path = os.path.join(root, "data", "kitti", "gt_depth", gt_rel)
    os.makedirs(os.path.dirname(gt_path))
    cv2.imwrite(gt_path, synth_gt())
    open(os.path.join(root, "data", "kitti", "kitti_eigen_train.txt"), "w").write(
        f"{DATE}/{DRIVE}/image_02/data/0000000005.png {gt_rel} 721.5377\n")
    # the script does `from IPython import embed` (unused): an empty stand-in package on PYTHONPATH
    os.makedirs(os.path.join(root, "_stubs", "IPython"))
    open(os.path.join(root, "_stubs", "IPython", "__init__.py"), "w").write("def embed(*a, **k):\n    pass\n")
    return gt_rel


def main():
    root = tempfile.mkdtemp(prefix="ged_ref_pre_")
    try:
        gt_rel = build_tree(root)
        env = dict(os.environ, PYTHONPATH=os.path.join(root, "_stubs"))
        r = subprocess.run([sys.executable, REF_SCRIPT], cwd=root, env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        pe = np.asarray(np.load(os.path.join(root, "data", "kitti", "input", DATE, "pe", "pe_165.npy")))
        k_path = os.path.join(root, "data", "kitti", "slope_range_5_5_interval_1", gt_rel.replace(".png", ".npz"))
        k = np.load(k_path)["k_img"]
        assert pe.shape == (H, W) and pe.dtype == np.float64 and k.shape == (H, W)
        np.savez_compressed(OUT, pe_sha256=np.array(hashlib.sha256(np.ascontiguousarray(pe).tobytes()).hexdigest()),
                            pe_rows=pe[::25].copy(), pe_f32_sha256=np.array(hashlib.sha256(pe.astype(np.float32).tobytes()).hexdigest()),
                            k_img=k.astype(np.int16), k_dtype=np.array(str(k.dtype)),
                            gt_sha256=np.array(hashlib.sha256(synth_gt().tobytes()).hexdigest()))
        print(f"{OUT}: pe {pe.dtype} {pe.shape}, k values {np.unique(k).tolist()}, {os.path.getsize(OUT) / 1024:.0f} KiB")
        print(r.stdout[-300:])
    finally:
        shutil.rmtree(root, ignore_errors=True)


if __name__ == "__main__":
    main()
