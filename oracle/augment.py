"""ORACLE (test infrastructure: only tests/, smoke() and bench.py's cpu_baseline leg may import this).

CPU / numpy restatement of the reference's KITTI TRAIN-time augmentation of the 5-channel input:

    configs/depthformer/depthformer_v.py:13-33   KBCrop -> Resize(ratio_range=(0.5, 2.0)) -> Padding -> RandomRotate(0.5, 2.5)
                                                 -> RandomFlip(0.5) -> RandomCrop(352, 704) -> ColorAug(0.5) -> Normalize
    depth/datasets/pipelines/transforms.py       KBCrop :149-205, Resize :484-732, Padding :64-109, RandomRotate :208-296,
                                                 RandomFlip :299-353, RandomCrop :356-417, ColorAug :420-481, Normalize :12-61

The reference runs these through mmcv 1.3.13 (image/geometric.py: imrescale / imresize / imrotate / imflip,
image/photometric.py: imnormalize), i.e. through OpenCV.  OpenCV is a third-party dependency that is absent from
/root/reference; its algorithms for the calls on this path are restated here from the published sources
(modules/imgproc/src/resize.cpp, imgwarp.cpp) and pinned bit for bit against cv2 4.13 run in the build container
(tests/test_augment_host.py, oracle/make_golden_aug.py):

  * cv2.resize(float32, INTER_LINEAR): per destination column fx = float((dx + 0.5) * scale - 0.5), sx = floor(fx),
    fx -= sx, clamped with fx = 0 at both borders; horizontal pass S[sx] * (1 - fx) + S[sx + 1] * fx, then the vertical
    pass with the same formula but UNCLAMPED weights (both taps fall on the border row there); each product and sum
    rounded to float32 (no FMA).  Exactly 2x down-scaling switches to INTER_AREA: (a + b + c + d) * 0.25.
  * cv2.resize(INTER_NEAREST): sx = min(floor(dx * scale), w - 1).
  * cv2.warpAffine(float32, INTER_LINEAR, BORDER_CONSTANT): the matrix is inverted in double; coordinates in fixed point
    with 10 fractional bits, adelta[x] = rint(M0 x 1024), X0 = rint((M1 y + M2) 1024) + 16, X = (X0 + adelta[x]) >> 5:
    sx = X >> 5, fx = X & 31 (1/32 pixel); weights (1 - fy)(1 - fx) ... from a table of float products; value =
    ((v0 w0 + v1 w1) + v2 w2) + v3 w3 in float32, taps outside the image replaced by the border value.
    INTER_NEAREST: X = (X0 + adelta[x]) >> 10 with the rounding constant 512.
  * mmcv.imnormalize: uint8 -> float32, BGR -> RGB, (x - mean) * (1 / std) with the constants held in double.
  * ColorAug's `image ** gamma` is numpy's float32 power, whose rounding depends on the numpy build / CPU; the oracle uses
    the correctly rounded power (see train_augment).  Everything else is bit-exact against the reference.

The random draws of the transforms are taken in the reference's order from numpy's global RandomState and python's
`random` module (`draw_params`), so that the same seeds give the same augmentation.
"""
from __future__ import annotations

import random as _pyrandom

import numpy as np

F32 = np.float32
KB_H, KB_W = 352, 1216          # KBCrop / Padding target (transforms.py:159,66)
CROP_H, CROP_W = 352, 704       # RandomCrop (depthformer_v.py:23)
MEAN = np.array([123.675, 116.28, 103.53], np.float32)   # RGB order (depthformer_v.py:11)
STD = np.array([58.395, 57.12, 57.375], np.float32)


# ----------------------------------------------------------------------------------------------------------------------
# random parameters, in the order the transforms draw them
# ----------------------------------------------------------------------------------------------------------------------
def draw_params(h: int = KB_H, w: int = KB_W, crop=(CROP_H, CROP_W), scale=None) -> dict:
    """Consumes numpy's global RandomState and python's `random` exactly as the pipeline does for a (h, w) KB-cropped frame.
    `scale` = a preset results['scale'] (Resize.__call__ :724 then draws nothing)."""
    p = {}
    if scale is None:
        ratio = np.random.random_sample() * (2.0 - 0.5) + 0.5                   # Resize.random_sample_ratio :611
        scale = (int(w * ratio), int(h * ratio))                                # (w, h) tuple :612
    sf = min(max(scale) / max(h, w), min(scale) / min(h, w))                    # mmcv rescale_size
    nw, nh = int(w * float(sf) + 0.5), int(h * float(sf) + 0.5)
    p["new_w"], p["new_h"] = nw, nh
    if nh < KB_H or nw < KB_W:                                                  # Padding :82-87
        p["pad_y"] = _pyrandom.randint(0, KB_H - nh)
        p["pad_x"] = _pyrandom.randint(0, KB_W - nw)
        ch, cw = KB_H, KB_W
    else:
        p["pad_y"] = p["pad_x"] = 0
        ch, cw = nh, nw
    p["canvas_h"], p["canvas_w"] = ch, cw
    p["rotate"] = bool(np.random.rand() < 0.5)                                  # RandomRotate :262-263
    p["degree"] = float(np.random.uniform(-2.5, 2.5))
    p["flip"] = bool(np.random.rand() < 0.5)                                    # RandomFlip :333
    p["crop_y"] = int(np.random.randint(0, max(ch - crop[0], 0) + 1))           # RandomCrop :371-374
    p["crop_x"] = int(np.random.randint(0, max(cw - crop[1], 0) + 1))
    p["color"] = bool(np.random.rand() < 0.5)                                   # ColorAug :447
    if p["color"]:
        p["gamma"] = float(np.random.uniform(0.9, 1.1))
        p["brightness"] = float(np.random.uniform(0.9, 1.1))
        p["colors"] = [float(c) for c in np.random.uniform(0.9, 1.1, size=3)]
    else:
        p["gamma"], p["brightness"], p["colors"] = 1.0, 1.0, [1.0, 1.0, 1.0]
    return p


# ----------------------------------------------------------------------------------------------------------------------
# OpenCV restatements
# ----------------------------------------------------------------------------------------------------------------------
def _linear_taps(dsize: int, ssize: int, clamp: bool):
    scale = 1.0 / (dsize / ssize)
    f = ((np.arange(dsize, dtype=np.float64) + 0.5) * scale - 0.5).astype(F32)
    s = np.floor(f).astype(np.int64)
    a = (f - s.astype(F32)).astype(F32)
    if clamp:
        lo, hi = s < 0, s >= ssize - 1
        a[lo | hi] = 0
        s[lo] = 0
        s[hi] = ssize - 1
    return s, a


def resize_linear(src: np.ndarray, nw: int, nh: int) -> np.ndarray:
    """cv2.resize(src, (nw, nh), interpolation=INTER_LINEAR) for float32 (h, w, c)."""
    sh, sw = src.shape[:2]
    if sw == 2 * nw and sh == 2 * nh:
        s = src.reshape(nh, 2, nw, 2, -1)
        return ((s[:, 0, :, 0] + s[:, 0, :, 1] + s[:, 1, :, 0] + s[:, 1, :, 1]) * F32(0.25)).astype(F32)
    xi, xa = _linear_taps(nw, sw, True)
    x1 = np.minimum(xi + 1, sw - 1)
    a0, a1 = (F32(1) - xa)[None, :, None], xa[None, :, None]
    rows = (src[:, xi] * a0).astype(F32) + (src[:, x1] * a1).astype(F32)
    yi, ya = _linear_taps(nh, sh, False)
    y0, y1 = np.clip(yi, 0, sh - 1), np.clip(yi + 1, 0, sh - 1)
    b0, b1 = (F32(1) - ya)[:, None, None], ya[:, None, None]
    return (rows[y0] * b0).astype(F32) + (rows[y1] * b1).astype(F32)


def resize_nearest(src: np.ndarray, nw: int, nh: int) -> np.ndarray:
    sh, sw = src.shape[:2]
    ix = np.minimum(np.floor(np.arange(nw) * (1.0 / (nw / sw))).astype(np.int64), sw - 1)
    iy = np.minimum(np.floor(np.arange(nh) * (1.0 / (nh / sh))).astype(np.int64), sh - 1)
    return src[iy][:, ix]


def rotation_matrix(w: int, h: int, degree: float) -> np.ndarray:
    """mmcv.imrotate -> cv2.getRotationMatrix2D(((w-1)/2, (h-1)/2), -degree, 1.0), in double."""
    ang = -degree * (np.pi / 180.0)                     # cv2: angle *= CV_PI / 180
    al, be = np.cos(ang), np.sin(ang)
    cx, cy = (w - 1) * 0.5, (h - 1) * 0.5
    return np.array([[al, be, (1 - al) * cx - be * cy], [-be, al, be * cx + (1 - al) * cy]], np.float64)


def invert_affine(M: np.ndarray) -> np.ndarray:
    m = M.astype(np.float64).reshape(-1).copy()
    D = m[0] * m[4] - m[1] * m[3]
    D = 1.0 / D if D != 0 else 0.0
    A11, A22 = m[4] * D, m[0] * D
    m[0] = A11; m[1] *= -D; m[3] *= -D; m[4] = A22
    b1 = -m[0] * m[2] - m[1] * m[5]
    b2 = -m[3] * m[2] - m[4] * m[5]
    m[2] = b1; m[5] = b2
    return m


def warp_affine(src: np.ndarray, M: np.ndarray, nearest: bool, border: float) -> np.ndarray:
    """cv2.warpAffine(src, M, (w, h), flags=INTER_LINEAR | INTER_NEAREST, borderValue=border) for float32."""
    h, w = src.shape[:2]
    s = src.reshape(h, w, -1)
    m = invert_affine(M)
    rd = 512 if nearest else 16
    x, y = np.arange(w), np.arange(h)
    ad, bd = np.rint(m[0] * x * 1024).astype(np.int64), np.rint(m[3] * x * 1024).astype(np.int64)
    X0 = np.rint((m[1] * y + m[2]) * 1024).astype(np.int64) + rd
    Y0 = np.rint((m[4] * y + m[5]) * 1024).astype(np.int64) + rd
    if nearest:
        X, Y = (X0[:, None] + ad[None]) >> 10, (Y0[:, None] + bd[None]) >> 10
        ok = (X >= 0) & (X < w) & (Y >= 0) & (Y < h)
        out = np.full(s.shape, border, F32)
        out[ok] = s[Y[ok], X[ok]]
        return out.reshape(src.shape)
    X, Y = (X0[:, None] + ad[None]) >> 5, (Y0[:, None] + bd[None]) >> 5
    sx, sy = X >> 5, Y >> 5
    fx, fy = (X & 31).astype(F32) / F32(32), (Y & 31).astype(F32) / F32(32)
    w0, w1 = (F32(1) - fy) * (F32(1) - fx), (F32(1) - fy) * fx
    w2, w3 = fy * (F32(1) - fx), fy * fx

    def tap(yy, xx):
        ok = (xx >= 0) & (xx < w) & (yy >= 0) & (yy < h)
        v = s[np.clip(yy, 0, h - 1), np.clip(xx, 0, w - 1)].copy()
        v[~ok] = border
        return v

    t = (tap(sy, sx) * w0[..., None]).astype(F32) + (tap(sy, sx + 1) * w1[..., None]).astype(F32)
    t = t + (tap(sy + 1, sx) * w2[..., None]).astype(F32)
    t = t + (tap(sy + 1, sx + 1) * w3[..., None]).astype(F32)
    return t.astype(F32).reshape(src.shape)


def normalize(img5: np.ndarray, depth_scale=200) -> np.ndarray:
    """Normalize :38-48 + mmcv.imnormalize: RGB channels truncated to uint8, BGR -> RGB, (x - mean) / std; ch3 / depth_scale
    where positive; ch4 raw."""
    rgb = img5[:, :, 0:3].astype(np.uint8).astype(F32)[:, :, ::-1]
    rgb = ((rgb.astype(np.float64) - MEAN.astype(np.float64)).astype(F32).astype(np.float64) * (1 / STD.astype(np.float64))).astype(F32)
    pe = img5[:, :, 3].copy()
    pe[pe > 0] = pe[pe > 0] / depth_scale
    return np.concatenate([rgb, pe[:, :, None], img5[:, :, 4:5]], axis=-1).astype(F32)


# ----------------------------------------------------------------------------------------------------------------------
# the pipeline
# ----------------------------------------------------------------------------------------------------------------------
def kb_crop(a: np.ndarray) -> np.ndarray:
    h, w = a.shape[:2]
    top, left = int(h - KB_H), int((w - KB_W) / 2)
    return a[top:top + KB_H, left:left + KB_W]


def train_augment(img5: np.ndarray, depth_gt: np.ndarray, pe_k_gt: np.ndarray, p: dict):
    """img5 (H, W, 5) float32 as the loader emits it (BGR as float, clamped plane, raw plane); depth_gt, pe_k_gt (H, W)
    float32.  Returns img (5, 352, 704), depth_gt (1, 352, 704), pe_k_gt (352, 704) - what DefaultFormatBundle collates."""
    img, dep, lab = kb_crop(img5), kb_crop(depth_gt), kb_crop(pe_k_gt)
    nw, nh = p["new_w"], p["new_h"]
    img, dep, lab = resize_linear(img, nw, nh), resize_nearest(dep, nw, nh), resize_nearest(lab, nw, nh)
    ch, cw = p["canvas_h"], p["canvas_w"]
    if nh < KB_H or nw < KB_W:
        ci, cd, cl = np.zeros((ch, cw, 5), F32), np.zeros((ch, cw), F32), np.full((ch, cw), 255, F32)
        y0, x0 = p["pad_y"], p["pad_x"]
        ci[y0:y0 + nh, x0:x0 + nw] = img; cd[y0:y0 + nh, x0:x0 + nw] = dep; cl[y0:y0 + nh, x0:x0 + nw] = lab
        img, dep, lab = ci, cd, cl
    if p["rotate"]:
        M = rotation_matrix(cw, ch, p["degree"])
        img = warp_affine(img, M, False, 0.0)
        dep = warp_affine(dep, M, True, 0.0)          # depth_pad_val = 0 (RandomRotate default; the config passes none)
        lab = warp_affine(lab, M, True, 255.0)        # "pe" in key -> 255 (:285)
    if p["flip"]:
        img, dep, lab = img[:, ::-1], dep[:, ::-1], lab[:, ::-1]
    y0, x0 = p["crop_y"], p["crop_x"]
    img, dep, lab = (a[y0:y0 + CROP_H, x0:x0 + CROP_W] for a in (img, dep, lab))
    img = np.ascontiguousarray(img).copy()
    if p["color"]:
        # float32 ** python float -> float32 powf.  numpy's float32 power is platform-dependent (SVML on AVX-512 builds, libm
        # elsewhere: they differ from each other and from the correctly rounded value in ~20 % of the arguments, by one
        # ulp), so the reference itself is not bit-reproducible here; the oracle takes the CORRECTLY ROUNDED power (through
        # float64).  After the uint8 truncation of Normalize the two differ on ~1 pixel value in 10^6, by one grey level.
        a = (img[:, :, 0:3].astype(np.float64) ** np.float64(F32(p["gamma"]))).astype(F32)
        a = a * F32(p["brightness"])
        a = (a.astype(np.float64) * np.array(p["colors"], np.float64)[None, None, :]).astype(F32)   # float32 *= float64 array
        img[:, :, 0:3] = np.clip(a, 0, 255)
    out = normalize(img)
    return (np.ascontiguousarray(out.transpose(2, 0, 1)), np.ascontiguousarray(dep)[None].astype(F32),
            np.ascontiguousarray(lab).astype(F32))


def synth_frame(seed: int, h: int = 375, w: int = 1242):
    """Deterministic KITTI-shaped frame: smooth BGR image with texture, the KITTI ground plane (clamped / raw), sparse depth,
    slope labels in {0..10} (255 where there is no LiDAR return)."""
    from oracle import ground as og
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    base = np.stack([127 + 90 * np.sin(xx / (37.0 + 5 * c) + c) * np.cos(yy / (23.0 + 3 * c)) for c in range(3)], -1)
    bgr = np.clip(base + rng.normal(0, 12, (h, w, 3)), 0, 255).astype(np.uint8)
    coef = og.plane_coefficients(og.kitti_projection(), og.KITTI_CAM_HEIGHT)
    pe = og.ground_plane(coef, h, w).astype(F32)
    pe_c = pe.copy()
    pe_c[pe_c > 200] = 0
    pe_c[pe_c < 0] = 0                                                        # loading.py:398-401
    img5 = np.concatenate([bgr.astype(F32), pe_c[:, :, None], pe[:, :, None]], -1).astype(F32)
    keep = rng.random((h, w)) < 0.06
    depth = np.where(keep, rng.uniform(2, 80, (h, w)), 0).astype(F32)
    lab = np.where(keep, rng.integers(0, 11, (h, w)), 255).astype(F32)
    return img5, depth, lab


def cv2_reference_seconds(img5, depth_gt, pe_k_gt, p, reps: int = 3) -> float:
    """Wall time of the same chain through cv2 / numpy as the reference's CPU workers run it (mmcv.imrescale -> cv2.resize,
    mmcv.imrotate -> cv2.warpAffine, numpy for the rest): the CPU baseline of bench.py's `train_augment` block."""
    import time
    import cv2
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        img, dep, lab = kb_crop(img5), kb_crop(depth_gt), kb_crop(pe_k_gt)
        size = (p["new_w"], p["new_h"])
        img = cv2.resize(img, size, interpolation=cv2.INTER_LINEAR)
        dep = cv2.resize(dep, size, interpolation=cv2.INTER_NEAREST)
        lab = cv2.resize(lab, size, interpolation=cv2.INTER_NEAREST)
        ch, cw = p["canvas_h"], p["canvas_w"]
        if (ch, cw) != img.shape[:2]:
            ci, cd, cl = np.zeros((ch, cw, 5), F32), np.zeros((ch, cw), F32), np.full((ch, cw), 255, F32)
            y0, x0 = p["pad_y"], p["pad_x"]
            ci[y0:y0 + size[1], x0:x0 + size[0]] = img; cd[y0:y0 + size[1], x0:x0 + size[0]] = dep; cl[y0:y0 + size[1], x0:x0 + size[0]] = lab
            img, dep, lab = ci, cd, cl
        if p["rotate"]:
            M = rotation_matrix(cw, ch, p["degree"])
            img = cv2.warpAffine(img, M, (cw, ch), flags=cv2.INTER_LINEAR, borderValue=0)
            dep = cv2.warpAffine(dep, M, (cw, ch), flags=cv2.INTER_NEAREST, borderValue=0)
            lab = cv2.warpAffine(lab, M, (cw, ch), flags=cv2.INTER_NEAREST, borderValue=255)
        if p["flip"]:
            img, dep, lab = np.flip(img, 1), np.flip(dep, 1).copy(), np.flip(lab, 1).copy()
        y0, x0 = p["crop_y"], p["crop_x"]
        img = np.ascontiguousarray(img[y0:y0 + CROP_H, x0:x0 + CROP_W])
        if p["color"]:
            a = img[:, :, 0:3] ** p["gamma"] * p["brightness"]
            a *= np.array(p["colors"])[None, None, :]
            img[:, :, 0:3] = np.clip(a, 0, 255)
        out = normalize(img)
        np.ascontiguousarray(out.transpose(2, 0, 1))
        best = min(best, time.perf_counter() - t0)
    return best
