"""ORACLE tooling (build container only): run the reference's OWN train-time transform classes
(/root/reference/depth/datasets/pipelines/transforms.py, loaded verbatim by path) on seeded synthetic KITTI frames and
write tests/golden/train_aug.npz - SHA-256 of every output array, the drawn parameters, and a strided sample of the
outputs for diagnosis.

mmcv (1.3.13, absent) is replaced by the five image functions the transforms call, restated from
mmcv/image/geometric.py (imresize l. 50-100, rescale_size l. 190-225, imrescale l. 228-257, imflip l. 260-275,
imrotate l. 300-350) and mmcv/image/photometric.py (imnormalize l. 9-43) - thin wrappers over cv2, which IS in the image.

    python -m oracle.make_golden_aug
"""
from __future__ import annotations

import hashlib
import importlib.util
import os
import random
import sys
import types

import cv2
import numpy as np

REF = os.environ.get("GEDEPTH_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
SEEDS = (0, 2, 6, 7, 9, 13, 21, 77, 5)
PRESET_SCALE = {5: (608, 176)}          # exactly half: cv2.resize switches INTER_LINEAR to the 2 x 2 INTER_AREA average
_INTERP = {"nearest": cv2.INTER_NEAREST, "bilinear": cv2.INTER_LINEAR}


def _scale_size(size, scale):
    w, h = size
    return int(w * float(scale) + 0.5), int(h * float(scale) + 0.5)


def imresize(img, size, return_scale=False, interpolation="bilinear", out=None, backend=None):
    h, w = img.shape[:2]
    resized = cv2.resize(img, size, dst=out, interpolation=_INTERP[interpolation])
    if not return_scale:
        return resized
    return resized, size[0] / w, size[1] / h


def rescale_size(old_size, scale, return_scale=False):
    w, h = old_size
    if isinstance(scale, (float, int)):
        scale_factor = scale
    else:
        max_long_edge, max_short_edge = max(scale), min(scale)
        scale_factor = min(max_long_edge / max(h, w), max_short_edge / min(h, w))
    new_size = _scale_size((w, h), scale_factor)
    return (new_size, scale_factor) if return_scale else new_size


def imrescale(img, scale, return_scale=False, interpolation="bilinear", backend=None):
    h, w = img.shape[:2]
    new_size, scale_factor = rescale_size((w, h), scale, return_scale=True)
    rescaled = imresize(img, new_size, interpolation=interpolation)
    return (rescaled, scale_factor) if return_scale else rescaled


def imflip(img, direction="horizontal"):
    assert direction == "horizontal"
    return np.flip(img, axis=1)


def imrotate(img, angle, center=None, scale=1.0, border_value=0, interpolation="bilinear", auto_bound=False):
    assert not auto_bound
    h, w = img.shape[:2]
    if center is None:
        center = ((w - 1) * 0.5, (h - 1) * 0.5)
    matrix = cv2.getRotationMatrix2D(center, -angle, scale)
    return cv2.warpAffine(img, matrix, (w, h), flags=_INTERP[interpolation], borderValue=border_value)


def imnormalize(img, mean, std, to_rgb=True):
    img = img.copy().astype(np.float32)
    mean = np.float64(mean.reshape(1, -1))
    stdinv = 1 / np.float64(std.reshape(1, -1))
    if to_rgb:
        cv2.cvtColor(img, cv2.COLOR_BGR2RGB, img)
    cv2.subtract(img, mean, img)
    cv2.multiply(img, stdinv, img)
    return img


def load_reference_transforms():
    """The reference's transforms.py, executed verbatim under stub parents for its relative / mmcv imports."""
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class _Registry:
        def register_module(self, *a, **k):
            return lambda cls: cls

    def deprecated_api_warning(name_dict, cls_name=None):
        return lambda f: f

    mod("mmcv", imresize=imresize, imrescale=imrescale, imflip=imflip, imrotate=imrotate, imnormalize=imnormalize,
        is_list_of=lambda seq, t: isinstance(seq, list) and all(isinstance(x, t) for x in seq))
    mod("mmcv.utils", deprecated_api_warning=deprecated_api_warning)
    for pkg in ("depth", "depth.datasets", "depth.datasets.pipelines"):
        mod(pkg).__path__ = []
    mod("depth.ops", resize=None)
    mod("depth.datasets.builder", PIPELINES=_Registry())
    path = os.path.join(REF, "depth", "datasets", "pipelines", "transforms.py")
    spec = importlib.util.spec_from_file_location("depth.datasets.pipelines.transforms", path)
    m = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = m
    spec.loader.exec_module(m)
    return m


def reference_pipeline(T):
    """configs/depthformer/depthformer_v.py:13-33, the geometric / photometric transforms between the loaders and the
    format bundle, with the config's arguments."""
    return [T.KBCrop(depth=True, pe_k=True), T.Resize(ratio_range=(0.5, 2.0)),
            T.Padding(img_padding_value=(0, 0, 0), depth_padding_value=255, pe_k=True),
            T.RandomRotate(prob=0.5, degree=2.5), T.RandomFlip(prob=0.5), T.RandomCrop(crop_size=(352, 704)),
            T.ColorAug(prob=0.5, gamma_range=[0.9, 1.1], brightness_range=[0.9, 1.1], color_range=[0.9, 1.1]),
            T.Normalize(mean=[123.675, 116.28, 103.53], std=[58.395, 57.12, 57.375], to_rgb=True)]


def run_reference(T, seed: int):
    from oracle import augment as oa
    img5, depth, lab = oa.synth_frame(seed)
    results = dict(img=img5.copy(), depth_gt=depth.copy(), pe_k_gt=lab.copy(), img_shape=img5.shape, ori_shape=img5.shape,
                   depth_fields=["depth_gt", "pe_k_gt"])           # as DepthLoadAnnotations registers them (loading.py:143-150)
    if seed in PRESET_SCALE:
        results["scale"] = PRESET_SCALE[seed]
    np.random.seed(seed)
    random.seed(seed)
    for t in reference_pipeline(T):
        results = t(results)
    img = np.ascontiguousarray(results["img"].transpose(2, 0, 1)).astype(np.float32)      # DefaultFormatBundle
    dep = np.ascontiguousarray(results["depth_gt"])[None].astype(np.float32)
    lab = np.ascontiguousarray(results["pe_k_gt"]).astype(np.float32)
    return img, dep, lab


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    from oracle import augment as oa
    T = load_reference_transforms()
    out = {"seeds": np.array(SEEDS)}
    for seed in SEEDS:
        img, dep, lab = run_reference(T, seed)
        np.random.seed(seed)
        random.seed(seed)
        p = oa.draw_params(scale=PRESET_SCALE.get(seed))
        o_img, o_dep, o_lab = oa.train_augment(*oa.synth_frame(seed), p)
        same = (np.array_equal(img, o_img, equal_nan=True), np.array_equal(dep, o_dep), np.array_equal(lab, o_lab))
        nbad = int((img != o_img).sum())
        level = float((np.abs(img[:3] - o_img[:3]).max(axis=(1, 2)) * oa.STD).max())     # in grey levels
        print(seed, {k: (round(v, 4) if isinstance(v, float) else v) for k, v in p.items()}, "oracle == reference:", same,
              "mismatches", nbad, "max grey levels", round(level, 3))
        # geometry, the two plane channels, depth and labels are bit-identical; with ColorAug a handful of RGB values sit one
        # grey level apart (numpy's platform-dependent float32 power vs the oracle's correctly rounded one)
        assert same[1] and same[2] and np.array_equal(img[3:], o_img[3:]) and (p["color"] or nbad == 0)
        assert nbad <= 8 and level <= 1.01
        out[f"s{seed}_ref_img_sha"] = sha(img); out[f"s{seed}_ref_mismatch"] = np.array([nbad, img.size])
        out[f"s{seed}_img_sha"] = sha(o_img); out[f"s{seed}_dep_sha"] = sha(dep); out[f"s{seed}_lab_sha"] = sha(lab)
        out[f"s{seed}_img_probe"] = o_img[:, ::16, ::16].copy()
        out[f"s{seed}_dep_probe"] = dep[:, ::8, ::8].copy()
        out[f"s{seed}_lab_probe"] = lab[::8, ::8].copy()
        out[f"s{seed}_params"] = np.array([p["new_w"], p["new_h"], p["pad_x"], p["pad_y"], p["canvas_w"], p["canvas_h"], int(p["rotate"]),
                                           p["degree"], int(p["flip"]), p["crop_x"], p["crop_y"], int(p["color"]), p["gamma"],
                                           p["brightness"], *p["colors"]], np.float64)
    dst = os.path.join(os.path.dirname(HERE), "tests", "golden", "train_aug.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
