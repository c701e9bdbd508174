"""Train-time augmentation on the device: the CUDA kernels (csrc/augment.cu through the C ABI) against the
numpy oracle (oracle/augment.py) bit for bit, and against the fixture hashes (tests/golden/train_aug.npz: the oracle's output,
which equals what the reference's own transform classes produced except for <= 1 RGB value per case when ColorAug is drawn -
see tests/test_augment_host.py)."""
import hashlib
import os
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLD = os.path.join(os.path.dirname(__file__), "golden", "train_aug.npz")
PRESET = {5: (608, 176)}


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _device_frame(aug, img5, depth, lab):
    bgr = torch.from_numpy(np.ascontiguousarray(img5[:, :, 0:3].astype(np.uint8))).to(DEV)
    frame = aug.frame_planes(bgr, torch.from_numpy(np.ascontiguousarray(img5[:, :, 3])).to(DEV),
                             torch.from_numpy(np.ascontiguousarray(img5[:, :, 4])).to(DEV))
    return frame, torch.from_numpy(depth).to(DEV), torch.from_numpy(lab).to(DEV)


@pytest.mark.parametrize("seed", [0, 2, 6, 7, 9, 13, 21, 77, 5])
def test_device_augmentation_is_bit_identical_to_the_reference(seed):
    from gedepth_b200 import augment as ga
    from oracle import augment as oa
    g = np.load(GOLD)
    aug = ga.TrainAugmenter(DEV)
    img5, depth, lab = oa.synth_frame(seed)
    np.random.seed(seed)
    random.seed(seed)
    p = ga.draw_params(scale=PRESET.get(seed))
    f, d, l = _device_frame(aug, img5, depth, lab)
    img_g, dep_g, lab_g = aug([f], [d], [l], [p])
    img_g, dep_g, lab_g = img_g[0].cpu().numpy(), dep_g[0].cpu().numpy(), lab_g[0].cpu().numpy()
    o_img, o_dep, o_lab = oa.train_augment(img5, depth, lab, p)
    assert int((img_g != o_img).sum()) == 0, (int((img_g != o_img).sum()), float(np.abs(img_g - o_img).max()), p)
    assert np.array_equal(dep_g, o_dep) and np.array_equal(lab_g, o_lab)
    assert _sha(img_g) == str(g[f"s{seed}_img_sha"]) and _sha(dep_g) == str(g[f"s{seed}_dep_sha"]) and _sha(lab_g) == str(g[f"s{seed}_lab_sha"])


def test_device_augmentation_batch_of_random_draws():
    """A batch of frames with freshly drawn parameters (padding, rotation, flips, colour in every combination)."""
    from gedepth_b200 import augment as ga
    from oracle import augment as oa
    aug = ga.TrainAugmenter(DEV)
    np.random.seed(1234)
    random.seed(1234)
    frames, deps, labs, params, ref = [], [], [], [], []
    for i in range(6):
        img5, depth, lab = oa.synth_frame(100 + i)
        p = ga.draw_params()
        f, d, l = _device_frame(aug, img5, depth, lab)
        frames.append(f); deps.append(d); labs.append(l); params.append(p)
        ref.append(oa.train_augment(img5, depth, lab, p))
    img_g, dep_g, lab_g = aug(frames, deps, labs, params)
    for i, (o_img, o_dep, o_lab) in enumerate(ref):
        assert np.array_equal(img_g[i].cpu().numpy(), o_img), (i, params[i])
        assert np.array_equal(dep_g[i].cpu().numpy(), o_dep) and np.array_equal(lab_g[i].cpu().numpy(), o_lab)
