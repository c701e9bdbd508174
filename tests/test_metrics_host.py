"""Host side of the on-device metrics (gedepth_b200/metrics.py): crop rectangles and the sums -> metrics step,
against the numpy restatement of depth/core/evaluation/metrics.py (oracle.ground.depth_metrics).  No GPU."""
import numpy as np
import torch

from gedepth_b200 import metrics as Mx
from oracle import ground as og


def _sums(gt, pred):
    d = gt - pred
    lg, lp = np.log(gt), np.log(pred)
    th = np.maximum(gt / pred, pred / gt)
    return [gt.size, (th < 1.25).sum(), (th < 1.25 ** 2).sum(), (th < 1.25 ** 3).sum(), (np.abs(d) / gt).sum(),
            (d * d / gt).sum(), (d * d).sum(), ((lg - lp) ** 2).sum(), (lp - lg).sum(),
            np.abs(np.log10(gt) - np.log10(pred)).sum()]


def test_finalize_matches_reference_formulas():
    rng = np.random.default_rng(0)
    rows, refs = [], []
    for n in (1000, 17, 0):
        gt = rng.uniform(1, 70, n).astype(np.float32)
        pred = (gt * (1 + 0.1 * rng.standard_normal(n))).clip(1e-3, 80).astype(np.float32)
        rows.append(_sums(gt.astype(np.float64), pred.astype(np.float64)) if n else [0.0] * 10)
        refs.append(og.depth_metrics(gt, pred, 1e-3, 80.0))
    out = Mx.finalize(torch.tensor(rows, dtype=torch.float64)).numpy()
    assert np.isnan(out[2]).all() and np.isnan(np.array(refs[2], dtype=float)).all()
    np.testing.assert_allclose(out[:2], np.array(refs[:2], dtype=np.float64), rtol=2e-5, atol=1e-6)


def test_crop_rectangles_follow_kitti_eval():
    # kitti.py:373-383 on the 352x1216 KB-cropped ground truth
    assert Mx.crop_rect(352, 1216, True) == (143, 349, 43, 1172)
    assert Mx.crop_rect(352, 1216, False, True) == (117, 321, 43, 1172)
    assert Mx.crop_rect(352, 1216, False, False) is None
    assert Mx.kb_crop_window(375, 1242) == (23, 13)
