"""The oracle (oracle/model.py, oracle/ground.py) against the fixtures produced by the REFERENCE's
own modules (oracle/make_golden.py via oracle/ref_harness.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import ground as og
from oracle import model as om
from tests.golden_util import ZERO_GRAD_KEYS, build_host_model, load_case, state_sha

# *_k8: BASELINE shape 352 x 1120 (Swin stage 3 = 11 x 35 padded to 14 x 35, 98 560 cross-attention queries)
TRAIN = ["vanilla_train", "adaptive_train", "adaptive_ddad_train", "adaptive_train_k8"]
EVAL = ["vanilla_eval_ragged", "adaptive_eval", "vanilla_eval_k8"]


def _path_cfg(name, train):
    ddad = "ddad" in name
    return om.PathConfig(adaptive="adaptive" in name, train_bn=train,
                         depth_scale=250.0 if ddad else 200.0, max_depth=200.0 if ddad else 80.0)


@pytest.fixture(scope="module")
def weights():
    cache = {}

    def get(name):
        key = "a" if "adaptive" in name else "v"
        if key not in cache:
            case, _, _ = load_case(name)
            _, sd = build_host_model(case)
            cache[key] = sd
        return cache[key]
    return get


@pytest.mark.parametrize("name", TRAIN)
def test_oracle_train_matches_reference(name, weights):
    case, g, b = load_case(name)
    buffers = ("running_mean", "running_var", "num_batches_tracked", "relative_position_index")
    sd = {k: v.clone().requires_grad_(not k.endswith(buffers)) for k, v in weights(name).items()}
    assert state_sha(sd) == str(g["state_sha"])       # same state_dict keys/shapes/bytes as the reference
    kw = {}
    if "height" in b:
        kw["height"] = torch.from_numpy(b["height"])
    r = om.forward_train(sd, _path_cfg(name, True), torch.from_numpy(b["img"]),
                         torch.from_numpy(b["depth_gt"]),
                         torch.from_numpy(b["pe_k_gt"]) if "pe_k_gt" in b else None, **kw)
    np.testing.assert_allclose(r["depth"].detach().float().numpy(), g["depth"], rtol=2e-5, atol=2e-5)
    sub = case.get("sub", 1)          # the full-resolution side outputs of the k8 fixtures are stored every `sub`-th pixel
    np.testing.assert_allclose(r["y"].detach().float().numpy()[..., ::sub, ::sub], g["y"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(r["pe_mask"].detach().float().numpy()[..., ::sub, ::sub], g["pe_mask"], rtol=2e-5, atol=2e-5)
    assert abs(float(r["loss"]) - float(g["loss"])) < 2e-6 * max(1.0, abs(float(g["loss"])))
    r["loss"].backward()
    gn = dict(zip(g["grad_names"].tolist(), g["grad_norms"].tolist()))
    top = max(gn.values())
    for n, ref in gn.items():
        if n in ZERO_GRAD_KEYS:
            continue
        got = float(sd[n].grad.double().norm())
        assert abs(got - ref) <= 2e-5 * ref + 1e-7 * top, (n, got, ref)
    for k in g.files:
        if k.startswith("grad.") :
            np.testing.assert_allclose(sd[k[5:]].grad.float().numpy(), g[k], rtol=1e-4,
                                       atol=2e-6 * float(np.abs(g[k]).max()) + 1e-9)


@pytest.mark.parametrize("name", EVAL)
def test_oracle_eval_matches_reference(name, weights):
    case, g, b = load_case(name)
    with torch.no_grad():
        pred = om.forward_test(weights(name), _path_cfg(name, False), torch.from_numpy(b["img"]))
    np.testing.assert_allclose(pred.numpy(), g["pred"], rtol=2e-5, atol=2e-5)


def test_relative_position_index_equals_reference_buffer(weights):
    sd = weights("vanilla_train")
    ref = sd["backbone.stages.0.blocks.0.attn.w_msa.relative_position_index"]
    assert torch.equal(om.relative_position_index(7), ref)
    assert int(ref.sum()) == 201684 and ref[0, :3].tolist() == [84, 83, 82]


def test_ground_plane_known_answers():
    """KATs recorded in SURVEY.md §8(a) a1 for the public KITTI 2011_09_26 calibration."""
    coef = og.plane_coefficients(og.kitti_projection(), og.KITTI_CAM_HEIGHT)
    num, cu, cv, c1 = coef
    assert abs(cu - (-1.464e-5)) < 2e-8 and abs(cv - (-1.386e-3)) < 2e-6 and abs(c1 - 0.2589) < 2e-4
    assert abs(num - (-1.578)) < 2e-3
    pe = og.ground_plane(coef, 375, 1242)
    assert pe.dtype == np.float64 and pe.shape == (375, 1242)
    assert abs(pe[-1, -1] - 5.69) < 0.01
    frac = np.mean((pe > 0) & (pe <= 200))
    assert 0.45 < frac < 0.55
    u, v = og.pixel_grid(375, 1242)
    assert u.dtype == np.int64 and v.dtype == np.int64 and u[3, 7] == 7 and v[3, 7] == 3
    # crop window == slice of the full grid, bit-exact
    sub = og.ground_plane(coef, 352, 1120, u0=61, v0=23)
    assert np.array_equal(sub, pe[23:375, 61:61 + 1120])


def test_oracle_equals_the_reference_preprocessing_script():
    """oracle/ground.py against what tools/preprocess_data_kitti.py itself wrote (executed verbatim by
    oracle/run_ref_preprocess.py on a synthetic tree): pe_165.npy in float64 and the slope labels, bit for bit."""
    import hashlib
    import os
    from oracle.run_ref_preprocess import H, W, synth_gt
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_preprocess_kitti.npz"))
    pe = og.ground_plane(og.plane_coefficients(og.kitti_projection(), og.KITTI_CAM_HEIGHT), H, W)
    assert hashlib.sha256(np.ascontiguousarray(pe).tobytes()).hexdigest() == str(g["pe_sha256"])
    assert np.array_equal(pe[::25], g["pe_rows"])
    gt16 = synth_gt()
    assert hashlib.sha256(gt16.tobytes()).hexdigest() == str(g["gt_sha256"])
    k = og.find_k_kitti(gt16.astype(np.float64) / 256, pe.astype(np.float32))
    assert np.array_equal(k, g["k_img"]) and str(g["k_dtype"]) == str(k.dtype)


def test_loader_channels_and_find_k():
    coef = og.plane_coefficients(og.kitti_projection(), og.KITTI_CAM_HEIGHT)
    pe = og.ground_plane(coef, 375, 1242)
    ch3, ch4 = og.load_channels(pe, 200.0)
    assert ch3.dtype == np.float32 and ch3.min() == 0 and ch3.max() <= 200
    assert np.array_equal(ch4, pe.astype(np.float32))
    n3 = og.normalize_pe(ch3, 200.0)
    assert n3.max() <= 1.0 and np.array_equal(n3 == 0, ch3 == 0)
    rng = np.random.default_rng(0)
    gt = np.where(rng.random(pe.shape) < 0.1, np.abs(pe) * (1 + 0.1 * rng.standard_normal(pe.shape)), 0.0)
    k = og.find_k_kitti(gt, pe)
    assert set(np.unique(k)).issubset(set(range(-5, 6)) | {255}) and np.all(k[gt == 0] == 255)
    # flat ground -> slope 0
    k0 = og.find_k_kitti(np.where(pe > 0, pe, 0.0), pe)
    assert np.all(k0[pe > 0] == 0)
    kd = og.find_k_ddad(gt, pe, 1.56)
    assert kd.dtype == np.int64 and np.all(kd[gt == 0] == 255)


def test_abs_rel_metric():
    gt = np.array([10.0, 20.0, 0.0, 100.0])
    pred = np.array([11.0, 18.0, 5.0, 50.0])
    assert abs(og.abs_rel(gt, pred) - np.mean([0.1, 0.1])) < 1e-12     # 0 and 100 masked by (1e-3, 80)
    m = og.depth_metrics(gt, pred)
    assert abs(m[3] - 0.1) < 1e-12 and len(m) == 9
