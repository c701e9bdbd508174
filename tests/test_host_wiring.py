"""The host-side mirror (registry modules in gedepth_b200/) against the REFERENCE's goldens, on CPU,
with the library statements of each op injected (conftest.host_ops_on_cpu): checks the wiring,
state_dict compatibility and the train_step contract without a GPU."""
import numpy as np
import pytest
import torch

from tests.golden_util import ZERO_GRAD_KEYS, build_host_model, load_case, metas_for, state_sha


@pytest.mark.parametrize("name", ["vanilla_train", "adaptive_ddad_train"])
def test_host_train_step_matches_reference(name, host_ops_on_cpu):
    case, g, b = load_case(name)
    model, sd = build_host_model(case)
    assert state_sha(sd) == str(g["state_sha"])
    model.train()
    data = dict(img=torch.from_numpy(b["img"]), img_metas=metas_for(case),
                depth_gt=torch.from_numpy(b["depth_gt"]))
    if "pe_k_gt" in b:
        data["pe_k_gt"] = torch.from_numpy(b["pe_k_gt"])
    if "height" in b:
        data["height"] = torch.from_numpy(b["height"]).float()   # fp32 by design
    out = model.train_step(data, None)
    assert set(out) == {"loss", "log_vars", "num_samples", "log_imgs"} and out["num_samples"] == case["B"]
    assert abs(out["log_vars"]["loss"] - float(g["loss"])) < 5e-6 * max(1, float(g["loss"]))
    for k in g.files:
        if k.startswith("loss.decode"):
            assert abs(out["log_vars"][k[5:]] - float(g[k])) < 5e-6 * max(1, abs(float(g[k])))
    out["loss"].backward()
    gn = dict(zip(g["grad_names"].tolist(), g["grad_norms"].tolist()))
    top = max(gn.values())
    for n, p in model.named_parameters():
        if n in ZERO_GRAD_KEYS:
            continue
        assert abs(float(p.grad.double().norm()) - gn[n]) <= 5e-5 * gn[n] + 1e-6 * top, n


@pytest.mark.parametrize("name", ["vanilla_eval_ragged", "adaptive_eval"])
def test_host_inference_matches_reference(name, host_ops_on_cpu):
    case, g, b = load_case(name)
    model, _ = build_host_model(case)
    model.eval()
    with torch.no_grad():
        res = model(img=[torch.from_numpy(b["img"])], img_metas=[metas_for(case)], return_loss=False)
    assert isinstance(res, list) and res[0].shape == (1, case["H"], case["W"])
    np.testing.assert_allclose(res[0], g["pred"][0], rtol=3e-5, atol=3e-5)


def test_host_flip_tta_matches_reference_aug_test(host_ops_on_cpu):
    """The reference's own aug_test on (image, mirrored image) (encoder_decoder.py:249-274) vs the host mirror."""
    case, g, b = load_case("vanilla_eval_tta")
    model, _ = build_host_model(case)
    model.eval()
    img = torch.from_numpy(b["img"])
    metas0 = metas_for(case)
    metas1 = [dict(m, flip=True, flip_direction="horizontal") for m in metas0]
    with torch.no_grad():
        res = model(img=[img, img.flip(3)], img_metas=[metas0, metas1], return_loss=False,
                    pe_ori_point=[torch.zeros(1), torch.zeros(1)])
    assert isinstance(res, list) and res[0].shape == (1, case["H"], case["W"])
    np.testing.assert_allclose(res[0], g["pred"][0], rtol=3e-5, atol=3e-5)


def test_product_refuses_cpu_tensors():
    case, _, b = load_case("vanilla_eval_ragged")
    model, _ = build_host_model(case)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model.encode_decode(torch.from_numpy(b["img"]), metas_for(case))
