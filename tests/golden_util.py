"""Shared helpers for the golden-fixture tests (inputs are regenerated, hashes are checked)."""
import hashlib
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

# gradients that are analytically zero (a bias in front of a train-mode BatchNorm): pure round-off
ZERO_GRAD_KEYS = ("backbone.norm0.bias", "backbone.norm1.bias", "backbone.norm2.bias", "backbone.norm3.bias")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def load_case(name):
    from oracle.make_golden import CASES, case_inputs
    g = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    b = case_inputs(name)
    assert sha(b["img"]) == str(g["img_sha"]), "synthetic input generator drifted from the fixture"
    return CASES[name], g, b


def build_host_model(case, device="cpu"):
    """The repo's DepthEncoderDecoder from the reference's config file + deterministic weights."""
    import gedepth_b200.models  # noqa: F401  (registers everything)
    from gedepth_b200 import builder
    from gedepth_b200.synth import synth_state_dict
    from oracle.make_golden import model_cfg_for
    m = builder.build_depther(model_cfg_for(case))
    sd = synth_state_dict(m.state_dict(), 0)
    m.load_state_dict(sd)
    for mod in m.modules():
        if isinstance(getattr(mod, "dropout", None), torch.nn.Dropout):
            mod.dropout.p = 0.0
    return m.to(device), sd


def state_sha(sd):
    return sha(np.concatenate([sd[k].detach().float().cpu().numpy().ravel()[:16] for k in sorted(sd)]))


def metas_for(case):
    return [dict(ori_shape=(case["H"], case["W"], 3), img_shape=(case["H"], case["W"], 3),
                 pad_shape=(case["H"], case["W"], 3), flip=False, flip_direction=None)] * case["B"]
