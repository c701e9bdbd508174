"""ORACLE tooling: generate tests/golden/*.npz by running the REFERENCE's own modules
(oracle/ref_harness.py, /root/reference, build container only) on deterministic weights/inputs.

    python -m oracle.make_golden            # rewrites tests/golden/

The fixtures are what pins oracle/model.py (tests/test_oracle_golden.py) and, through it, the CUDA
path (tests/test_model_gpu.py compares against the same files on the GPU box, where /root/reference
does not exist).  Case definitions live in CASES so that tests rebuild identical inputs.
"""
from __future__ import annotations

import copy
import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from gedepth_b200.synth import synth_batch, synth_state_dict  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

# Swin-T widths with the L-width neck - BASELINE config 2's architecture.
BACKBONE_T = dict(embed_dims=96, depths=[2, 2, 6, 2], num_heads=[3, 6, 12, 24], drop_path_rate=0.0)
NECK_IN_T = [64, 96, 192, 384, 768]

CASES = {
    # name: (config file, B, H, W, mode, extras)
    # seeds are explicit so that adding a case never changes the inputs of the others
    "vanilla_train": dict(cfg="depthformer_v.py", B=2, H=64, W=160, train=True, seed=1238),
    "adaptive_train": dict(cfg="depthformer_a.py", B=2, H=64, W=160, train=True, seed=1236),
    "vanilla_eval_ragged": dict(cfg="depthformer_v.py", B=1, H=70, W=166, train=False, seed=1237),
    "adaptive_eval": dict(cfg="depthformer_a.py", B=1, H=64, W=160, train=False, seed=1235),
    "adaptive_ddad_train": dict(cfg="depthformer_a_ddad.py", B=2, H=96, W=160, train=True, ddad=True, seed=1234),
    # the reference's own aug_test (encoder_decoder.py:249-274) on the two views of its test pipeline
    "vanilla_eval_tta": dict(cfg="depthformer_v.py", B=1, H=64, W=160, train=False, tta=True, seed=1239),
    # BASELINE shapes (352 x 1120: Swin stage 3 is 11 x 35 -> padded to 14 x 35, 98 560 cross-attention queries).  The
    # full-resolution side outputs are stored every 4th pixel (`sub`) to keep the fixtures small; pred / depth are complete.
    "vanilla_eval_k8": dict(cfg="depthformer_v.py", B=1, H=352, W=1120, train=False, seed=1240, sub=4),
    "adaptive_train_k8": dict(cfg="depthformer_a.py", B=1, H=352, W=1120, train=True, seed=1241, sub=4),
}
FULL_GRADS = ["decode_head.conv_depth.weight", "pe_mask_neck.convfinal.weight",
              "backbone.patch_embed.projection.weight", "neck.level_embed",
              "backbone.stages.0.blocks.1.attn.w_msa.relative_position_bias_table",
              "neck.reference_points.weight", "backbone.bn1.weight"]


def model_cfg_for(case: dict, config_dir: str = None) -> dict:
    """Model dict of a case: the reference's config FILE when ``config_dir`` is given (golden
    generation, build container), else the equal programmatic preset (tests, GPU box)."""
    if config_dir is not None:
        from gedepth_b200.compat import Config
        m = copy.deepcopy(dict(Config.fromfile(os.path.join(config_dir, case["cfg"])).model))
    else:
        from gedepth_b200.presets import model_cfg
        stem = case["cfg"][:-3].split("_")
        m = model_cfg(stem[1], "ddad" if "ddad" in stem else "kitti")
    m["pretrained"] = None
    m["backbone"].update(copy.deepcopy(BACKBONE_T))
    m["neck"]["in_channels"] = list(NECK_IN_T)
    return m


def case_inputs(name: str) -> dict:
    c = CASES[name]
    ddad = c.get("ddad", False)
    b = synth_batch(c["B"], c["H"], c["W"], seed=c["seed"],
                    depth_scale=250.0 if ddad else 200.0, max_depth=200.0 if ddad else 80.0,
                    adaptive="adaptive" in name)
    if ddad:
        b["height"] = np.array([1.56, 1.57, 1.53, 1.53][: c["B"]], dtype=np.float64)
    return b


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def run_case(name: str) -> dict:
    from oracle import ref_harness as rh
    c = CASES[name]
    model = rh.build_reference_model(model_cfg_for(c, os.path.join(rh.REF, "configs", "depthformer")))
    sd = synth_state_dict(model.state_dict(), seed=0)
    model.load_state_dict(sd)
    for mod in model.modules():
        if mod.__class__.__name__ == "MultiScaleDeformableAttention":
            mod.dropout.p = 0.0          # parity runs disable the stochastic pieces
    b = case_inputs(name)
    img = torch.from_numpy(b["img"])
    metas = [dict(img_norm_cfg=dict(mean=[0, 0, 0], std=[1, 1, 1], to_rgb=True),
                  ori_shape=(c["H"], c["W"], 3), flip=False, flip_direction=None)] * c["B"]
    kw = {}
    if "height" in b:
        kw["height"] = torch.from_numpy(b["height"])
    out = {"img_sha": np.array(sha(b["img"])), "state_sha": np.array(
        sha(np.concatenate([sd[k].float().numpy().ravel()[:16] for k in sorted(sd)])))}
    if c["train"]:
        model.train()
        gt = torch.from_numpy(b["depth_gt"])
        if "pe_k_gt" in b:
            kw["pe_k_gt"] = torch.from_numpy(b["pe_k_gt"])
        # forward_train, with the intermediates captured the way extract_feat computes them
        x, y, pe_mask, pe_off = model.extract_feat(img, metas, **kw)
        depth, y_h = model.decode_head.forward(x, metas, pe_mask, y)
        model.zero_grad()
        losses = model.forward_train(img, metas, gt, **kw)
        loss = sum(v for k, v in losses.items() if "loss" in k)
        loss.backward()
        sub = c.get("sub", 1)
        out.update(depth=depth.detach().float().numpy(), y=y.detach().float().numpy()[..., ::sub, ::sub],
                   pe_mask=pe_mask.detach().float().numpy()[..., ::sub, ::sub], loss=np.array(float(loss)),
                   depth_gt_sha=np.array(sha(b["depth_gt"])))
        for k, v in losses.items():
            if "loss" in k:
                out["loss." + k] = np.array(float(v))
        names, norms, sums = [], [], []
        for n, p in model.named_parameters():
            names.append(n)
            g = p.grad if p.grad is not None else torch.zeros_like(p)
            norms.append(float(g.double().norm()))
            sums.append(float(g.double().sum()))
        out.update(grad_names=np.array(names), grad_norms=np.array(norms), grad_sums=np.array(sums))
        for n in FULL_GRADS:
            out["grad." + n] = dict(model.named_parameters())[n].grad.float().numpy()
        for i, f in enumerate(x):
            out[f"neck{i}_stats"] = np.array([float(f.mean()), float(f.std())])
    elif c.get("tta"):
        model.eval()
        metas_f = [dict(m, flip=True, flip_direction="horizontal") for m in metas]
        with torch.no_grad():
            res = model.aug_test([img, img.flip(3)], [metas, metas_f], rescale=True,
                                 pe_ori_point=[torch.zeros(1), torch.zeros(1)])
        out.update(pred=np.stack([np.asarray(r, dtype=np.float32) for r in res]))
    else:
        model.eval()
        with torch.no_grad():
            if "height" in kw:
                kw["height"] = [kw["height"]]
                kw["test"] = True
            x, y, pe_mask, _ = model.extract_feat(img, metas, **kw)
            pred = model.encode_decode(img, metas, True, **kw)
        sub = c.get("sub", 1)
        out.update(pred=pred.float().numpy(), y=y.float().numpy()[..., ::sub, ::sub],
                   pe_mask=pe_mask.float().numpy()[..., ::sub, ::sub])
        for i, f in enumerate(x):
            out[f"neck{i}_stats"] = np.array([float(f.mean()), float(f.std())])
    return out


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    torch.set_num_threads(8)
    only = sys.argv[1:]
    for name in CASES:
        if only and name not in only:
            continue
        out = run_case(name)
        path = os.path.join(GOLDEN_DIR, f"{name}.npz")
        np.savez_compressed(path, **out)
        print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB",
              {k: float(v) for k, v in out.items() if k.startswith("loss")})


if __name__ == "__main__":
    main()
