"""Pretrained-checkpoint interop for the DepthFormerSwin backbone.

What the reference does when ``backbone.pretrained`` is a path
(depth/models/backbones/depthformer_swin.py:1059-1125 + depth/models/utils/ckpt_convert.py:5-56):

1. unwrap ``state_dict`` / ``model``; drop a leading ``module.``;
2. official Swin keys -> this backbone's keys (``layers.i`` -> ``stages.i``, ``attn.`` ->
   ``attn.w_msa.``, ``mlp.fc1/fc2`` -> ``ffn.layers.0.0`` / ``ffn.layers.1``, ``patch_embed.proj``
   -> ``patch_embed.projection``; classifier ``head.*`` dropped);
3. PatchMerging feature order: the official model concatenates the 2x2 neighbours as
   [x(0,0), x(1,0), x(0,1), x(1,1)] blocks of C channels; ``nn.Unfold`` (and ged_merge_patches)
   emits feature = c*4 + ky*2 + kx.  ``downsample.reduction.weight`` columns and
   ``downsample.norm.{weight,bias}`` are permuted accordingly;
4. relative-position-bias tables of another window size are resized bicubically;
5. USEPE: a 4-D weight whose input-channel count is one short of the model's (the RGB
   ``patch_embed.projection.weight`` [C,3,4,4] vs the RGB+PE model's [C,4,4,4]) is zero-padded on the
   last input channel, so the ground-embedding channel starts with no influence;
6. ``load_state_dict(strict=False)``.

Everything here is host-side tensor bookkeeping on CPU tensors (a one-off at start-up).
"""
from __future__ import annotations

import re
from collections import OrderedDict
from typing import Dict, Mapping

import torch
import torch.nn.functional as F

_RENAMES = (
    (re.compile(r"^layers\.(\d+)\."), r"stages.\1."),
    (re.compile(r"\.attn\."), ".attn.w_msa."),
    (re.compile(r"\.mlp\.fc1\."), ".ffn.layers.0.0."),
    (re.compile(r"\.mlp\.fc2\."), ".ffn.layers.1."),
    (re.compile(r"\.mlp\."), ".ffn."),
    (re.compile(r"^patch_embed\.proj\."), "patch_embed.projection."),
)


def _unfold_feature_order(n_features: int) -> torch.Tensor:
    """index[j] = official feature that lands at Unfold position j, j = c*4 + ky*2 + kx.
    Official block order over (ky,kx) is (0,0),(1,0),(0,1),(1,1)  ->  block = kx*2 + ky."""
    C = n_features // 4
    j = torch.arange(n_features)
    c, ky, kx = j // 4, (j % 4) // 2, j % 2
    return (kx * 2 + ky) * C + c


def convert_official_swin(ckpt: Mapping[str, torch.Tensor]) -> "OrderedDict[str, torch.Tensor]":
    """Official (microsoft/Swin-Transformer) backbone state_dict -> DepthFormerSwin keys / layouts."""
    out = OrderedDict()
    for k, v in ckpt.items():
        if k.startswith("head"):
            continue
        nk = k
        if k.startswith("layers"):
            if ".downsample." in k:
                if k.endswith("reduction.weight"):
                    v = v[:, _unfold_feature_order(v.shape[1])]
                elif ".norm." in k:
                    v = v[_unfold_feature_order(v.shape[0])]
            for pat, rep in _RENAMES[:5]:      # after the fc1/fc2 rules no ".mlp." is left for the generic one
                nk = pat.sub(rep, nk)
        elif k.startswith("patch_embed"):
            nk = _RENAMES[5][0].sub(_RENAMES[5][1], k)
        out[nk] = v
    return out


def resize_rel_pos_bias(table: torch.Tensor, target_len: int) -> torch.Tensor:
    """(L1, nH) -> (L2, nH) by bicubic interpolation of the (2w-1)x(2w-1) grid (align_corners=False)."""
    L1, nH = table.shape
    if L1 == target_len:
        return table
    S1, S2 = int(L1 ** 0.5), int(target_len ** 0.5)
    t = F.interpolate(table.permute(1, 0).reshape(1, nH, S1, S1).float(), size=(S2, S2), mode="bicubic",
                      align_corners=False)
    return t.view(nH, target_len).permute(1, 0).contiguous().to(table.dtype)


def adapt_to_model(state: Dict[str, torch.Tensor], model_state: Mapping[str, torch.Tensor], usepe: bool,
                   warn=None) -> Dict[str, torch.Tensor]:
    """Steps 4 and 5 above against the shapes of ``model_state``."""
    state = dict(state)
    for k in [k for k in state if "relative_position_bias_table" in k and k in model_state]:
        have, want = state[k], model_state[k]
        if have.shape[1] != want.shape[1]:
            if warn:
                warn(f"Error in loading {k}, pass")
            continue
        state[k] = resize_rel_pos_bias(have, want.shape[0])
    if usepe:
        for k, v in list(state.items()):
            if k not in model_state or model_state[k].shape == v.shape:
                continue
            want = model_state[k].shape
            if v.dim() == 4 and len(want) == 4 and want[1] == v.shape[1] + 1 and want[0] == v.shape[0] \
                    and want[2:] == v.shape[2:]:
                tgt = torch.zeros(want, dtype=v.dtype)
                tgt[:, 0:tgt.shape[1] - 1, :, :] = v
                state[k] = tgt
    return state


def load_swin_pretrained(backbone, path_or_state, map_location="cpu", logger=None):
    """``DepthFormerSwin.init_weights`` with ``pretrained`` set.  Returns load_state_dict's result."""
    ckpt = torch.load(path_or_state, map_location=map_location, weights_only=False) \
        if isinstance(path_or_state, str) else path_or_state
    state = ckpt.get("state_dict", ckpt.get("model", ckpt)) if isinstance(ckpt, Mapping) else ckpt
    if backbone.pretrain_style == "official":
        state = convert_official_swin(state)
    if state and next(iter(state)).startswith("module."):
        state = OrderedDict((k[7:], v) for k, v in state.items())
    warn = (logger.warning if logger is not None else None)
    state = adapt_to_model(state, backbone.state_dict(), getattr(backbone, "USEPE", False), warn)
    return backbone.load_state_dict(state, strict=False)
