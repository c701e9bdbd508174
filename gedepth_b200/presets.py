"""Model dictionaries of the four GEDepth configs, built programmatically.

The reference's own ``configs/depthformer/*.py`` load unchanged through ``compat.Config.fromfile``
(tests/test_config_registry.py checks that, against /root/reference when present, and that the
result equals these presets); the presets exist because the reference tree is not on the GPU box
and its files are not copied into this repo.  Values: configs/_base_/models/depthformer_swin.py:2-41,
configs/depthformer/depthformer_{v,a}.py:94-126, depthformer_{v,a}_ddad.py:88-119.
"""
from __future__ import annotations

import copy

SWIN_L = dict(embed_dims=192, depths=[2, 2, 18, 2], num_heads=[6, 12, 24, 48])
SWIN_T = dict(embed_dims=96, depths=[2, 2, 6, 2], num_heads=[3, 6, 12, 24])
KITTI_PRETRAINED = "ckpt/swin_large_patch4_window7_224_22k.pth"


def model_cfg(variant: str = "v", dataset: str = "kitti", backbone: str = "swin_l",
              pretrained="config", drop_path_rate: float = 0.3) -> dict:
    """variant 'v' (Vanilla) | 'a' (Adaptive); dataset 'kitti' | 'ddad'; backbone 'swin_l' (what
    the reference ships) | 'swin_t' (BASELINE config 2: Swin-T widths into the L-width neck,
    SURVEY.md §0.4).  ``pretrained='config'`` keeps the config's value."""
    assert variant in ("v", "a") and dataset in ("kitti", "ddad") and backbone in ("swin_l", "swin_t")
    ddad = dataset == "ddad"
    bb = dict(type="DepthFormerSwin", pretrain_img_size=224, patch_size=4, window_size=7, mlp_ratio=4,
              strides=(4, 2, 2, 2), out_indices=(0, 1, 2, 3), qkv_bias=True, qk_scale=None,
              patch_norm=True, drop_rate=0.0, attn_drop_rate=0.0, drop_path_rate=drop_path_rate,
              use_abs_pos_embed=False, act_cfg=dict(type="GELU"),
              norm_cfg=dict(type="LN", requires_grad=True), pretrain_style="official",
              conv_norm_cfg=dict(type="BN", requires_grad=True), depth=50, num_stages=0, USEPE=True)
    bb.update(copy.deepcopy(SWIN_L if backbone == "swin_l" else SWIN_T))
    widths = [64, 192, 384, 768, 1536]
    neck_in = widths if backbone == "swin_l" else [64, 96, 192, 384, 768]
    m = dict(
        type="DepthEncoderDecoder",
        pretrained=(None if ddad else KITTI_PRETRAINED) if pretrained == "config" else pretrained,
        backbone=bb,
        neck=dict(type="HAHIHeteroNeck",
                  positional_encoding=dict(type="SinePositionalEncoding", num_feats=256),
                  in_channels=list(neck_in), out_channels=list(widths), embedding_dim=512,
                  scales=[1, 1, 1, 1, 1]),
        pe_mask_neck=dict(type="LightPEMASKNeck"),
        decode_head=dict(type="DenseDepthHead", act_cfg=dict(type="LeakyReLU", inplace=True),
                         in_channels=list(widths), up_sample_channels=list(widths), channels=64,
                         align_corners=True, min_depth=1e-3, max_depth=200 if ddad else 80,
                         loss_decode=dict(type="SigLoss", valid_mask=True, loss_weight=1.0)),
        train_cfg=dict(), test_cfg=dict(mode="whole"))
    if variant == "a":
        m["dynamic_pe_neck"] = dict(type="DynamicPENeckSOFT")
    if ddad:
        m["depth_scale"] = 250
    return m
