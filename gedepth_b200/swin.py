"""DepthFormerSwin backbone - host-side mirror of depth/models/backbones/depthformer_swin.py.

Same registry name, constructor keywords and state_dict keys as the reference
(``DepthFormerSwin`` :751-1184, ``WindowMSA`` :125-230, ``ShiftWindowMSA`` :233-393, ``SwinBlock``
:396-472, ``SwinBlockSequence`` :475-551, ``PatchMerging`` :56-122, ``PatchEmbedSwin``
utils/embed.py:201-302); the arithmetic is issued through gedepth_b200.ops (sm_100a kernels behind
the C-ABI).  Differences in *how*, not *what*:

* QKV / proj / FFN linears commute with window partition + cyclic shift, so they run once on the
  image-ordered token matrix; only the 49x49 attention core sees windows, and it gathers its
  tokens by coordinate (pad, roll, partition, reverse, un-roll, crop are index math, not copies).
* LayerNorm, bias, GELU, residual add and DropPath scaling are GEMM prologues/epilogues.
"""
from __future__ import annotations

import warnings
from copy import deepcopy

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .builder import BACKBONES
from .compat import (BaseModule, ModuleList, Sequential, build_activation_layer, build_conv_layer,
                     build_dropout, build_norm_layer, trunc_normal_init)


def to_2tuple(x):
    return tuple(x) if isinstance(x, (tuple, list)) else (x, x)


class FFN(BaseModule):
    """mmcv FFN(num_fcs=2, add_identity=True) [external]: keys ``layers.0.0`` and ``layers.1``."""

    def __init__(self, embed_dims, feedforward_channels, num_fcs=2, act_cfg=dict(type="GELU"),
                 ffn_drop=0.0, dropout_layer=None, add_identity=True, init_cfg=None):
        super().__init__(init_cfg)
        assert num_fcs == 2 and ffn_drop == 0.0
        self.layers = Sequential(
            Sequential(nn.Linear(embed_dims, feedforward_channels), build_activation_layer(act_cfg),
                       nn.Dropout(ffn_drop)),
            nn.Linear(feedforward_channels, embed_dims), nn.Dropout(ffn_drop))
        self.dropout_layer = build_dropout(dropout_layer) if dropout_layer else nn.Identity()
        self.add_identity = add_identity
        self.act = act_cfg["type"].lower()

    def forward(self, x, identity=None):
        fc1, fc2 = self.layers[0][0], self.layers[1]
        h = ops.linear(x, fc1.weight, fc1.bias, act=self.act)
        identity = x if identity is None else identity
        scale = ops.drop_path_scale(self.dropout_layer, x)
        return ops.linear(h, fc2.weight, fc2.bias, residual=identity if self.add_identity else None,
                          row_scale=scale)


class PatchEmbedSwin(BaseModule):
    def __init__(self, in_channels=3, embed_dims=768, conv_type=None, kernel_size=16, stride=16,
                 padding=0, dilation=1, pad_to_patch_size=True, norm_cfg=None, init_cfg=None):
        super().__init__(init_cfg)
        self.embed_dims = embed_dims
        stride = kernel_size if stride is None else stride
        self.pad_to_patch_size = pad_to_patch_size
        self.patch_size = to_2tuple(kernel_size)
        self.projection = build_conv_layer(dict(type=conv_type or "Conv2d"), in_channels, embed_dims,
                                           kernel_size=kernel_size, stride=stride, padding=padding,
                                           dilation=dilation)
        self.norm = build_norm_layer(norm_cfg, embed_dims)[1] if norm_cfg is not None else None

    def forward(self, x):
        """x: (B, Cin, H, W) -> tokens (B, DH*DW, C).  Padding to a patch multiple goes bottom/right
        with zeros (embed.py:286-294)."""
        x, (self.DH, self.DW) = ops.patch_embed(x, self.projection.weight, self.projection.bias,
                                                self.patch_size[0])
        if self.norm is not None:
            x = ops.layer_norm(x, self.norm.weight, self.norm.bias, self.norm.eps)
        return x


class PatchMerging(BaseModule):
    def __init__(self, in_channels, out_channels, stride=2, bias=False, norm_cfg=dict(type="LN"),
                 init_cfg=None):
        super().__init__(init_cfg)
        self.in_channels, self.out_channels, self.stride = in_channels, out_channels, stride
        assert stride == 2
        sample_dim = stride ** 2 * in_channels
        self.norm = build_norm_layer(norm_cfg, sample_dim)[1] if norm_cfg is not None else None
        self.reduction = nn.Linear(sample_dim, out_channels, bias=bias)

    def forward(self, x, hw_shape):
        """2x2 gather in nn.Unfold channel order (c-major, then kh, kw; :86,115), LN(4C), Linear."""
        H, W = hw_shape
        x = ops.merge_patches(x, H, W)                       # (B, H/2*W/2, 4C)
        if self.norm is not None:
            x = ops.layer_norm(x, self.norm.weight, self.norm.bias, self.norm.eps)
        x = ops.linear(x, self.reduction.weight, self.reduction.bias)
        return x, ((H + 1) // 2, (W + 1) // 2)


class WindowMSA(BaseModule):
    def __init__(self, embed_dims, num_heads, window_size, qkv_bias=True, qk_scale=None,
                 attn_drop_rate=0.0, proj_drop_rate=0.0, init_cfg=None):
        super().__init__(init_cfg)
        assert attn_drop_rate == 0.0 and proj_drop_rate == 0.0
        self.embed_dims, self.window_size, self.num_heads = embed_dims, window_size, num_heads
        head_embed_dims = embed_dims // num_heads
        self.scale = qk_scale or head_embed_dims ** -0.5
        Wh, Ww = self.window_size
        self.relative_position_bias_table = nn.Parameter(
            torch.zeros((2 * Wh - 1) * (2 * Ww - 1), num_heads))
        seq1 = torch.arange(0, (2 * Ww - 1) * Wh, 2 * Ww - 1)
        seq2 = torch.arange(0, Ww, 1)
        coords = (seq1[:, None] + seq2[None, :]).reshape(1, -1)
        self.register_buffer("relative_position_index", (coords + coords.T).flip(1).contiguous())
        self.qkv = nn.Linear(embed_dims, embed_dims * 3, bias=qkv_bias)
        self.proj = nn.Linear(embed_dims, embed_dims)

    def init_weights(self):
        trunc_normal_init(self.relative_position_bias_table, std=0.02)


class ShiftWindowMSA(BaseModule):
    def __init__(self, embed_dims, num_heads, window_size, shift_size=0, qkv_bias=True,
                 qk_scale=None, attn_drop_rate=0, proj_drop_rate=0,
                 dropout_layer=dict(type="DropPath", drop_prob=0.0), init_cfg=None):
        super().__init__(init_cfg)
        self.window_size, self.shift_size = window_size, shift_size
        assert 0 <= shift_size < window_size
        self.w_msa = WindowMSA(embed_dims, num_heads, to_2tuple(window_size), qkv_bias, qk_scale,
                               attn_drop_rate, proj_drop_rate)
        self.drop = build_dropout(dropout_layer)

    def forward(self, query, hw_shape, identity=None):
        """query: LN1(x) (B, L, C).  Returns identity + DropPath(attn(query)) when identity is given
        (the residual of SwinBlock.forward :466 rides in the proj GEMM's epilogue)."""
        m = self.w_msa
        qkv = ops.linear(query, m.qkv.weight, m.qkv.bias)                  # (B, L, 3C), image order
        ctx = ops.window_attention(qkv, m.qkv.bias, m.relative_position_bias_table,
                                   m.relative_position_index, hw_shape, m.num_heads,
                                   self.window_size, self.shift_size, m.scale)  # (B, L, C)
        scale = ops.drop_path_scale(self.drop, query)
        return ops.linear(ctx, m.proj.weight, m.proj.bias, residual=identity, row_scale=scale)


class SwinBlock(BaseModule):
    def __init__(self, embed_dims, num_heads, feedforward_channels, window_size=7, shift=False,
                 qkv_bias=True, qk_scale=None, drop_rate=0.0, attn_drop_rate=0.0, drop_path_rate=0.0,
                 act_cfg=dict(type="GELU"), norm_cfg=dict(type="LN"), init_cfg=None):
        super().__init__(init_cfg)
        self.norm1 = build_norm_layer(norm_cfg, embed_dims)[1]
        self.attn = ShiftWindowMSA(embed_dims, num_heads, window_size,
                                   shift_size=window_size // 2 if shift else 0, qkv_bias=qkv_bias,
                                   qk_scale=qk_scale, attn_drop_rate=attn_drop_rate,
                                   proj_drop_rate=drop_rate,
                                   dropout_layer=dict(type="DropPath", drop_prob=drop_path_rate))
        self.norm2 = build_norm_layer(norm_cfg, embed_dims)[1]
        self.ffn = FFN(embed_dims, feedforward_channels, num_fcs=2, ffn_drop=drop_rate,
                       dropout_layer=dict(type="DropPath", drop_prob=drop_path_rate),
                       act_cfg=act_cfg, add_identity=True)

    def forward(self, x, hw_shape):
        h, x = ops.layer_norm_fork(x, self.norm1.weight, self.norm1.bias, self.norm1.eps)
        x = self.attn(h, hw_shape, identity=x)
        h, x = ops.layer_norm_fork(x, self.norm2.weight, self.norm2.bias, self.norm2.eps)
        return self.ffn(h, identity=x)


class SwinBlockSequence(BaseModule):
    def __init__(self, embed_dims, num_heads, feedforward_channels, depth, window_size=7,
                 qkv_bias=True, qk_scale=None, drop_rate=0.0, attn_drop_rate=0.0, drop_path_rate=0.0,
                 downsample=None, act_cfg=dict(type="GELU"), norm_cfg=dict(type="LN"), init_cfg=None):
        super().__init__(init_cfg)
        dpr = drop_path_rate if isinstance(drop_path_rate, list) else [deepcopy(drop_path_rate)] * depth
        self.blocks = ModuleList([
            SwinBlock(embed_dims, num_heads, feedforward_channels, window_size, shift=bool(i % 2),
                      qkv_bias=qkv_bias, qk_scale=qk_scale, drop_rate=drop_rate,
                      attn_drop_rate=attn_drop_rate, drop_path_rate=dpr[i], act_cfg=act_cfg,
                      norm_cfg=norm_cfg) for i in range(depth)])
        self.downsample = downsample

    def forward(self, x, hw_shape):
        for block in self.blocks:
            x = block(x, hw_shape)
        if self.downsample:
            x_down, down_hw = self.downsample(x, hw_shape)
            return x_down, down_hw, x, hw_shape
        return x, hw_shape, x, hw_shape


@BACKBONES.register_module()
class DepthFormerSwin(BaseModule):
    """Conv stem (7x7 s2 conv + BN + ReLU on RGB) || Swin on the 4-channel RGB+PE patch embedding.
    Constructor keywords as depthformer_swin.py:833-866; ``num_stages`` must be 0 (all GE configs)."""

    def __init__(self, pretrain_img_size=224, in_channels=3, embed_dims=96, patch_size=4,
                 window_size=7, mlp_ratio=4, depths=(2, 2, 6, 2), num_heads=(3, 6, 12, 24),
                 strides=(4, 2, 2, 2), out_indices=(0, 1, 2, 3), qkv_bias=True, qk_scale=None,
                 patch_norm=True, drop_rate=0.0, attn_drop_rate=0.0, drop_path_rate=0.1,
                 use_abs_pos_embed=False, act_cfg=dict(type="GELU"), norm_cfg=dict(type="LN"),
                 pretrain_style="official", pretrained=None, init_cfg=None, conv_cfg=None,
                 conv_norm_cfg=None, depth=None, num_stages=None, with_cp=False,
                 conv_strides=(1, 2, 2, 2), conv_dilations=(1, 1, 1, 1), style="pytorch",
                 conv_pretrained=None, USEPE=False, USE_PARAM_PE=False):
        super().__init__(init_cfg)
        if num_stages not in (0, None):
            raise NotImplementedError("the GEDepth configs use the stem only (num_stages=0)")
        if use_abs_pos_embed or USE_PARAM_PE:
            raise NotImplementedError("use_abs_pos_embed / USE_PARAM_PE are off in every GE config")
        if not (isinstance(pretrained, str) or pretrained is None):
            raise TypeError("pretrained must be a str or None")
        assert pretrain_style in ("official", "mmcls")
        assert strides[0] == patch_size, "Use non-overlapping patch embed."
        self.USEPE, self.out_indices, self.pretrained = USEPE, out_indices, pretrained
        self.pretrain_style, self.num_stages = pretrain_style, 0
        self.patch_embed = PatchEmbedSwin(in_channels=4 if USEPE else in_channels,
                                          embed_dims=embed_dims, conv_type="Conv2d",
                                          kernel_size=patch_size, stride=strides[0],
                                          pad_to_patch_size=True,
                                          norm_cfg=norm_cfg if patch_norm else None)
        total_depth = sum(depths)
        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, total_depth, device="cpu")]
        self.stages = ModuleList()
        c = embed_dims
        for i in range(len(depths)):
            down = None
            if i < len(depths) - 1:
                down = PatchMerging(c, 2 * c, stride=strides[i + 1],
                                    norm_cfg=norm_cfg if patch_norm else None)
            self.stages.append(SwinBlockSequence(c, num_heads[i], mlp_ratio * c, depths[i],
                                                 window_size, qkv_bias, qk_scale, drop_rate,
                                                 attn_drop_rate, dpr[:depths[i]], down, act_cfg,
                                                 norm_cfg))
            dpr = dpr[depths[i]:]
            if down:
                c = down.out_channels
        self.num_features = [int(embed_dims * 2 ** i) for i in range(len(depths))]
        for i in out_indices:
            self.add_module(f"norm{i}", build_norm_layer(norm_cfg, self.num_features[i])[1])
        # conv stem (:1031-1043): conv1 7x7 s2 p3 (no bias) + bn1 + ReLU
        self.conv1 = build_conv_layer(conv_cfg, 3, 64, kernel_size=7, stride=2, padding=3, bias=False)
        name, bn = build_norm_layer(conv_norm_cfg or dict(type="BN"), 64, postfix=1)
        self._stem_norm_name = name
        self.add_module(name, bn)

    def init_weights(self):
        if self.pretrained is None:
            super().init_weights()
            for m in self.modules():
                if isinstance(m, nn.Linear):
                    trunc_normal_init(m.weight, std=0.02)
                    if m.bias is not None:
                        nn.init.constant_(m.bias, 0)
                elif isinstance(m, nn.LayerNorm):
                    nn.init.constant_(m.bias, 0)
                    nn.init.constant_(m.weight, 1.0)
        else:
            from .checkpoint import load_swin_pretrained
            load_swin_pretrained(self, self.pretrained)

    def forward(self, x_ori):
        """x_ori: (B, 5, H, W) -> [stem 64@H/2, C@H/4, 2C@H/8, 4C@H/16, 8C@H/32] (logical NCHW,
        channels_last memory)."""
        bn = getattr(self, self._stem_norm_name)
        outs = [ops.conv_bn_act(x_ori[:, 0:3] if self.USEPE else x_ori, self.conv1.weight, None, bn,
                                stride=2, padding=3, act="relu")]
        x = self.patch_embed(x_ori[:, 0:4] if self.USEPE else x_ori)
        hw_shape = (self.patch_embed.DH, self.patch_embed.DW)
        for i, stage in enumerate(self.stages):
            x, hw_shape, out, out_hw = stage(x, hw_shape)
            if i in self.out_indices:
                n = getattr(self, f"norm{i}")
                out = ops.layer_norm(out, n.weight, n.bias, n.eps)
                outs.append(ops.tokens_to_map(out, out_hw))
        return outs
