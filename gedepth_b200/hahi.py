"""HAHIHeteroNeck + MultiScaleDeformableAttention + SinePositionalEncoding - host-side mirror of
depth/models/necks/hahi.py:84-356, depth/utils/position_encoding.py:10-99 and mmcv 1.3.x
``mmcv.ops.MultiScaleDeformableAttention`` [external; restated from its published semantics,
SURVEY.md §8(c)].  Same registry names, constructor keywords and state_dict keys.

B200-first differences: feature maps are channels-last, so ``flatten(2).transpose(1, 2)`` is a
view; the sine encodings and self-attention reference points are constants cached per shape; a
deformable-attention module is ONE autograd node (ops.msda_module): ``query + pos`` (+ the level
embedding), the three input GEMMs, the sampling (softmax over the 32 (level, point) weights,
``ref + offset / (W_l, H_l)``, bilinear gather) and ``output_proj`` + dropout + identity, with the
query gradient's fan-in summed inside the backward GEMMs' epilogues.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import ops
from .builder import NECKS
from .compat import (POSITIONAL_ENCODING, BaseModule, ConvModule, build_positional_encoding,
                     xavier_init)


@POSITIONAL_ENCODING.register_module()
class SinePositionalEncoding(BaseModule):
    """DETR sine encoding from a (bs, h, w) mask (position_encoding.py:54-89)."""

    def __init__(self, num_feats, temperature=10000, normalize=False, scale=2 * math.pi, eps=1e-6,
                 offset=0.0, init_cfg=None):
        super().__init__(init_cfg)
        self.num_feats, self.temperature, self.normalize = num_feats, temperature, normalize
        self.scale, self.eps, self.offset = scale, eps, offset
        self._cache = {}

    def forward(self, mask):
        mask = mask.to(torch.int)
        not_mask = 1 - mask
        y_embed = not_mask.cumsum(1, dtype=torch.float32)
        x_embed = not_mask.cumsum(2, dtype=torch.float32)
        if self.normalize:
            y_embed = (y_embed + self.offset) / (y_embed[:, -1:, :] + self.eps) * self.scale
            x_embed = (x_embed + self.offset) / (x_embed[:, :, -1:] + self.eps) * self.scale
        dim_t = torch.arange(self.num_feats, dtype=torch.float32, device=mask.device)
        dim_t = self.temperature ** (2 * (dim_t // 2) / self.num_feats)
        pos_x = x_embed[:, :, :, None] / dim_t
        pos_y = y_embed[:, :, :, None] / dim_t
        B, H, W = mask.size()
        pos_x = torch.stack((pos_x[..., 0::2].sin(), pos_x[..., 1::2].cos()), dim=4).view(B, H, W, -1)
        pos_y = torch.stack((pos_y[..., 0::2].sin(), pos_y[..., 1::2].cos()), dim=4).view(B, H, W, -1)
        return torch.cat((pos_y, pos_x), dim=3).permute(0, 3, 1, 2)

    def tokens(self, h: int, w: int, device) -> torch.Tensor:
        """(1, h*w, 2*num_feats) encoding of an all-valid mask, cached per (h, w, device)."""
        key = (h, w, str(device))
        if key not in self._cache:
            m = torch.zeros(1, h, w, dtype=torch.bool, device=device)
            self._cache[key] = self.forward(m).flatten(2).transpose(1, 2).contiguous()
        return self._cache[key]


class MultiScaleDeformableAttention(BaseModule):
    """mmcv.ops.MultiScaleDeformableAttention(batch_first=True) [external]."""

    def __init__(self, embed_dims=256, num_heads=8, num_levels=4, num_points=4, im2col_step=64,
                 dropout=0.1, batch_first=False, norm_cfg=None, init_cfg=None):
        super().__init__(init_cfg)
        if embed_dims % num_heads != 0:
            raise ValueError(f"embed_dims must be divisible by num_heads, but got {embed_dims} and {num_heads}")
        assert batch_first, "the GEDepth path constructs it with batch_first=True (hahi.py:179-188)"
        self.embed_dims, self.num_heads = embed_dims, num_heads
        self.num_levels, self.num_points = num_levels, num_points
        self.dropout = nn.Dropout(dropout)
        self.sampling_offsets = nn.Linear(embed_dims, num_heads * num_levels * num_points * 2)
        self.attention_weights = nn.Linear(embed_dims, num_heads * num_levels * num_points)
        self.value_proj = nn.Linear(embed_dims, embed_dims)
        self.output_proj = nn.Linear(embed_dims, embed_dims)
        self.init_weights()

    def init_weights(self):
        nn.init.constant_(self.sampling_offsets.weight, 0.0)
        thetas = torch.arange(self.num_heads, dtype=torch.float32) * (2.0 * math.pi / self.num_heads)
        grid = torch.stack([thetas.cos(), thetas.sin()], -1)
        grid = (grid / grid.abs().max(-1, keepdim=True)[0]).view(self.num_heads, 1, 1, 2).repeat(
            1, self.num_levels, self.num_points, 1)
        for i in range(self.num_points):
            grid[:, :, i, :] *= i + 1
        with torch.no_grad():
            self.sampling_offsets.bias.copy_(grid.view(-1))
        nn.init.constant_(self.attention_weights.weight, 0.0)
        nn.init.constant_(self.attention_weights.bias, 0.0)
        xavier_init(self.value_proj, distribution="uniform", bias=0.0)
        xavier_init(self.output_proj, distribution="uniform", bias=0.0)
        self._is_init = True

    def forward(self, query, key=None, value=None, identity=None, query_pos=None,
                key_padding_mask=None, reference_points=None, spatial_shapes=None,
                level_start_index=None, level_embed=None, **kwargs):
        """query (B, Q, E); value (B, S, E) or None (= query before pos); reference_points
        (B|1, Q, 2) shared by all levels (valid_ratios == 1 on this path); spatial_shapes: list of
        (h, w).  query_pos: the CONSTANT part of the positional encoding, (1, Q, E); a learnable
        per-level embedding (hahi.py:252-270 adds ``level_embed[i]`` to the sine encoding) is passed
        as level_embed (4, E) + level_start_index (5 token offsets) so its gradient comes out of the
        same node.  Returns dropout(output_proj(sampled)) + identity (identity = query)."""
        assert key_padding_mask is None and identity is None
        if query_pos is None:
            query_pos = torch.zeros(1, query.shape[1], query.shape[2], dtype=query.dtype, device=query.device)
        p = self.dropout.p if self.training else 0.0
        return ops.msda_module(query, value, query_pos, level_embed, level_start_index, reference_points,
                               spatial_shapes, self, p)


@NECKS.register_module()
class HAHIHeteroNeck(BaseModule):
    def __init__(self, in_channels, out_channels, embedding_dim, scales=[1, 1, 1, 1],
                 norm_cfg=dict(type="BN", requires_grad=True), act_cfg=dict(type="ReLU", inplace=True),
                 cross_att=True, self_att=True, constrain=False, positional_encoding=None,
                 num_points=8):
        super().__init__()
        assert isinstance(in_channels, list)
        assert all(s == 1 for s in scales), "every GE config uses scales=[1]*5"
        self.cross_att, self.self_att = cross_att, self_att
        self.in_channels, self.out_channels = in_channels, out_channels
        self.scales, self.num_outs = scales, len(scales)
        self.embedding_dim, self.constrain = embedding_dim, constrain
        kw = dict(norm_cfg=norm_cfg, act_cfg=act_cfg)
        self.lateral_convs = nn.ModuleList(
            ConvModule(i, o, kernel_size=1, **kw) for i, o in zip(in_channels, out_channels))
        self.trans_proj = nn.ModuleList(
            ConvModule(o, embedding_dim, kernel_size=1, **kw) for o in out_channels[1:])
        self.trans_fusion = nn.ModuleList(
            ConvModule(o + embedding_dim, o, kernel_size=3, padding=1, stride=1, **kw)
            for o in out_channels[1:])
        self.conv_proj = nn.Sequential(ConvModule(in_channels[0], embedding_dim, kernel_size=1, **kw))
        self.conv_fusion = nn.Sequential(
            ConvModule(in_channels[0] + embedding_dim, out_channels[0], kernel_size=3, padding=1,
                       stride=1, **kw))
        self.trans_positional_encoding = build_positional_encoding(positional_encoding)
        self.conv_positional_encoding = build_positional_encoding(positional_encoding)
        self.reference_points = nn.Linear(embedding_dim, 2)
        self.level_embed = nn.Parameter(torch.Tensor(4, embedding_dim))
        nn.init.normal_(self.level_embed)
        self.multi_att = MultiScaleDeformableAttention(embedding_dim, num_levels=4, num_heads=8,
                                                       num_points=num_points, batch_first=True)
        self.self_attn = MultiScaleDeformableAttention(embedding_dim, num_levels=4, num_heads=8,
                                                       num_points=num_points, batch_first=True)
        self._ref_cache = {}

    def init_weights(self):
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)
        nn.init.xavier_uniform_(self.reference_points.weight.data, gain=1.0)
        nn.init.constant_(self.reference_points.bias.data, 0.0)
        nn.init.normal_(self.level_embed)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                xavier_init(m, distribution="uniform")
            if isinstance(m, MultiScaleDeformableAttention):
                m.init_weights()

    def _self_ref(self, shapes, device):
        """Pixel centres ((x+.5)/W_l, (y+.5)/H_l) concatenated over levels (hahi.py:221-233)."""
        key = (tuple(shapes), str(device))
        if key not in self._ref_cache:
            refs = []
            for h, w in shapes:
                ry, rx = torch.meshgrid(torch.linspace(0.5, h - 0.5, h, device=device),
                                        torch.linspace(0.5, w - 0.5, w, device=device), indexing="ij")
                refs.append(torch.stack((rx.reshape(-1) / w, ry.reshape(-1) / h), -1))
            self._ref_cache[key] = torch.cat(refs, 0)[None].contiguous()
        return self._ref_cache[key]

    def _self_pos(self, shapes, device):
        """Sine encodings of the four levels, concatenated: the constant part of the self-attention query_pos."""
        key = ("pos", tuple(shapes), str(device))
        if key not in self._ref_cache:
            self._ref_cache[key] = torch.cat([self.trans_positional_encoding.tokens(h, w, device) for h, w in shapes], 1).contiguous()
        return self._ref_cache[key]

    @staticmethod
    def _cm(m: ConvModule, x, padding=0):
        return ops.conv_bn_act(x, m.conv.weight, m.conv.bias, m.norm, stride=1, padding=padding,
                               act="relu")

    def forward(self, inputs):
        assert len(inputs) == len(self.in_channels)
        lat = [self._cm(c, inputs[i]) for i, c in enumerate(self.lateral_convs)]
        feat_conv, feats_trans = lat[0], lat[1:]
        dev = feat_conv.device
        shapes = [(int(t.shape[2]), int(t.shape[3])) for t in feats_trans]

        # HI: deformable self-attention over the four Swin levels
        src = ops.cat_tokens([ops.map_to_tokens(self._cm(self.trans_proj[i], t)) for i, t in enumerate(feats_trans)])
        if self.self_att:
            starts = [0]
            for hh, ww in shapes:
                starts.append(starts[-1] + hh * ww)
            src = self.self_attn(src, value=None, identity=None, query_pos=self._self_pos(shapes, dev),
                                 reference_points=self._self_ref(shapes, dev), spatial_shapes=shapes,
                                 level_start_index=starts, level_embed=self.level_embed)

        # HA: deformable cross-attention from the stem grid into the Swin levels
        skip = self._cm(self.conv_proj[0], feat_conv)
        bs, c, h, w = skip.shape
        query = ops.map_to_tokens(skip)
        qpe = self.conv_positional_encoding.tokens(h, w, dev)
        ref = ops.linear_small(qpe, self.reference_points.weight, self.reference_points.bias, act="sigmoid")
        if self.cross_att:
            fused = self.multi_att(query, value=src, identity=None, query_pos=qpe,
                                   reference_points=ref, spatial_shapes=shapes)
        else:
            fused = query
        fused = ops.tokens_to_map(fused, (h, w))
        cf = self.conv_fusion[0]
        outs = [ops.conv_bn_act_cat(fused, feat_conv, cf.conv.weight, cf.conv.bias, cf.norm, act="relu")]
        levels = ops.split_levels(src, [hh * ww for hh, ww in shapes])
        for i, t in enumerate(feats_trans):
            f = ops.tokens_to_map(levels[i], shapes[i])
            tf = self.trans_fusion[i]
            outs.append(ops.conv_bn_act_cat(t, f, tf.conv.weight, tf.conv.bias, tf.norm, act="relu"))
        return tuple(outs)
