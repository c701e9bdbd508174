"""PE necks, DenseDepth decode head and losses - host-side mirror of
depth/models/necks/pemask_neck.py:29-64 (``LightPEMASKNeck``), necks/dynamicpe_neck.py:490-539
(``DynamicPENeckSOFT``), decode_heads/decode_head.py:268-648 (``DepthBaseDecodeHead``),
decode_heads/densedepth_head.py:14-131 (``UpSample``, ``DenseDepthHead``),
losses/sigloss.py:9-69 (``SigLoss``), losses/celoss.py:355-413 (``CrossEntropyLoss``) and
losses/bceloss.py (``BinaryCrossEntropyLoss``, constructed by default, never called).
Same registry names, constructor keywords and state_dict keys.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from .builder import HEADS, LOSSES, NECKS, build_loss
from .compat import BaseModule, ConvModule, xavier_init


class _PENeck(BaseModule):
    """Five 3x3 convs (C_i -> 64) summed on the stem grid (bilinear, align_corners=True), then a
    3x3 ``convfinal``.  Channel widths are hard-coded to the Swin-L neck (pemask_neck.py:38-42)."""
    out_ch = 1

    def __init__(self):
        super().__init__()
        self.convfinal = nn.Conv2d(64, self.out_ch, kernel_size=3, padding=1, stride=1)
        self.conv0 = nn.Conv2d(1536, 64, kernel_size=3, padding=1, stride=1)
        self.conv1 = nn.Conv2d(768, 64, kernel_size=3, padding=1, stride=1)
        self.conv2 = nn.Conv2d(384, 64, kernel_size=3, padding=1, stride=1)
        self.conv3 = nn.Conv2d(192, 64, kernel_size=3, padding=1, stride=1)
        self.conv4 = nn.Conv2d(64, 64, kernel_size=3, padding=1, stride=1)

    def init_weights(self):
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                xavier_init(m, distribution="uniform")

    def _sum(self, inputs):
        x0, x1, x2, x3, x4 = inputs[::-1]
        size = (x4.shape[2], x4.shape[3])
        acc = ops.conv2d(x4, self.conv4.weight, self.conv4.bias, padding=1)
        for conv, t in ((self.conv0, x0), (self.conv1, x1), (self.conv2, x2), (self.conv3, x3)):
            t = ops.conv2d(t, conv.weight, conv.bias, padding=1)
            acc = ops.resize_add(t, size, acc)         # acc += bilinear(t -> size, align_corners=True)
        return acc


@NECKS.register_module()
class LightPEMASKNeck(_PENeck):
    out_ch = 1

    def __init__(self):
        super().__init__()
        self.sigmoid = nn.Sigmoid()

    def forward(self, inputs):
        x = self._sum(inputs)
        y = ops.conv2d(x, self.convfinal.weight, self.convfinal.bias, padding=1, act="sigmoid")
        return y, x


@NECKS.register_module()
class DynamicPENeckSOFT(_PENeck):
    out_ch = 11

    def forward(self, inputs):
        return ops.conv2d(self._sum(inputs), self.convfinal.weight, self.convfinal.bias, padding=1)


# ---------------------------------------------------------------------------------------------
# losses
# ---------------------------------------------------------------------------------------------
@LOSSES.register_module()
class SigLoss(nn.Module):
    """SiLog: g = log(pred+eps) - log(gt+eps) over gt>0; sqrt(var_unbiased(g) + 0.15 mean(g)^2)."""

    def __init__(self, loss_name="loss_sig", valid_mask=True, loss_weight=1.0, max_depth=None,
                 warm_up=False, warm_iter=100):
        super().__init__()
        if warm_up or not valid_mask:
            raise NotImplementedError("GE configs use valid_mask=True, warm_up=False (sigloss.py:16-22)")
        self._loss_name, self.valid_mask, self.loss_weight = loss_name, valid_mask, loss_weight
        self.max_depth, self.eps = max_depth, 0.001

    def forward(self, depth_pred, depth_gt, **kwargs):
        """depth_pred already at gt resolution (reference contract)."""
        return self.loss_weight * ops.silog(depth_pred, depth_gt, self.eps, 0.15, self.max_depth,
                                            upsample=False)

    def forward_fused(self, depth_half, depth_gt):
        """SiLog of bilinear(depth_half -> gt size, align_corners=True) without materialising the
        full-resolution prediction (decode_head.py:586-599 + sigloss.py:36-53 in one pass)."""
        return self.loss_weight * ops.silog(depth_half, depth_gt, self.eps, 0.15, self.max_depth,
                                            upsample=True)

    @property
    def loss_name(self):
        return self._loss_name


@LOSSES.register_module()
class CrossEntropyLoss(nn.Module):
    """nn.CrossEntropyLoss(ignore_index=255) * loss_weight (celoss.py:355-413)."""

    def __init__(self, loss_weight=1.0):
        super().__init__()
        self.loss_weight = loss_weight

    def forward(self, input, target):
        return self.loss_weight * ops.cross_entropy(input, target, ignore_index=255)


@LOSSES.register_module()
class BinaryCrossEntropyLoss(nn.Module):
    """Constructed by DepthBaseDecodeHead's default ``loss_pe`` (decode_head.py:310-312); never
    called on the GE path."""

    def __init__(self, loss_weight=1.0):
        super().__init__()
        self.loss_weight = loss_weight

    def forward(self, input, target):
        raise NotImplementedError("loss_pe is dead code on the GEDepth path")


# ---------------------------------------------------------------------------------------------
# decode head
# ---------------------------------------------------------------------------------------------
class DepthBaseDecodeHead(BaseModule):
    def __init__(self, in_channels, channels=96, conv_cfg=None, act_cfg=dict(type="ReLU"),
                 loss_decode=dict(type="SigLoss", valid_mask=True, loss_weight=10),
                 loss_pe=dict(type="BinaryCrossEntropyLoss", loss_weight=10),
                 loss_dynamic_pe=dict(type="CrossEntropyLoss", loss_weight=0.08),
                 loss_surface_norm=None, sampler=None, align_corners=False, min_depth=1e-3,
                 max_depth=None, norm_cfg=None, classify=False, n_bins=256, bins_strategy="UD",
                 norm_strategy="linear", scale_up=False, depth2norm=False):
        super().__init__()
        if classify or scale_up or depth2norm:
            raise NotImplementedError("classify / scale_up / depth2norm are off in every GE config")
        self.in_channels, self.channels = in_channels, channels
        self.conv_cfg, self.act_cfg, self.norm_cfg = conv_cfg, act_cfg, norm_cfg
        self.loss_decode = build_loss(loss_decode)
        self.loss_pe = build_loss(loss_pe)
        self.loss_dynamic_pe = build_loss(loss_dynamic_pe)
        self.align_corners, self.min_depth, self.max_depth = align_corners, min_depth, max_depth
        self.conv_depth = nn.Conv2d(channels, 1, kernel_size=3, padding=1, stride=1)
        self.fp16_enabled = False

    def extra_repr(self):
        return f"align_corners={self.align_corners}"

    def depth_pred(self, feat, pe, depth_y):
        """out = relu(conv_depth(feat)) * (1 - y_h) + pe_h + min_depth with pe_h, y_h the
        align_corners bilinear resamples of the full-resolution maps (decode_head.py:489-508)."""
        d = ops.conv2d(feat, self.conv_depth.weight, self.conv_depth.bias, padding=1, act="relu")
        if pe is None:
            return d + self.min_depth, depth_y
        if not self.align_corners:
            raise NotImplementedError("GE configs set align_corners=True (_base_/models/depthformer_swin.py:36)")
        return ops.fuse_head(d, pe, depth_y, self.min_depth)

    def forward_train(self, img, inputs, img_metas, depth_gt, train_cfg, pe_mask, y, pe_offset, **kwargs):
        depth_pred, _ = self.forward(inputs, img_metas, pe_mask, y)
        if pe_offset is not None:
            losses = self.losses_dynamic_pe(depth_pred, depth_gt, pe_offset, kwargs["pe_k_gt"], None, None)
        else:
            losses = self.losses(depth_pred, depth_gt)
        # The reference also returns three log images here via .cpu() every step
        # (decode_head.py:438,628-648); that device->host sync is logging, not the path.
        return losses

    def forward_test(self, img, inputs, img_metas, test_cfg, pe_mask, y, **kwargs):
        depth_pred, _ = self.forward(inputs, img_metas, pe_mask, y)
        return depth_pred

    def losses(self, depth_pred, depth_gt, **kwargs):
        return {"loss_depth": self._loss_depth(depth_pred, depth_gt)}

    def losses_dynamic_pe(self, depth_pred, depth_gt, dynamic_pe, pe_k_gt, attn_pred=None, attn_gt=None):
        loss = {"loss_dynamic_pe": self.loss_dynamic_pe(dynamic_pe, pe_k_gt)}
        loss["loss_depth"] = self._loss_depth(depth_pred, depth_gt)
        return loss

    def _loss_depth(self, depth_pred, depth_gt):
        if not self.align_corners:
            raise NotImplementedError("GE configs set align_corners=True")
        if isinstance(self.loss_decode, SigLoss):
            return self.loss_decode.forward_fused(depth_pred, depth_gt)
        pred = ops.resize(depth_pred, depth_gt.shape[2:], align_corners=True)
        return self.loss_decode(pred, depth_gt)


class UpSample(nn.Sequential):
    """bilinear(align_corners=True) to the skip size, concat, two 3x3 conv+act
    (densedepth_head.py:14-27)."""

    def __init__(self, skip_input, output_features, conv_cfg=None, norm_cfg=None, act_cfg=None):
        super().__init__()
        if norm_cfg is not None:
            raise NotImplementedError("decoder norm_cfg is None in every GE config")
        self.convA = ConvModule(skip_input, output_features, kernel_size=3, stride=1, padding=1,
                                conv_cfg=conv_cfg, norm_cfg=norm_cfg, act_cfg=act_cfg)
        self.convB = ConvModule(output_features, output_features, kernel_size=3, stride=1, padding=1,
                                conv_cfg=conv_cfg, norm_cfg=norm_cfg, act_cfg=act_cfg)

    def forward(self, x, concat_with):
        a, b = self.convA, self.convB
        t = ops.conv2d_cat(x, concat_with, a.conv.weight, a.conv.bias, act=_act_name(a), slope=a.act_slope)
        return ops.conv2d(t, b.conv.weight, b.conv.bias, padding=1, act=_act_name(b), slope=b.act_slope)


def _act_name(m: ConvModule):
    return None if m.act_kind is None else {"ReLU": "relu", "LeakyReLU": "leaky_relu"}[m.act_kind]


@HEADS.register_module()
class DenseDepthHead(DepthBaseDecodeHead):
    def __init__(self, up_sample_channels, fpn=False, conv_dim=256, **kwargs):
        super().__init__(**kwargs)
        if fpn:
            raise NotImplementedError("fpn=False in every GE config")
        self.up_sample_channels = up_sample_channels[::-1]
        self.in_channels = self.in_channels[::-1]
        self.fpn = fpn
        self.conv_list = nn.ModuleList()
        prev = 0
        for index, (cin, cup) in enumerate(zip(self.in_channels, self.up_sample_channels)):
            if index == 0:
                self.conv_list.append(ConvModule(cin, cup, kernel_size=1, stride=1, padding=0, act_cfg=None))
            else:
                self.conv_list.append(UpSample(cin + prev, cup, norm_cfg=self.norm_cfg, act_cfg=self.act_cfg))
            prev = cup

    def forward(self, inputs, img_metas, pe_mask, depth_mask_y):
        t = None
        for index, feat in enumerate(inputs[::-1]):
            if index == 0:
                c = self.conv_list[0]
                t = ops.conv2d(feat, c.conv.weight, c.conv.bias)
            else:
                t = self.conv_list[index](t, feat)
        return self.depth_pred(t, pe_mask, depth_mask_y)
