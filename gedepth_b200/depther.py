"""DepthEncoderDecoder + GroundEmbedding - host-side mirror of
depth/models/depther/encoder_decoder.py:22-274 and depth/models/depther/base.py:15-247.

``GroundEmbedding`` is new and additive: the reference smears the ground embedding
over an offline numpy script (tools/preprocess_data_kitti.py:47-56), ``extract_feat`` /
``dynamic_pe`` (encoder_decoder.py:79-124) and the head's fusion line (decode_head.py:489-508);
here one registered module owns the generator and the model-side embedding, each a single
HBM-bound sm_100a kernel.  ``DepthEncoderDecoder`` keeps the reference's constructor, call
signatures, kwargs (pe_ori_point, pe_k_gt, height, test) and state_dict keys.
"""
from __future__ import annotations

from collections import OrderedDict

import torch
import torch.distributed as dist
import torch.nn as nn

from . import builder, ops
from .builder import DEPTHER, MODELS
from .compat import BaseModule


def add_prefix(inputs, prefix):
    return {f"{prefix}.{k}": v for k, v in inputs.items()}


@MODELS.register_module()
class GroundEmbedding(nn.Module):
    """Ground-plane depth prior and its fusion with the predicted attention / slope maps.

    * ``ground_plane(coef, H, W, ...)``  - a1: pe[v,u] = num / (c_u*u + c_v*v + c_1) on the integer
      pixel grid, emitted as the two input channels the loader would attach (ch3 clamped and
      divided by depth_scale, ch4 raw; loading.py:388-403, transforms.py:40-48).
    * ``forward(img, y_half)``           - a14 (Vanilla): y = up(y_half), pe_mask = img[:,3]*y*200.
    * ``forward(img, y_half, logits_half, height)`` - a15 (Adaptive): softmax-expected slope,
      tan, inverse-depth shift, range mask, times y; also returns the full-resolution logits.
    """

    def __init__(self, depth_scale=200.0, adaptive=False, cam_height=1.65):
        super().__init__()
        self.depth_scale, self.adaptive, self.cam_height = float(depth_scale), adaptive, cam_height

    @staticmethod
    def ground_plane(coef, H, W, device, batch=1, u0=0, v0=0, depth_scale=200.0, clamp_max=None,
                     su=1.0, sv=1.0):
        return ops.ground_plane(coef, H, W, device, batch, u0, v0, depth_scale,
                                depth_scale if clamp_max is None else clamp_max, su, sv)

    def forward(self, img, y_half, logits_half=None, height=None):
        if logits_half is None:
            y, pe_mask = ops.ge_vanilla(img, y_half)
            return y, pe_mask, None
        h = self.cam_height if height is None else height
        # the full-resolution logits feed the CE loss of forward_train: keyed on training mode, not on autograd's grad mode
        # (forward_train under torch.no_grad() still reports loss_dynamic_pe)
        return ops.ge_adaptive(img, y_half, logits_half, h, self.depth_scale, want_logits=self.training or torch.is_grad_enabled())


class BaseDepther(BaseModule):
    """forward dispatch / train_step / _parse_losses of depther/base.py:97-204."""

    def __init__(self, init_cfg=None):
        super().__init__(init_cfg)
        self.fp16_enabled = False

    @property
    def with_neck(self):
        return hasattr(self, "neck") and self.neck is not None

    @property
    def with_decode_head(self):
        return hasattr(self, "decode_head") and self.decode_head is not None

    def forward_test(self, imgs, img_metas, **kwargs):
        for var, name in [(imgs, "imgs"), (img_metas, "img_metas")]:
            if not isinstance(var, list):
                raise TypeError(f"{name} must be a list, but got {type(var)}")
        if len(imgs) != len(img_metas):
            raise ValueError(f"num of augmentations ({len(imgs)}) != num of image meta ({len(img_metas)})")
        for img_meta in img_metas:
            for key in ("ori_shape", "img_shape", "pad_shape"):
                vals = [m[key] for m in img_meta if key in m]
                assert all(v == vals[0] for v in vals)
        if len(imgs) == 1:
            return self.simple_test(imgs[0], img_metas[0], **kwargs)
        return self.aug_test(imgs, img_metas, **kwargs)

    def forward(self, img, img_metas, return_loss=True, **kwargs):
        if return_loss:
            return self.forward_train(img, img_metas, **kwargs)
        return self.forward_test(img, img_metas, **kwargs)

    def train_step(self, data_batch, optimizer=None, **kwargs):
        losses = self(**data_batch)
        real_losses = {k: v for k, v in losses.items() if "img" not in k}
        log_imgs = {k: v for k, v in losses.items() if "img" in k}
        loss, log_vars = self._parse_losses(real_losses)
        return dict(loss=loss, log_vars=log_vars, num_samples=len(data_batch["img_metas"]),
                    log_imgs=log_imgs)

    def val_step(self, data_batch, **kwargs):
        return self(**data_batch, **kwargs)

    @staticmethod
    def _parse_losses(losses, sync=True):
        """Same contract as base.py:170-204 (loss = sum of keys containing 'loss'; log_vars are
        rank-averaged floats) but ONE all-reduce and ONE device->host copy for all scalars instead
        of one of each per key.  ``sync=False`` leaves log_vars as device tensors (no host sync)."""
        log_vars = OrderedDict()
        for name, value in losses.items():
            if isinstance(value, torch.Tensor):
                log_vars[name] = value.mean()
            elif isinstance(value, list):
                log_vars[name] = sum(v.mean() for v in value)
            else:
                raise TypeError(f"{name} is not a tensor or list of tensors")
        loss = sum(v for k, v in log_vars.items() if "loss" in k)
        log_vars["loss"] = loss
        if not sync:
            return loss, OrderedDict((k, v.detach()) for k, v in log_vars.items())
        packed = torch.stack([v.detach().float() for v in log_vars.values()])
        if dist.is_available() and dist.is_initialized():
            packed = packed / dist.get_world_size()
            dist.all_reduce(packed)
        for k, v in zip(list(log_vars.keys()), packed.tolist()):
            log_vars[k] = v
        return loss, log_vars


@DEPTHER.register_module()
class DepthEncoderDecoder(BaseDepther):
    def __init__(self, backbone, decode_head, neck=None, pe_mask_neck=None, dynamic_pe_neck=None,
                 train_cfg=None, test_cfg=None, pretrained=None, init_cfg=None, depth_scale=200):
        super().__init__(init_cfg)
        if pretrained is not None:
            assert backbone.get("pretrained") is None, "both backbone and depther set pretrained weight"
            backbone["pretrained"] = pretrained
        self.backbone = builder.build_backbone(backbone)
        self.decode_head = builder.build_head(decode_head)
        self.align_corners = self.decode_head.align_corners
        self.pe_mask_neck_FLAGS = self.dynamic_pe_neck_FLAGS = False
        self.depth_scale = depth_scale
        if neck is not None:
            self.neck = builder.build_neck(neck)
        if pe_mask_neck is not None:
            self.pe_mask_neck = builder.build_neck(pe_mask_neck)
            self.pe_mask_neck_FLAGS = True
        if dynamic_pe_neck is not None:
            self.dynamic_pe_neck = builder.build_neck(dynamic_pe_neck)
            self.dynamic_pe_neck_FLAGS = True
        self.ground_embedding = GroundEmbedding(depth_scale, adaptive=self.dynamic_pe_neck_FLAGS)
        self.train_cfg, self.test_cfg = train_cfg, test_cfg
        assert self.with_decode_head

    def dynamic_pe(self, x, y_half, img, img_metas, **kwargs):
        """encoder_decoder.py:79-102; ``y_half`` is the half-resolution attention map (its bilinear
        upsample is folded into the kernel)."""
        height = None
        if "height" in kwargs:
            height = kwargs["height"][0] if "test" in kwargs else kwargs["height"]
        logits_half = self.dynamic_pe_neck(x)
        return self.ground_embedding(img, y_half, logits_half, height)

    def extract_feat(self, img, img_metas, **kwargs):
        ops.require_cuda(img)
        x = self.backbone(img)
        if self.with_neck:
            x = self.neck(x)
            if self.pe_mask_neck_FLAGS:
                y_half, _ = self.pe_mask_neck(x)
                if self.dynamic_pe_neck_FLAGS:
                    y, pe_mask, logits = self.dynamic_pe(x, y_half, img, img_metas, **kwargs)
                    return x, y, pe_mask, logits
                y, pe_mask, _ = self.ground_embedding(img, y_half)
                return x, y, pe_mask, None
        return x, None, None, None

    def encode_decode(self, img, img_metas, rescale=True, **kwargs):
        x, y, pe_mask, _ = self.extract_feat(img, img_metas, **kwargs)
        out = self.decode_head.forward_test(img, x, img_metas, self.test_cfg, pe_mask, y, **kwargs)
        return ops.clamp_resize(out, self.decode_head.min_depth, self.decode_head.max_depth,
                                img.shape[2:] if rescale else None, self.align_corners)

    def forward_dummy(self, img):
        return self.encode_decode(img, None)

    def forward_train(self, img, img_metas, depth_gt, **kwargs):
        x, y, pe_mask, pe_offset = self.extract_feat(img, img_metas, **kwargs)
        loss_decode = self.decode_head.forward_train(img, x, img_metas, depth_gt, self.train_cfg,
                                                     pe_mask, y, pe_offset, **kwargs)
        return add_prefix(loss_decode, "decode")

    def whole_inference(self, img, img_meta, rescale, **kwargs):
        return self.encode_decode(img, img_meta, rescale, **kwargs)

    def inference(self, img, img_meta, rescale, **kwargs):
        assert self.test_cfg["mode"] in ["slide", "whole"]
        ori_shape = img_meta[0]["ori_shape"]
        assert all(m["ori_shape"] == ori_shape for m in img_meta)
        if self.test_cfg["mode"] == "slide":
            raise NotImplementedError
        output = self.whole_inference(img, img_meta, rescale, **kwargs)
        if img_meta[0]["flip"]:
            d = img_meta[0]["flip_direction"]
            assert d in ["horizontal", "vertical"]
            output = output.flip(dims=(3,) if d == "horizontal" else (2,))
        return output

    def simple_test(self, img, img_meta, rescale=True, **kwargs):
        return list(self.inference(img, img_meta, rescale, **kwargs).cpu().numpy())

    def aug_test(self, imgs, img_metas, rescale=True, **kwargs):
        assert rescale
        kwargs["pe_ori_point_test"] = kwargs["pe_ori_point"][0]
        if "pe_k_gt" in kwargs:
            kwargs["pe_k_gt_test"] = kwargs["pe_k_gt"][0]
        m0, m1 = img_metas[0][0], (img_metas[1][0] if len(imgs) == 2 else None)
        if (m1 is not None and not m0.get("flip") and m1.get("flip") and m1.get("flip_direction") == "horizontal"
                and self.test_cfg["mode"] == "whole" and ops.use_native("tta_merge")):
            # the two-view flip TTA of every GE config: un-flip + average in one kernel (ops.tta_merge)
            p0 = self.whole_inference(imgs[0], img_metas[0], rescale, **kwargs)
            kwargs.update({"pe_ori_point_test": kwargs["pe_ori_point"][1]})
            if "pe_k_gt" in kwargs:
                kwargs.update({"pe_k_gt_test": kwargs["pe_k_gt"][1]})
            p1 = self.whole_inference(imgs[1], img_metas[1], rescale, **kwargs)
            return list(ops.tta_merge(p0, p1).cpu().numpy())
        depth_pred = self.inference(imgs[0], img_metas[0], rescale, **kwargs)
        for i in range(1, len(imgs)):
            kwargs.update({"pe_ori_point_test": kwargs["pe_ori_point"][i]})
            if "pe_k_gt" in kwargs:
                kwargs.update({"pe_k_gt_test": kwargs["pe_k_gt"][i]})
            depth_pred = depth_pred + self.inference(imgs[i], img_metas[i], rescale, **kwargs)
        depth_pred = depth_pred / len(imgs)
        return list(depth_pred.cpu().numpy())
