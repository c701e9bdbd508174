"""Operator layer of the GEDepth path: every function takes CUDA tensors and launches hand-written
sm_100a kernels through the C-ABI in ``libgedepth_sm100.so`` (include/gedepth.h) on the current
stream.  There is no CPU path: CPU tensors raise, a missing extension raises.

``native_table()`` lists, op by op, whether the sm_100a kernel is in place or the op still goes to
the library statement in ops_lib.py (cuDNN/cuBLAS/ATen on the same CUDA tensors).
"""
from __future__ import annotations

import os
from typing import Optional, Sequence

import torch

from . import ops_lib as L

_NATIVE = {}          # op name -> bool, filled by kernels.py when the extension is loaded
_FORCE_LIB = set(filter(None, os.environ.get("GEDEPTH_FORCE_LIB", "").split(",")))


def require_cuda(*tensors):
    for t in tensors:
        if torch.is_tensor(t) and not t.is_cuda:
            raise RuntimeError(
                "gedepth_b200 runs on sm_100a only: got a CPU tensor. There is no CPU fallback "
                "(the CPU restatement lives in oracle/ and is test infrastructure).")


def _k():
    from . import kernels
    return kernels


def use_native(name: str) -> bool:
    if name in _FORCE_LIB or "all" in _FORCE_LIB:
        return False
    return _k().has(name)


def native_table():
    k = _k()
    return {name: (k.has(name) and name not in _FORCE_LIB and "all" not in _FORCE_LIB) for name in OPS}


# ops reached by the four GE configs (``resize`` alone is not: it only serves a non-SiLog loss, heads.py `_loss_depth`)
OPS = ["linear", "layer_norm", "patch_embed", "merge_patches", "window_attention", "conv2d",
       "conv_bn_act", "conv2d_cat", "batch_norm", "resize_add", "msda_sample", "ground_plane", "ge_vanilla",
       "ge_adaptive", "fuse_head", "silog", "cross_entropy", "clamp_resize", "find_k", "depth_metrics", "tta_merge",
       "adamw"]


# ---- GEMM-shaped ---------------------------------------------------------------------------
def linear(x, w, b=None, act=None, residual=None, row_scale=None, dropout_p: float = 0.0):
    """residual + dropout_p(act(x w^T + b) * row_scale); dropout only when dropout_p > 0 (training)."""
    require_cuda(x, w)
    if use_native("linear"):
        return _k().linear(x, w, b, act, residual, row_scale, dropout_p)
    y = L.linear(x, w, b, act, None if dropout_p > 0 else residual, row_scale)
    if dropout_p > 0:
        y = torch.nn.functional.dropout(y, dropout_p, True)
        y = y if residual is None else y + residual
    return y


def conv2d(x, w, b=None, stride=1, padding=0, act=None, slope=0.01):
    require_cuda(x, w)
    if use_native("conv2d") and _k().conv2d_supported(x, w, stride, padding):
        return _k().conv2d(x, w, b, stride, padding, act, slope)
    return L.conv2d(x, w, b, stride, padding, act, slope)


def conv_bn_act(x, w, b, bn, stride=1, padding=0, act=None):
    require_cuda(x, w)
    if use_native("conv_bn_act") and _k().conv2d_supported(x, w, stride, padding):
        return _k().conv_bn_act(x, w, b, bn, stride, padding, act)
    if use_native("conv_bn_act") and _k().conv_im2col_supported(x, w, stride, padding):
        # stem 7x7/s2 conv on the RGB planes (K = 147): im2col gather + tcgen05 GEMM, then BN (+ReLU)
        if bn is None:
            return _k().conv_im2col(x, w, b, stride, padding, act)
        if not bn.training:
            s = bn.weight * torch.rsqrt(bn.running_var + bn.eps)
            bf = bn.bias - bn.running_mean * s + (b * s if b is not None else 0)
            return _k().conv_im2col(x, w * s.view(-1, 1, 1, 1), bf, stride, padding, act)
        y = _k().conv_im2col(x, w, b, stride, padding, None)
        if act in (None, "relu") and bn.momentum is not None and bn.track_running_stats and use_native("batch_norm"):
            return _k().bn_act_train(y, bn, relu=act == "relu")
        return L._act(bn(y), act)
    if (bn is not None and bn.training and use_native("batch_norm") and w.shape[0] % 4 == 0 and act in (None, "relu")
            and bn.momentum is not None and bn.track_running_stats):
        return _k().bn_act_train(L.conv2d(x, w, b, stride, padding), bn, relu=act == "relu")
    return L.conv_bn_act(x, w, b, bn, stride, padding, act)


def conv2d_cat(x_low, x_skip, w, b=None, act=None, slope=0.01):
    """3x3 conv (+bias, +act) over cat([bilinear(x_low -> skip size, align_corners=True), x_skip], 1):
    the UpSample block of densedepth_head.py:24-27 without materialising the resize or the concat."""
    require_cuda(x_low, x_skip, w)
    if use_native("conv2d_cat") and _k().conv2d_cat_supported(x_low, x_skip, w):
        return _k().conv2d_cat(x_low, x_skip, w, b, act, slope)
    up = L.resize(x_low, (x_skip.shape[2], x_skip.shape[3]), True) if x_low.shape[2:] != x_skip.shape[2:] else x_low
    return L.conv2d(L.cat_channels([up, x_skip]), w, b, 1, 1, act, slope)


def conv_bn_act_cat(x0, x1, w, b, bn, act=None):
    """ConvModule(3x3, BN, act) over cat([x0, x1], 1) (hahi.py:329-353) without the concat copy."""
    require_cuda(x0, x1, w)
    if use_native("conv2d_cat") and _k().conv2d_cat_supported(x0, x1, w) and x0.shape[2:] == x1.shape[2:]:
        return _k().conv_bn_act_cat(x0, x1, w, b, bn, act)
    return L.conv_bn_act(L.cat_channels([x0, x1]), w, b, bn, 1, 1, act)


def patch_embed(x, w, b, patch):
    require_cuda(x, w)
    if use_native("patch_embed") and x.dtype == torch.float32 and x.stride(3) == 1 and x.stride(2) == x.shape[3] \
            and x.stride(1) == x.shape[2] * x.shape[3] and (x.shape[1] * patch * patch) % 32 == 0:
        return _k().patch_embed(x, w, b, patch)
    return L.patch_embed(x, w, b, patch)


# ---- token-shaped --------------------------------------------------------------------------
def layer_norm(x, w, b, eps):
    require_cuda(x)
    if use_native("layer_norm"):
        return _k().layer_norm(x, w, b, eps)
    return L.layer_norm(x, w, b, eps)


def layer_norm_fork(x, w, b, eps):
    """(LN(x), x): the second output is x itself, to be used as the residual identity of the sub-block that follows, so
    the LayerNorm backward can add the residual-branch gradient in its own pass."""
    require_cuda(x)
    if use_native("layer_norm"):
        return _k().layer_norm_fork(x, w, b, eps)
    return L.layer_norm(x, w, b, eps), x


def merge_patches(x, H, W):
    require_cuda(x)
    if use_native("merge_patches"):
        return _k().merge_patches(x, H, W)
    return L.merge_patches(x, H, W)


def window_attention(qkv, qkv_bias, table, index, hw, nH, ws, shift, scale):
    require_cuda(qkv)
    if use_native("window_attention") and ws == 7 and qkv.shape[-1] // 3 // nH == 32:
        return _k().window_attention(qkv, qkv_bias, table, index, hw, nH, ws, shift, scale)
    return L.window_attention(qkv, qkv_bias, table, index, hw, nH, ws, shift, scale)


def drop_path_scale(drop, x) -> Optional[torch.Tensor]:
    """(B,) per-sample factor floor(keep + U[0,1)) / keep, or None when DropPath is inactive; it is
    applied in the epilogue of the GEMM that produces the residual branch."""
    p = getattr(drop, "drop_prob", 0.0)
    if not getattr(drop, "training", False) or p == 0.0:
        return None
    keep = 1.0 - p
    return (keep + torch.rand(x.shape[0], dtype=x.dtype, device=x.device)).floor_().div_(keep)


def tokens_to_map(x, hw):
    """(B, L, C) -> logical (B, C, h, w) in channels-last memory: a view, no copy."""
    return L.tokens_to_map(x, hw)


def map_to_tokens(x):
    """logical (B, C, h, w) -> (B, h*w, C); a view when x is channels-last."""
    return L.map_to_tokens(x)


def cat_channels(xs):
    return L.cat_channels(xs)


def add_bcast(a, b):
    return L.add_bcast(a, b)


# ---- resampling ----------------------------------------------------------------------------
def resize(x, size, align_corners=True):
    require_cuda(x)
    if use_native("resize") and align_corners:
        return _k().resize(x, size)
    return L.resize(x, size, align_corners)


def resize_add(t, size, acc):
    require_cuda(t)
    if use_native("resize_add") and t.shape[1] % 4 == 0:
        return _k().resize_add(t, size, acc)
    return L.resize_add(t, size, acc)


def clamp_resize(x, lo, hi, size, align_corners=True):
    require_cuda(x)
    if use_native("clamp_resize") and align_corners and x.shape[1] == 1 and not torch.is_grad_enabled():
        return _k().clamp_resize(x, lo, hi, size)
    return L.clamp_resize(x, lo, hi, size, align_corners)


def msda_sample(v, shapes, ref, off, logit, nH, P):
    require_cuda(v)
    if use_native("msda_sample") and v.shape[-1] // nH == 64 and len(shapes) * P == 32:
        return _k().msda_sample(v, shapes, ref, off, logit, nH, P)
    return L.msda_sample(v, shapes, ref, off, logit, nH, P)


# ---- ground embedding ----------------------------------------------------------------------
def ground_plane(coef, H, W, device, batch=1, u0=0, v0=0, depth_scale=200.0, clamp_max=200.0,
                 su=1.0, sv=1.0):
    if torch.device(device).type != "cuda":
        raise RuntimeError("gedepth_b200 runs on sm_100a only (ground_plane on a non-CUDA device)")
    if use_native("ground_plane"):
        return _k().ground_plane(coef, H, W, device, batch, u0, v0, depth_scale, clamp_max, su, sv)
    return L.ground_plane(coef, H, W, device, batch, u0, v0, depth_scale, clamp_max, su, sv)


def ge_vanilla(img, y_half):
    require_cuda(img, y_half)
    if use_native("ge_vanilla"):
        return _k().ge_vanilla(img, y_half)
    return L.ge_vanilla(img, y_half)


def ge_adaptive(img, y_half, logits_half, height, depth_scale):
    require_cuda(img, y_half, logits_half)
    if use_native("ge_adaptive"):
        return _k().ge_adaptive(img, y_half, logits_half, height, depth_scale)
    return L.ge_adaptive(img, y_half, logits_half, height, depth_scale)


def fuse_head(d, pe_mask, y, min_depth):
    require_cuda(d, pe_mask, y)
    if use_native("fuse_head"):
        return _k().fuse_head(d, pe_mask, y, min_depth)
    return L.fuse_head(d, pe_mask, y, min_depth)


def silog(pred, gt, eps=1e-3, lam=0.15, max_depth=None, upsample=False):
    require_cuda(pred, gt)
    if use_native("silog"):
        return _k().silog(pred, gt, eps, lam, max_depth, upsample)
    return L.silog(pred, gt, eps, lam, max_depth, upsample)


def cross_entropy(logits, target, ignore_index=255):
    require_cuda(logits, target)
    if use_native("cross_entropy"):
        return _k().cross_entropy(logits, target, ignore_index)
    return L.cross_entropy(logits, target, ignore_index)


def depth_metric_sums(pred, gt, rect, min_depth, max_depth, sums=None):
    """(B,10) fp64 sums of the nine-metric reduction (metrics.py:8-45) over the evaluation mask."""
    require_cuda(pred, gt)
    return _k().depth_metric_sums(pred, gt, rect, min_depth, max_depth, sums)


def tta_merge(a, b_flipped):
    require_cuda(a, b_flipped)
    return _k().tta_merge(a, b_flipped)


def find_k(gt, pe, h, truncate=False):
    """Slope labels: float32 (B,H,W) in {-5..5} | 255."""
    require_cuda(gt, pe)
    return _k().find_k(gt, pe, h, truncate)
