"""Operator layer of the GEDepth path: every function takes CUDA tensors and launches hand-written
sm_100a kernels through the C-ABI in ``libgedepth_sm100.so`` (include/gedepth.h) on the current
stream.  There is ONE implementation per op: no CPU path, no library (cuDNN / cuBLAS / ATen) dispatch - CPU
tensors raise, a missing extension raises, a shape the kernels do not cover raises ``NotImplementedError``.
The library statement of every op lives in tests/ops_lib.py: the checker of the per-op tests and the
``gpu_library_baseline`` of bench.py, never part of the product.
"""
from __future__ import annotations

from typing import Optional

import torch


def require_cuda(*tensors):
    for t in tensors:
        if torch.is_tensor(t) and not t.is_cuda:
            raise RuntimeError(
                "gedepth_b200 runs on sm_100a only: got a CPU tensor. There is no CPU fallback "
                "(the CPU restatement lives in oracle/ and is test infrastructure).")


def _k():
    from . import kernels
    kernels.load()
    return kernels


def _unsupported(op: str, why: str):
    raise NotImplementedError(f"gedepth_b200.ops.{op}: {why} - outside what the four GE configs use; there is no "
                              f"library fallback in the product (tests/ops_lib.py holds the reference statement)")


# ops reached by the four GE configs
OPS = ["linear", "layer_norm", "patch_embed", "merge_patches", "window_attention", "conv2d",
       "conv_bn_act", "conv2d_cat", "batch_norm", "resize_add", "msda_module", "ground_plane", "ge_vanilla",
       "ge_adaptive", "fuse_head", "silog", "cross_entropy", "clamp_resize", "find_k", "depth_metrics", "tta_merge",
       "adamw"]


def use_native(name: str) -> bool:
    """Kept for callers that used to ask; every op is native (or raises)."""
    _k()
    return True


def native_table():
    _k()
    return {name: True for name in OPS}


# ---- GEMM-shaped ---------------------------------------------------------------------------
def linear(x, w, b=None, act=None, residual=None, row_scale=None, dropout_p: float = 0.0):
    """residual + dropout_p(act(x w^T + b) * row_scale); dropout only when dropout_p > 0 (training)."""
    require_cuda(x, w)
    return _k().linear(x, w, b, act, residual, row_scale, dropout_p)


def conv2d(x, w, b=None, stride=1, padding=0, act=None, slope=0.01):
    require_cuda(x, w)
    k = _k()
    if not k.conv2d_supported(x, w, stride, padding):
        _unsupported("conv2d", f"weight {tuple(w.shape)}, stride {stride}, padding {padding}")
    return k.conv2d(x, w, b, stride, padding, act, slope)


def conv_bn_act(x, w, b, bn, stride=1, padding=0, act=None):
    require_cuda(x, w)
    k = _k()
    if k.conv2d_supported(x, w, stride, padding):
        return k.conv_bn_act(x, w, b, bn, stride, padding, act)
    if k.conv_im2col_supported(x, w, stride, padding):
        # stem 7x7/s2 conv on the RGB planes (K = 147): im2col gather + tcgen05 GEMM, then BN (+ReLU)
        if bn is None:
            return k.conv_im2col(x, w, b, stride, padding, act)
        if not bn.training:
            s = bn.weight * torch.rsqrt(bn.running_var + bn.eps)
            bf = bn.bias - bn.running_mean * s + (b * s if b is not None else 0)
            return k.conv_im2col(x, w * s.view(-1, 1, 1, 1), bf, stride, padding, act)
        if act not in (None, "relu") or bn.momentum is None or not bn.track_running_stats:
            _unsupported("conv_bn_act", "train-mode BatchNorm with this activation / momentum setting")
        return k.bn_act_train(k.conv_im2col(x, w, b, stride, padding, None), bn, relu=act == "relu")
    _unsupported("conv_bn_act", f"weight {tuple(w.shape)}, stride {stride}, padding {padding}")


def conv2d_cat(x_low, x_skip, w, b=None, act=None, slope=0.01):
    """3x3 conv (+bias, +act) over cat([bilinear(x_low -> skip size, align_corners=True), x_skip], 1):
    the UpSample block of densedepth_head.py:24-27 without materialising the resize or the concat."""
    require_cuda(x_low, x_skip, w)
    k = _k()
    if not k.conv2d_cat_supported(x_low, x_skip, w):
        _unsupported("conv2d_cat", f"weight {tuple(w.shape)} over {tuple(x_low.shape)} | {tuple(x_skip.shape)}")
    return k.conv2d_cat(x_low, x_skip, w, b, act, slope)


def conv_bn_act_cat(x0, x1, w, b, bn, act=None):
    """ConvModule(3x3, BN, act) over cat([x0, x1], 1) (hahi.py:329-353) without the concat copy."""
    require_cuda(x0, x1, w)
    k = _k()
    if not (k.conv2d_cat_supported(x0, x1, w) and x0.shape[2:] == x1.shape[2:]):
        _unsupported("conv_bn_act_cat", f"weight {tuple(w.shape)} over {tuple(x0.shape)} | {tuple(x1.shape)}")
    return k.conv_bn_act_cat(x0, x1, w, b, bn, act)


def patch_embed(x, w, b, patch):
    require_cuda(x, w)
    if not (x.dtype == torch.float32 and x.stride(3) == 1 and x.stride(2) == x.shape[3]
            and x.stride(1) == x.shape[2] * x.shape[3] and (x.shape[1] * patch * patch) % 32 == 0):
        _unsupported("patch_embed", f"input {tuple(x.shape)} strides {x.stride()} patch {patch}")
    return _k().patch_embed(x, w, b, patch)


# ---- token-shaped --------------------------------------------------------------------------
def layer_norm(x, w, b, eps):
    require_cuda(x)
    return _k().layer_norm(x, w, b, eps)


def layer_norm_fork(x, w, b, eps):
    """(LN(x), x): the second output is x itself, to be used as the residual identity of the sub-block that follows, so
    the LayerNorm backward can add the residual-branch gradient in its own pass."""
    require_cuda(x)
    return _k().layer_norm_fork(x, w, b, eps)


def merge_patches(x, H, W):
    require_cuda(x)
    return _k().merge_patches(x, H, W)


def window_attention(qkv, qkv_bias, table, index, hw, nH, ws, shift, scale):
    require_cuda(qkv)
    if ws != 7 or qkv.shape[-1] // 3 // nH != 32:
        _unsupported("window_attention", f"window {ws}, head dim {qkv.shape[-1] // 3 // nH} (kernels: 7, 32)")
    return _k().window_attention(qkv, qkv_bias, table, index, hw, nH, ws, shift, scale)


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
    B, L, C = x.shape
    return x.reshape(B, hw[0], hw[1], C).permute(0, 3, 1, 2)


def map_to_tokens(x):
    """logical (B, C, h, w) -> (B, h*w, C); a view when x is channels-last."""
    B, C, H, W = x.shape
    return x.permute(0, 2, 3, 1).reshape(B, H * W, C)


def cat_tokens(xs):
    """Concatenate token sequences (B, L_i, C) along L (one copy; not arithmetic)."""
    return torch.cat(xs, dim=1)


def split_levels(src, sizes):
    """(B, S, C) -> list of (B, n_l, C) views; their gradients land in one buffer (kernels._SplitLevels)."""
    require_cuda(src)
    return _k().split_levels(src, sizes)


# ---- resampling ----------------------------------------------------------------------------
def resize(x, size, align_corners=True):
    _unsupported("resize", "a materialised bilinear resize is not on the GE path (SiLog / conv inputs / fuse_head resize "
                           "inside their own kernels)")


def resize_add(t, size, acc):
    require_cuda(t)
    if t.shape[1] % 4:
        _unsupported("resize_add", f"{t.shape[1]} channels (multiple of 4 needed)")
    return _k().resize_add(t, size, acc)


def clamp_resize(x, lo, hi, size, align_corners=True):
    require_cuda(x)
    if not (align_corners and x.shape[1] == 1) or torch.is_grad_enabled() and x.requires_grad:
        _unsupported("clamp_resize", "inference-only, single channel, align_corners=True (encoder_decoder.py:132-138)")
    return _k().clamp_resize(x, lo, hi, size)


# ---- deformable attention -------------------------------------------------------------------
def msda_module(query, value, pos, level_embed, level_start, ref, shapes, mod, dropout_p):
    """mmcv MultiScaleDeformableAttention.forward(batch_first=True) as ONE autograd node (kernels._MSDAModule):
    q = query + pos (+ level embedding); value_proj / sampling_offsets / attention_weights GEMMs; sampling;
    output_proj + dropout + identity.  The backward sums the query gradient's fan-in inside GEMM epilogues."""
    require_cuda(query)
    if mod.embed_dims // mod.num_heads != 64 or mod.num_levels * mod.num_points != 32 or len(shapes) != 4:
        _unsupported("msda_module", "kernels cover 4 levels x 8 points, head dim 64 (hahi.py:179-188)")
    return _k().msda_module(query, value, pos, level_embed, level_start, ref, shapes, mod, dropout_p)


def linear_small(x, w, b, act=None):
    """Linear with <= 4 outputs (+sigmoid): HAHIHeteroNeck.reference_points (hahi.py:299-300)."""
    require_cuda(x, w)
    return _k().linear_small(x, w, b, act)


# ---- ground embedding ----------------------------------------------------------------------
def ground_plane(coef, H, W, device, batch=1, u0=0, v0=0, depth_scale=200.0, clamp_max=200.0,
                 su=1.0, sv=1.0):
    if torch.device(device).type != "cuda":
        raise RuntimeError("gedepth_b200 runs on sm_100a only (ground_plane on a non-CUDA device)")
    return _k().ground_plane(coef, H, W, device, batch, u0, v0, depth_scale, clamp_max, su, sv)


def ge_vanilla(img, y_half):
    require_cuda(img, y_half)
    return _k().ge_vanilla(img, y_half)


def ge_adaptive(img, y_half, logits_half, height, depth_scale, want_logits=None):
    require_cuda(img, y_half, logits_half)
    return _k().ge_adaptive(img, y_half, logits_half, height, depth_scale, want_logits)


def fuse_head(d, pe_mask, y, min_depth):
    require_cuda(d, pe_mask, y)
    return _k().fuse_head(d, pe_mask, y, min_depth)


def silog(pred, gt, eps=1e-3, lam=0.15, max_depth=None, upsample=False):
    require_cuda(pred, gt)
    return _k().silog(pred, gt, eps, lam, max_depth, upsample)


def cross_entropy(logits, target, ignore_index=255):
    require_cuda(logits, target)
    return _k().cross_entropy(logits, target, ignore_index)


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
