"""ctypes binding of libgedepth_sm100.so (include/gedepth.h) + the torch.autograd.Function wrappers
that own tensors around it.  PyTorch is plumbing here: it allocates device memory and provides the
current stream; every number is produced by the sm_100a kernels behind the C ABI.

Fails loudly: a missing / unloadable library or a non-zero status raises RuntimeError; a shape the kernels do not
cover raises NotImplementedError (there is no cuDNN / cuBLAS / ATen fallback in the product).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import torch
from torch.autograd import Function


_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgedepth_sm100.so")

_P, _I, _I64, _F, _D = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double

# name -> argtypes (mirrors include/gedepth.h; tests/test_cabi.py checks the header against this)
SIGNATURES = {
    "ged_ground_plane": [_P, _P, _I64, _I64, _I, _I, _I, _P, _D, _D, _D, _D, _F, _F, _P],
    "ged_pixel_grid": [_P, _P, _I, _I, _I, _I, _P],
    "ged_ge_vanilla_fwd": [_P, _I64, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "ged_ge_vanilla_bwd": [_P, _I64, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "ged_ge_adaptive_fwd": [_P, _I64, _P, _P, _P, _F, _F, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "ged_ge_adaptive_bwd": [_P, _I64, _P, _P, _P, _F, _F, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "ged_fuse_head_fwd": [_P, _P, _P, _P, _P, _F, _I, _I, _I, _I, _I, _P],
    "ged_fuse_head_bwd": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "ged_find_k": [_P, _P, _I64, _P, _I, _I, _I, _D, _I, _P],
    "ged_silog_fwd": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _F, _F, _F, _I, _P],
    "ged_silog_bwd": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _F, _F, _F, _I, _P],
    "ged_ce_fwd": [_P, _P, _P, _P, _I, _I, _I, _I, _F, _P],
    "ged_ce_bwd": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _P],
    "ged_layernorm_fwd": [_P, _P, _P, _P, _P, _P, _I64, _I, _F, _P],
    "ged_layernorm_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _I, _P],
    "ged_winattn_fwd": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _F, _P],
    "ged_winattn_tc_fwd": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _F, _P],
    "ged_winattn_tc_bwd": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _F, _P],
    "ged_winattn_bwd_mma": [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _F, _I, _P],
    "ged_winattn_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _F, _P],
    "ged_gemm_tf32": [_P, _I, _P, _I, _P, _I, _I, _I, _I, _P, _I, _F, _P, _P, _I, _P, _F, C.c_uint, _P, _P],
    "ged_dropout_bwd": [_P, _I64, _P, _P, _I64, _I, _F, C.c_uint, _P, _P],
    "ged_prep_conv_input": [_P, _I, _I, _I, _P, _I, _P, _I, _I, _I, _I64, _I64, _P],
    "ged_upsample_nhwc_bwd": [_P, _I, _P, _I, _I, _I, _I, _I, _I, _P],
    "ged_resize_add_nhwc": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "ged_act_bwd": [_P, _I64, _P, _P, _P, _P, _I, _I64, _I, _I, _F, _P],
    "ged_bn_train_fwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _I, _F, _F, _I, _P],
    "ged_bn_train_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _I, _I, _P],
    "ged_patchify": [_P, _I64, _P, _I, _I, _I, _I, _I, _P],
    "ged_merge_patches": [_P, _P, _I, _I, _I, _I, _I, _P],
    "ged_im2col": [_P, _I64, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    "ged_clamp_resize": [_P, _P, _I, _I, _I, _I, _I, _F, _F, _P],
    "ged_conv3x3_tf32": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P, _I, _F, _P],
    "ged_msda_fwd": [_P, _P, _I, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "ged_msda_bwd": [_P, _P, _I, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "ged_msda_sort_queries": [_P, _I, _I, _I, _P, _P, _I64, _P],
    "ged_msda_tile_fwd": [_P, _P, _I, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "ged_msda_tile_bwd": [_P, _P, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "ged_msda_tc_fwd": [_P, _P, _P, _I, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "ged_msda_tc_bwd": [_P, _P, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "ged_gemm_tf32_bt": [_P, _I, _P, _I, _P, _I, _I, _I, _I, _P],
    "ged_gemm_tf32_bt_acc": [_P, _I, _P, _I, _P, _I, _I, _I, _I, _P, _I, _P],
    "ged_conv3x3_small_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _F, _P],
    "ged_linear_small_fwd": [_P, _P, _P, _P, _I64, _I, _I, _I, _P],
    "ged_linear_small_bwd": [_P, _P, _P, _P, _P, _P, _P, _I64, _I, _I, _I, _P],
    "ged_add_pos_fwd": [_P, _P, _P, _P, _P, _I, _I, _I, _P],
    "ged_add_pos_bwd": [_P, _P, _P, _P, _I, _I, _I, _P],
    "ged_conv3x3_dx_tf32": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "ged_set_gemm_precision": [_I],
    "ged_set_gemm_wide_tiles": [_I],
    "ged_set_gemm_pair": [_I],
    "ged_set_gemm_pair_dw": [_I],
    "ged_set_gemm_a_tmem": [_I],
    "ged_set_ge_x2": [_I],
    "ged_set_layernorm_reg": [_I],
    "ged_set_layout_rows": [_I],
    "ged_set_msda_variant": [_I],
    "ged_gemm_dw_tf32": [_P, _I, _P, _I, _P, _I, _I, _I, _I64, _I64, _I, _P, _I, _P],
    "ged_depth_metrics": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _F, _F, _P],
    "ged_tta_merge": [_P, _P, _P, _I, _I, _I, _P],
    "ged_rgb_crop_normalize": [_P, _I, _I, _I, _I, _I, _I, _P, _P, _P, _I, _I, _P],
    "ged_aug_u8_to_planes": [_P, _P, _I, _I, _P],
    "ged_aug_resize_pad": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P],
    "ged_aug_warp_crop_norm": [_P, _P, _P, _I, _I, _P, _I, _I, _I, _I, _I, _I, _I, _F, _F, _P, _P, _P, _F, _P, _P, _P, _P],
    "ged_aug_frame_bytes": [],
    "ged_aug_pack_frame": [_P, _I, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P, _I, _I, _I, _I, _I, _I,
                           _I, _F, _F, _P, _P, _P, _F],
    "ged_aug_train_batch": [_P, _I, _P, _P],
    "ged_sumsq": [_P, _I64, _P, _P],
    "ged_adamw_step": [_P, _P, _P, _P, _P, _I64, _P, _F, _F, _F, _F, _F, _F, _F, _I, _P, _P, _P],
}

_lib = None
_load_error: Optional[str] = None

# ops whose sm_100a kernel is wired in (ops.py consults has()); everything else is a library call
NATIVE_OPS = {"ground_plane", "ge_vanilla", "ge_adaptive", "fuse_head", "silog", "cross_entropy",
              "layer_norm", "window_attention", "msda_sample", "linear", "conv2d", "conv_bn_act",
              "conv2d_cat", "resize_add", "batch_norm", "patch_embed", "merge_patches", "clamp_resize", "find_k", "depth_metrics", "tta_merge", "adamw"}


def load():
    """dlopen the extension once.  Raises RuntimeError when it is missing (no fallback)."""
    global _lib, _load_error
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        _load_error = (f"{LIB_PATH} is missing: build it with `python -m gedepth_b200.build` "
                       "(__graft_entry__.build()). gedepth_b200 has no fallback path.")
        raise RuntimeError(_load_error)
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.argtypes = argtypes
        fn.restype = C.c_int
    lib.ged_version.restype = C.c_int
    lib.ged_arch.restype = C.c_char_p
    _lib = lib
    if os.environ.get("GEDEPTH_GEMM_PASSES"):
        lib.ged_set_gemm_precision(int(os.environ["GEDEPTH_GEMM_PASSES"]))
    if os.environ.get("GEDEPTH_GEMM_PAIR"):
        lib.ged_set_gemm_pair(int(os.environ["GEDEPTH_GEMM_PAIR"]))
    if os.environ.get("GEDEPTH_MSDA_VARIANT"):
        lib.ged_set_msda_variant(int(os.environ["GEDEPTH_MSDA_VARIANT"]))
    return lib


def set_gemm_precision(passes: int) -> int:
    """3 = error-compensated 3xTF32 (fp32-accurate, default), 1 = single-pass TF32.  Returns the previous mode."""
    return load().ged_set_gemm_precision(int(passes))


# GEMM arithmetic of the backward (dX) kernels.  The forward default (3xTF32) is what the depth-map parity bar
# needs; gradients are computed in single-pass TF32 exactly as the reference's PyTorch 1.8 does on Ampere+
# (torch.backends.{cuda.matmul,cudnn}.allow_tf32 default True there).  Set to 3 for fp32-accurate gradients.
BACKWARD_PASSES = int(os.environ.get("GEDEPTH_BWD_GEMM_PASSES", "1"))


class _bwd_precision:
    def __enter__(self):
        self.prev = load().ged_set_gemm_precision(BACKWARD_PASSES)

    def __exit__(self, *a):
        load().ged_set_gemm_precision(self.prev)


def has(name: str) -> bool:
    load()
    return name in NATIVE_OPS


_ERR = {-1: "GED_ERR_ARG", -2: "GED_ERR_SHAPE", -3: "GED_ERR_ALIGN", -4: "GED_ERR_LAUNCH", -5: "GED_ERR_WORKSPACE"}
LAUNCHES = 0      # number of C-ABI calls issued (bench.py reports it as gpu_launches)


def _call(name, *args):
    global LAUNCHES
    rc = getattr(load(), name)(*args)
    LAUNCHES += 1
    if rc != 0:
        raise RuntimeError(f"{name} failed: {_ERR.get(rc, rc)}")


def _sink(p):
    """The running-gradient buffer of a parameter that lives in train.FlatArena (marked ``_ged_sink``), or None.
    Backward kernels that ACCUMULATE (weight-gradient GEMM, bias/LayerNorm/relative-position column sums) then add
    straight into the arena and return no gradient to autograd: no per-parameter zero-fill, no `grad += dw` pass."""
    if p is None or not getattr(p, "_ged_sink", False) or p.grad is None or not torch.is_grad_enabled():
        return None
    return p.grad


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _rows(g: torch.Tensor, N: int) -> torch.Tensor:
    """g (..., N) as a 2-D [rows, N] tensor with unit inner stride and a 16-byte aligned, uniform row pitch - a VIEW when
    g is e.g. a channel slice of a wider NHWC gradient (the kernels take the pitch), a dense copy otherwise."""
    if g.dtype == torch.float32 and g.dim() >= 2 and g.stride(-1) == 1:
        try:
            g2 = g.view(-1, N)
            if g2.stride(0) % 4 == 0 and g2.data_ptr() % 16 == 0:
                return g2
        except RuntimeError:
            pass
    return _f32c(g).reshape(-1, N)


def _f32c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


# =============================================================================================
# ground embedding
# =============================================================================================
def ground_plane(coef, H, W, device, batch, u0, v0, depth_scale, clamp_max, su, sv):
    out = torch.empty(batch, 2, H, W, dtype=torch.float32, device=device)
    c = (C.c_double * 4)(*[float(x) for x in coef])
    with torch.cuda.device(out.device):
        _call("ged_ground_plane", _p(out[:, 0]), _p(out[:, 1]), 2 * H * W, 2 * H * W, batch, H, W, c,
              float(u0), float(v0), float(su), float(sv), float(depth_scale), float(clamp_max), _stream())
    return out


def ground_plane_into(img: torch.Tensor, coef, u0=0, v0=0, depth_scale=200.0, clamp_max=200.0, su=1.0, sv=1.0):
    """Write channels 3 and 4 of a (B,5,H,W) input batch in place (what the loader would attach)."""
    B, Cc, H, W = img.shape
    assert Cc == 5 and img.is_contiguous() and img.dtype == torch.float32
    c = (C.c_double * 4)(*[float(x) for x in coef])
    with torch.cuda.device(img.device):
        _call("ged_ground_plane", _p(img[:, 3]), _p(img[:, 4]), 5 * H * W, 5 * H * W, B, H, W, c, float(u0),
              float(v0), float(su), float(sv), float(depth_scale), float(clamp_max), _stream())
    return img


def pixel_grid(H, W, device, u0=0, v0=0):
    u = torch.empty(H, W, dtype=torch.int64, device=device)
    v = torch.empty(H, W, dtype=torch.int64, device=device)
    with torch.cuda.device(u.device):
        _call("ged_pixel_grid", _p(u), _p(v), H, W, int(u0), int(v0), _stream())
    return u, v


def _plane(img: torch.Tensor, ch: int):
    """(pointer tensor, batch stride) of one channel plane of an NCHW batch without copying."""
    B, Cc, H, W = img.shape
    if img.dtype == torch.float32 and img.stride(3) == 1 and img.stride(2) == W:
        return img[:, ch], img.stride(0)
    pl = img[:, ch].float().contiguous()
    return pl, H * W


class _GEVanilla(Function):
    @staticmethod
    def forward(ctx, img, y_half):
        B, _, H, W = img.shape
        yh = _f32c(y_half)
        h2, w2 = yh.shape[2], yh.shape[3]
        pe, bs = _plane(img, 3)
        y = torch.empty(B, 1, H, W, dtype=torch.float32, device=img.device)
        pm = torch.empty_like(y)
        _call("ged_ge_vanilla_fwd", _p(pe), bs, _p(yh), _p(y), _p(pm), B, H, W, h2, w2, _stream())
        ctx.save_for_backward(img)
        ctx.dims = (B, H, W, h2, w2)
        return y, pm

    @staticmethod
    def backward(ctx, g_y, g_pm):
        (img,) = ctx.saved_tensors
        B, H, W, h2, w2 = ctx.dims
        pe, bs = _plane(img, 3)
        g_half = torch.empty(B, 1, h2, w2, dtype=torch.float32, device=img.device)
        _call("ged_ge_vanilla_bwd", _p(pe), bs, _p(None if g_y is None else _f32c(g_y)),
              _p(None if g_pm is None else _f32c(g_pm)), _p(g_half), B, H, W, h2, w2, _stream())
        return None, g_half


def ge_vanilla(img, y_half):
    return _GEVanilla.apply(img, y_half)


class _GEAdaptive(Function):
    @staticmethod
    def forward(ctx, img, y_half, logits_half, height, height_scalar, depth_scale, want_logits):
        B, _, H, W = img.shape
        yh, lh = _f32c(y_half), _f32c(logits_half)
        h2, w2 = yh.shape[2], yh.shape[3]
        pe, bs = _plane(img, 4)
        y = torch.empty(B, 1, H, W, dtype=torch.float32, device=img.device)
        pm = torch.empty_like(y)
        lf = torch.empty(B, 11, H, W, dtype=torch.float32, device=img.device) if want_logits else None
        ht = None if height is None else _f32c(height.to(img.device))
        _call("ged_ge_adaptive_fwd", _p(pe), bs, _p(yh), _p(lh), _p(ht), float(height_scalar),
              float(depth_scale), _p(y), _p(pm), _p(lf), B, H, W, h2, w2, _stream())
        ctx.save_for_backward(img, yh, lh, ht if ht is not None else torch.empty(0, device=img.device))
        ctx.cfg = (B, H, W, h2, w2, float(height_scalar), float(depth_scale), ht is not None)
        if lf is None:
            lf = torch.empty(0, device=img.device)
        return y, pm, lf

    @staticmethod
    def backward(ctx, g_y, g_pm, g_lf):
        img, yh, lh, ht = ctx.saved_tensors
        B, H, W, h2, w2, hs, ds, has_h = ctx.cfg
        pe, bs = _plane(img, 4)
        g_yh = torch.empty_like(yh)
        g_lh = torch.empty_like(lh)
        if g_lf is not None and g_lf.numel() == 0:
            g_lf = None
        _call("ged_ge_adaptive_bwd", _p(pe), bs, _p(yh), _p(lh), _p(ht if has_h else None), hs, ds,
              _p(None if g_y is None else _f32c(g_y)), _p(None if g_pm is None else _f32c(g_pm)),
              _p(None if g_lf is None else _f32c(g_lf)), _p(g_yh), _p(g_lh), B, H, W, h2, w2, _stream())
        return None, g_yh, g_lh, None, None, None, None


def ge_adaptive(img, y_half, logits_half, height, depth_scale, want_logits=None):
    """want_logits: also return the full-resolution logits (the CE loss input); default: whenever autograd records."""
    if torch.is_tensor(height):
        ht, hs = height.reshape(-1), 0.0
    else:
        ht, hs = None, float(height)
    want_logits = torch.is_grad_enabled() if want_logits is None else bool(want_logits)
    y, pm, lf = _GEAdaptive.apply(img, y_half, logits_half, ht, hs, depth_scale, want_logits)
    return y, pm, (lf if want_logits else None)


class _FuseHead(Function):
    @staticmethod
    def forward(ctx, d, pe_mask, y, min_depth):
        d, pe_mask, y = _f32c(d), _f32c(pe_mask), _f32c(y)
        B, _, h2, w2 = d.shape
        H, W = y.shape[2], y.shape[3]
        out, y_h = torch.empty_like(d), torch.empty_like(d)
        _call("ged_fuse_head_fwd", _p(d), _p(pe_mask), _p(y), _p(out), _p(y_h), float(min_depth), B, H, W, h2,
              w2, _stream())
        ctx.save_for_backward(d, y_h)
        ctx.dims = (B, H, W, h2, w2)
        return out, y_h

    @staticmethod
    def backward(ctx, g_out, g_yh):
        d, y_h = ctx.saved_tensors
        B, H, W, h2, w2 = ctx.dims
        g_d = torch.empty_like(d)
        g_pm = torch.empty(B, 1, H, W, dtype=torch.float32, device=d.device)
        g_y = torch.empty_like(g_pm)
        g_out = _f32c(g_out) if g_out is not None else torch.zeros_like(d)
        _call("ged_fuse_head_bwd", _p(g_out), _p(None if g_yh is None else _f32c(g_yh)), _p(d), _p(y_h),
              _p(g_d), _p(g_pm), _p(g_y), B, H, W, h2, w2, _stream())
        return g_d, g_pm, g_y, None


def fuse_head(d, pe_mask, y, min_depth):
    return _FuseHead.apply(d, pe_mask, y, min_depth)


def find_k(gt, pe, h, truncate=False):
    gt = _f32c(gt)
    B = gt.shape[0]
    H, W = gt.shape[-2], gt.shape[-1]
    pe = _f32c(pe)
    bs = 0 if pe.numel() == H * W else H * W
    out = torch.empty(B, H, W, dtype=torch.float32, device=gt.device)
    _call("ged_find_k", _p(gt), _p(pe), bs, _p(out), B, H, W, float(h), int(bool(truncate)), _stream())
    return out


# =============================================================================================
# losses
# =============================================================================================
class _SiLog(Function):
    @staticmethod
    def forward(ctx, pred, gt, eps, lam, max_depth, upsample):
        pred, gt = _f32c(pred), _f32c(gt)
        B, _, H, W = gt.shape
        hp, wp = pred.shape[2], pred.shape[3]
        stats = torch.empty(8, dtype=torch.float64, device=gt.device)
        loss = torch.empty((), dtype=torch.float32, device=gt.device)
        md = float(max_depth) if max_depth is not None else 0.0
        _call("ged_silog_fwd", _p(pred), _p(gt), _p(stats), _p(loss), B, H, W, hp, wp, float(eps), float(lam),
              md, int(upsample), _stream())
        ctx.save_for_backward(pred, gt, stats)
        ctx.cfg = (B, H, W, hp, wp, float(eps), float(lam), md, int(upsample))
        return loss

    @staticmethod
    def backward(ctx, g):
        pred, gt, stats = ctx.saved_tensors
        B, H, W, hp, wp, eps, lam, md, up = ctx.cfg
        g_pred = torch.empty_like(pred)
        _call("ged_silog_bwd", _p(pred), _p(gt), _p(stats), _p(_f32c(g)), _p(g_pred), B, H, W, hp, wp, eps, lam,
              md, up, _stream())
        return g_pred, None, None, None, None, None


def silog(pred, gt, eps, lam, max_depth, upsample):
    return _SiLog.apply(pred, gt, eps, lam, max_depth, bool(upsample))


class _CE(Function):
    @staticmethod
    def forward(ctx, logits, target, ignore_index):
        logits, target = _f32c(logits), _f32c(target)
        B, Cc, H, W = logits.shape
        stats = torch.empty(8, dtype=torch.float64, device=logits.device)
        loss = torch.empty((), dtype=torch.float32, device=logits.device)
        _call("ged_ce_fwd", _p(logits), _p(target), _p(stats), _p(loss), B, Cc, H, W, float(ignore_index), _stream())
        ctx.save_for_backward(logits, target, stats)
        ctx.ignore = float(ignore_index)
        return loss

    @staticmethod
    def backward(ctx, g):
        logits, target, stats = ctx.saved_tensors
        B, Cc, H, W = logits.shape
        gl = torch.empty_like(logits)
        _call("ged_ce_bwd", _p(logits), _p(target), _p(stats), _p(_f32c(g)), _p(gl), B, Cc, H, W, ctx.ignore, _stream())
        return gl, None, None


def cross_entropy(logits, target, ignore_index=255):
    if logits.shape[1] != 11:
        raise NotImplementedError("cross_entropy: the kernel covers the 11 slope classes of the GE path (encoder_decoder.py:68)")
    return _CE.apply(logits, target, ignore_index)


# =============================================================================================
# Swin pieces
# =============================================================================================
class _LayerNorm(Function):
    """y = LN(x).  fork=True additionally hands x back as a second output: the caller uses THAT as the block's identity,
    so both gradient paths of x arrive here and the residual one is added inside the dx kernel (no autograd sum pass)."""

    @staticmethod
    def forward(ctx, x, w, b, eps, w_sink=None, b_sink=None, fork=False):
        ctx.sinks = (w_sink, b_sink) if (w_sink is not None and b_sink is not None) else None
        xc = _f32c(x)
        Cc = xc.shape[-1]
        rows = xc.numel() // Cc
        y = torch.empty_like(xc)
        mean = torch.empty(rows, dtype=torch.float32, device=x.device)
        rstd = torch.empty_like(mean)
        _call("ged_layernorm_fwd", _p(xc), _p(w), _p(b), _p(y), _p(mean), _p(rstd), rows, Cc, float(eps), _stream())
        ctx.save_for_backward(xc, w, mean, rstd)
        ctx.fork = fork
        return (y, x) if fork else y

    @staticmethod
    def backward(ctx, g, g_id=None):
        xc, w, mean, rstd = ctx.saved_tensors
        Cc = xc.shape[-1]
        rows = xc.numel() // Cc
        g = _f32c(g)
        g_add = None if g_id is None else _f32c(g_id)
        dx = torch.empty_like(xc)
        need_w = ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        tail = (None,) * 4
        if need_w and ctx.sinks is not None:
            _call("ged_layernorm_bwd", _p(g), _p(xc), _p(w), _p(mean), _p(rstd), _p(g_add), _p(dx), _p(ctx.sinks[0]),
                  _p(ctx.sinks[1]), rows, Cc, _stream())
            return (dx, None, None) + tail
        dw = torch.zeros(Cc, dtype=torch.float32, device=xc.device) if need_w else None
        db = torch.zeros_like(dw) if need_w else None
        _call("ged_layernorm_bwd", _p(g), _p(xc), _p(w), _p(mean), _p(rstd), _p(g_add), _p(dx), _p(dw), _p(db), rows, Cc,
              _stream())
        return (dx, dw, db) + tail


def layer_norm(x, w, b, eps):
    if x.shape[-1] % 4:
        raise NotImplementedError("layer_norm: channel count must be a multiple of 4")
    return _LayerNorm.apply(x, w, b, eps, _sink(w), _sink(b), False)


def layer_norm_fork(x, w, b, eps):
    """(LN(x), x_identity): use x_identity as the residual of the block that follows (see _LayerNorm)."""
    if x.shape[-1] % 4 or not torch.is_grad_enabled() or not x.requires_grad:
        return layer_norm(x, w, b, eps), x
    return _LayerNorm.apply(x, w, b, eps, _sink(w), _sink(b), True)


# Forward of the 49 x 49 core on tcgen05 (csrc/winattn_tc.cu, 3xTF32) unless GEDEPTH_WINATTN_TC=0.  The backward has a
# tcgen05 implementation too (one pass TF32 like the other backward GEMMs; GEDEPTH_WINATTN_TC_BWD=1), parity-tested but
# NOT the default: measured at Swin-L / B = 16 / 352 x 1120 it takes 59.8 ms per step against 38.3 ms for the fp32 SIMT
# kernel of csrc/winattn.cu (tools/ab_winattn_bwd.py) - its 220 KB of operand tiles and 255 registers allow one CTA per SM
# and the serial stage -> MMA -> softmax/dS -> MMA -> store chain of a work item leaves the SM 91 % idle (ncu:
# profiles/r02_ncu_winattn_tc_bwd.csv).
WINATTN_TC = os.environ.get("GEDEPTH_WINATTN_TC", "1") != "0"
WINATTN_TC_BWD = os.environ.get("GEDEPTH_WINATTN_TC_BWD", "0") != "0"
# Default backward (one-pass TF32 arithmetic): the SIMT kernel's CTA-per-(window, head) structure with its five products on
# warp-level mma.sync (csrc/winattn.cu::winattn_bwd_mma_kernel): 19.8 ms per step against 38.3 (SIMT) and 53-60 (tcgen05);
# GEDEPTH_WINATTN_BWD_MMA=0 or GEDEPTH_BWD_GEMM_PASSES=3 select the fp32 SIMT kernel.
WINATTN_BWD_MMA = os.environ.get("GEDEPTH_WINATTN_BWD_MMA", "1") != "0"


def _standard_rel_index(index: torch.Tensor) -> bool:
    """True when `index` is Swin's relative-position index (dy + 6) * 13 + (dx + 6) (depthformer_swin.py:168-172), which the
    tensor-core kernel evaluates in closed form.  Checked once per buffer (device -> host copy on first use)."""
    # cached on the tensor object itself (never keyed by address: freed addresses are reused).  The buffer is filled once at
    # construction (depthformer_swin.py:168-172); in-place copies of the same content (state_dict loads, Trainer.capture()'s
    # snapshot restore) bump its version, so the version is deliberately not part of the key.
    cached = getattr(index, "_ged_std_index", None)
    if cached is not None:
        return cached[1]
    c = torch.arange(7)
    yy, xx = torch.meshgrid(c, c, indexing="ij")
    y, x = yy.reshape(-1), xx.reshape(-1)
    want = (y[:, None] - y[None, :] + 6) * 13 + (x[:, None] - x[None, :] + 6)
    std = tuple(index.shape) == (49, 49) and bool(torch.equal(index.detach().cpu().long(), want))
    try:
        index._ged_std_index = (index._version, std)
    except Exception:
        pass
    return std


class _WinAttn(Function):
    @staticmethod
    def forward(ctx, qkv, qkv_bias, table, index, H, W, nH, ws, shift, scale, bias_sink=None, table_sink=None):
        ctx.sinks = (bias_sink, table_sink)
        qkv = _f32c(qkv)
        B, Lt, C3 = qkv.shape
        Cc = C3 // 3
        ctx_out = torch.empty(B, Lt, Cc, dtype=torch.float32, device=qkv.device)
        idx = index if index.dtype == torch.int64 and index.is_contiguous() else index.long().contiguous()
        table_c = _f32c(table)
        if WINATTN_TC and _standard_rel_index(index):
            _call("ged_winattn_tc_fwd", _p(qkv), _p(qkv_bias), _p(table_c), _p(ctx_out), B, H, W, Cc, nH, ws, shift,
                  float(scale), _stream())
        else:
            _call("ged_winattn_fwd", _p(qkv), _p(qkv_bias), _p(table_c), _p(idx), _p(ctx_out), B, H, W, Cc, nH, ws,
                  shift, float(scale), _stream())
        ctx.save_for_backward(qkv, qkv_bias if qkv_bias is not None else torch.empty(0, device=qkv.device),
                              table_c, idx)
        ctx.cfg = (B, H, W, Cc, nH, ws, shift, float(scale), qkv_bias is not None)
        ctx.std_index = _standard_rel_index(index)
        return ctx_out

    @staticmethod
    def backward(ctx, g):
        qkv, bias, table, idx = ctx.saved_tensors
        B, H, W, Cc, nH, ws, shift, scale, has_bias = ctx.cfg
        g = _f32c(g)
        g_qkv = torch.empty_like(qkv)
        bias_sink, table_sink = ctx.sinks
        g_table = table_sink if table_sink is not None else torch.zeros_like(table)
        g_bias = None
        if has_bias:
            g_bias = bias_sink if bias_sink is not None else torch.zeros(3 * Cc, dtype=torch.float32, device=qkv.device)
        if WINATTN_TC and WINATTN_TC_BWD and BACKWARD_PASSES == 1 and ctx.std_index:
            _call("ged_winattn_tc_bwd", _p(qkv), _p(bias if has_bias else None), _p(table), _p(g), _p(g_qkv),
                  _p(g_bias), _p(g_table), B, H, W, Cc, nH, ws, shift, scale, _stream())
        elif WINATTN_BWD_MMA and BACKWARD_PASSES == 1:
            _call("ged_winattn_bwd_mma", _p(qkv), _p(bias if has_bias else None), _p(table), _p(idx), _p(g), _p(g_qkv),
                  _p(g_bias), _p(g_table), B, H, W, Cc, nH, ws, shift, scale, int(ctx.std_index), _stream())
        else:
            _call("ged_winattn_bwd", _p(qkv), _p(bias if has_bias else None), _p(table), _p(idx), _p(g), _p(g_qkv),
                  _p(g_bias), _p(g_table), B, H, W, Cc, nH, ws, shift, scale, _stream())
        return (g_qkv, None if bias_sink is not None else g_bias, None if table_sink is not None else g_table,
                None, None, None, None, None, None, None, None, None)


def window_attention(qkv, qkv_bias, table, index, hw, nH, ws, shift, scale):
    ts = _sink(table)
    ts = ts if (ts is not None and ts.is_contiguous() and table.is_contiguous()) else None
    return _WinAttn.apply(qkv, qkv_bias, table, index, int(hw[0]), int(hw[1]), nH, ws, shift, scale, _sink(qkv_bias), ts)


# =============================================================================================
# tcgen05 GEMM: linear / 1x1 conv
# =============================================================================================
_ACT = {None: 0, "relu": 1, "leaky_relu": 2, "gelu": 3, "sigmoid": 4}


# Dropout inside the GEMM epilogue: (p, seed, step) with `step` an optional int32 device tensor mixed into the seed
# (train.Trainer registers its step counter here so a replayed CUDA graph draws a new mask every step).
RNG_STEP: Optional[torch.Tensor] = None
_DROP_CALLS = 0


def next_dropout_seed() -> int:
    """Per call site, mixed with torch's seed and the data-parallel rank: ranks draw different masks, torch.manual_seed
    changes the stream (the reference's nn.Dropout uses the per-rank torch generator)."""
    global _DROP_CALLS
    _DROP_CALLS += 1
    rank = 0
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        rank = torch.distributed.get_rank()
    base = (torch.initial_seed() * 0x9E3779B1 + rank * 0x85EBCA77) & 0xFFFFFFFF
    return (0x2545F491 * _DROP_CALLS + 0x1234567 + base) & 0xFFFFFFFF


def gemm(a2d: torch.Tensor, w: torch.Tensor, bias=None, act=None, slope=0.01, residual=None,
         row_scale=None, rows_per_batch=1, out: Optional[torch.Tensor] = None,
         pre_out: Optional[torch.Tensor] = None, dropout=None) -> torch.Tensor:
    """out[M,N] = epi(a2d[M,K] @ w[N,K]^T).  a2d / w: last dim contiguous, 16B-aligned pitches.
    dropout = (p, seed, step_tensor|None): applied after the activation, before row scale / residual."""
    M, K = a2d.shape
    N = w.shape[0]
    assert w.shape[1] == K and a2d.stride(1) == 1 and w.stride(1) == 1
    if out is None:
        out = torch.empty(M, N, dtype=torch.float32, device=a2d.device)
    dp, dseed, dstep = dropout if dropout is not None else (0.0, 0, None)
    _call("ged_gemm_tf32", _p(a2d), a2d.stride(0), _p(w), w.stride(0), _p(out), out.stride(0), M, N, K,
          _p(bias), _ACT[act], float(slope), _p(residual), _p(row_scale), int(rows_per_batch), _p(pre_out),
          float(dp), int(dseed), _p(dstep), _stream())
    return out


# Weight gradients: 1 = tcgen05 on the operands as stored (MN-major UMMA descriptors, split-K partials accumulated
# with vector atomics), 0 = library (cuBLAS / cuDNN wgrad; kept as an A/B switch for the tests).
DW_MODE = int(os.environ.get("GEDEPTH_DW_MODE", "1"))


def set_ge_x2(on) -> int:
    """Closed-form x2 ground-embedding kernels: 1 = on, TMA-staged forward where W % 8 == 0 (default), 2 = on with per-thread
    asynchronous copies only, 0 = generic bilinear kernels; returns the previous setting."""
    return load().ged_set_ge_x2(int(on))


def set_layernorm_reg(on: bool) -> int:
    """LayerNorm rows of <= 768 channels held in registers, dw / db fused into the dx kernel (default on); returns the
    previous setting."""
    return load().ged_set_layernorm_reg(int(bool(on)))


def set_layout_rows(on: bool) -> int:
    """Row-structured prep_conv_input / upsample adjoint kernels (default on); returns the previous setting."""
    return load().ged_set_layout_rows(int(bool(on)))


def set_gemm_pair(on) -> int:
    """CTA-pair (cta_group::2) GEMM kernel for large problems: 2 = both arithmetic modes (default), 1 = 3xTF32 only,
    0 = off; returns the previous setting."""
    return load().ged_set_gemm_pair(int(on))


def set_gemm_a_tmem(on: bool) -> int:
    """3xTF32 single-CTA kernels with the A operand's hi / lo in tensor memory (default on); returns the
    previous setting."""
    return load().ged_set_gemm_a_tmem(int(bool(on)))


def set_gemm_pair_dw(on: bool) -> int:
    """Weight-gradient GEMMs on the CTA-pair kernel where the output has >= 256 rows (default on); returns the previous setting."""
    return load().ged_set_gemm_pair_dw(int(bool(on)))


def set_gemm_wide_tiles(on: bool) -> int:
    return load().ged_set_gemm_wide_tiles(int(bool(on)))


# Weight-gradient GEMMs have no consumer before the optimizer, so train.Trainer can move them off the backward's
# critical path: DW_SIDE = {"stream": side stream, "keep": [operands kept alive until the join]} (GEDEPTH_DW_STREAM=1).
DW_SIDE = None


def gemm_dw(g2d: torch.Tensor, x2d: torch.Tensor, out: Optional[torch.Tensor] = None, tap_off=None) -> torch.Tensor:
    """out[n, t, k] += sum_p g2d[p, n] * x2d[p + tap_off[t], k]  (rows of x2d outside the tensor count as zero).
    g2d [P, N], x2d [Px, K] with unit inner stride.  out: [N, K] (no taps) or [N, T, K]; allocated zeroed if None."""
    P, N = g2d.shape
    Px, K = x2d.shape
    T = 1 if tap_off is None else len(tap_off)
    sunk = out is not None
    if out is None:
        out = torch.zeros((N, K) if tap_off is None else (N, T, K), dtype=torch.float32, device=g2d.device)
    taps = None if tap_off is None else (C.c_int * T)(*[int(t) for t in tap_off])
    assert g2d.stride(1) == 1 and x2d.stride(1) == 1
    side = DW_SIDE
    if side is not None and sunk:
        # accumulating into the arena: fork to the side stream (joined by Trainer.step before the all-reduce)
        side["stream"].wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side["stream"]):
            _call("ged_gemm_dw_tf32", _p(g2d), g2d.stride(0), _p(x2d), x2d.stride(0), _p(out), T * K, N, K, P, Px, T, taps,
                  K, _stream())
        side["keep"].append((g2d, x2d))
        return out
    _call("ged_gemm_dw_tf32", _p(g2d), g2d.stride(0), _p(x2d), x2d.stride(0), _p(out), T * K, N, K, P, Px, T, taps, K,
          _stream())
    return out


def gemm_bt(a2d: torch.Tensor, wt: torch.Tensor, residual: Optional[torch.Tensor] = None,
            out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """a2d[M,K] @ wt[K,N] (+ residual[M,N]) with wt read in place (the dX GEMM: grad_output @ weight).  `out` may be
    the residual itself: the product is then accumulated in place (gradient fan-in summed in the epilogue)."""
    M, K = a2d.shape
    N = wt.shape[1]
    assert wt.shape[0] == K and a2d.stride(1) == 1 and wt.stride(1) == 1
    if out is None:
        out = torch.empty(M, N, dtype=torch.float32, device=a2d.device)
    if residual is None:
        _call("ged_gemm_tf32_bt", _p(a2d), a2d.stride(0), _p(wt), wt.stride(0), _p(out), N, M, N, K, _stream())
    else:
        # the residual may be a channel slice of a wider gradient (its own row pitch); the output is dense
        assert residual.shape == (M, N) and residual.stride(1) == 1 and residual.stride(0) % 4 == 0 and out.is_contiguous()
        _call("ged_gemm_tf32_bt_acc", _p(a2d), a2d.stride(0), _p(wt), wt.stride(0), _p(out), N, M, N, K, _p(residual),
              residual.stride(0), _stream())
    return out


def conv3x3_dx(gzp: torch.Tensor, wk: torch.Tensor) -> torch.Tensor:
    """gzp (B,H+2,W+2,Cout) zero-bordered dY, wk [Cout,3,3,Cin] forward weights -> dX (B,H,W,Cin)."""
    B, Hp, Wp, Cout = gzp.shape
    Cin = wk.shape[3]
    dx = torch.empty(B, Hp - 2, Wp - 2, Cin, dtype=torch.float32, device=gzp.device)
    _call("ged_conv3x3_dx_tf32", _p(gzp), _p(wk), _p(dx), Cin, B, Hp - 2, Wp - 2, Cin, Cout, _stream())
    return dx


def _dw_ok(N, K, *tensors):
    return DW_MODE == 1 and N % 4 == 0 and K % 4 == 0 and N >= 16 and K >= 16 and all(
        t.dtype == torch.float32 and t.data_ptr() % 16 == 0 for t in tensors)


def act_bwd(g2d: torch.Tensor, ref: Optional[torch.Tensor], act, slope=0.01, row_scale=None, rows_per_batch=1,
            want_db=False, db_sink: Optional[torch.Tensor] = None):
    """gz = g * act'(ref) * row_scale and (optionally) db = column sums of gz, in ONE pass.  With no
    activation and no scale gz is g itself and only the column sums are computed."""
    rows, N = g2d.shape
    ident = act is None and row_scale is None
    if ident and not want_db:
        return g2d, None
    gz = g2d if ident else torch.empty(rows, N, dtype=torch.float32, device=g2d.device)
    db = None
    if want_db:
        db = db_sink if db_sink is not None else torch.zeros(N, dtype=torch.float32, device=g2d.device)
    _call("ged_act_bwd", _p(g2d), g2d.stride(0), _p(ref), _p(None if ident else gz), _p(db), _p(row_scale),
          int(rows_per_batch), rows, N, _ACT[act], float(slope), _stream())
    return gz, (None if db_sink is not None else db)


def _gemm_ok(M, N, K, *tensors):
    if K % 4 or N % 4 or N < 16 or K < 32:
        return False
    return all(t is None or (t.dtype == torch.float32 and t.data_ptr() % 16 == 0) for t in tensors)


class _Linear(Function):
    """y = [residual +] [row_scale *] act(x @ w^T + b).  Forward and dX on tcgen05 (GELU keeps its
    pre-activation through the epilogue's second store); the activation derivative, DropPath scale and
    the bias gradient are one fused pass; dX reads the weight in place (MN-major operand); dW = gz^T @ x runs on the
    activations as stored and accumulates into the arena when the parameter lives there."""

    @staticmethod
    def forward(ctx, x, w, b, act, residual, row_scale, w_sink=None, b_sink=None, dropout=None):
        ctx.w_sink, ctx.b_sink, ctx.dropout = w_sink, b_sink, dropout
        K = x.shape[-1]
        x2 = _f32c(x).reshape(-1, K)
        M, N = x2.shape[0], w.shape[0]
        wc = _f32c(w)
        res2 = None if residual is None else _f32c(residual).reshape(M, N)
        rpb = M // x.shape[0] if row_scale is not None else 1
        need_pre = act == "gelu" and (ctx.needs_input_grad[0] or ctx.needs_input_grad[1])
        pre = torch.empty(M, N, dtype=torch.float32, device=x.device) if need_pre else None
        out = gemm(x2, wc, b, act, 0.01, res2, row_scale, rpb, pre_out=pre, dropout=dropout)
        ctx.act, ctx.rpb, ctx.has_res, ctx.has_bias = act, rpb, residual is not None, b is not None
        empty = torch.empty(0, device=x.device)
        # relu/leaky/sigmoid derive from the output - only valid when nothing was added after the activation
        post = out if (act in ("relu", "leaky_relu", "sigmoid") and residual is None and row_scale is None) else empty
        if act in ("relu", "leaky_relu", "sigmoid") and post is empty:
            raise NotImplementedError("activation + residual/row_scale in one linear is only wired for GELU/none")
        ctx.save_for_backward(x2, wc, pre if pre is not None else empty, post,
                              row_scale if row_scale is not None else empty)
        ctx.xshape = x.shape
        return out.reshape(*x.shape[:-1], N)

    @staticmethod
    def backward(ctx, g):
        x2, w, pre, post, row_scale = ctx.saved_tensors
        N, K = w.shape
        g2 = _rows(g, N)
        want_db = ctx.has_bias and ctx.needs_input_grad[2]
        ref = pre if ctx.act == "gelu" else (post if ctx.act is not None else None)
        if ctx.dropout is not None:
            # y = drop(x w^T + b) + residual: re-draw the epilogue's mask (no activation / row scale on this path)
            dp, dseed, dstep = ctx.dropout
            gz = torch.empty(g2.shape[0], N, dtype=torch.float32, device=g2.device)
            db = None
            if want_db:
                db = ctx.b_sink if ctx.b_sink is not None else torch.zeros(N, dtype=torch.float32, device=g2.device)
            _call("ged_dropout_bwd", _p(g2), g2.stride(0), _p(gz), _p(db), g2.shape[0], N, float(dp), int(dseed), _p(dstep),
                  _stream())
            if ctx.b_sink is not None:
                db = None
        else:
            gz, db = act_bwd(g2, ref, ctx.act, 0.01, row_scale if row_scale.numel() else None, ctx.rpb, want_db,
                             ctx.b_sink)
        dx = dw = None
        if ctx.needs_input_grad[0]:
            if not (_gemm_ok(gz.shape[0], K, N, gz, w) and w.stride(1) == 1):
                raise NotImplementedError(f"linear backward: dX GEMM for weight {tuple(w.shape)} is not covered")
            with _bwd_precision():
                dx = gemm_bt(gz, w).reshape(ctx.xshape)      # w [N][K] read in place as the [K'][N'] operand
        if ctx.needs_input_grad[1]:
            if not _dw_ok(N, K, gz, x2):
                raise NotImplementedError(f"linear backward: dW GEMM for weight {tuple(w.shape)} is not covered")
            with _bwd_precision():
                if ctx.w_sink is not None:
                    gemm_dw(gz, x2, out=ctx.w_sink)
                else:
                    dw = gemm_dw(gz, x2)
        dres = g if ctx.has_res else None
        return dx, dw, db, None, dres, None, None, None, None


def linear(x, w, b=None, act=None, residual=None, row_scale=None, dropout_p: float = 0.0):
    """dropout_p > 0 (training): residual + dropout(x w^T + b) with the mask drawn in the GEMM epilogue."""
    K, N = x.shape[-1], w.shape[0]
    M = x.numel() // K
    if not _gemm_ok(M, N, K, x, w, b, residual) or act not in _ACT or K < 16:
        if N <= 4 and K % 4 == 0 and residual is None and row_scale is None and dropout_p == 0 and act in (None, "sigmoid"):
            return linear_small(x, w, b, act)
        raise NotImplementedError(f"linear: x {tuple(x.shape)} w {tuple(w.shape)} act {act} is outside the tcgen05 GEMM's "
                                  "(N % 4 == 0, N >= 16, K % 4 == 0, K >= 32, 16-byte aligned) and the small-N kernel's range")
    ws = _sink(w)
    ws = ws if (ws is not None and ws.is_contiguous()) else None
    dropout = None
    if dropout_p > 0:
        assert act is None and row_scale is None and N % 4 == 0 and M * N < 2 ** 32
        dropout = (float(dropout_p), next_dropout_seed(), RNG_STEP)
    return _Linear.apply(x, w, b, act, residual, row_scale, ws, _sink(b), dropout)


# =============================================================================================
# tcgen05 implicit-GEMM 3x3 conv (NHWC) and 1x1 conv
# =============================================================================================
def _nhwc(x: torch.Tensor) -> torch.Tensor:
    """logical NCHW -> contiguous (B,H,W,C) tensor (a view when x is channels_last)."""
    return _f32c(x.permute(0, 2, 3, 1))


def conv2d_supported(x, w, stride, padding) -> bool:
    Cout, Cin, kh, kw = w.shape
    if stride != 1 or Cin % 32 or x.dtype != torch.float32:
        return False
    if kh == 1 and kw == 1 and padding == 0:
        return Cout >= 16
    return kh == 3 and kw == 3 and padding == 1


def _batch_dense(x: torch.Tensor) -> bool:
    """(B, h, w, C) fp32 whose samples are dense NHWC blocks at a uniform, 16-byte aligned batch stride (e.g. one level's
    slice of the (B, S, C) token tensor): the kernels take the stride, no copy."""
    if x.dtype != torch.float32 or x.dim() != 4:
        return False
    _, h, w, c = x.shape
    return (x.stride(3) == 1 and x.stride(2) == c and x.stride(1) == w * c and x.stride(0) % 4 == 0 and x.stride(0) >= h * w * c
            and x.data_ptr() % 16 == 0)


def prep_conv_input(x0: torch.Tensor, x1: Optional[torch.Tensor], H: int, W: int) -> torch.Tensor:
    """Zero-bordered NHWC input [B,H+2,W+2,C0+C1] = [bilinear(x0 -> HxW, align_corners=True) | x1] in one pass."""
    x0 = x0 if _batch_dense(x0) else _f32c(x0)
    if x1 is not None:
        x1 = x1 if _batch_dense(x1) else _f32c(x1)
    B, h0, w0, C0 = x0.shape
    C1 = 0 if x1 is None else x1.shape[3]
    xp = torch.empty(B, H + 2, W + 2, C0 + C1, dtype=torch.float32, device=x0.device)
    _call("ged_prep_conv_input", _p(x0), C0, h0, w0, _p(x1), C1, _p(xp), B, H, W, x0.stride(0), 0 if x1 is None else x1.stride(0),
          _stream())
    return xp


def conv3x3_padded(xp: torch.Tensor, wk: torch.Tensor, bias, act, slope) -> torch.Tensor:
    """xp (B,H+2,W+2,Cin) zero-bordered, wk [Cout,3,3,Cin] -> (B,H,W,Cout)."""
    B, Hp, Wp, Cin = xp.shape
    H, W, Cout = Hp - 2, Wp - 2, wk.shape[0]
    y = torch.empty(B, H, W, Cout, dtype=torch.float32, device=xp.device)
    _call("ged_conv3x3_tf32", _p(xp), _p(wk), _p(y), Cout, B, H, W, Cin, Cout, _p(bias), _ACT[act], float(slope),
          _stream())
    return y


def conv3x3_raw(x_nhwc: torch.Tensor, wk: torch.Tensor, bias, act, slope) -> torch.Tensor:
    """x (B,H,W,Cin) contiguous, wk [Cout,3,3,Cin] contiguous -> (B,H,W,Cout)."""
    return conv3x3_padded(prep_conv_input(x_nhwc, None, x_nhwc.shape[1], x_nhwc.shape[2]), wk, bias, act, slope)


class _Conv(Function):
    """3x3/s1/p1 or 1x1 conv + bias + activation over [resize(x0) | x1] (x1 optional, resize only when
    x0 is smaller).  Forward and dX on tcgen05 (dX of a 3x3 conv is the 3x3 conv of dY with the flipped,
    transposed kernel, read in place from the forward weights); activation derivative + bias gradient one fused pass;
    dW = nine row-shifted contractions of bordered dY against bordered X on tcgen05.  The narrow heads (Cout in
    {1, 2, 11}) take the SIMT kernels of csrc/small.cu."""

    @staticmethod
    def forward(ctx, x0, x1, w, b, act, slope, w_sink=None, b_sink=None):
        ctx.w_sink, ctx.b_sink = w_sink, b_sink
        Cout, Cin, kh, kw = w.shape
        if kh == 3:      # the staging kernel takes batch-strided sources (a level's slice of the token tensor): no copy
            a0 = x0.permute(0, 2, 3, 1) if _batch_dense(x0.permute(0, 2, 3, 1)) else _nhwc(x0)
            a1 = None if x1 is None else (x1.permute(0, 2, 3, 1) if _batch_dense(x1.permute(0, 2, 3, 1)) else _nhwc(x1))
        else:
            a0 = _nhwc(x0)
            a1 = None if x1 is None else _nhwc(x1)
        B = a0.shape[0]
        H, W = (a0.shape[1], a0.shape[2]) if a1 is None else (a1.shape[1], a1.shape[2])
        if kh == 3:
            xin = prep_conv_input(a0, a1, H, W)                          # padded, concatenated
            y = conv3x3_padded(xin, w.permute(0, 2, 3, 1).contiguous(), b, act, slope)
        else:
            assert a1 is None
            xin = a0
            y = gemm(a0.reshape(-1, Cin), w.reshape(Cout, Cin), b, act, slope).reshape(B, H, W, Cout)
        ctx.save_for_backward(xin, w, y if act is not None else torch.empty(0, device=x0.device))
        ctx.cfg = (act, slope, kh, b is not None, a0.shape, None if a1 is None else a1.shape[3])
        return y.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, g):
        xin, w, y = ctx.saved_tensors
        act, slope, kh, has_bias, shape0, C1 = ctx.cfg
        Cout, Cin = w.shape[0], w.shape[1]
        B, h0, w0, C0 = shape0
        gh = _nhwc(g)
        H, W = gh.shape[1], gh.shape[2]
        want_db = has_bias and ctx.needs_input_grad[3]
        need_dx = ctx.needs_input_grad[0] or (C1 is not None and ctx.needs_input_grad[1])
        need_dw = ctx.needs_input_grad[2]
        db = dxc = dw = None
        if Cout in SMALL_COUT and kh == 3:
            # narrow heads (conv_depth 64->1, convfinal 64->1 / 64->11): SIMT kernels, activation derivative inside
            wk = w.permute(0, 2, 3, 1)
            wk = wk if wk.is_contiguous() else wk.contiguous()
            gz = torch.empty_like(gh)
            dxc = torch.empty(B, H, W, Cin, dtype=torch.float32, device=gh.device) if need_dx else None
            sink = None if ctx.w_sink is None else ctx.w_sink.permute(0, 2, 3, 1)
            dwk = None
            if need_dw:
                dwk = sink if sink is not None else torch.zeros(Cout, 3, 3, Cin, dtype=torch.float32, device=gh.device)
            dbb = None
            if want_db:
                dbb = ctx.b_sink if ctx.b_sink is not None else torch.zeros(Cout, dtype=torch.float32, device=gh.device)
            _call("ged_conv3x3_small_bwd", _p(gh), _p(y if act is not None else None), _p(gz), _p(xin), _p(wk), _p(dxc),
                  _p(dwk), _p(dbb), B, H, W, Cin, Cout, _ACT[act], float(slope), _stream())
            dw = None if (sink is not None or dwk is None) else dwk.permute(0, 3, 1, 2)
            db = None if ctx.b_sink is not None else dbb
        else:
            if Cout % 4 or (kh == 3 and need_dx and Cout % 32) or (need_dw and not _dw_ok(Cout, Cin, gh, xin)):
                raise NotImplementedError(f"conv backward: weight {tuple(w.shape)} is outside the tcgen05 dX / dW kernels' range")
            gz2, db = act_bwd(gh.reshape(-1, Cout), y.reshape(-1, Cout) if act is not None else None, act, slope,
                              None, 1, want_db, ctx.b_sink)
            gz = gz2.reshape(B, H, W, Cout)
            gzp = prep_conv_input(gz, None, H, W) if kh == 3 else None      # zero-bordered dY, shared by dX and dW
            if need_dx:
                with _bwd_precision():
                    if kh == 3:
                        wk = w.permute(0, 2, 3, 1)                           # [Cout,3,3,Cin]: a view inside the arena
                        dxc = conv3x3_dx(gzp, wk if wk.is_contiguous() else wk.contiguous())   # (B,H,W,Cin)
                    else:
                        if not _gemm_ok(B * H * W, Cin, Cout, gz):
                            raise NotImplementedError(f"1x1 conv backward: weight {tuple(w.shape)} not covered")
                        dxc = gemm_bt(gz.reshape(-1, Cout), w.reshape(Cout, Cin)).reshape(B, H, W, Cin)
            if need_dw:
                with _bwd_precision():
                    # the arena keeps conv weights / gradients channels-last ([Cout][kh][kw][Cin]): accumulate in place
                    sink = None if ctx.w_sink is None else ctx.w_sink.permute(0, 2, 3, 1)
                    if kh == 3:
                        Wp = W + 2
                        taps = [(ky - 1) * Wp + (kx - 1) for ky in range(3) for kx in range(3)]
                        dwk = gemm_dw(gzp.reshape(-1, Cout), xin.reshape(-1, Cin),
                                      None if sink is None else sink.view(Cout, 9, Cin), taps)    # [Cout, 9, Cin]
                        dw = None if sink is not None else dwk.reshape(Cout, 3, 3, Cin).permute(0, 3, 1, 2)
                    else:
                        dwk = gemm_dw(gz.reshape(-1, Cout), xin.reshape(-1, Cin), None if sink is None else sink.view(Cout, Cin))
                        dw = None if sink is not None else dwk.reshape(Cout, Cin, 1, 1)
        dx0 = dx1 = None
        if need_dx:
            if C1 is not None and ctx.needs_input_grad[1]:
                dx1 = dxc[..., C0:].permute(0, 3, 1, 2)
            if ctx.needs_input_grad[0]:
                if (h0, w0) != (H, W):
                    d0 = torch.empty(B, h0, w0, C0, dtype=torch.float32, device=gh.device)
                    dcc = dxc if dxc.is_contiguous() else dxc.contiguous()
                    _call("ged_upsample_nhwc_bwd", _p(dcc), dcc.shape[3], _p(d0), C0, B, H, W, h0, w0, _stream())
                    dx0 = d0.permute(0, 3, 1, 2)
                else:
                    dx0 = (dxc[..., :C0] if C1 is not None else dxc).permute(0, 3, 1, 2)
        return dx0, dx1, dw, db, None, None, None, None


SMALL_COUT = (1, 2, 11)


def _conv_apply(x0, x1, w, b, act, slope):
    ws = _sink(w)
    if ws is not None and not ws.permute(0, 2, 3, 1).is_contiguous():
        ws = None
    return _Conv.apply(x0, x1, w, b, act, slope, ws, _sink(b))


def conv2d(x, w, b=None, stride=1, padding=0, act=None, slope=0.01):
    return _conv_apply(x, None, w, b, act, slope)


def conv2d_cat(x0, x1, w, b=None, act=None, slope=0.01):
    """3x3 conv over cat([bilinear(x0 -> size of x1, align_corners=True), x1], channels)."""
    return _conv_apply(x0, x1, w, b, act, slope)


def conv2d_cat_supported(x0, x1, w) -> bool:
    Cout, Cin, kh, kw = w.shape
    return (kh == 3 and kw == 3 and x0.shape[1] % 4 == 0 and x1.shape[1] % 4 == 0 and Cin % 32 == 0
            and x0.dtype == torch.float32 and x1.dtype == torch.float32)


def conv_bn_act(x, w, b, bn, stride=1, padding=0, act=None):
    """ConvModule(conv -> BN -> act).  Eval mode: BN folds into the conv's weight/bias and the
    activation into its epilogue.  Train mode needs batch statistics between conv and activation:
    conv (tcgen05) -> ged_bn_train (per-GPU batch statistics as in the reference) (+ReLU)."""
    return conv_bn_act_cat(x, None, w, b, bn, act)


def conv_bn_act_cat(x0, x1, w, b, bn, act=None):
    if bn is None:
        return _conv_apply(x0, x1, w, b, act, 0.01)
    if not bn.training:
        s = bn.weight * torch.rsqrt(bn.running_var + bn.eps)
        wf = w * s.view(-1, 1, 1, 1)
        bf = bn.bias - bn.running_mean * s + (b * s if b is not None else 0)
        return _Conv.apply(x0, x1, wf, bf, act, 0.01)
    y = _conv_apply(x0, x1, w, b, None, 0.0)
    if act in (None, "relu") and y.shape[1] % 4 == 0 and bn.momentum is not None and bn.track_running_stats:
        return bn_act_train(y, bn, relu=act == "relu")
    raise NotImplementedError("conv + train-mode BatchNorm: only (none | relu) activations with tracked running statistics")


class _BNTrain(Function):
    """Train-mode BatchNorm2d (+ReLU) on a channels-last map; batch statistics per GPU."""

    @staticmethod
    def forward(ctx, x, w, b, running_mean, running_var, eps, momentum, relu):
        xh = _nhwc(x)
        B, H, W, Cc = xh.shape
        rows = B * H * W
        y = torch.empty_like(xh)
        mean = torch.empty(Cc, dtype=torch.float32, device=x.device)
        rstd = torch.empty_like(mean)
        sums = torch.empty(2 * Cc, dtype=torch.float64, device=x.device)
        _call("ged_bn_train_fwd", _p(xh), _p(w), _p(b), _p(running_mean), _p(running_var), _p(y), _p(mean), _p(rstd),
              _p(sums), rows, Cc, float(eps), float(momentum), int(relu), _stream())
        ctx.save_for_backward(xh, y if relu else torch.empty(0, device=x.device), w, mean, rstd)
        ctx.relu = relu
        ctx.mark_non_differentiable(running_mean, running_var)
        return y.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, g):
        xh, y, w, mean, rstd = ctx.saved_tensors
        B, H, W, Cc = xh.shape
        gh = _nhwc(g)
        dx = torch.empty_like(xh)
        dw = torch.empty(Cc, dtype=torch.float32, device=g.device)
        db = torch.empty_like(dw)
        sums = torch.empty(2 * Cc, dtype=torch.float64, device=g.device)
        _call("ged_bn_train_bwd", _p(gh), _p(xh), _p(y if ctx.relu else None), _p(w), _p(mean), _p(rstd), _p(dx), _p(dw),
              _p(db), _p(sums), B * H * W, Cc, int(ctx.relu), _stream())
        return dx.permute(0, 3, 1, 2), dw, db, None, None, None, None, None


def bn_act_train(x, bn, relu: bool):
    y = _BNTrain.apply(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps, bn.momentum, relu)
    bn.num_batches_tracked.add_(1)
    return y


class _ResizeAdd(Function):
    """acc + bilinear(t -> size of acc, align_corners=True), NHWC, in place on a fresh copy of acc."""

    @staticmethod
    def forward(ctx, t, acc):
        th, ah = _nhwc(t), _nhwc(acc)
        out = torch.empty_like(ah)
        B, H, W, Cc = out.shape
        _call("ged_resize_add_nhwc", _p(th), _p(ah), _p(out), Cc, B, H, W, th.shape[1], th.shape[2], _stream())
        ctx.shape = th.shape
        return out.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, g):
        gh = _nhwc(g)
        B, h0, w0, Cc = ctx.shape
        gt = torch.empty(B, h0, w0, Cc, dtype=torch.float32, device=g.device)
        _call("ged_upsample_nhwc_bwd", _p(gh), Cc, _p(gt), Cc, B, gh.shape[1], gh.shape[2], h0, w0, _stream())
        return gt.permute(0, 3, 1, 2), g


def resize_add(t, size, acc):
    return _ResizeAdd.apply(t, acc)


# =============================================================================================
# patch embedding / patch merging / inference resize
# =============================================================================================
def patch_embed(x, w, b, patch):
    """4x4/s4 conv of embed.py:282-297 as patchify (gather) + tcgen05 GEMM.  x: NCHW view of the input batch."""
    B, Cin, H, W = x.shape
    assert x.stride(3) == 1 and x.stride(2) == W and x.stride(1) == H * W and x.dtype == torch.float32
    DH, DW = -(-H // patch), -(-W // patch)
    K = Cin * patch * patch
    tok = torch.empty(B, DH * DW, K, dtype=torch.float32, device=x.device)
    _call("ged_patchify", _p(x), x.stride(0), _p(tok), B, Cin, H, W, patch, _stream())
    return linear(tok, w.reshape(w.shape[0], K), b), (DH, DW)


def conv_im2col_supported(x, w, stride, padding) -> bool:
    """Strided / large-kernel convs on an input that needs no gradient (the 7x7/s2 RGB stem)."""
    Cout, Cin, kh, kw = w.shape
    return (x.dtype == torch.float32 and not x.requires_grad and x.stride(3) == 1 and x.stride(2) == x.shape[3]
            and x.stride(1) == x.shape[2] * x.shape[3] and Cout % 4 == 0 and Cout >= 16 and Cin * kh * kw >= 32)


def conv_im2col(x, w, b, stride, padding, act=None, slope=0.01):
    """conv2d as im2col (one gather pass) + tcgen05 GEMM; returns logical NCHW in channels-last memory.
    The weight gradient comes back through the GEMM's dW kernel; x gets no gradient."""
    B, Cin, H, W = x.shape
    Cout, _, kh, kw = w.shape
    K = Cin * kh * kw
    Kp = (K + 3) // 4 * 4
    Ho, Wo = (H + 2 * padding - kh) // stride + 1, (W + 2 * padding - kw) // stride + 1
    tok = torch.empty(B * Ho * Wo, Kp, dtype=torch.float32, device=x.device)
    _call("ged_im2col", _p(x), x.stride(0), _p(tok), B, Cin, H, W, kh, kw, stride, padding, Kp, _stream())
    w2 = w.reshape(Cout, K)
    if Kp != K:
        w2 = torch.nn.functional.pad(w2, (0, Kp - K))
    y = linear(tok, w2, b, act)
    return y.view(B, Ho, Wo, Cout).permute(0, 3, 1, 2)


class _MergePatches(Function):
    @staticmethod
    def forward(ctx, x, H, W):
        xc = _f32c(x)
        B, Lt, Cc = xc.shape
        H2, W2 = (H + 1) // 2, (W + 1) // 2
        out = torch.empty(B, H2 * W2, 4 * Cc, dtype=torch.float32, device=x.device)
        _call("ged_merge_patches", _p(xc), _p(out), B, H, W, Cc, 0, _stream())
        ctx.dims = (B, H, W, Cc)
        return out

    @staticmethod
    def backward(ctx, g):
        B, H, W, Cc = ctx.dims
        g = _f32c(g)
        gx = torch.empty(B, H * W, Cc, dtype=torch.float32, device=g.device)
        _call("ged_merge_patches", _p(g), _p(gx), B, H, W, Cc, 1, _stream())
        return gx, None, None


def merge_patches(x, H, W):
    return _MergePatches.apply(x, int(H), int(W))


def clamp_resize(x, lo, hi, size):
    x = _f32c(x)
    B, Cc, h0, w0 = x.shape
    assert Cc == 1
    if size is None:
        return torch.clamp(x, min=lo, max=hi)
    out = torch.empty(B, 1, int(size[0]), int(size[1]), dtype=torch.float32, device=x.device)
    _call("ged_clamp_resize", _p(x), _p(out), B, h0, w0, int(size[0]), int(size[1]), float(lo), float(hi), _stream())
    return out


# =============================================================================================
# deformable attention sampling
# =============================================================================================
# Three implementations of the sampling, selectable per direction (A/B partners of each other in the tests):
#   "round1": one warp per (query, head) straight from / to L2 (csrc/msda.cu)
#   "tile"  : queries sorted by reference point, 32-query tiles, shared-memory windows, fp32 SIMT (csrc/msda_tile.cu)
#   "tc"    : the same tiles on tcgen05 - forward 3xTF32 (fp32-accurate), backward one-pass TF32 (csrc/msda_tc.cu)
# Measured at B = 8, 352 x 1120 (tools/ab_msda_tile.py): forward round1 13.2 / tile 15.8 / tc 19.7 ms, backward
# round1 48.1 / tile 48.5 / tc 17.3 ms -> forward stays on round1, the backward runs on the tensor cores (the fp32 tile
# kernels when GEDEPTH_BWD_GEMM_PASSES=3 asks for fp32-accurate gradients).
MSDA_FWD = os.environ.get("GEDEPTH_MSDA_FWD", "round1")
MSDA_BWD = os.environ.get("GEDEPTH_MSDA_BWD", "tc")
MSDA_TILE_Q = 32


def _msda_bwd_impl() -> str:
    return "tile" if (MSDA_BWD == "tc" and BACKWARD_PASSES != 1) else MSDA_BWD


def msda_query_order(ref: torch.Tensor, shapes) -> torch.Tensor:
    """int32 (Q,) permutation of the queries grouped by reference point (band of y, bucket of x) so that a tile of
    32 consecutive entries samples a compact patch of every level.  Only locality depends on it, never the result.
    Constant reference points (no grad) keep their order cached on the tensor."""
    cached = getattr(ref, "_ged_order", None)
    if cached is not None and not ref.requires_grad and cached[0] == ref._version:
        return cached[1]
    Q = ref.shape[1]
    H0, W0 = int(shapes[0][0]), int(shapes[0][1])
    side = max(1.0, (MSDA_TILE_Q * H0 * W0 / max(Q, 1)) ** 0.5)       # tile side in level-0 pixels
    bands = max(1, min(H0, int(round(H0 / side))))
    xb = max(1, min(8 * W0, 65536 // bands))
    order = torch.empty(Q, dtype=torch.int32, device=ref.device)
    work = torch.empty(bands * xb, dtype=torch.int32, device=ref.device)
    _call("ged_msda_sort_queries", _p(ref), Q, bands, xb, _p(order), _p(work), work.numel(), _stream())
    if not ref.requires_grad:
        ref._ged_order = (ref._version, order)
    return order


class _MSDA(Function):
    @staticmethod
    def forward(ctx, v, ref, off, logit, shapes, nH, P, fwd_impl, bwd_impl):
        v, ref, off, logit = _f32c(v), _f32c(ref), _f32c(off), _f32c(logit)
        B, S, E = v.shape
        Q = off.shape[1]
        hw = (C.c_int * (2 * len(shapes)))(*[int(a) for s in shapes for a in s])
        out = torch.empty(B, Q, E, dtype=torch.float32, device=v.device)
        need_bwd = any(ctx.needs_input_grad[:4])
        order = None
        if fwd_impl != "round1" or (need_bwd and bwd_impl != "round1"):
            order = msda_query_order(ref, shapes)
        if fwd_impl == "tc":
            v_lo = torch.empty_like(v)          # value - tf32(value), the lo operand of the 3xTF32 forward
            _call("ged_msda_tc_fwd", _p(v), _p(v_lo), _p(ref), ref.shape[0], _p(off), _p(logit), _p(order), _p(out),
                  hw, len(shapes), B, S, Q, nH, E // nH, P, _stream())
        elif fwd_impl == "tile":
            _call("ged_msda_tile_fwd", _p(v), _p(ref), ref.shape[0], _p(off), _p(logit), _p(order), _p(out), hw,
                  len(shapes), B, S, Q, nH, E // nH, P, _stream())
        else:
            _call("ged_msda_fwd", _p(v), _p(ref), ref.shape[0], _p(off), _p(logit), _p(out), hw, len(shapes), B, S, Q,
                  nH, E // nH, P, _stream())
        ctx.save_for_backward(v, ref, off, logit, order if order is not None else torch.empty(0, device=v.device))
        ctx.cfg = (tuple(tuple(s) for s in shapes), nH, P, bwd_impl)
        return out

    @staticmethod
    def backward(ctx, g):
        v, ref, off, logit, order = ctx.saved_tensors
        shapes, nH, P, bwd_impl = ctx.cfg
        B, S, E = v.shape
        Q = off.shape[1]
        hw = (C.c_int * (2 * len(shapes)))(*[int(a) for s in shapes for a in s])
        g = _f32c(g)
        g_v = torch.zeros_like(v)
        g_off, g_logit = torch.empty_like(off), torch.empty_like(logit)
        tail = (None,) * 5
        if bwd_impl != "round1":
            # reference points shared by the batch (ref_batch == 1): their gradient is summed over the batch in place
            g_ref = torch.zeros_like(ref) if ctx.needs_input_grad[1] else None
            _call("ged_msda_tc_bwd" if bwd_impl == "tc" else "ged_msda_tile_bwd", _p(v), _p(ref), ref.shape[0], _p(off),
                  _p(logit), _p(order), _p(g), _p(g_v), _p(g_ref), _p(g_off), _p(g_logit), hw, len(shapes), B, S, Q, nH,
                  E // nH, P, _stream())
            return (g_v, g_ref, g_off, g_logit) + tail
        need_ref = ctx.needs_input_grad[1] and ref.shape[0] == B
        g_ref = torch.zeros_like(ref) if need_ref else None
        _call("ged_msda_bwd", _p(v), _p(ref), ref.shape[0], _p(off), _p(logit), _p(g), _p(g_v), _p(g_ref),
              _p(g_off), _p(g_logit), hw, len(shapes), B, S, Q, nH, E // nH, P, _stream())
        return (g_v, g_ref, g_off, g_logit) + tail


def msda_atomic_probe(rows: int = 32725, heads: int = 8, iters: int = 2000, device="cuda") -> float:
    """GB/s of atomic payload the L2 sustains for the MSDA backward's scatter pattern (roofline denominator)."""
    lib = load()
    lib.ged_msda_atomic_probe.argtypes = [_P, _I, _I, _I, _P]
    lib.ged_msda_atomic_probe.restype = C.c_int
    buf = torch.zeros(rows, heads * 64, dtype=torch.float32, device=device)
    lib.ged_msda_atomic_probe(_p(buf), rows, heads, 50, _stream())
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    warps = lib.ged_msda_atomic_probe(_p(buf), rows, heads, iters, _stream())
    e.record()
    torch.cuda.synchronize()
    if warps <= 0:
        raise RuntimeError("ged_msda_atomic_probe failed")
    return warps * iters * 512 / (s.elapsed_time(e) * 1e-3) / 1e9


def set_msda_variant(v: int) -> int:
    return load().ged_set_msda_variant(int(v))


def msda_sample(v, shapes, ref, off, logit, nH, P):
    B = v.shape[0]
    if ref.shape[0] not in (1, B):
        raise ValueError("reference_points batch must be 1 or B")
    bwd = _msda_bwd_impl()
    if ref.shape[0] == 1 and ref.requires_grad and B > 1 and bwd == "round1":
        ref = ref.expand(B, -1, -1)        # round-1 backward: learnable reference points need one copy per sample
    return _MSDA.apply(v, ref, off, logit, shapes, nH, P, MSDA_FWD, bwd)


class _SplitLevels(Function):
    """(B, S, C) token tensor -> one (B, n_l, C) view per level; the backward writes the four gradients into ONE buffer
    instead of autograd's four zero-filled full-size tensors and three additions."""

    @staticmethod
    def forward(ctx, src, *sizes):
        ctx.sizes = sizes
        ctx.shape = src.shape
        outs, start = [], 0
        for n in sizes:
            outs.append(src[:, start:start + n])
            start += n
        return tuple(outs)

    @staticmethod
    def backward(ctx, *gs):
        out = torch.empty(ctx.shape, dtype=torch.float32, device=gs[0].device)
        start = 0
        for g, n in zip(gs, ctx.sizes):
            out[:, start:start + n].copy_(g.reshape(ctx.shape[0], n, ctx.shape[2]))
            start += n
        return (out,) + (None,) * len(ctx.sizes)


def split_levels(src, sizes):
    if not (torch.is_grad_enabled() and src.requires_grad):
        outs, start = [], 0
        for n in sizes:
            outs.append(src[:, start:start + n])
            start += n
        return outs
    return list(_SplitLevels.apply(src, *[int(n) for n in sizes]))


class _LinearSmall(Function):
    """y = act(x w^T + b) with N <= 4 outputs (csrc/small.cu): HAHIHeteroNeck.reference_points, 512 -> 2 + sigmoid."""

    @staticmethod
    def forward(ctx, x, w, b, act, w_sink=None, b_sink=None):
        K, N = x.shape[-1], w.shape[0]
        x2 = _f32c(x).reshape(-1, K)
        wc = _f32c(w)
        y = torch.empty(x2.shape[0], N, dtype=torch.float32, device=x.device)
        _call("ged_linear_small_fwd", _p(x2), _p(wc), _p(b), _p(y), x2.shape[0], N, K, _ACT[act], _stream())
        ctx.save_for_backward(x2, wc, y)
        ctx.cfg = (act, b is not None, x.shape, w_sink, b_sink)
        return y.reshape(*x.shape[:-1], N)

    @staticmethod
    def backward(ctx, g):
        x2, w, y = ctx.saved_tensors
        act, has_bias, xshape, w_sink, b_sink = ctx.cfg
        N, K = w.shape
        g2 = _f32c(g).reshape(-1, N)
        dx = torch.empty_like(x2) if ctx.needs_input_grad[0] else None
        dw = db = None
        if ctx.needs_input_grad[1]:
            dw = w_sink if w_sink is not None else torch.zeros_like(w)
        if has_bias and ctx.needs_input_grad[2]:
            db = b_sink if b_sink is not None else torch.zeros(N, dtype=torch.float32, device=w.device)
        _call("ged_linear_small_bwd", _p(g2), _p(y), _p(x2), _p(w), _p(dx), _p(dw), _p(db), x2.shape[0], N, K, _ACT[act],
              _stream())
        return (None if dx is None else dx.reshape(xshape), None if w_sink is not None else dw,
                None if b_sink is not None else db, None, None, None)


def linear_small(x, w, b, act=None):
    if w.shape[0] > 4 or x.shape[-1] % 4 or act not in (None, "sigmoid"):
        raise NotImplementedError(f"linear_small: weight {tuple(w.shape)} act {act}")
    ws = _sink(w)
    ws = ws if (ws is not None and ws.is_contiguous() and w.is_contiguous()) else None
    return _LinearSmall.apply(x, w, b, act, ws, _sink(b))


def _level_starts(level_start):
    return None if level_start is None else (C.c_int * 5)(*[int(v) for v in level_start])


class _MSDAModule(Function):
    """mmcv MultiScaleDeformableAttention.forward(batch_first=True) [external; hahi.py:179-188,280-289,316-325] as one
    autograd node:  q = query + pos (+ level embedding);  v = value_proj(value);  off / logit = Linear(q);  sampling;
    out = dropout(output_proj(sampled)) + query.
    The backward keeps every sum inside a kernel: the query gradient's fan-in (identity + the two Linear(q) + - when
    value is the query - value_proj) is accumulated through the residual input of the dX GEMMs, the level-embedding
    gradient is the per-level column sum of dq, and parameter gradients go straight to the arena when there is one."""

    @staticmethod
    def forward(ctx, query, value, pos, level_embed, ref, w_v, b_v, w_so, b_so, w_aw, b_aw, w_o, b_o, cfg):
        qc = _f32c(query)
        B, Q, E = qc.shape
        shapes, nH, P = cfg["shapes"], cfg["nH"], cfg["P"]
        ls = _level_starts(cfg["level_start"]) if level_embed is not None else None
        posc = _f32c(pos).reshape(-1, E)
        assert posc.shape[0] == Q, "query_pos must be (1, Q, E)"
        q = torch.empty_like(qc)
        _call("ged_add_pos_fwd", _p(qc), _p(posc), _p(None if level_embed is None else _f32c(level_embed)), ls, _p(q),
              B, Q, E, _stream())
        vin = qc if value is None else _f32c(value)
        S = vin.shape[1]
        v = gemm(vin.reshape(-1, E), _f32c(w_v), b_v).reshape(B, S, E)
        q2 = q.reshape(-1, E)
        off = gemm(q2, _f32c(w_so), b_so).reshape(B, Q, -1)
        logit = gemm(q2, _f32c(w_aw), b_aw).reshape(B, Q, -1)
        refc = _f32c(ref)
        hw = (C.c_int * (2 * len(shapes)))(*[int(a) for s in shapes for a in s])
        fwd_impl, bwd_impl = MSDA_FWD, _msda_bwd_impl()
        samp = torch.empty(B, Q, E, dtype=torch.float32, device=qc.device)
        order = None
        if fwd_impl != "round1" or bwd_impl != "round1":
            order = msda_query_order(refc, shapes)
        if fwd_impl == "tc":
            v_lo = torch.empty_like(v)
            _call("ged_msda_tc_fwd", _p(v), _p(v_lo), _p(refc), refc.shape[0], _p(off), _p(logit), _p(order), _p(samp),
                  hw, len(shapes), B, S, Q, nH, E // nH, P, _stream())
        elif fwd_impl == "tile":
            _call("ged_msda_tile_fwd", _p(v), _p(refc), refc.shape[0], _p(off), _p(logit), _p(order), _p(samp), hw,
                  len(shapes), B, S, Q, nH, E // nH, P, _stream())
        else:
            _call("ged_msda_fwd", _p(v), _p(refc), refc.shape[0], _p(off), _p(logit), _p(samp), hw, len(shapes), B, S, Q,
                  nH, E // nH, P, _stream())
        dropout = None
        if cfg["dropout_p"] > 0:
            assert B * Q * E < 2 ** 32
            dropout = (float(cfg["dropout_p"]), next_dropout_seed(), RNG_STEP)
        out = gemm(samp.reshape(-1, E), _f32c(w_o), b_o, None, 0.01, qc.reshape(-1, E), dropout=dropout).reshape(B, Q, E)
        empty = torch.empty(0, device=qc.device)
        ctx.save_for_backward(q, vin, v, refc, off, logit, samp,
                              order if order is not None else empty, w_v, w_so, w_aw, w_o,
                              level_embed if level_embed is not None else empty)
        ctx.cfg = dict(cfg, dropout=dropout, bwd_impl=bwd_impl, shared_value=value is None, has_le=level_embed is not None,
                       hw=hw, dims=(B, S, Q, E))
        return out

    @staticmethod
    def backward(ctx, g):
        q, vin, v, ref, off, logit, samp, order, w_v, w_so, w_aw, w_o, level_embed = ctx.saved_tensors
        cfg = ctx.cfg
        B, S, Q, E = cfg["dims"]
        shapes, nH, P, sinks = cfg["shapes"], cfg["nH"], cfg["P"], cfg["sinks"]
        dev = q.device
        # g may be a channel slice of the fusion conv's dX (row pitch 512 + C1): its consumers below take the pitch, except
        # the level-embedding adjoint, which wants the dense matrix
        g2 = _rows(g, E) if not cfg["has_le"] else _f32c(g).reshape(-1, E)

        def acc_or_new(name, shape):
            return sinks[name] if sinks.get(name) is not None else torch.zeros(shape, dtype=torch.float32, device=dev)

        def ret(name, t):
            return None if sinks.get(name) is not None else t

        # ---- output_proj: out = dropout(samp w_o^T + b_o) + query ----------------------------------------------------
        db_o = acc_or_new("b_o", (E,))
        if cfg["dropout"] is not None:
            dp, dseed, dstep = cfg["dropout"]
            gz = torch.empty_like(g2)
            _call("ged_dropout_bwd", _p(g2), g2.stride(0), _p(gz), _p(db_o), g2.shape[0], E, float(dp), int(dseed), _p(dstep),
                  _stream())
        else:
            gz = g2
            act_bwd(g2, None, None, want_db=True, db_sink=db_o)
        with _bwd_precision():
            d_samp = gemm_bt(gz, w_o)
            dw_o = gemm_dw(gz, samp.reshape(-1, E), out=sinks.get("w_o"))
        del gz
        # ---- sampling ------------------------------------------------------------------------------------------------
        g_v = torch.zeros_like(v)
        g_off, g_logit = torch.empty_like(off), torch.empty_like(logit)
        need_ref = ctx.needs_input_grad[4]
        bwd_impl = cfg["bwd_impl"]
        if bwd_impl != "round1":
            g_ref = torch.zeros_like(ref) if need_ref else None
            _call("ged_msda_tc_bwd" if bwd_impl == "tc" else "ged_msda_tile_bwd", _p(v), _p(ref), ref.shape[0], _p(off),
                  _p(logit), _p(order), _p(d_samp), _p(g_v), _p(g_ref), _p(g_off), _p(g_logit), cfg["hw"], len(shapes), B, S,
                  Q, nH, E // nH, P, _stream())
        else:
            refb = ref if (ref.shape[0] == B or not need_ref) else ref.expand(B, -1, -1).contiguous()
            g_refb = torch.zeros_like(refb) if need_ref else None
            _call("ged_msda_bwd", _p(v), _p(refb), refb.shape[0], _p(off), _p(logit), _p(d_samp), _p(g_v), _p(g_refb),
                  _p(g_off), _p(g_logit), cfg["hw"], len(shapes), B, S, Q, nH, E // nH, P, _stream())
            g_ref = None if g_refb is None else (g_refb if ref.shape[0] == B else g_refb.sum(0, keepdim=True))
        del d_samp
        # ---- sampling_offsets / attention_weights: Linear(q) ------------------------------------------------------------
        q2, go2, gl2 = q.reshape(-1, E), g_off.reshape(B * Q, -1), g_logit.reshape(B * Q, -1)
        db_so, db_aw = acc_or_new("b_so", (go2.shape[1],)), acc_or_new("b_aw", (gl2.shape[1],))
        act_bwd(go2, None, None, want_db=True, db_sink=db_so)
        act_bwd(gl2, None, None, want_db=True, db_sink=db_aw)
        with _bwd_precision():
            dw_so = gemm_dw(go2, q2, out=sinks.get("w_so"))
            dw_aw = gemm_dw(gl2, q2, out=sinks.get("w_aw"))
            # g_query = g (identity) + g_off W_so + g_logit W_aw (+ g_v W_v when the value is the query itself)
            g_le = None
            if cfg["has_le"]:
                dq = gemm_bt(go2, w_so)
                gemm_bt(gl2, w_aw, residual=dq, out=dq)
                g_le = acc_or_new("level_embed", tuple(level_embed.shape))
                _call("ged_add_pos_bwd", _p(dq), _p(g2), _p(g_le), _level_starts(cfg["level_start"]), B, Q, E, _stream())
            else:
                dq = gemm_bt(go2, w_so, residual=g2)
                gemm_bt(gl2, w_aw, residual=dq, out=dq)
            # ---- value_proj ------------------------------------------------------------------------------------------
            gv2 = g_v.reshape(-1, E)
            db_v = acc_or_new("b_v", (E,))
            act_bwd(gv2, None, None, want_db=True, db_sink=db_v)
            dw_v = gemm_dw(gv2, vin.reshape(-1, E), out=sinks.get("w_v"))
            g_value = None
            if cfg["shared_value"]:          # value is the query itself (self-attention): one more term of its fan-in
                gemm_bt(gv2, w_v, residual=dq, out=dq)
            else:
                g_value = gemm_bt(gv2, w_v).reshape(B, S, E)
        g_query = dq.reshape(B, Q, E)
        return (g_query, g_value, None, ret("level_embed", g_le), g_ref, ret("w_v", dw_v), ret("b_v", db_v),
                ret("w_so", dw_so), ret("b_so", db_so), ret("w_aw", dw_aw), ret("b_aw", db_aw), ret("w_o", dw_o),
                ret("b_o", db_o), None)


def msda_module(query, value, pos, level_embed, level_start, ref, shapes, mod, dropout_p):
    B = query.shape[0]
    if ref.shape[0] not in (1, B):
        raise ValueError("reference_points batch must be 1 or B")
    if torch.is_tensor(pos) and pos.requires_grad:
        raise NotImplementedError("msda_module: query_pos must be constant (a learnable level embedding goes in level_embed)")
    named = dict(w_v=mod.value_proj.weight, b_v=mod.value_proj.bias, w_so=mod.sampling_offsets.weight,
                 b_so=mod.sampling_offsets.bias, w_aw=mod.attention_weights.weight, b_aw=mod.attention_weights.bias,
                 w_o=mod.output_proj.weight, b_o=mod.output_proj.bias, level_embed=level_embed)
    sinks = {}
    for k, p in named.items():
        sk = _sink(p)
        sinks[k] = sk if (sk is not None and sk.is_contiguous() and p.is_contiguous()) else None
    qc = _f32c(query)
    cfg = dict(shapes=tuple(tuple(int(a) for a in s) for s in shapes), nH=mod.num_heads, P=mod.num_points,
               level_start=None if level_start is None else tuple(int(v) for v in level_start), dropout_p=float(dropout_p),
               sinks=sinks)
    return _MSDAModule.apply(qc, value, pos, level_embed, ref, named["w_v"], named["b_v"], named["w_so"], named["b_so"],
                             named["w_aw"], named["b_aw"], named["w_o"], named["b_o"], cfg)


# =============================================================================================
# evaluation
# =============================================================================================
def depth_metric_sums(pred: torch.Tensor, gt: torch.Tensor, rect, min_depth: float, max_depth: float,
                      sums: Optional[torch.Tensor] = None) -> torch.Tensor:
    """(B,10) fp64 per-image sums (see include/gedepth.h); pred / gt (B,H,W) or (B,1,H,W) fp32 on the device."""
    pred, gt = _f32c(pred), _f32c(gt)
    H, W = gt.shape[-2], gt.shape[-1]
    B = gt.numel() // (H * W)
    assert pred.numel() == gt.numel()
    if sums is None:
        sums = torch.zeros(B, 10, dtype=torch.float64, device=gt.device)
    y0, y1, x0, x1 = (0, H, 0, W) if rect is None else [int(v) for v in rect]
    _call("ged_depth_metrics", _p(pred), _p(gt), _p(sums), B, H, W, y0, y1, x0, x1, float(min_depth), float(max_depth), _stream())
    return sums


def rgb_crop_normalize_into(img5: torch.Tensor, bgr_u8: torch.Tensor, top: int, left: int, flip: bool, mean, std,
                            to_rgb: bool = True):
    """Planes 0..2 of one (5,H,W) slice of the input batch from a uint8 (H0,W0,3) BGR image on the device."""
    assert bgr_u8.dtype == torch.uint8 and bgr_u8.is_contiguous() and bgr_u8.dim() == 3 and bgr_u8.shape[2] == 3
    assert img5.dtype == torch.float32 and img5.is_contiguous() and img5.shape[0] == 5
    H, W = img5.shape[1], img5.shape[2]
    m = (C.c_float * 3)(*[float(v) for v in mean])
    sd = (C.c_float * 3)(*[float(v) for v in std])
    _call("ged_rgb_crop_normalize", _p(bgr_u8), bgr_u8.shape[0], bgr_u8.shape[1], int(top), int(left), int(bool(flip)),
          int(bool(to_rgb)), m, sd, _p(img5), H, W, _stream())
    return img5


def tta_merge(a: torch.Tensor, b_flipped: torch.Tensor) -> torch.Tensor:
    a, b_flipped = _f32c(a), _f32c(b_flipped)
    H, W = a.shape[-2], a.shape[-1]
    out = torch.empty_like(a)
    _call("ged_tta_merge", _p(a), _p(b_flipped), _p(out), a.numel() // (H * W), H, W, _stream())
    return out


# =============================================================================================
# optimizer
# =============================================================================================
def sumsq(flat_grad: torch.Tensor, out: torch.Tensor):
    _call("ged_sumsq", _p(flat_grad), flat_grad.numel(), _p(out), _stream())
    return out


def adamw_step(p, g, m, v, wd_mask, sumsq_buf, max_norm, grad_scale, lr, beta1, beta2, eps, wd, step, step_dev=None,
               lr_dev=None):
    """step: host-side 1-based count, or step_dev: int32 device tensor holding it (CUDA-graph replay); lr_dev: fp32
    device scalar overriding lr (so that a schedule keeps working when the step is a replayed graph)."""
    _call("ged_adamw_step", _p(p), _p(g), _p(m), _p(v), _p(wd_mask), p.numel(), _p(sumsq_buf), float(max_norm),
          float(grad_scale), float(lr), float(beta1), float(beta2), float(eps), float(wd), int(step), _p(step_dev),
          _p(lr_dev), _stream())
