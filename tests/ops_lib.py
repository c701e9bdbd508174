"""Library (cuDNN / cuBLAS / ATen) statements of every op in gedepth_b200.ops - TEST INFRASTRUCTURE, not product.

Three uses, all explicit:
  1. tests: the "plain PyTorch fp32 reference of the same op" each hand-written kernel is compared
     with, op by op, on the GPU box;
  2. tests/conftest.py ``host_ops_on_cpu``: ``install(ops)`` swaps these in so the host-side mirror's wiring can be
     checked against the reference goldens on CPU;
  3. bench.py ``gpu_library_baseline``: the same step through cuDNN / cuBLAS-TF32 / grid_sample - the reference's own
     design on the same B200 - as the comparator of the hand-written path.
gedepth_b200 itself never imports this module: ops.py has one implementation per op and raises otherwise.
Signatures match ops.py one for one.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


def _act(x: Tensor, act: Optional[str], slope: float = 0.01) -> Tensor:
    if act is None:
        return x
    if act == "relu":
        return F.relu(x)
    if act == "leaky_relu":
        return F.leaky_relu(x, slope)
    if act == "gelu":
        return F.gelu(x)
    if act == "sigmoid":
        return torch.sigmoid(x)
    raise KeyError(act)


def linear(x, w, b=None, act=None, residual=None, row_scale=None):
    y = _act(F.linear(x, w, b), act)
    if row_scale is not None:
        y = y * row_scale.view(-1, *([1] * (y.dim() - 1)))
    return y if residual is None else residual + y


def layer_norm(x, w, b, eps):
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


def patch_embed(x, w, b, patch):
    H, W = x.shape[2:]
    if H % patch:
        x = F.pad(x, (0, 0, 0, patch - H % patch))
    if W % patch:
        x = F.pad(x, (0, patch - W % patch, 0, 0))
    x = F.conv2d(x, w, b, stride=patch)
    hw = (x.shape[2], x.shape[3])
    return x.flatten(2).transpose(1, 2), hw


def merge_patches(x, H, W):
    B, L, C = x.shape
    x = x.view(B, H, W, C)
    if H % 2 or W % 2:
        x = F.pad(x, (0, 0, 0, W % 2, 0, H % 2))
    Hp, Wp = x.shape[1], x.shape[2]
    # nn.Unfold(2,2) on NCHW orders features (c, kh, kw)
    x = x.view(B, Hp // 2, 2, Wp // 2, 2, C).permute(0, 1, 3, 5, 2, 4)
    return x.reshape(B, (Hp // 2) * (Wp // 2), 4 * C)


def _shift_mask(Hp, Wp, ws, shift, device):
    img_mask = torch.zeros(1, Hp, Wp, 1, device=device)
    cnt = 0
    for h in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
        for w in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
            img_mask[:, h, w, :] = cnt
            cnt += 1
    mw = img_mask.view(1, Hp // ws, ws, Wp // ws, ws, 1).permute(0, 1, 3, 2, 4, 5).reshape(-1, ws * ws)
    m = mw.unsqueeze(1) - mw.unsqueeze(2)
    return m.masked_fill(m != 0, -100.0).masked_fill(m == 0, 0.0)


def window_attention(qkv, qkv_bias, table, index, hw, nH, ws, shift, scale):
    """qkv (B, L, 3C) in image order -> context (B, L, C).  Zero-padded tokens (bottom/right, to a
    multiple of ws) are real keys whose q=k=v equal the qkv bias (depthformer_swin.py:292-294)."""
    B, L, C3 = qkv.shape
    C = C3 // 3
    H, W = hw
    hd = C // nH
    x = qkv.view(B, H, W, C3)
    pad_r, pad_b = (ws - W % ws) % ws, (ws - H % ws) % ws
    if pad_r or pad_b:
        Hp, Wp = H + pad_b, W + pad_r
        bias = qkv_bias if qkv_bias is not None else qkv.new_zeros(C3)
        xp = bias.view(1, 1, 1, C3).expand(B, Hp, Wp, C3).clone()
        xp[:, :H, :W] = x
        x = xp
    Hp, Wp = x.shape[1], x.shape[2]
    mask = None
    if shift > 0:
        x = torch.roll(x, shifts=(-shift, -shift), dims=(1, 2))
        mask = _shift_mask(Hp, Wp, ws, shift, qkv.device).to(qkv.dtype)
    win = x.view(B, Hp // ws, ws, Wp // ws, ws, C3).permute(0, 1, 3, 2, 4, 5).reshape(-1, ws * ws, C3)
    Bw, N = win.shape[0], ws * ws
    q, k, v = win.view(Bw, N, 3, nH, hd).permute(2, 0, 3, 1, 4)
    attn = (q * scale) @ k.transpose(-2, -1)
    rpb = table[index.view(-1).long()].view(N, N, nH).permute(2, 0, 1)
    attn = attn + rpb.unsqueeze(0)
    if mask is not None:
        nW = mask.shape[0]
        attn = (attn.view(Bw // nW, nW, nH, N, N) + mask.unsqueeze(1).unsqueeze(0)).view(-1, nH, N, N)
    attn = attn.softmax(-1)
    o = (attn @ v).transpose(1, 2).reshape(Bw, N, C)
    o = o.view(B, Hp // ws, Wp // ws, ws, ws, C).permute(0, 1, 3, 2, 4, 5).reshape(B, Hp, Wp, C)
    if shift > 0:
        o = torch.roll(o, shifts=(shift, shift), dims=(1, 2))
    return o[:, :H, :W, :].reshape(B, H * W, C)


def tokens_to_map(x, hw):
    B, L, C = x.shape
    return x.reshape(B, hw[0], hw[1], C).permute(0, 3, 1, 2)


def map_to_tokens(x):
    B, C, H, W = x.shape
    return x.permute(0, 2, 3, 1).reshape(B, H * W, C)


def conv2d(x, w, b=None, stride=1, padding=0, act=None, slope=0.01):
    return _act(F.conv2d(x, w, b, stride=stride, padding=padding), act, slope)


def conv_bn_act(x, w, b, bn, stride=1, padding=0, act=None):
    y = F.conv2d(x, w, b, stride=stride, padding=padding)
    if bn is not None:
        y = bn(y)
    return _act(y, act)


def resize(x, size, align_corners):
    return F.interpolate(x, size=tuple(size), mode="bilinear", align_corners=align_corners)


def resize_add(t, size, acc):
    return acc + F.interpolate(t, size=tuple(size), mode="bilinear", align_corners=True)


def cat_channels(xs):
    return torch.cat(xs, dim=1)


def add_bcast(a, b):
    return a + b


def msda_sample(v, shapes, ref, off, logit, nH, P):
    """v (B,S,E); ref (1|B,Q,2); off (B,Q,nH*L*P*2); logit (B,Q,nH*L*P) -> (B,Q,E)."""
    B, S, E = v.shape
    Q = off.shape[1]
    L = len(shapes)
    hd = E // nH
    value = v.view(B, S, nH, hd)
    off = off.view(B, Q, nH, L, P, 2)
    w = logit.view(B, Q, nH, L * P).softmax(-1).view(B, Q, nH, L, P)
    norm = torch.tensor([[w_, h] for h, w_ in shapes], dtype=v.dtype, device=v.device)
    loc = ref[:, :, None, None, None, :] + off / norm[None, None, None, :, None, :]
    vals = value.split([h * w_ for h, w_ in shapes], dim=1)
    grids = 2 * loc - 1
    samp = []
    for lvl, (h, w_) in enumerate(shapes):
        vl = vals[lvl].flatten(2).transpose(1, 2).reshape(B * nH, hd, h, w_)
        g = grids[:, :, :, lvl].transpose(1, 2).flatten(0, 1)
        samp.append(F.grid_sample(vl, g, mode="bilinear", padding_mode="zeros", align_corners=False))
    w = w.transpose(1, 2).reshape(B * nH, 1, Q, L * P)
    out = (torch.stack(samp, dim=-2).flatten(-2) * w).sum(-1).view(B, nH * hd, Q)
    return out.transpose(1, 2).contiguous()


def ground_plane(coef, H, W, device, batch, u0, v0, depth_scale, clamp_max, su, sv):
    """(batch, 2, H, W): ch0 = clamp(pe,0,clamp_max)/depth_scale (the loader's ch3 after
    Normalize), ch1 = raw pe.  fp64 math on the integer grid like the reference script."""
    num, cu, cv, c1 = coef
    u = (torch.arange(W, device=device, dtype=torch.float64) * su + u0).view(1, W)
    v = (torch.arange(H, device=device, dtype=torch.float64) * sv + v0).view(H, 1)
    pe = (num / (cu * u + cv * v + c1)).to(torch.float32)
    ch3 = pe.clone()
    ch3[ch3 > clamp_max] = 0
    ch3[ch3 < 0] = 0
    ch3 = torch.where(ch3 > 0, ch3 / depth_scale, ch3)
    return torch.stack([ch3, pe], 0).unsqueeze(0).expand(batch, -1, -1, -1).contiguous()


def ge_vanilla(img, y_half):
    y = F.interpolate(y_half, size=img.shape[2:], mode="bilinear", align_corners=False)
    return y, img[:, 3:4] * y * 200


def ge_adaptive(img, y_half, logits_half, height, depth_scale):
    y = F.interpolate(y_half, size=img.shape[2:], mode="bilinear", align_corners=False)
    logits = F.interpolate(logits_half, size=img.shape[2:], mode="bilinear", align_corners=False)
    idx = torch.linspace(-5, 5, 11, device=img.device).view(1, 11, 1, 1)
    k = torch.tan(torch.deg2rad((logits.softmax(1) * idx).sum(1, keepdim=True)))
    h = height.to(img.dtype).view(-1, 1, 1, 1) if torch.is_tensor(height) else height
    pe = img[:, 4:5]
    a = -h / (pe + 1e-8)
    off = -h / ((a - k) + 1e-8)
    m = ((off > 0) & (off <= depth_scale)).to(off.dtype)
    return y, off * m * y, logits


def fuse_head(d, pe_mask, y, min_depth):
    pe_h = F.interpolate(pe_mask, size=d.shape[2:], mode="bilinear", align_corners=True)
    y_h = F.interpolate(y, size=d.shape[2:], mode="bilinear", align_corners=True)
    return d * (1 - y_h) + pe_h + min_depth, y_h


def silog(pred, gt, eps, lam, max_depth, upsample):
    if upsample:
        pred = F.interpolate(pred, size=gt.shape[2:], mode="bilinear", align_corners=True)
    m = gt > 0
    if max_depth is not None:
        m = m & (gt <= max_depth)
    g = torch.log(pred[m] + eps) - torch.log(gt[m] + eps)
    return torch.sqrt(torch.var(g) + lam * torch.mean(g) ** 2)


def cross_entropy(logits, target, ignore_index=255):
    return F.cross_entropy(logits, target.long(), ignore_index=ignore_index)


def clamp_resize(x, lo, hi, size, align_corners):
    x = torch.clamp(x, min=lo, max=hi)
    return x if size is None else F.interpolate(x, size=tuple(size), mode="bilinear",
                                                align_corners=align_corners)


# ---- ops added with the fused / narrow kernels ---------------------------------------------------------------
def linear_full(x, w, b=None, act=None, residual=None, row_scale=None, dropout_p: float = 0.0):
    y = linear(x, w, b, act, None if dropout_p > 0 else residual, row_scale)
    if dropout_p > 0:
        y = F.dropout(y, dropout_p, True)
        y = y if residual is None else y + residual
    return y


def linear_small(x, w, b, act=None):
    return _act(F.linear(x, w, b), act)


def layer_norm_fork(x, w, b, eps):
    return layer_norm(x, w, b, eps), x


def conv2d_cat(x_low, x_skip, w, b=None, act=None, slope=0.01):
    up = resize(x_low, (x_skip.shape[2], x_skip.shape[3]), True) if x_low.shape[2:] != x_skip.shape[2:] else x_low
    return conv2d(cat_channels([up, x_skip]), w, b, 1, 1, act, slope)


def conv_bn_act_cat(x0, x1, w, b, bn, act=None):
    return conv_bn_act(cat_channels([x0, x1]), w, b, bn, 1, 1, act)


def cat_tokens(xs):
    return torch.cat(xs, dim=1)


def split_levels(src, sizes):
    return list(src.split([int(n) for n in sizes], dim=1))


def ge_adaptive_full(img, y_half, logits_half, height, depth_scale, want_logits=None):
    y, pm, lf = ge_adaptive(img, y_half, logits_half, height, depth_scale)
    want = torch.is_grad_enabled() if want_logits is None else want_logits
    return y, pm, (lf if want else None)


def msda_module(query, value, pos, level_embed, level_start, ref, shapes, mod, dropout_p):
    """mmcv MultiScaleDeformableAttention.forward(batch_first=True) [external, mmcv-full 1.3.13] with the level embedding
    of hahi.py:252-270 added to the constant part of query_pos."""
    q = query + pos
    if level_embed is not None:
        q = q + torch.cat([level_embed[i].view(1, 1, -1).expand(1, level_start[i + 1] - level_start[i], -1)
                           for i in range(len(level_start) - 1)], 1)
    value = query if value is None else value
    v = F.linear(value, mod.value_proj.weight, mod.value_proj.bias)
    off = F.linear(q, mod.sampling_offsets.weight, mod.sampling_offsets.bias)
    logit = F.linear(q, mod.attention_weights.weight, mod.attention_weights.bias)
    out = msda_sample(v, shapes, ref, off, logit, mod.num_heads, mod.num_points)
    y = F.linear(out, mod.output_proj.weight, mod.output_proj.bias)
    if dropout_p > 0:
        y = F.dropout(y, dropout_p, True)
    return y + query


def install(ops):
    """Route every op of ``gedepth_b200.ops`` to its library statement (tests and bench's comparator only).  Returns a
    callable that restores the product's functions."""
    table = dict(linear=linear_full, conv2d=conv2d, conv_bn_act=conv_bn_act, conv2d_cat=conv2d_cat,
                 conv_bn_act_cat=conv_bn_act_cat, patch_embed=patch_embed, layer_norm=layer_norm,
                 layer_norm_fork=layer_norm_fork, merge_patches=merge_patches, window_attention=window_attention,
                 cat_tokens=cat_tokens, split_levels=split_levels, resize=resize, resize_add=resize_add, clamp_resize=clamp_resize,
                 msda_module=msda_module, linear_small=linear_small, ground_plane=ground_plane, ge_vanilla=ge_vanilla,
                 ge_adaptive=ge_adaptive_full, fuse_head=fuse_head, silog=silog, cross_entropy=cross_entropy,
                 require_cuda=lambda *a, **k: None, use_native=lambda name: False)
    saved = {k: getattr(ops, k) for k in table}
    for k, v in table.items():
        setattr(ops, k, v)

    def restore():
        for k, v in saved.items():
            setattr(ops, k, v)
    return restore
