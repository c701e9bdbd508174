"""ORACLE (test infrastructure, not product code) - plain PyTorch CPU restatement of the GEDepth
hot path as pure functions over a reference-keyed ``state_dict``.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this.  Pinned: oracle/ref_harness.py executes the reference's own files (under mmcv stubs,
build container only) on the same weights/inputs; tests/golden/*.npz hold the reference's outputs
and tests/test_oracle_golden.py checks this file against them.  mmcv leaf layers (ConvModule, FFN,
MultiScaleDeformableAttention) are restated from mmcv 1.3.x semantics - unpinned by any reference
test.

Reference lines followed (relative to /root/reference):
  stem                depth/models/backbones/depthformer_swin.py:1127-1139,1152-1154
  patch_embed         depth/models/utils/embed.py:282-302
  window_msa          depth/models/backbones/depthformer_swin.py:184-224
  shift_window_msa    :285-360 (partition :379-393, reverse :362-377)
  swin_block          :461-472 ; patch_merging :98-122 ; backbone :1149-1184
  sine_pe             depth/utils/position_encoding.py:54-89
  msda                mmcv.ops.multi_scale_deform_attn (1.3.x) [external]
  hahi                depth/models/necks/hahi.py:235-356
  pe_neck             depth/models/necks/pemask_neck.py:52-64, dynamicpe_neck.py:512-539
  ground_embed_*      depth/models/depther/encoder_decoder.py:79-124
  dense_depth_head    depth/models/decode_heads/densedepth_head.py:14-27,100-131
  fuse_head           depth/models/decode_heads/decode_head.py:489-508
  silog / losses      depth/models/losses/sigloss.py:36-53, decode_head.py:512-542,583-599
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


@dataclass
class PathConfig:
    embed_dims: int = 96
    depths: Sequence[int] = (2, 2, 6, 2)
    num_heads: Sequence[int] = (3, 6, 12, 24)
    window_size: int = 7
    adaptive: bool = False
    depth_scale: float = 200.0
    min_depth: float = 1e-3
    max_depth: float = 80.0
    msda_heads: int = 8
    msda_levels: int = 4
    msda_points: int = 8
    embedding_dim: int = 512
    train_bn: bool = False       # BatchNorm uses batch statistics (train mode)
    leaky_slope: float = 0.01
    ce_weight: float = 0.08
    sig_weight: float = 1.0
    bn_eps: float = 1e-5
    ln_eps: float = 1e-5


# ---------------------------------------------------------------------------------------------
# leaf helpers
# ---------------------------------------------------------------------------------------------
def _bn(x: Tensor, sd, p: str, cfg: PathConfig) -> Tensor:
    return F.batch_norm(x, sd[p + "running_mean"].clone(), sd[p + "running_var"].clone(),
                        sd[p + "weight"], sd[p + "bias"], training=cfg.train_bn, momentum=0.1,
                        eps=cfg.bn_eps)


def _conv_bn_relu(x, sd, p, cfg, padding=0):
    """mmcv ConvModule(norm=BN, act=ReLU): conv has no bias (bias='auto')."""
    x = F.conv2d(x, sd[p + "conv.weight"], None, padding=padding)
    return F.relu(_bn(x, sd, p + "bn.", cfg))


def _ln(x, sd, p, cfg):
    return F.layer_norm(x, (x.shape[-1],), sd[p + "weight"], sd[p + "bias"], cfg.ln_eps)


def _lin(x, sd, p):
    return F.linear(x, sd[p + "weight"], sd.get(p + "bias"))


# ---------------------------------------------------------------------------------------------
# backbone
# ---------------------------------------------------------------------------------------------
def stem(img: Tensor, sd, cfg: PathConfig) -> Tensor:
    x = F.conv2d(img[:, 0:3], sd["backbone.conv1.weight"], None, stride=2, padding=3)
    return F.relu(_bn(x, sd, "backbone.bn1.", cfg))


def patch_embed(img: Tensor, sd, cfg: PathConfig) -> Tuple[Tensor, Tuple[int, int]]:
    x = img[:, 0:4]
    H, W = x.shape[2:]
    if H % 4:
        x = F.pad(x, (0, 0, 0, 4 - H % 4))
    if W % 4:
        x = F.pad(x, (0, 4 - W % 4, 0, 0))
    x = F.conv2d(x, sd["backbone.patch_embed.projection.weight"],
                 sd["backbone.patch_embed.projection.bias"], stride=4)
    hw = (x.shape[2], x.shape[3])
    x = x.flatten(2).transpose(1, 2)
    return _ln(x, sd, "backbone.patch_embed.norm.", cfg), hw


def relative_position_index(ws: int = 7) -> Tensor:
    """Official Swin formula; equals the reference buffer (depthformer_swin.py:168-172)."""
    coords = torch.stack(torch.meshgrid(torch.arange(ws), torch.arange(ws), indexing="ij")).flatten(1)
    rel = (coords[:, :, None] - coords[:, None, :]).permute(1, 2, 0).contiguous()
    rel[:, :, 0] += ws - 1
    rel[:, :, 1] += ws - 1
    rel[:, :, 0] *= 2 * ws - 1
    return rel.sum(-1)


def shift_mask(Hp: int, Wp: int, ws: int, shift: int) -> Tensor:
    """(nW, N, N) additive mask with 0 / -100 (depthformer_swin.py:304-326)."""
    img_mask = torch.zeros(1, Hp, Wp, 1)
    cnt = 0
    for h in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
        for w in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
            img_mask[:, h, w, :] = cnt
            cnt += 1
    mw = _partition(img_mask, ws).view(-1, ws * ws)
    m = mw.unsqueeze(1) - mw.unsqueeze(2)
    return m.masked_fill(m != 0, -100.0).masked_fill(m == 0, 0.0)


def _partition(x: Tensor, ws: int) -> Tensor:
    B, H, W, C = x.shape
    x = x.view(B, H // ws, ws, W // ws, ws, C)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(-1, ws, ws, C)


def _reverse(win: Tensor, ws: int, H: int, W: int) -> Tensor:
    B = win.shape[0] // ((H // ws) * (W // ws))
    x = win.view(B, H // ws, W // ws, ws, ws, -1)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(B, H, W, -1)


def window_msa(x: Tensor, sd, p: str, nH: int, mask: Optional[Tensor]) -> Tensor:
    Bw, N, C = x.shape
    hd = C // nH
    qkv = _lin(x, sd, p + "qkv.").reshape(Bw, N, 3, nH, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0] * hd ** -0.5, qkv[1], qkv[2]
    attn = q @ k.transpose(-2, -1)
    idx = sd[p + "relative_position_index"].view(-1).long()
    bias = sd[p + "relative_position_bias_table"][idx].view(N, N, nH).permute(2, 0, 1)
    attn = attn + bias.unsqueeze(0)
    if mask is not None:
        nW = mask.shape[0]
        attn = attn.view(Bw // nW, nW, nH, N, N) + mask.unsqueeze(1).unsqueeze(0)
        attn = attn.view(-1, nH, N, N)
    attn = attn.softmax(-1)
    x = (attn @ v).transpose(1, 2).reshape(Bw, N, C)
    return _lin(x, sd, p + "proj.")


def shift_window_msa(x: Tensor, hw, sd, p: str, nH: int, ws: int, shift: int) -> Tensor:
    B, L, C = x.shape
    H, W = hw
    x = x.view(B, H, W, C)
    pad_r, pad_b = (ws - W % ws) % ws, (ws - H % ws) % ws
    x = F.pad(x, (0, 0, 0, pad_r, 0, pad_b))
    Hp, Wp = x.shape[1], x.shape[2]
    mask = None
    if shift > 0:
        x = torch.roll(x, shifts=(-shift, -shift), dims=(1, 2))
        mask = shift_mask(Hp, Wp, ws, shift).to(x.dtype)
    win = _partition(x, ws).view(-1, ws * ws, C)
    win = window_msa(win, sd, p + "w_msa.", nH, mask)
    x = _reverse(win.view(-1, ws, ws, C), ws, Hp, Wp)
    if shift > 0:
        x = torch.roll(x, shifts=(shift, shift), dims=(1, 2))
    return x[:, :H, :W, :].contiguous().view(B, H * W, C)


def swin_block(x, hw, sd, p, nH, ws, shift, cfg):
    x = x + shift_window_msa(_ln(x, sd, p + "norm1.", cfg), hw, sd, p + "attn.", nH, ws, shift)
    h = _ln(x, sd, p + "norm2.", cfg)
    h = _lin(F.gelu(_lin(h, sd, p + "ffn.layers.0.0.")), sd, p + "ffn.layers.1.")
    return x + h


def patch_merging(x, hw, sd, p, cfg):
    B, L, C = x.shape
    H, W = hw
    x = x.view(B, H, W, C).permute(0, 3, 1, 2)
    if H % 2 or W % 2:
        x = F.pad(x, (0, W % 2, 0, H % 2))
    x = F.unfold(x, kernel_size=2, stride=2).transpose(1, 2)   # channel order: c-major then (kh,kw)
    x = _lin(_ln(x, sd, p + "norm.", cfg), sd, p + "reduction.")
    return x, ((H + 1) // 2, (W + 1) // 2)


def backbone(img: Tensor, sd, cfg: PathConfig) -> List[Tensor]:
    outs = [stem(img, sd, cfg)]
    x, hw = patch_embed(img, sd, cfg)
    C = cfg.embed_dims
    for i, depth in enumerate(cfg.depths):
        for j in range(depth):
            x = swin_block(x, hw, sd, f"backbone.stages.{i}.blocks.{j}.", cfg.num_heads[i],
                           cfg.window_size, cfg.window_size // 2 if j % 2 else 0, cfg)
        out, out_hw = x, hw
        if i < len(cfg.depths) - 1:
            x, hw = patch_merging(x, hw, sd, f"backbone.stages.{i}.downsample.", cfg)
        o = _ln(out, sd, f"backbone.norm{i}.", cfg)
        outs.append(o.view(-1, out_hw[0], out_hw[1], C * 2 ** i).permute(0, 3, 1, 2).contiguous())
    return outs


# ---------------------------------------------------------------------------------------------
# HAHI neck
# ---------------------------------------------------------------------------------------------
def sine_pe(H: int, W: int, num_feats: int = 256, temperature: float = 10000.0) -> Tensor:
    """(1, 2*num_feats, H, W); normalize=False so y_embed = 1..H, x_embed = 1..W."""
    y = torch.arange(1, H + 1, dtype=torch.float32).view(H, 1).expand(H, W)
    x = torch.arange(1, W + 1, dtype=torch.float32).view(1, W).expand(H, W)
    dim_t = torch.arange(num_feats, dtype=torch.float32)
    dim_t = temperature ** (2 * (dim_t // 2) / num_feats)
    px, py = x[:, :, None] / dim_t, y[:, :, None] / dim_t
    px = torch.stack((px[:, :, 0::2].sin(), px[:, :, 1::2].cos()), dim=3).view(H, W, -1)
    py = torch.stack((py[:, :, 0::2].sin(), py[:, :, 1::2].cos()), dim=3).view(H, W, -1)
    return torch.cat((py, px), dim=2).permute(2, 0, 1).unsqueeze(0)


def msda_core(value: Tensor, shapes: Sequence[Tuple[int, int]], loc: Tensor, w: Tensor) -> Tensor:
    """multi_scale_deformable_attn_pytorch [mmcv 1.3.x, external].
    value (B, S, nH, hd); loc (B, Q, nH, L, P, 2) in [0,1]; w (B, Q, nH, L, P) -> (B, Q, nH*hd)."""
    B, S, nH, hd = value.shape
    _, Q, _, L, P, _ = loc.shape
    vals = value.split([h * w_ for h, w_ in shapes], dim=1)
    grids = 2 * loc - 1
    samp = []
    for lvl, (h, w_) in enumerate(shapes):
        v = vals[lvl].flatten(2).transpose(1, 2).reshape(B * nH, hd, h, w_)
        g = grids[:, :, :, lvl].transpose(1, 2).flatten(0, 1)          # (B*nH, Q, P, 2)
        samp.append(F.grid_sample(v, g, mode="bilinear", padding_mode="zeros", align_corners=False))
    w = w.transpose(1, 2).reshape(B * nH, 1, Q, L * P)
    out = (torch.stack(samp, dim=-2).flatten(-2) * w).sum(-1).view(B, nH * hd, Q)
    return out.transpose(1, 2).contiguous()


def msda(query, query_pos, value, ref, shapes, sd, p, cfg: PathConfig) -> Tensor:
    """MultiScaleDeformableAttention.forward(batch_first=True), dropout disabled (parity runs),
    identity = query.  ref: (B, Q, L, 2)."""
    B, Q, E = query.shape
    nH, L, P = cfg.msda_heads, cfg.msda_levels, cfg.msda_points
    identity = query
    q = query + query_pos
    v = _lin(value, sd, p + "value_proj.").view(B, -1, nH, E // nH)
    off = _lin(q, sd, p + "sampling_offsets.").view(B, Q, nH, L, P, 2)
    aw = _lin(q, sd, p + "attention_weights.").view(B, Q, nH, L * P).softmax(-1).view(B, Q, nH, L, P)
    norm = torch.tensor([[w_, h] for h, w_ in shapes], dtype=query.dtype)
    loc = ref[:, :, None, :, None, :] + off / norm[None, None, None, :, None, :]
    out = msda_core(v, shapes, loc, aw)
    return _lin(out, sd, p + "output_proj.") + identity


def hahi(feats: List[Tensor], sd, cfg: PathConfig) -> List[Tensor]:
    lat = [_conv_bn_relu(f, sd, f"neck.lateral_convs.{i}.", cfg) for i, f in enumerate(feats)]
    conv_f, trans = lat[0], lat[1:]
    B = conv_f.shape[0]
    shapes = [(t.shape[2], t.shape[3]) for t in trans]
    src, pos, refs = [], [], []
    for i, t in enumerate(trans):
        h, w = shapes[i]
        pe = sine_pe(h, w, cfg.embedding_dim // 2).flatten(2).transpose(1, 2)
        pos.append((pe + sd["neck.level_embed"][i].view(1, 1, -1)).expand(B, -1, -1))
        src.append(_conv_bn_relu(t, sd, f"neck.trans_proj.{i}.", cfg).flatten(2).transpose(1, 2))
        ry, rx = torch.meshgrid(torch.linspace(0.5, h - 0.5, h), torch.linspace(0.5, w - 0.5, w),
                                indexing="ij")
        refs.append(torch.stack((rx.reshape(-1) / w, ry.reshape(-1) / h), -1))
    src, pos = torch.cat(src, 1), torch.cat(pos, 1)
    ref = torch.cat(refs, 0)[None, :, None, :].expand(B, -1, len(shapes), -1)
    src = msda(src, pos, src, ref, shapes, sd, "neck.self_attn.", cfg)

    skip = _conv_bn_relu(conv_f, sd, "neck.conv_proj.0.", cfg)
    _, c, h, w = skip.shape
    q = skip.flatten(2).transpose(1, 2)
    qpe = sine_pe(h, w, cfg.embedding_dim // 2).flatten(2).transpose(1, 2).expand(B, -1, -1)
    rp = _lin(qpe, sd, "neck.reference_points.").sigmoid()
    rp = rp[:, :, None, :].expand(-1, -1, len(shapes), -1)
    fused = msda(q, qpe, src, rp, shapes, sd, "neck.multi_att.", cfg)
    fused = fused.permute(0, 2, 1).reshape(B, c, h, w)
    outs = [_conv_bn_relu(torch.cat([fused, conv_f], 1), sd, "neck.conv_fusion.0.", cfg, padding=1)]
    start = 0
    for i, t in enumerate(trans):
        hh, ww = shapes[i]
        f = src[:, start:start + hh * ww].permute(0, 2, 1).reshape(B, cfg.embedding_dim, hh, ww)
        start += hh * ww
        outs.append(_conv_bn_relu(torch.cat([t, f], 1), sd, f"neck.trans_fusion.{i}.", cfg, padding=1))
    return outs


# ---------------------------------------------------------------------------------------------
# PE necks, ground embedding, decoder, losses
# ---------------------------------------------------------------------------------------------
def pe_neck(x: List[Tensor], sd, p: str) -> Tensor:
    """Sum of 3x3 convs upsampled (align_corners=True) to the stem grid, then convfinal."""
    x0, x1, x2, x3, x4 = x[::-1]
    size = x4.shape[2:]
    acc = None
    for i, t in enumerate((x0, x1, x2, x3)):
        t = F.conv2d(t, sd[f"{p}conv{i}.weight"], sd[f"{p}conv{i}.bias"], padding=1)
        t = F.interpolate(t, size=size, mode="bilinear", align_corners=True)
        acc = t if acc is None else acc + t
    acc = acc + F.conv2d(x4, sd[f"{p}conv4.weight"], sd[f"{p}conv4.bias"], padding=1)
    return F.conv2d(acc, sd[f"{p}convfinal.weight"], sd[f"{p}convfinal.bias"], padding=1)


def ground_embed_vanilla(img: Tensor, y_half: Tensor) -> Tuple[Tensor, Tensor]:
    """encoder_decoder.py:112-123.  Note the literal 200 (not depth_scale)."""
    y = F.interpolate(y_half, size=img.shape[2:], mode="bilinear")
    return y, img[:, 3:4] * y * 200


def ground_embed_adaptive(img: Tensor, y_half: Tensor, logits_half: Tensor, depth_scale: float,
                          height=1.65) -> Tuple[Tensor, Tensor, Tensor]:
    """encoder_decoder.py:79-102,112-117.  height: float or (B,) tensor.  Returns
    (y, pe_mask, logits_full)."""
    y = F.interpolate(y_half, size=img.shape[2:], mode="bilinear")
    pe = img[:, 4:5]
    logits = F.interpolate(logits_half, size=img.shape[2:], mode="bilinear")
    idx = torch.linspace(-5, 5, 11).view(1, 11, 1, 1)
    k = torch.tan(torch.deg2rad((logits.softmax(1) * idx).sum(1, keepdim=True)))
    h = height.view(-1, 1, 1, 1) if torch.is_tensor(height) else height
    a = -h / (pe + 1e-8)
    off = -h / ((a - k) + 1e-8)
    m = off.detach().clone()
    m[m < 0] = 0
    m[m > depth_scale] = 0
    m[m > 0] = 1
    return y, (off * m) * y, logits


def dense_depth_head(x: List[Tensor], sd, cfg: PathConfig) -> Tensor:
    feats = x[::-1]
    t = F.conv2d(feats[0], sd["decode_head.conv_list.0.conv.weight"],
                 sd["decode_head.conv_list.0.conv.bias"])
    for i in range(1, len(feats)):
        skip = feats[i]
        up = F.interpolate(t, size=skip.shape[2:], mode="bilinear", align_corners=True)
        t = torch.cat([up, skip], 1)
        for n in ("convA", "convB"):
            p = f"decode_head.conv_list.{i}.{n}.conv."
            t = F.leaky_relu(F.conv2d(t, sd[p + "weight"], sd[p + "bias"], padding=1), cfg.leaky_slope)
    return t


def fuse_head(feat: Tensor, pe_mask: Tensor, y: Tensor, sd, cfg: PathConfig) -> Tuple[Tensor, Tensor]:
    """decode_head.py:489-508: out = relu(conv)*(1-y_h) + pe_h + min_depth."""
    d = F.relu(F.conv2d(feat, sd["decode_head.conv_depth.weight"], sd["decode_head.conv_depth.bias"],
                        padding=1))
    pe_h = F.interpolate(pe_mask, size=d.shape[2:], mode="bilinear", align_corners=True)
    y_h = F.interpolate(y, size=d.shape[2:], mode="bilinear", align_corners=True)
    return d * (1 - y_h) + pe_h + cfg.min_depth, y_h


def silog(pred: Tensor, gt: Tensor, eps: float = 1e-3, lam: float = 0.15) -> Tensor:
    m = gt > 0
    g = torch.log(pred[m] + eps) - torch.log(gt[m] + eps)
    return torch.sqrt(torch.var(g) + lam * torch.mean(g) ** 2)


# ---------------------------------------------------------------------------------------------
# whole path
# ---------------------------------------------------------------------------------------------
def forward_features(sd: Dict[str, Tensor], cfg: PathConfig, img: Tensor, height=1.65):
    x = hahi(backbone(img, sd, cfg), sd, cfg)
    y_half = torch.sigmoid(pe_neck(x, sd, "pe_mask_neck."))
    logits = None
    if cfg.adaptive:
        y, pe_mask, logits = ground_embed_adaptive(img, y_half, pe_neck(x, sd, "dynamic_pe_neck."),
                                                   cfg.depth_scale, height)
    else:
        y, pe_mask = ground_embed_vanilla(img, y_half)
    out, y_h = fuse_head(dense_depth_head(x, sd, cfg), pe_mask, y, sd, cfg)
    return dict(x=x, y_half=y_half, y=y, pe_mask=pe_mask, logits=logits, depth=out, y_h=y_h)


def forward_train(sd, cfg: PathConfig, img, depth_gt, pe_k_gt=None, height=1.65):
    r = forward_features(sd, cfg, img, height)
    pred = F.interpolate(r["depth"], size=depth_gt.shape[2:], mode="bilinear", align_corners=True)
    losses = {"decode.loss_depth": cfg.sig_weight * silog(pred, depth_gt)}
    if cfg.adaptive:
        losses["decode.loss_dynamic_pe"] = cfg.ce_weight * F.cross_entropy(
            r["logits"], pe_k_gt.long(), ignore_index=255)
    r["losses"] = losses
    r["loss"] = sum(losses.values())
    return r


def forward_test(sd, cfg: PathConfig, img, height=1.65) -> Tensor:
    """encode_decode (encoder_decoder.py:126-139): clamp then resize to the input size."""
    r = forward_features(sd, cfg, img, height)
    out = torch.clamp(r["depth"], min=cfg.min_depth, max=cfg.max_depth)
    return F.interpolate(out, size=img.shape[2:], mode="bilinear", align_corners=True)
