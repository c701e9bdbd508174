"""ORACLE tooling (build container only): execute the reference's OWN hot-path files from
/root/reference on CPU, under stubs for the absent mmcv / IPython / matplotlib packages.

Nothing here is copied from the reference: its files are loaded by path with importlib, verbatim,
from where they lie.  The stubs are ours (they restate mmcv 1.3.x leaf behaviour), so this pins the
oracle to the reference's wiring and arithmetic above the mmcv leaves.
/root/reference does not exist on the GPU box; only oracle/make_golden.py (run here) uses this.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import torch
import torch.nn as nn

REF = os.environ.get("GEDEPTH_REFERENCE", "/root/reference")
_LOADED = {}


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "depth", "models"))


def _mod(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _install_stubs():
    from gedepth_b200 import compat
    from oracle import model as om

    # arithmetic-bearing mmcv leaves: independent restatements (oracle/mmcv_stubs.py), NOT the product's compat.py
    from oracle import mmcv_stubs as ms
    FFN = ms.make_ffn(compat.BaseModule, compat.Sequential)

    class MultiScaleDeformableAttention(compat.BaseModule):
        """Pure-PyTorch (grid_sample) stand-in for mmcv.ops.MultiScaleDeformableAttention."""

        def __init__(self, embed_dims=256, num_heads=8, num_levels=4, num_points=4, im2col_step=64,
                     dropout=0.1, batch_first=False, norm_cfg=None, init_cfg=None):
            super().__init__(init_cfg)
            self.embed_dims, self.num_heads = embed_dims, num_heads
            self.num_levels, self.num_points = num_levels, num_points
            self.batch_first = batch_first
            self.dropout = nn.Dropout(dropout)
            self.sampling_offsets = nn.Linear(embed_dims, num_heads * num_levels * num_points * 2)
            self.attention_weights = nn.Linear(embed_dims, num_heads * num_levels * num_points)
            self.value_proj = nn.Linear(embed_dims, embed_dims)
            self.output_proj = nn.Linear(embed_dims, embed_dims)

        def init_weights(self):
            pass

        def forward(self, query, key=None, value=None, identity=None, query_pos=None,
                    key_padding_mask=None, reference_points=None, spatial_shapes=None,
                    level_start_index=None, **kw):
            assert self.batch_first
            value = query if value is None else value
            identity = query if identity is None else identity
            if query_pos is not None:
                query = query + query_pos
            bs, nq, _ = query.shape
            nH, L, P = self.num_heads, self.num_levels, self.num_points
            v = self.value_proj(value).view(bs, value.shape[1], nH, -1)
            off = self.sampling_offsets(query).view(bs, nq, nH, L, P, 2)
            aw = self.attention_weights(query).view(bs, nq, nH, L * P).softmax(-1).view(bs, nq, nH, L, P)
            norm = torch.stack([spatial_shapes[..., 1], spatial_shapes[..., 0]], -1).to(query.dtype)
            loc = reference_points[:, :, None, :, None, :] + off / norm[None, None, None, :, None, :]
            shapes = [(int(h), int(w)) for h, w in spatial_shapes.tolist()]
            out = om.msda_core(v, shapes, loc, aw)
            return self.dropout(self.output_proj(out)) + identity

    def _identity_deco(*a, **k):
        def deco(f):
            return f
        return deco

    mmcv = _mod("mmcv", __version__="1.3.13",
                imdenormalize=lambda img, mean, std, to_bgr=True: img)
    cnn = _mod("mmcv.cnn", MODELS=compat.MMCV_MODELS, ConvModule=ms.ConvModule,
               build_norm_layer=ms.build_norm_layer, build_conv_layer=ms.build_conv_layer,
               build_activation_layer=ms.build_activation_layer,
               trunc_normal_init=compat.trunc_normal_init, xavier_init=compat.xavier_init,
               constant_init=compat.constant_init, kaiming_init=compat.kaiming_init)
    bricks = _mod("mmcv.cnn.bricks")
    registry = _mod("mmcv.cnn.bricks.registry", ATTENTION=compat.Registry("attention"))
    transformer = _mod("mmcv.cnn.bricks.transformer", FFN=FFN, build_dropout=ms.build_dropout,
                       POSITIONAL_ENCODING=compat.POSITIONAL_ENCODING,
                       build_positional_encoding=compat.build_positional_encoding)
    utils_ = _mod("mmcv.cnn.utils")
    winit = _mod("mmcv.cnn.utils.weight_init", constant_init=compat.constant_init,
                 trunc_normal_init=compat.trunc_normal_init, xavier_init=compat.xavier_init)
    runner = _mod("mmcv.runner", BaseModule=compat.BaseModule, ModuleList=compat.ModuleList,
                  Sequential=compat.Sequential, auto_fp16=_identity_deco, force_fp32=_identity_deco,
                  _load_checkpoint=lambda *a, **k: (_ for _ in ()).throw(RuntimeError("no ckpt")))
    base_module = _mod("mmcv.runner.base_module", BaseModule=compat.BaseModule,
                       ModuleList=compat.ModuleList, Sequential=compat.Sequential)
    mutils = _mod("mmcv.utils", Registry=compat.Registry)
    ops = _mod("mmcv.ops")
    msda = _mod("mmcv.ops.multi_scale_deform_attn",
                MultiScaleDeformableAttention=MultiScaleDeformableAttention)
    mmcv.cnn, mmcv.runner, mmcv.utils, mmcv.ops = cnn, runner, mutils, ops
    cnn.bricks, cnn.utils = bricks, utils_
    bricks.registry, bricks.transformer = registry, transformer
    utils_.weight_init = winit
    runner.base_module = base_module
    ops.multi_scale_deform_attn = msda
    _mod("IPython", embed=lambda *a, **k: None)
    mpl = _mod("matplotlib")
    mpl.cm = _mod("matplotlib.cm", get_cmap=lambda *a, **k: None)
    mpl.pyplot = _mod("matplotlib.pyplot")

    # package shells of the reference (their __init__ files import far too much)
    import torch.nn.functional as F

    def resize(input, size=None, scale_factor=None, mode="nearest", align_corners=None, warning=True):
        return F.interpolate(input, size, scale_factor, mode, align_corners)

    def add_prefix(inputs, prefix):
        return {f"{prefix}.{k}": v for k, v in inputs.items()}

    depth = _mod("depth", __path__=[])
    _mod("depth.ops", resize=resize)
    _mod("depth.utils", get_root_logger=lambda *a, **k: None, colorize=None, __path__=[])
    _mod("depth.core", add_prefix=add_prefix)
    _mod("depth.models", __path__=[])
    _mod("depth.models.utils", __path__=[])
    _mod("depth.models.backbones", __path__=[])
    _mod("depth.models.backbones.resnet", BasicBlock=None, Bottleneck=None, ResNet=None)
    _mod("depth.models.backbones.hrnet", Bottleneck=None, ResLayer=None)
    _mod("depth.models.necks", __path__=[])
    _mod("depth.models.losses", __path__=[])
    _mod("depth.models.decode_heads", __path__=[])
    _mod("depth.models.decode_heads.pac", __path__=[])
    _mod("depth.models.depther", __path__=[])
    torch.cuda.current_device = lambda: "cpu"   # encoder_decoder.py:68


def _load(modname: str, relpath: str):
    if modname in _LOADED:
        return _LOADED[modname]
    spec = importlib.util.spec_from_file_location(modname, os.path.join(REF, relpath))
    m = importlib.util.module_from_spec(spec)
    sys.modules[modname] = m
    spec.loader.exec_module(m)
    _LOADED[modname] = m
    parent, _, child = modname.rpartition(".")
    if parent in sys.modules:
        setattr(sys.modules[parent], child, m)
    return m


def load_reference():
    """Returns the reference's ``builder`` module with every hot-path class registered."""
    if "builder" in _LOADED:
        return _LOADED["builder"]
    if not available():
        raise RuntimeError(f"reference tree not found at {REF}")
    _install_stubs()
    builder = _load("depth.models.builder", "depth/models/builder.py")
    sys.modules["depth.models"].builder = builder
    sys.modules["depth.models"].depther = sys.modules["depth.models.depther"]
    embed = _load("depth.models.utils.embed", "depth/models/utils/embed.py")
    ckpt = _load("depth.models.utils.ckpt_convert", "depth/models/utils/ckpt_convert.py")
    mu = sys.modules["depth.models.utils"]
    mu.PatchEmbedSwin, mu.swin_convert = embed.PatchEmbedSwin, ckpt.swin_convert
    mu.ResLayer = mu.UpConvBlock = mu.BasicConvBlock = None
    pac = sys.modules["depth.models.decode_heads.pac"]
    pac.packernel2d = None
    _load("depth.utils.position_encoding", "depth/utils/position_encoding.py")
    _load("depth.models.backbones.depthformer_swin", "depth/models/backbones/depthformer_swin.py")
    _load("depth.models.necks.hahi", "depth/models/necks/hahi.py")
    _load("depth.models.necks.pemask_neck", "depth/models/necks/pemask_neck.py")
    _load("depth.models.necks.dynamicpe_neck", "depth/models/necks/dynamicpe_neck.py")
    _load("depth.models.losses.sigloss", "depth/models/losses/sigloss.py")
    _load("depth.models.losses.celoss", "depth/models/losses/celoss.py")
    _load("depth.models.losses.bceloss", "depth/models/losses/bceloss.py")
    _load("depth.models.decode_heads.decode_head", "depth/models/decode_heads/decode_head.py")
    _load("depth.models.decode_heads.densedepth_head", "depth/models/decode_heads/densedepth_head.py")
    _load("depth.models.depther.base", "depth/models/depther/base.py")
    _load("depth.models.depther.encoder_decoder", "depth/models/depther/encoder_decoder.py")
    _LOADED["builder"] = builder
    return builder


def build_reference_model(model_cfg: dict):
    """``build_depther`` of the reference on a config dict (e.g. Config.fromfile(...).model)."""
    from gedepth_b200.compat import _to_cfgdict
    builder = load_reference()
    return builder.build_depther(_to_cfgdict(model_cfg))
