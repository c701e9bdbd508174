"""mmcv-free shell for the one GEDepth path: Registry, Config, ConvModule, FFN, DropPath.

The reference is an mmcv-1.3.x code base (`docs/install.md:26`) and mmcv is absent here, so the
handful of mmcv behaviours the path touches are restated ("mmcv pieces"):

* ``Registry.register_module()/build()``     <- depth/models/builder.py:8-44 (mmcv.utils.Registry)
* ``Config.fromfile`` with ``_base_`` merge   <- configs/depthformer/*.py (mmcv.utils.Config)
* ``ConvModule`` conv -> norm -> act          <- hahi.py:122-165, densedepth_head.py:21-22,82-89
* ``FFN`` / ``DropPath``                      <- depthformer_swin.py:283,451-459

Only what ``configs/depthformer/*.py`` exercises is implemented.
"""
from __future__ import annotations

import copy
import importlib.util
import math
import os
import sys
import types
from typing import Any, Callable, Dict, Optional

import torch
import torch.nn as nn


# --------------------------------------------------------------------------------------------
# ConfigDict / Config
# --------------------------------------------------------------------------------------------
class ConfigDict(dict):
    """dict with attribute access (mmcv.utils.ConfigDict semantics used at
    encoder_decoder.py:44 ``backbone.pretrained = ...`` and :218 ``test_cfg.mode``)."""

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError as e:
            raise AttributeError(name) from e

    def __setattr__(self, name, value):
        self[name] = value

    def __deepcopy__(self, memo):
        return ConfigDict({k: copy.deepcopy(v, memo) for k, v in self.items()})


def _to_cfgdict(obj):
    if isinstance(obj, dict):
        return ConfigDict({k: _to_cfgdict(v) for k, v in obj.items()})
    if isinstance(obj, list):
        return [_to_cfgdict(v) for v in obj]
    if isinstance(obj, tuple):
        return tuple(_to_cfgdict(v) for v in obj)
    return obj


DELETE_KEY = "_delete_"
BASE_KEY = "_base_"


def _merge_a_into_b(a: dict, b: dict) -> dict:
    """Child ``a`` over base ``b``: dicts merge recursively, ``_delete_=True`` replaces
    (used at configs/depthformer/depthformer_v.py:160-161)."""
    b = dict(b)
    for k, v in a.items():
        if isinstance(v, dict) and k in b and isinstance(b[k], dict) and not v.get(DELETE_KEY, False):
            b[k] = _merge_a_into_b(v, b[k])
        else:
            if isinstance(v, dict):
                v = {kk: vv for kk, vv in v.items() if kk != DELETE_KEY}
            b[k] = v
    return b


class Config:
    """Python-file configs with ``_base_`` inheritance. Files are exec'd (list comprehensions and
    ``**img_norm_cfg`` must work), every non-dunder, non-module top-level name becomes a key."""

    def __init__(self, cfg_dict: Optional[dict] = None, filename: Optional[str] = None):
        object.__setattr__(self, "_cfg_dict", _to_cfgdict(cfg_dict or {}))
        object.__setattr__(self, "filename", filename)

    @staticmethod
    def _file2dict(filename: str) -> dict:
        filename = os.path.abspath(os.path.expanduser(filename))
        if not os.path.isfile(filename):
            raise FileNotFoundError(filename)
        with open(filename, "r") as f:
            src = f.read()
        ns: Dict[str, Any] = {"__file__": filename, "__name__": "_gedepth_cfg_"}
        exec(compile(src, filename, "exec"), ns)
        cfg = {
            k: v
            for k, v in ns.items()
            if not k.startswith("__") and not isinstance(v, (types.ModuleType, types.FunctionType))
        }
        if BASE_KEY in cfg:
            bases = cfg.pop(BASE_KEY)
            bases = bases if isinstance(bases, (list, tuple)) else [bases]
            base_cfg: dict = {}
            for b in bases:
                sub = Config._file2dict(os.path.join(os.path.dirname(filename), b))
                dup = base_cfg.keys() & sub.keys()
                if dup:
                    raise KeyError(f"Duplicate key is not allowed among bases: {dup}")
                base_cfg.update(sub)
            cfg = _merge_a_into_b(cfg, base_cfg)
        return cfg

    @staticmethod
    def fromfile(filename: str) -> "Config":
        return Config(Config._file2dict(filename), filename=filename)

    def merge_from_dict(self, options: dict) -> None:
        """``--options a.b.c=v`` style overrides (tools/train.py:87-88)."""
        nested: dict = {}
        for full_key, v in options.items():
            d = nested
            keys = full_key.split(".")
            for k in keys[:-1]:
                d = d.setdefault(k, {})
            d[keys[-1]] = v
        merged = _merge_a_into_b(nested, self._cfg_dict)
        object.__setattr__(self, "_cfg_dict", _to_cfgdict(merged))

    def __getattr__(self, name):
        return getattr(self._cfg_dict, name)

    def __getitem__(self, name):
        return self._cfg_dict[name]

    def __setattr__(self, name, value):
        self._cfg_dict[name] = _to_cfgdict(value)

    def __contains__(self, name):
        return name in self._cfg_dict

    def get(self, key, default=None):
        return self._cfg_dict.get(key, default)

    def keys(self):
        return self._cfg_dict.keys()

    def to_dict(self):
        return copy.deepcopy(dict(self._cfg_dict))


# --------------------------------------------------------------------------------------------
# Registry
# --------------------------------------------------------------------------------------------
class Registry:
    """Name -> class map with parent fallback (mmcv.utils.Registry as used at
    depth/models/builder.py:8-44)."""

    def __init__(self, name: str, parent: Optional["Registry"] = None):
        self._name = name
        self._module_dict: Dict[str, type] = {}
        self.parent = parent

    @property
    def name(self):
        return self._name

    @property
    def module_dict(self):
        return self._module_dict

    def __len__(self):
        return len(self._module_dict)

    def __contains__(self, key):
        return self.get(key) is not None

    def get(self, key: str):
        if key in self._module_dict:
            return self._module_dict[key]
        if self.parent is not None:
            return self.parent.get(key)
        return None

    def _register(self, cls, name=None, force=False):
        names = [name or cls.__name__] if not isinstance(name, (list, tuple)) else list(name)
        for n in names:
            if not force and n in self._module_dict:
                raise KeyError(f"{n} is already registered in {self._name}")
            self._module_dict[n] = cls

    def register_module(self, name=None, force=False, module=None):
        if module is not None:
            self._register(module, name, force)
            return module

        def deco(cls):
            self._register(cls, name, force)
            return cls

        return deco

    def build(self, cfg, default_args: Optional[dict] = None):
        if not isinstance(cfg, dict):
            raise TypeError(f"cfg must be a dict, but got {type(cfg)}")
        if "type" not in cfg and not (default_args and "type" in default_args):
            raise KeyError(f'`cfg` or `default_args` must contain the key "type", but got {cfg}')
        args = dict(cfg)
        if default_args is not None:
            for k, v in default_args.items():
                args.setdefault(k, v)
        obj_type = args.pop("type")
        if isinstance(obj_type, str):
            cls = self.get(obj_type)
            if cls is None:
                raise KeyError(f"{obj_type} is not in the {self._name} registry")
        elif isinstance(obj_type, type):
            cls = obj_type
        else:
            raise TypeError(f"type must be a str or valid type, but got {type(obj_type)}")
        try:
            return cls(**args)
        except Exception as e:  # same re-raise flavour as mmcv.build_from_cfg
            raise type(e)(f"{cls.__name__}: {e}") from e


MMCV_MODELS = Registry("model")
POSITIONAL_ENCODING = Registry("position encoding")
ACTIVATION_LAYERS = Registry("activation layer")
for _act in (nn.ReLU, nn.LeakyReLU, nn.GELU, nn.Sigmoid, nn.Tanh):
    ACTIVATION_LAYERS.register_module(module=_act)


def build_positional_encoding(cfg):
    return POSITIONAL_ENCODING.build(cfg)


def build_activation_layer(cfg):
    return ACTIVATION_LAYERS.build(cfg)


# --------------------------------------------------------------------------------------------
# BaseModule and small builders
# --------------------------------------------------------------------------------------------
class BaseModule(nn.Module):
    """mmcv.runner.BaseModule: ``init_weights`` recurses into children that define it."""

    def __init__(self, init_cfg=None):
        super().__init__()
        self.init_cfg = copy.deepcopy(init_cfg)
        self._is_init = False

    @property
    def is_init(self):
        return self._is_init

    def init_weights(self):
        for m in self.children():
            if hasattr(m, "init_weights"):
                m.init_weights()
        self._is_init = True


class ModuleList(BaseModule, nn.ModuleList):
    def __init__(self, modules=None, init_cfg=None):
        BaseModule.__init__(self, init_cfg)
        nn.ModuleList.__init__(self, modules)


class Sequential(BaseModule, nn.Sequential):
    def __init__(self, *args, init_cfg=None):
        BaseModule.__init__(self, init_cfg)
        nn.Sequential.__init__(self, *args)


def build_norm_layer(cfg: dict, num_features: int, postfix=""):
    """(name, layer). ``LN`` -> ('ln', LayerNorm eps 1e-5); ``BN``/``SyncBN`` -> ('bn', BatchNorm2d
    eps 1e-5 momentum 0.1). The reference never activates SyncBN."""
    cfg = dict(cfg)
    t = cfg.pop("type")
    requires_grad = cfg.pop("requires_grad", True)
    cfg.setdefault("eps", 1e-5)
    if t == "LN":
        layer, abbr = nn.LayerNorm(num_features, **cfg), "ln"
    elif t in ("BN", "BN2d", "SyncBN"):
        layer, abbr = nn.BatchNorm2d(num_features, **cfg), "bn"
    else:
        raise KeyError(f"Unrecognized norm type {t}")
    for p in layer.parameters():
        p.requires_grad = requires_grad
    return abbr + str(postfix), layer


def build_conv_layer(cfg, *args, **kwargs):
    if cfg is not None and cfg.get("type", "Conv2d") not in ("Conv2d", "Conv"):
        raise KeyError(f"Unrecognized conv type {cfg['type']}")
    return nn.Conv2d(*args, **kwargs)


def xavier_init(module, gain=1, bias=0, distribution="normal"):
    if hasattr(module, "weight") and module.weight is not None:
        if distribution == "uniform":
            nn.init.xavier_uniform_(module.weight, gain=gain)
        else:
            nn.init.xavier_normal_(module.weight, gain=gain)
    if hasattr(module, "bias") and module.bias is not None:
        nn.init.constant_(module.bias, bias)


def constant_init(module, val, bias=0):
    if hasattr(module, "weight") and module.weight is not None:
        nn.init.constant_(module.weight, val)
    if hasattr(module, "bias") and module.bias is not None:
        nn.init.constant_(module.bias, bias)


def trunc_normal_init(module, mean=0.0, std=1.0, a=-2.0, b=2.0, bias=0.0):
    t = module if isinstance(module, torch.Tensor) else getattr(module, "weight", None)
    if t is not None:
        nn.init.trunc_normal_(t, mean, std, a, b)
    if not isinstance(module, torch.Tensor) and getattr(module, "bias", None) is not None:
        nn.init.constant_(module.bias, bias)


def kaiming_init(module, a=0, mode="fan_out", nonlinearity="relu", bias=0):
    nn.init.kaiming_normal_(module.weight, a=a, mode=mode, nonlinearity=nonlinearity)
    if module.bias is not None:
        nn.init.constant_(module.bias, bias)


class ConvModule(nn.Module):
    """conv -> norm -> act; ``bias='auto'`` means bias iff there is no norm. Attribute names
    ``conv`` / ``bn`` / ``activate`` give the reference's state_dict keys."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1,
                 groups=1, bias="auto", conv_cfg=None, norm_cfg=None, act_cfg=dict(type="ReLU"),
                 inplace=True):
        super().__init__()
        self.with_norm = norm_cfg is not None
        self.with_activation = act_cfg is not None
        if bias == "auto":
            bias = not self.with_norm
        self.conv = build_conv_layer(conv_cfg, in_channels, out_channels, kernel_size, stride=stride,
                                     padding=padding, dilation=dilation, groups=groups, bias=bias)
        self.norm_name = None
        if self.with_norm:
            self.norm_name, norm = build_norm_layer(norm_cfg, out_channels)
            self.add_module(self.norm_name, norm)
        self.act_kind, self.act_slope = None, 0.0
        if self.with_activation:
            act_cfg_ = dict(act_cfg)
            if act_cfg_["type"] not in ("Tanh", "PReLU", "Sigmoid", "HSigmoid", "Swish", "GELU"):
                act_cfg_.setdefault("inplace", inplace)
            self.activate = build_activation_layer(act_cfg_)
            self.act_kind = act_cfg_["type"]
            self.act_slope = getattr(self.activate, "negative_slope", 0.0)
        self.init_weights()

    @property
    def norm(self):
        return getattr(self, self.norm_name) if self.norm_name else None

    def init_weights(self):
        if self.with_activation and self.act_kind == "LeakyReLU":
            nonlinearity, a = "leaky_relu", self.act_slope
        else:
            nonlinearity, a = "relu", 0
        kaiming_init(self.conv, a=a, nonlinearity=nonlinearity)
        if self.with_norm:
            constant_init(self.norm, 1, bias=0)

    def forward(self, x):
        x = self.conv(x)
        if self.with_norm:
            x = self.norm(x)
        if self.with_activation:
            x = self.activate(x)
        return x


class DropPath(nn.Module):
    """Per-sample stochastic depth: ``x / keep * floor(keep + U[0,1))`` in train mode."""

    def __init__(self, drop_prob=0.1):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1 - self.drop_prob
        shape = (x.shape[0],) + (1,) * (x.ndim - 1)
        r = keep + torch.rand(shape, dtype=x.dtype, device=x.device)
        return x.div(keep) * r.floor()


def build_dropout(cfg):
    cfg = dict(cfg)
    t = cfg.pop("type")
    if t == "DropPath":
        return DropPath(**cfg)
    if t == "Dropout":
        return nn.Dropout(cfg.pop("drop_prob", 0.5))
    raise KeyError(t)
