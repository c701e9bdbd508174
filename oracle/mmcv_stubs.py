"""ORACLE tooling (test infrastructure, never imported by gedepth_b200): stand-ins for the ARITHMETIC-bearing leaves of
mmcv-full 1.3.13 that the reference's files import - written here on their own, independent of the product's
gedepth_b200/compat.py, so that oracle/make_golden.py does not generate the golden fixtures through the product's
restatement of the same classes.  mmcv itself is not installable here (no network); each class cites the mmcv 1.3.13
source file + definition it restates (line numbers as of tag v1.3.13; they cannot be re-checked offline).

  mmcv/cnn/bricks/conv_module.py  ConvModule.__init__ (l. 70-166), forward (l. 188-200)
  mmcv/cnn/bricks/norm.py         build_norm_layer (l. 73-144), infer_abbr (l. 24-70)
  mmcv/cnn/bricks/activation.py   registration of nn.ReLU / nn.LeakyReLU / nn.GELU ... (l. 7-12), build_activation_layer (l. 81-92)
  mmcv/cnn/bricks/drop.py         drop_path (l. 9-25), DropPath (l. 28-43), build_dropout (l. 63-65)
  mmcv/cnn/bricks/transformer.py  FFN.__init__ / forward (l. 349-426)

tests/test_third_party_pins.py checks these against plain-torch statements and against the product's compat.py.
"""
from __future__ import annotations

import torch
import torch.nn as nn

_ACTS = {"ReLU": nn.ReLU, "LeakyReLU": nn.LeakyReLU, "PReLU": nn.PReLU, "RReLU": nn.RReLU, "ReLU6": nn.ReLU6,
         "ELU": nn.ELU, "Sigmoid": nn.Sigmoid, "Tanh": nn.Tanh, "GELU": nn.GELU}


def build_activation_layer(cfg):
    """activation.py: ``build_from_cfg(cfg, ACTIVATION_LAYERS)`` - the registered torch.nn class with the cfg's kwargs."""
    cfg = dict(cfg)
    return _ACTS[cfg.pop("type")](**cfg)


def build_norm_layer(cfg, num_features, postfix=""):
    """norm.py build_norm_layer: (abbr + postfix, layer); eps defaults to 1e-5; requires_grad applied to the parameters.
    BN -> nn.BatchNorm2d ('bn'), SyncBN -> treated as BN on one device, LN -> nn.LayerNorm ('ln')."""
    cfg_ = dict(cfg)
    layer_type = cfg_.pop("type")
    requires_grad = cfg_.pop("requires_grad", True)
    cfg_.setdefault("eps", 1e-5)
    if layer_type in ("BN", "BN2d", "SyncBN"):
        abbr, layer = "bn", nn.BatchNorm2d(num_features, **cfg_)
    elif layer_type == "LN":
        abbr, layer = "ln", nn.LayerNorm(num_features, **cfg_)
    else:
        raise KeyError(f"Unrecognized norm type {layer_type}")
    for param in layer.parameters():
        param.requires_grad = requires_grad
    return abbr + str(postfix), layer


def build_conv_layer(cfg, *args, **kwargs):
    if cfg is not None and cfg.get("type", "Conv2d") not in ("Conv2d", "Conv"):
        raise KeyError(cfg["type"])
    return nn.Conv2d(*args, **kwargs)


class ConvModule(nn.Module):
    """conv_module.py: order ('conv', 'norm', 'act'); ``bias='auto'`` -> ``bias = not with_norm`` (l. 104-106); the norm
    layer is registered under the abbreviation returned by build_norm_layer (``bn``); activations other than Tanh /
    PReLU / Sigmoid / HSigmoid / Swish receive ``inplace`` (l. 150-155)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias="auto",
                 conv_cfg=None, norm_cfg=None, act_cfg=dict(type="ReLU"), inplace=True, with_spectral_norm=False,
                 padding_mode="zeros", order=("conv", "norm", "act")):
        super().__init__()
        assert order == ("conv", "norm", "act") and not with_spectral_norm and padding_mode == "zeros"
        self.with_norm = norm_cfg is not None
        self.with_activation = act_cfg is not None
        if bias == "auto":
            bias = not self.with_norm
        self.with_bias = bias
        self.conv = build_conv_layer(conv_cfg, in_channels, out_channels, kernel_size, stride=stride, padding=padding,
                                     dilation=dilation, groups=groups, bias=bias)
        if self.with_norm:
            self.norm_name, norm = build_norm_layer(norm_cfg, out_channels)
            self.add_module(self.norm_name, norm)
        else:
            self.norm_name = None
        if self.with_activation:
            act_cfg_ = dict(act_cfg)
            if act_cfg_["type"] not in ("Tanh", "PReLU", "Sigmoid", "HSigmoid", "Swish", "GELU"):
                act_cfg_.setdefault("inplace", inplace)
            self.activate = build_activation_layer(act_cfg_)

    @property
    def norm(self):
        return getattr(self, self.norm_name) if self.norm_name else None

    def init_weights(self):
        pass          # the golden runs load a name-keyed deterministic state_dict (gedepth_b200.synth)

    def forward(self, x, activate=True, norm=True):
        x = self.conv(x)
        if norm and self.with_norm:
            x = self.norm(x)
        if activate and self.with_activation:
            x = self.activate(x)
        return x


def drop_path(x, drop_prob=0.0, training=False):
    """drop.py drop_path: per-sample keep mask floor(keep_prob + U[0,1)), output x / keep_prob * mask."""
    if drop_prob == 0.0 or not training:
        return x
    keep_prob = 1 - drop_prob
    shape = (x.shape[0],) + (1,) * (x.ndim - 1)
    random_tensor = keep_prob + torch.rand(shape, dtype=x.dtype, device=x.device)
    return x.div(keep_prob) * random_tensor.floor()


class DropPath(nn.Module):
    def __init__(self, drop_prob=0.1):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        return drop_path(x, self.drop_prob, self.training)


def build_dropout(cfg):
    cfg = dict(cfg)
    t = cfg.pop("type")
    if t == "DropPath":
        return DropPath(**cfg)
    if t == "Dropout":
        return nn.Dropout(cfg.pop("drop_prob", 0.5))
    raise KeyError(t)


def make_ffn(base_module, sequential):
    """FFN needs mmcv's BaseModule / Sequential (scaffolding, no arithmetic): the harness passes the classes it installs."""

    class FFN(base_module):
        """transformer.py FFN: num_fcs - 1 blocks of (Linear, act, Dropout), then Linear and Dropout;
        forward: ``identity + dropout_layer(layers(x))`` when add_identity (identity defaults to x)."""

        def __init__(self, embed_dims=256, feedforward_channels=1024, num_fcs=2, act_cfg=dict(type="ReLU", inplace=True),
                     ffn_drop=0.0, dropout_layer=None, add_identity=True, init_cfg=None, **kwargs):
            super().__init__(init_cfg)
            assert num_fcs >= 2
            layers, in_channels = [], embed_dims
            for _ in range(num_fcs - 1):
                layers.append(sequential(nn.Linear(in_channels, feedforward_channels), build_activation_layer(act_cfg),
                                         nn.Dropout(ffn_drop)))
                in_channels = feedforward_channels
            layers.append(nn.Linear(feedforward_channels, embed_dims))
            layers.append(nn.Dropout(ffn_drop))
            self.layers = sequential(*layers)
            self.dropout_layer = build_dropout(dropout_layer) if dropout_layer else nn.Identity()
            self.add_identity = add_identity

        def forward(self, x, identity=None):
            out = self.layers(x)
            if not self.add_identity:
                return self.dropout_layer(out)
            if identity is None:
                identity = x
            return identity + self.dropout_layer(out)

    return FFN
