"""Pins for the arithmetic the reference takes from mmcv-full 1.3.13 (absent here): the oracle's and the
product's restatements are checked against independent statements - an unrelated third-party implementation found in
the image for deformable attention, plain torch for FFN / ConvModule / DropPath / build_norm_layer.  CPU only."""
import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F


def test_msda_core_equals_transformers_deformable_detr():
    """oracle.model.msda_core (what every MSDA kernel is compared with, through tests/ops_lib.msda_sample) against
    transformers' MultiScaleDeformableAttention - the Deformable-DETR reference implementation, written independently
    of this repo and of mmcv's copy."""
    mod = pytest.importorskip("transformers.models.deformable_detr.modeling_deformable_detr")
    from oracle import model as om
    from tests import ops_lib as L
    g = torch.Generator().manual_seed(0)
    shapes = [(12, 30), (6, 15), (3, 8), (2, 4)]
    S = sum(h * w for h, w in shapes)
    B, Q, nH, hd, Lv, P = 2, 77, 8, 64, 4, 8
    value = torch.randn(B, S, nH, hd, generator=g)
    loc = torch.rand(B, Q, nH, Lv, P, 2, generator=g) * 1.2 - 0.1          # some samples outside [0, 1]: zero padding
    w = torch.softmax(torch.randn(B, Q, nH, Lv * P, generator=g), -1).view(B, Q, nH, Lv, P)
    ours = om.msda_core(value, shapes, loc, w)
    starts = torch.tensor([0] + list(np.cumsum([h * w_ for h, w_ in shapes])[:-1]))
    theirs = mod.MultiScaleDeformableAttention()(value, torch.tensor(shapes), shapes, starts, loc, w, 64)
    assert torch.equal(ours, theirs) or float((ours - theirs).abs().max()) < 1e-6
    # and the (ref, offset, logit) parametrisation the kernels take
    ref = torch.rand(1, Q, 2, generator=g)
    off = torch.randn(B, Q, nH * Lv * P * 2, generator=g) * 2
    logit = torch.randn(B, Q, nH * Lv * P, generator=g)
    norm = torch.tensor([[w_, h] for h, w_ in shapes], dtype=torch.float32)
    loc2 = ref[:, :, None, None, None, :] + off.view(B, Q, nH, Lv, P, 2) / norm[None, None, None, :, None, :]
    w2 = logit.view(B, Q, nH, Lv * P).softmax(-1).view(B, Q, nH, Lv, P)
    a = L.msda_sample(value.view(B, S, nH * hd), shapes, ref, off, logit, nH, P)
    b = mod.MultiScaleDeformableAttention()(value, torch.tensor(shapes), shapes, starts, loc2.expand(B, -1, -1, -1, -1, -1), w2, 64)
    assert float((a - b).abs().max()) < 1e-5


def _both():
    from gedepth_b200 import compat
    from oracle import mmcv_stubs as ms
    return compat, ms


def test_ffn_is_identity_plus_droppath_of_two_linears():
    """mmcv FFN(num_fcs=2, add_identity=True) (transformer.py:349-426): x + DropPath(Linear(GELU(Linear(x))))."""
    compat, ms = _both()
    from gedepth_b200 import swin
    torch.manual_seed(0)
    ffn_o = ms.make_ffn(compat.BaseModule, compat.Sequential)(96, 384, 2, dict(type="GELU"), 0.0, dict(type="DropPath", drop_prob=0.2))
    sd = ffn_o.state_dict()
    assert sorted(sd) == ["layers.0.0.bias", "layers.0.0.weight", "layers.1.bias", "layers.1.weight"]   # ckpt_convert.py:26-34
    x = torch.randn(3, 50, 96)
    ffn_o.eval()
    plain = x + F.linear(F.gelu(F.linear(x, sd["layers.0.0.weight"], sd["layers.0.0.bias"])), sd["layers.1.weight"], sd["layers.1.bias"])
    assert torch.allclose(ffn_o(x), plain, atol=1e-6)
    ident = torch.randn_like(x)
    assert torch.allclose(ffn_o(x, identity=ident), plain - x + ident, atol=1e-6)
    # the product's FFN has the same parameters and constructor
    ffn_p = swin.FFN(96, 384, 2, dict(type="GELU"), 0.0, dict(type="DropPath", drop_prob=0.2))
    assert sorted(ffn_p.state_dict()) == sorted(sd)


@pytest.mark.parametrize("norm_cfg,act_cfg", [(dict(type="BN", requires_grad=True), dict(type="ReLU", inplace=True)),
                                              (None, dict(type="LeakyReLU", inplace=True)), (None, None)])
def test_conv_module_order_bias_and_names(norm_cfg, act_cfg):
    """mmcv ConvModule (conv_module.py:70-200): conv -> norm -> act, bias iff there is no norm, norm registered as `bn`,
    LeakyReLU slope 0.01 - the oracle's stub, the product's compat.ConvModule and plain torch agree."""
    compat, ms = _both()
    torch.manual_seed(1)
    mo = ms.ConvModule(8, 12, 3, padding=1, norm_cfg=norm_cfg, act_cfg=act_cfg)
    mp = compat.ConvModule(8, 12, 3, padding=1, norm_cfg=norm_cfg, act_cfg=act_cfg)
    assert sorted(mo.state_dict()) == sorted(mp.state_dict())
    assert ("conv.bias" in mo.state_dict()) == (norm_cfg is None)
    assert ("bn.weight" in mo.state_dict()) == (norm_cfg is not None)
    mp.load_state_dict(mo.state_dict())
    x = torch.randn(2, 8, 9, 11)
    for train in (True, False):
        mo.train(train); mp.train(train)
        y = F.conv2d(x, mo.conv.weight, mo.conv.bias, padding=1)
        if norm_cfg is not None:
            bn = nn.BatchNorm2d(12, eps=1e-5, momentum=0.1)
            bn.load_state_dict(mo.bn.state_dict())
            bn.train(train)
            y = bn(y)
        if act_cfg is not None:
            y = F.relu(y) if act_cfg["type"] == "ReLU" else F.leaky_relu(y, 0.01)
        assert torch.allclose(mo(x.clone()), y, atol=1e-6) and torch.allclose(mp(x.clone()), y, atol=1e-6)


def test_drop_path_and_norm_builders():
    compat, ms = _both()
    for mod in (ms, compat):
        name, ln = mod.build_norm_layer(dict(type="LN"), 16)
        assert name == "ln" and isinstance(ln, nn.LayerNorm) and ln.eps == 1e-5
        name, bn = mod.build_norm_layer(dict(type="BN", requires_grad=True), 16, postfix=1)
        assert name == "bn1" and isinstance(bn, nn.BatchNorm2d) and bn.eps == 1e-5 and bn.momentum == 0.1
    x = torch.ones(64, 3, 5)
    for mod in (ms, compat):
        torch.manual_seed(2)
        d = mod.build_dropout(dict(type="DropPath", drop_prob=0.25))
        d.train()
        y = d(x)
        kept = y[:, 0, 0]
        assert all(min(abs(float(v)), abs(float(v) - 1 / 0.75)) < 1e-6 for v in kept)          # per-sample, scaled by 1/keep
        assert torch.equal(y, kept.view(-1, 1, 1).expand_as(y))
        d.eval()
        assert torch.equal(d(x), x)
