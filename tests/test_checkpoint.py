"""Pretrained-checkpoint interop: official Swin keys/layouts -> DepthFormerSwin, on CPU.
Checked against the reference's own converter when /root/reference is mounted, and through the property the
conversion exists for: official-order patch merging with official weights == Unfold-order merging with the
converted weights."""
import importlib.util
import os

import pytest
import torch

from gedepth_b200 import checkpoint as ck

REF = "/root/reference/depth/models/utils/ckpt_convert.py"


def _official_like(C=8, heads=2, window=7, depth=2, seed=0, index_window=None):
    index_window = index_window or window
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)
    sd = {"patch_embed.proj.weight": r(C, 3, 4, 4), "patch_embed.proj.bias": r(C), "patch_embed.norm.weight": r(C),
          "patch_embed.norm.bias": r(C), "norm.weight": r(4 * C), "head.weight": r(10, 4 * C), "head.bias": r(10)}
    for j in range(depth):
        p = f"layers.0.blocks.{j}."
        sd.update({p + "norm1.weight": r(C), p + "attn.qkv.weight": r(3 * C, C), p + "attn.qkv.bias": r(3 * C),
                   p + "attn.proj.weight": r(C, C), p + "attn.relative_position_bias_table": r((2 * window - 1) ** 2, heads),
                   p + "attn.relative_position_index": torch.zeros(index_window ** 2, index_window ** 2, dtype=torch.long),
                   p + "mlp.fc1.weight": r(4 * C, C), p + "mlp.fc2.weight": r(C, 4 * C), p + "mlp.fc2.bias": r(C)})
    sd["layers.0.blocks.1.attn_mask"] = r(4, 49, 49)
    sd.update({"layers.0.downsample.reduction.weight": r(2 * C, 4 * C), "layers.0.downsample.norm.weight": r(4 * C),
               "layers.0.downsample.norm.bias": r(4 * C)})
    return sd


def test_key_mapping_and_head_dropped():
    out = ck.convert_official_swin(_official_like())
    assert "patch_embed.projection.weight" in out and not any(k.startswith("head") for k in out)
    assert "stages.0.blocks.0.attn.w_msa.qkv.weight" in out
    assert "stages.0.blocks.1.ffn.layers.0.0.weight" in out and "stages.0.blocks.1.ffn.layers.1.bias" in out
    assert "stages.0.blocks.1.attn_mask" in out              # left alone (ignored by strict=False), as the reference does
    assert "stages.0.downsample.reduction.weight" in out and "norm.weight" in out


@pytest.mark.skipif(not os.path.exists(REF), reason="reference tree not mounted")
def test_matches_reference_converter():
    spec = importlib.util.spec_from_file_location("ref_ckpt_convert", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    src = _official_like(seed=3)
    want, got = mod.swin_convert(dict(src)), ck.convert_official_swin(src)
    assert list(want.keys()) == list(got.keys())
    for k in want:
        assert torch.equal(want[k], got[k]), k


def test_patch_merging_order_property():
    """cat[x00,x10,x01,x11] -> LN -> Linear with official weights == Unfold(2,2) order with converted weights."""
    C, H, W = 6, 4, 6
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, H, W, C, generator=g)
    w, nw, nb = torch.randn(2 * C, 4 * C, generator=g), torch.randn(4 * C, generator=g), torch.randn(4 * C, generator=g)
    off = torch.cat([x[:, 0::2, 0::2], x[:, 1::2, 0::2], x[:, 0::2, 1::2], x[:, 1::2, 1::2]], -1).reshape(2, -1, 4 * C)
    y_off = torch.nn.functional.layer_norm(off, (4 * C,), nw, nb) @ w.t()
    conv = ck.convert_official_swin({"layers.0.downsample.reduction.weight": w, "layers.0.downsample.norm.weight": nw,
                                     "layers.0.downsample.norm.bias": nb})
    unf = torch.nn.Unfold(2, stride=2)(x.permute(0, 3, 1, 2)).transpose(1, 2)           # feature = c*4 + ky*2 + kx
    y_unf = torch.nn.functional.layer_norm(unf, (4 * C,), conv["stages.0.downsample.norm.weight"],
                                           conv["stages.0.downsample.norm.bias"]) @ conv["stages.0.downsample.reduction.weight"].t()
    torch.testing.assert_close(y_unf, y_off, rtol=1e-5, atol=1e-5)


def test_load_into_backbone_pads_pe_channel_and_resizes_bias_table():
    import gedepth_b200.models as M          # noqa: F401  (registers modules)
    from gedepth_b200.builder import build_backbone
    bb = build_backbone(dict(type="DepthFormerSwin", embed_dims=8, depths=(2, 2), num_heads=(2, 4), out_indices=(0, 1),
                             strides=(4, 2), window_size=7, USEPE=True, num_stages=0, pretrained=None))
    src = _official_like(C=8, heads=2, window=5, index_window=7)            # window-5 table: must be resized to 13x13
    before = {k: v.clone() for k, v in bb.state_dict().items()}
    res = ck.load_swin_pretrained(bb, {"model": src})
    sd = bb.state_dict()
    w = sd["patch_embed.projection.weight"]
    assert w.shape == (8, 4, 4, 4) and torch.equal(w[:, :3], src["patch_embed.proj.weight"]) and float(w[:, 3].abs().sum()) == 0
    assert sd["stages.0.blocks.0.attn.w_msa.relative_position_bias_table"].shape == (169, 2)
    assert not torch.equal(sd["stages.0.blocks.0.attn.w_msa.relative_position_bias_table"],
                           before["stages.0.blocks.0.attn.w_msa.relative_position_bias_table"])
    assert torch.equal(sd["stages.0.blocks.1.ffn.layers.1.bias"], src["layers.0.blocks.1.mlp.fc2.bias"])
    assert "stages.0.blocks.1.attn_mask" in res.unexpected_keys and any(k.startswith("stages.1") for k in res.missing_keys)
