"""Config loader + registry surface: the reference's configs load unchanged and resolve every name."""
import os

import pytest
import torch

REF_CFG = "/root/reference/configs/depthformer"


def _plain(o):
    if isinstance(o, dict):
        return {k: _plain(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return [_plain(v) for v in o]
    return o


@pytest.mark.skipif(not os.path.isdir(REF_CFG), reason="reference tree only exists in the build container")
@pytest.mark.parametrize("fname,variant,dataset", [("depthformer_v.py", "v", "kitti"), ("depthformer_a.py", "a", "kitti"),
                                                   ("depthformer_v_ddad.py", "v", "ddad"), ("depthformer_a_ddad.py", "a", "ddad")])
def test_reference_configs_load_unchanged(fname, variant, dataset):
    from gedepth_b200.compat import Config
    from gedepth_b200.presets import model_cfg
    cfg = Config.fromfile(os.path.join(REF_CFG, fname))
    assert _plain(dict(cfg.model)) == _plain(model_cfg(variant, dataset))
    # non-model keys the runner reads
    assert cfg.optimizer["type"] == "AdamW" and cfg.optimizer_config["grad_clip"]["max_norm"] == 35
    assert cfg.log_config["interval"] in (10, 50) and "TensorboardImageLoggerHook" not in str(cfg.log_config)  # _delete_
    assert cfg.dist_params["backend"] == "nccl"
    assert cfg.model.backbone.embed_dims == 192          # attribute access on nested dicts
    cfg.merge_from_dict({"model.backbone.drop_path_rate": 0.0})
    assert cfg.model.backbone.drop_path_rate == 0.0 and cfg.model.backbone.embed_dims == 192


def test_config_base_merge_and_delete(tmp_path):
    from gedepth_b200.compat import Config
    (tmp_path / "base.py").write_text("a = dict(x=1, y=dict(p=1, q=2))\nb = [1, 2]\nhooks = dict(k=1, l=2)\n")
    (tmp_path / "child.py").write_text(
        "_base_ = ['./base.py']\nn = 3\na = dict(y=dict(q=5), z=[i for i in range(n)])\n"
        "hooks = dict(_delete_=True, m=3)\n")
    cfg = Config.fromfile(str(tmp_path / "child.py"))
    assert cfg.a == dict(x=1, y=dict(p=1, q=5), z=[0, 1, 2]) and cfg.b == [1, 2] and cfg.hooks == dict(m=3)
    assert cfg.n == 3 and "_base_" not in cfg


def test_registry_builds_every_name_of_the_path():
    import gedepth_b200.models as M
    for name in ["DepthEncoderDecoder", "DepthFormerSwin", "HAHIHeteroNeck", "LightPEMASKNeck",
                 "DynamicPENeckSOFT", "DenseDepthHead", "SigLoss", "CrossEntropyLoss",
                 "BinaryCrossEntropyLoss", "GroundEmbedding"]:
        assert M.MODELS.get(name) is not None, name
    from gedepth_b200.compat import POSITIONAL_ENCODING
    assert POSITIONAL_ENCODING.get("SinePositionalEncoding") is not None
    assert M.BACKBONES is M.NECKS is M.HEADS is M.LOSSES is M.DEPTHER is M.MODELS
    with pytest.raises(KeyError):
        M.MODELS.build(dict(type="NoSuchModule"))
    with pytest.raises(TypeError):
        M.MODELS.build(dict(type="SigLoss", not_an_argument=1))


@pytest.mark.parametrize("variant,n_expected", [("v", 275), ("a", 277)])
def test_swin_l_state_dict_surface(variant, n_expected):
    """Parameter totals and key names recorded in SURVEY.md §8(b)."""
    import gedepth_b200.models as M
    from gedepth_b200.presets import model_cfg
    with torch.device("meta"):
        m = M.build_depther(model_cfg(variant, "kitti", pretrained=None))
    n = sum(p.numel() for p in m.parameters()) / 1e6
    assert abs(n - n_expected) < 1.5, n
    keys = set(m.state_dict().keys())
    for k in ["backbone.patch_embed.projection.weight", "backbone.stages.2.blocks.17.attn.w_msa.relative_position_index",
              "backbone.stages.0.blocks.0.ffn.layers.0.0.weight", "backbone.stages.0.blocks.0.ffn.layers.1.bias",
              "backbone.stages.1.downsample.reduction.weight", "backbone.norm3.weight", "backbone.conv1.weight",
              "backbone.bn1.running_mean", "neck.lateral_convs.0.conv.weight", "neck.trans_fusion.3.bn.weight",
              "neck.conv_proj.0.conv.weight", "neck.level_embed", "neck.multi_att.sampling_offsets.weight",
              "neck.self_attn.output_proj.bias", "neck.reference_points.weight", "pe_mask_neck.convfinal.bias",
              "decode_head.conv_list.0.conv.weight", "decode_head.conv_list.4.convB.conv.bias",
              "decode_head.conv_depth.weight"]:
        assert k in keys, k
    assert m.state_dict()["backbone.patch_embed.projection.weight"].shape == (192, 4, 4, 4)
    assert ("dynamic_pe_neck.convfinal.weight" in keys) == (variant == "a")
