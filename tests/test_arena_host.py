"""train.FlatArena on CPU: one flat parameter / gradient buffer, conv weights channels-last inside it but with the
reference's logical shapes (state_dict compatibility), decay exemptions of depthformer_v.py:128-139."""
import torch

import gedepth_b200.models as M
from gedepth_b200.presets import model_cfg
from gedepth_b200.synth import synth_state_dict
from gedepth_b200.train import FlatArena, NO_DECAY_KEYS


def _model():
    m = M.build_depther(model_cfg("v", "kitti", "swin_t", pretrained=None))
    m.load_state_dict(synth_state_dict(m.state_dict(), 0))
    return m


def test_arena_views_layout_and_state_dict_surface():
    m = _model()
    before = {k: v.clone() for k, v in m.state_dict().items()}
    arena = FlatArena(m)
    lo, hi = arena.flat_p.data_ptr(), arena.flat_p.data_ptr() + arena.flat_p.numel() * 4
    glo, ghi = arena.flat_g.data_ptr(), arena.flat_g.data_ptr() + arena.flat_g.numel() * 4
    n4 = 0
    for n, p in m.named_parameters():
        assert lo <= p.data_ptr() < hi and glo <= p.grad.data_ptr() < ghi, n
        assert p.data_ptr() % 16 == 0 and p.grad.data_ptr() % 16 == 0, n          # TMA / float4 alignment of every segment
        assert getattr(p, "_ged_sink", False), n
        assert p.grad.shape == p.shape and p.grad.stride() == p.stride(), n
        if p.dim() == 4:
            n4 += 1
            assert p.permute(0, 2, 3, 1).is_contiguous(), n                       # [Cout][kh][kw][Cin] in memory
        else:
            assert p.is_contiguous(), n
    assert n4 > 20
    after = m.state_dict()
    assert list(after) == list(before)
    for k in before:                                                               # same logical shapes and values
        assert after[k].shape == before[k].shape and torch.equal(after[k], before[k]), k
    # a reference-layout checkpoint loads into the arena views
    sd2 = synth_state_dict(before, 1)
    m.load_state_dict(sd2)
    w = "decode_head.conv_list.1.convA.conv.weight"
    assert torch.equal(dict(m.named_parameters())[w], sd2[w]) and lo <= dict(m.named_parameters())[w].data_ptr() < hi


def test_arena_decay_mask_follows_custom_keys():
    m = _model()
    arena = FlatArena(m)
    off = 0
    byname = dict(m.named_parameters())
    for n in arena.names:                       # arena order: completion groups, registration order inside a group
        p = byname[n]
        sz = (p.numel() + 3) // 4 * 4
        want = 0 if any(k in n for k in NO_DECAY_KEYS) else 1
        assert int(arena.wd_mask[off]) == want and int(arena.wd_mask[off + p.numel() - 1]) == want, n
        off += sz
    assert off == arena.total
    names = dict(m.named_parameters())
    assert "backbone.stages.0.blocks.0.norm1.weight" in names and "backbone.bn1.weight" in names
    # LayerNorms and relative-position tables are exempt; BatchNorm parameters are decayed
    idx = {n: i for i, n in enumerate(arena.names)}
    assert idx  # names recorded in arena order
    # completion groups are contiguous and ordered: head + PE necks, neck, Swin stages 3, 2, 1, rest of the backbone
    groups = [g for g, _, _ in arena.group_ranges]
    assert groups == sorted(groups) and groups[0] == 0 and groups[-1] == 5 and len(set(groups)) == len(groups)
    assert arena.group_ranges[0][1] == 0 and arena.group_ranges[-1][2] == arena.total
    from gedepth_b200.train import completion_group
    assert completion_group("decode_head.conv_depth.weight") == 0 and completion_group("neck.level_embed") == 1
    assert completion_group("backbone.stages.3.blocks.0.norm1.weight") == 2 and completion_group("backbone.norm2.bias") == 3
    assert completion_group("backbone.stages.0.blocks.0.norm1.weight") == 5 and completion_group("backbone.bn1.weight") == 5
