"""N>1 host logic on CPU: world_size-2 gloo run of Trainer.step (batch sharded across ranks, ONE
all-reduce of the flat gradient arena, rank-averaged log_vars).  The CUDA kernels are replaced by
their library statements for this test only (conftest-style injection inside the workers)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _inject_cpu_ops():
    """Returns an undo callable (the pytest process must not leak the injection into other tests)."""
    from gedepth_b200 import kernels, ops
    from tests import ops_lib
    saved = (kernels.sumsq, kernels.adamw_step)
    restore = ops_lib.install(ops)          # library statements of every op (CPU tensors allowed)

    def undo():
        restore()
        kernels.sumsq, kernels.adamw_step = saved

    def sumsq(flat_g, out):
        out.copy_(flat_g.double().pow(2).sum().reshape(1))
        return out

    def adamw_step(p, g, m, v, wd_mask, ss, max_norm, grad_scale, lr, b1, b2, eps, wd, step, step_dev=None, lr_dev=None):
        coef = grad_scale * min(max_norm / (float(ss.sqrt()) * grad_scale + 1e-6), 1.0)
        gi = g * coef
        p.mul_(torch.where(wd_mask.bool(), 1 - lr * wd, 1.0))
        m.mul_(b1).add_(gi, alpha=1 - b1)
        v.mul_(b2).addcmul_(gi, gi, value=1 - b2)
        denom = v.sqrt() / (1 - b2 ** step) ** 0.5 + eps
        p.addcdiv_(m, denom, value=-lr / (1 - b1 ** step))

    kernels.sumsq, kernels.adamw_step = sumsq, adamw_step
    return undo


def _build(seed_data):
    import gedepth_b200.models as M
    from gedepth_b200.presets import model_cfg
    from gedepth_b200.synth import synth_batch, synth_state_dict
    cfg = model_cfg("v", "kitti", "swin_t", pretrained=None, drop_path_rate=0.0)
    model = M.build_depther(cfg)
    model.load_state_dict(synth_state_dict(model.state_dict(), 0))
    for m in model.modules():
        if isinstance(getattr(m, "dropout", None), torch.nn.Dropout):
            m.dropout.p = 0.0
    model.train()
    b = synth_batch(1, 64, 160, seed=seed_data)
    data = dict(img=torch.from_numpy(b["img"]), img_metas=[{}], depth_gt=torch.from_numpy(b["depth_gt"]))
    return model, data


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    _inject_cpu_ops()
    from gedepth_b200.train import Trainer
    model, data = _build(100 + rank)
    tr = Trainer(model)
    loss, logs = tr.step(data, sync_logs=True)
    flat = tr.arena.flat_p.clone()
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    if rank == 0:
        out["same_params"] = bool(torch.equal(gathered[0], gathered[1]))
        out["flat_p"] = flat
        out["flat_g"] = tr.arena.flat_g.clone()
        out["loss_logged"] = logs["loss"]
        out["loss_local"] = float(loss)
        out["wd_frac"] = float(tr.arena.wd_mask.float().mean())
        out["early_groups"] = list(tr.early_groups)
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_rank_step_equals_averaged_gradients():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert out["same_params"], "ranks diverged after the all-reduced step"
    # the all-reduce is sliced by completion group and launched from autograd hooks: head + PE necks, neck and the Swin
    # stages 3..1 were already in flight when the backward returned (group 5 = rest of the backbone goes last)
    assert out["early_groups"] == [0, 1, 2, 3, 4], out["early_groups"]
    assert 0.9 < out["wd_frac"] < 1.0        # LayerNorm / rel-pos-bias tensors are exempt from decay
    # single-process reference: average the two shards' gradients, same update.  Same CPU thread count as
    # the workers: with B=1 the train-mode BatchNorms see <= 10 values per channel at the deepest level, so
    # even reduction-order round-off is visibly amplified in the gradients.
    threads = torch.get_num_threads()
    torch.set_num_threads(2)
    undo = _inject_cpu_ops()
    try:
        _check_against_single_process(world, out)
    finally:
        undo()
        torch.set_num_threads(threads)


def _check_against_single_process(world, out):
    from gedepth_b200 import kernels
    from gedepth_b200.train import FlatArena
    grads, losses = [], []
    for r in range(world):
        model, data = _build(100 + r)
        arena = FlatArena(model)
        l, _ = model._parse_losses(model(**data), sync=False)
        l.backward()
        grads.append(arena.flat_g.clone())
        losses.append(float(l))
    g = (grads[0] + grads[1])
    p, m, v = arena.flat_p.clone(), torch.zeros_like(g), torch.zeros_like(g)
    model0, _ = _build(100)
    p = FlatArena(model0).flat_p.clone()
    ss = torch.zeros(1, dtype=torch.float64)
    kernels.sumsq(g, ss)
    kernels.adamw_step(p, g, m, v, arena.wd_mask, ss, 35.0, 0.5, 1e-4, 0.9, 0.999, 1e-8, 0.01, 1)
    # the all-reduced arena holds the SUM of the two shards' gradients (1/world is folded into AdamW)
    assert torch.allclose(out["flat_g"], g, rtol=1e-4, atol=1e-6 * float(g.abs().max()))
    # Adam's first step is lr*sign(g): compare parameters where the gradient is above round-off
    # (CPU thread count changes reduction order, and sign(noise) flips)
    solid = g.abs() > 1e-4 * float(g.abs().max())
    diff = (p - out["flat_p"]).abs()
    assert float(diff.max()) <= 2.1e-4 and float(diff[solid].max()) < 2e-6 and float(solid.float().mean()) > 0.2
    assert abs(out["loss_logged"] - sum(losses) / 2) < 1e-5      # log_vars are rank means (base.py:197-202)
    assert abs(out["loss_local"] - losses[0]) < 1e-5
