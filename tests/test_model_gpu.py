"""Whole-path parity on the B200 against the fixtures produced by the REFERENCE's own modules
(tests/golden/*.npz, see oracle/make_golden.py): forward depth map, losses, gradients, inference,
Abs-Rel.  The CUDA path runs TF32 tensor-core GEMMs/convs (fp32 storage and accumulation); the
tolerances below are the north star's: depth within 1e-3 relative at EVERY pixel, Abs-Rel within 1e-4.  The *_k8
cases run at the BASELINE shape 352 x 1120."""
import numpy as np
import pytest
import torch

from tests.golden_util import ZERO_GRAD_KEYS, build_host_model, load_case, metas_for

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module", autouse=True)
def _gpu():
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    torch.backends.cuda.matmul.allow_tf32 = True      # the remaining library GEMMs (dW) match the kernels' TF32
    torch.backends.cudnn.allow_tf32 = True
    yield
    torch.cuda.synchronize()


def _rel(a, b):
    return (np.abs(a - b) / np.maximum(np.abs(b), 1e-3))


@pytest.mark.parametrize("name", ["vanilla_train", "adaptive_train", "adaptive_ddad_train", "adaptive_train_k8"])
def test_train_step_matches_reference(name):
    from gedepth_b200 import kernels, ops
    case, g, b = load_case(name)
    model, _ = build_host_model(case, DEV)
    model.train()
    data = dict(img=torch.from_numpy(b["img"]).to(DEV), img_metas=metas_for(case),
                depth_gt=torch.from_numpy(b["depth_gt"]).to(DEV))
    if "pe_k_gt" in b:
        data["pe_k_gt"] = torch.from_numpy(b["pe_k_gt"]).to(DEV)
    if "height" in b:
        data["height"] = torch.from_numpy(b["height"]).float().to(DEV)
    n0 = kernels.LAUNCHES
    x, y, pe_mask, _ = model.extract_feat(data["img"], data["img_metas"], **{k: v for k, v in data.items() if k in ("height",)})
    depth, _ = model.decode_head.forward(x, data["img_metas"], pe_mask, y)
    rel = _rel(depth.detach().cpu().numpy(), g["depth"])
    print(f"{name}: depth rel err p50 {np.percentile(rel, 50):.2e} p99.9 {np.percentile(rel, 99.9):.2e} max {rel.max():.2e}")
    assert rel.max() < 1e-3
    sub = case.get("sub", 1)
    np.testing.assert_allclose(y.detach().cpu().numpy()[..., ::sub, ::sub], g["y"], rtol=2e-3, atol=2e-4)
    out = model.train_step(data, None)
    assert kernels.LAUNCHES - n0 > 100, "the sm_100a kernels did not run"
    assert abs(out["log_vars"]["loss"] - float(g["loss"])) < 2e-3 * float(g["loss"])
    out["loss"].backward()
    gn = dict(zip(g["grad_names"].tolist(), g["grad_norms"].tolist()))
    top = max(gn.values())
    worst = 0.0
    for n, p in model.named_parameters():
        if n in ZERO_GRAD_KEYS:
            continue
        assert p.grad is not None, n
        err = abs(float(p.grad.double().norm()) - gn[n]) / (gn[n] + 1e-3 * top)
        worst = max(worst, err)
        assert err < 2e-2, (n, float(p.grad.double().norm()), gn[n])
    print(f"{name}: worst grad-norm rel err {worst:.2e}")
    for k in g.files:
        if k.startswith("grad."):
            got = dict(model.named_parameters())[k[5:]].grad.cpu().numpy()
            # reference_points.weight collects d(bilinear sample)/d(location) over 10^7 samples: piecewise
            # constant in the location, so fp32-level location differences at cell borders show up here
            tol = 1e-1 if "reference_points" in k else 2e-2
            assert np.abs(got - g[k]).max() < tol * np.abs(g[k]).max(), k
    assert all(ops.native_table()[k] for k in ("ge_vanilla", "ge_adaptive", "fuse_head", "silog", "linear", "conv2d"))


@pytest.mark.parametrize("name", ["vanilla_train", "adaptive_train"])
def test_fp32_accurate_backward_matches_reference_gradients(name, monkeypatch, request):
    """GEDEPTH_BWD_GEMM_PASSES=3: dX / dW GEMMs in 3xTF32 and the fp32 deformable-attention backward.  Every gradient norm
    must then agree with the reference's to 2e-3 (measured: <= 1.3e-3; the default one-pass TF32 backward is held to 2e-2 in
    test_train_step_matches_reference), the head-side full gradients to 1e-3 of their largest entry (measured 5e-5).
    Gradients that pass through the deformable-attention sampling LOCATIONS (offsets, reference points, level embedding,
    everything upstream of the queries) are differences of neighbouring value rows: the 2^-21 relative error of the
    3xTF32 value_proj is amplified by |v| / |v01 - v00| there, which bounds them at ~5e-3 of the largest entry."""
    from gedepth_b200 import kernels
    monkeypatch.setattr(kernels, "BACKWARD_PASSES", 3)
    prev = kernels.set_gemm_precision(3)               # forward in 3xTF32 as well: the activations the gradients multiply
    request.addfinalizer(lambda: kernels.set_gemm_precision(prev))
    case, g, b = load_case(name)
    model, _ = build_host_model(case, DEV)
    model.train()
    data = dict(img=torch.from_numpy(b["img"]).to(DEV), img_metas=metas_for(case),
                depth_gt=torch.from_numpy(b["depth_gt"]).to(DEV))
    if "pe_k_gt" in b:
        data["pe_k_gt"] = torch.from_numpy(b["pe_k_gt"]).to(DEV)
    out = model.train_step(data, None)
    out["loss"].backward()
    params = dict(model.named_parameters())
    worst = {}
    for k in g.files:
        if k.startswith("grad."):
            got = params[k[5:]].grad.cpu().numpy()
            worst[k[5:]] = float(np.abs(got - g[k]).max() / np.abs(g[k]).max())
    print(name, {k: f"{v:.1e}" for k, v in worst.items()})
    for k, v in worst.items():
        # reference_points.weight sums d(bilinear sample)/d(location) over 10^7 samples; fp32 location differences at cell
        # borders (a piecewise-constant derivative) are not a GEMM-precision effect
        assert v < (1e-3 if k.startswith(("decode_head", "pe_mask_neck")) else 1e-2), (k, v)
    gn = dict(zip(g["grad_names"].tolist(), g["grad_norms"].tolist()))
    top = max(gn.values())
    for n, p in params.items():
        if n in ZERO_GRAD_KEYS:
            continue
        err = abs(float(p.grad.double().norm()) - gn[n]) / (gn[n] + 1e-3 * top)
        assert err < 2e-3, (n, err)


@pytest.mark.parametrize("name", ["vanilla_eval_ragged", "adaptive_eval", "vanilla_eval_k8"])
def test_inference_matches_reference(name):
    from oracle import ground as og
    case, g, b = load_case(name)
    model, _ = build_host_model(case, DEV)
    model.eval()
    with torch.no_grad():
        res = model(img=[torch.from_numpy(b["img"]).to(DEV)], img_metas=[metas_for(case)], return_loss=False)
    pred, ref = res[0], g["pred"][0]
    rel = _rel(pred, ref)
    print(f"{name}: pred rel err p50 {np.percentile(rel, 50):.2e} p99.9 {np.percentile(rel, 99.9):.2e} max {rel.max():.2e}")
    assert rel.max() < 1e-3
    # Abs-Rel (metrics.py:17) of both predictions against the same synthetic ground truth
    from gedepth_b200.synth import synth_batch
    gt = synth_batch(case["B"], case["H"], case["W"], seed=99, sparsity=0.2)["depth_gt"][0]
    a, r = og.abs_rel(gt, pred), og.abs_rel(gt, ref)
    assert abs(a - r) < 1e-4, (a, r)


def test_flip_tta_fused_average_equals_reference_sequence(monkeypatch):
    """aug_test with the two views of every GE config (plain + horizontal flip, encoder_decoder.py:249-274): the fused
    un-flip + average kernel gives what the reference's flip / add / divide sequence gives."""
    from gedepth_b200 import ops
    case, g, b = load_case("vanilla_eval_ragged")
    model, _ = build_host_model(case, DEV)
    model.eval()
    img = torch.from_numpy(b["img"]).to(DEV)
    metas0 = metas_for(case)
    metas1 = [dict(m, flip=True, flip_direction="horizontal") for m in metas0]
    kw = dict(pe_ori_point=[torch.zeros(1), torch.zeros(1)])
    with torch.no_grad():
        fast = model(img=[img, img.flip(3)], img_metas=[metas0, metas1], return_loss=False, **kw)
        monkeypatch.setattr(ops, "use_native", lambda name: False)      # generic flip / add / divide path of aug_test
        slow = model(img=[img, img.flip(3)], img_metas=[metas0, metas1], return_loss=False, **kw)
    np.testing.assert_allclose(fast[0], slow[0], rtol=1e-6, atol=1e-6)
    assert np.abs(fast[0] - g["pred"][0]).max() > 1e-4          # it really is a two-view average, not view 0


def test_flip_tta_matches_reference_aug_test():
    """Fixture produced by the reference's own aug_test (oracle/make_golden.py, case vanilla_eval_tta)."""
    case, g, b = load_case("vanilla_eval_tta")
    model, _ = build_host_model(case, DEV)
    model.eval()
    img = torch.from_numpy(b["img"]).to(DEV)
    metas0 = metas_for(case)
    metas1 = [dict(m, flip=True, flip_direction="horizontal") for m in metas0]
    with torch.no_grad():
        res = model(img=[img, img.flip(3)], img_metas=[metas0, metas1], return_loss=False,
                    pe_ori_point=[torch.zeros(1), torch.zeros(1)])
    rel = _rel(res[0], g["pred"][0])
    print(f"tta: pred rel err p50 {np.percentile(rel, 50):.2e} p99.9 {np.percentile(rel, 99.9):.2e} max {rel.max():.2e}")
    assert rel.max() < 1e-3


def test_library_statement_path_equals_reference_on_gpu(monkeypatch):
    """Same host mirror with every op forced to its library statement (fp32, no TF32): isolates
    wiring errors from kernel numerics."""
    from gedepth_b200 import ops
    from tests import ops_lib
    restore = ops_lib.install(ops)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    try:
        case, g, b = load_case("vanilla_eval_ragged")
        model, _ = build_host_model(case, DEV)
        model.eval()
        with torch.no_grad():
            res = model(img=[torch.from_numpy(b["img"]).to(DEV)], img_metas=[metas_for(case)], return_loss=False)
        np.testing.assert_allclose(res[0], g["pred"][0], rtol=2e-4, atol=2e-4)
    finally:
        restore()
        torch.backends.cuda.matmul.allow_tf32 = True
        torch.backends.cudnn.allow_tf32 = True


def test_cuda_graph_step_equals_eager_steps():
    """Trainer.capture/step_graph (the whole step as one CUDA graph) reproduces eager steps bit for bit
    up to atomics order: same losses over 3 optimisation steps, same parameters afterwards."""
    from gedepth_b200.train import Trainer
    case, g, b = load_case("vanilla_train")
    data = dict(img=torch.from_numpy(b["img"]).to(DEV), img_metas=metas_for(case),
                depth_gt=torch.from_numpy(b["depth_gt"]).to(DEV))
    losses = {}
    params = {}
    for mode in ("eager", "graph"):
        model, _ = build_host_model(case, DEV)
        model.train()
        tr = Trainer(model)
        if mode == "graph":
            snapshot = (tr.arena.flat_p.clone(), [b_.clone() for b_ in model.buffers()])
            tr.capture(data, warmup=2)
            # capture() runs real warm-up steps but must hand back the state it found (parameters, Adam moments,
            # BatchNorm statistics, step counters)
            assert torch.equal(tr.arena.flat_p, snapshot[0]) and float(tr.m.abs().max()) == 0.0 and tr.step_idx == 0
            assert int(tr.step_dev) == 0
            for b_, s_ in zip(model.buffers(), snapshot[1]):
                assert torch.equal(b_, s_)
        out = []
        for i in range(3):
            loss = tr.step_graph(data) if mode == "graph" else tr.step(data)[0]
            out.append(float(loss.detach()))
        losses[mode] = out
        params[mode] = tr.arena.flat_p.clone()
    assert losses["eager"][0] == pytest.approx(float(g["loss"]), rel=2e-3)
    for a, c in zip(losses["eager"], losses["graph"]):
        assert a == pytest.approx(c, rel=2e-4), (losses,)
    assert losses["eager"][2] != losses["eager"][0]              # the optimizer really stepped
    d = (params["eager"] - params["graph"]).abs()
    assert float((d > 5e-5).float().mean()) < 1e-3                # Adam sign flips on round-off gradients only


def test_graph_replay_follows_the_lr_schedule():
    """The learning rate is a device scalar refreshed before every replay: changing it between replays changes the
    parameter update accordingly (lr = 0 freezes the parameters), and a schedule callable is honoured."""
    from gedepth_b200.train import Trainer, cosine_warmup_lr
    case, g, b = load_case("vanilla_train")
    data = dict(img=torch.from_numpy(b["img"]).to(DEV), img_metas=metas_for(case),
                depth_gt=torch.from_numpy(b["depth_gt"]).to(DEV))
    model, _ = build_host_model(case, DEV)
    model.train()
    tr = Trainer(model, lr=1e-4, weight_decay=0.0)
    tr.capture(data, warmup=1)
    p0 = tr.arena.flat_p.clone()
    tr.step_graph(data)
    d1 = float((tr.arena.flat_p - p0).abs().max())
    assert 0.5e-4 < d1 < 2e-4                         # first Adam step moves every touched parameter by ~lr
    p1 = tr.arena.flat_p.clone()
    tr.lr = 0.0
    tr.step_graph(data)
    assert torch.equal(tr.arena.flat_p, p1), "lr = 0 must freeze the parameters of a replayed step"
    tr.lr_schedule = lambda step: 1e-6 * (step + 1)
    tr.step_graph(data)
    d3 = float((tr.arena.flat_p - p1).abs().max())
    assert 1e-6 < d3 < 1e-5, d3                       # step_idx == 2 -> lr = 3e-6
    assert cosine_warmup_lr(0) == pytest.approx(1e-4 * 1e-3, rel=1e-6)
    assert cosine_warmup_lr(16 * 1600) == pytest.approx(1e-8 * 1e-4 + 0.5 * (1e-4 - 1e-12) * (1 + np.cos(np.pi / 3)), rel=1e-6)


@pytest.mark.parametrize("name", ["vanilla_train", "adaptive_train"])
def test_arena_gradient_sinks_equal_autograd_accumulation(name):
    """With train.FlatArena the backward kernels add weight / bias / LayerNorm / rel-pos-bias gradients straight
    into the flat gradient buffer (conv weights channels-last) and hand autograd nothing; the result must be the
    gradient plain autograd accumulation gives on the same model."""
    from gedepth_b200.train import FlatArena
    case, g, b = load_case(name)
    data = dict(img=torch.from_numpy(b["img"]).to(DEV), img_metas=metas_for(case),
                depth_gt=torch.from_numpy(b["depth_gt"]).to(DEV))
    if "pe_k_gt" in b:
        data["pe_k_gt"] = torch.from_numpy(b["pe_k_gt"]).to(DEV)
    grads = {}
    for mode in ("autograd", "arena"):
        model, _ = build_host_model(case, DEV)
        model.train()
        arena = FlatArena(model) if mode == "arena" else None
        if arena is not None:
            arena.zero_grad()
            w = dict(model.named_parameters())["decode_head.conv_list.1.convA.conv.weight"]
            assert w.shape[2:] == (3, 3) and w.permute(0, 2, 3, 1).is_contiguous()      # channels-last inside the arena
        model.train_step(data, None)["loss"].backward()
        grads[mode] = {n: p.grad.detach().clone().contiguous() for n, p in model.named_parameters()}
        if arena is not None:
            assert float(arena.flat_g.abs().sum()) > 0
    top = max(float(v.abs().max()) for v in grads["autograd"].values())
    for n, ga in grads["autograd"].items():
        if n in ZERO_GRAD_KEYS:
            continue
        gb = grads["arena"][n]
        assert ga.shape == gb.shape, n
        tol = 3e-3 * float(ga.abs().max()) + 1e-6 * top     # atomics-order round-off, amplified through 12 blocks of backward
        assert float((ga - gb).abs().max()) <= tol, (n, float((ga - gb).abs().max()), float(ga.abs().max()))


def test_training_reduces_the_loss_on_a_fixed_batch():
    """Functional check of the whole optimisation loop (forward, SiLog, backward with gradient sinks, clip, AdamW on
    the flat arena, CUDA-graph replay): 40 steps on one fixed batch must drive the loss down."""
    from gedepth_b200.train import Trainer
    case, g, b = load_case("vanilla_train")
    data = dict(img=torch.from_numpy(b["img"]).to(DEV), img_metas=metas_for(case),
                depth_gt=torch.from_numpy(b["depth_gt"]).to(DEV))
    model, _ = build_host_model(case, DEV)
    model.train()
    tr = Trainer(model, lr=3e-4)
    tr.capture(data, warmup=2)
    losses = [float(tr.step_graph(data).detach()) for _ in range(40)]
    assert all(np.isfinite(losses)), losses
    assert np.mean(losses[-5:]) < 0.8 * losses[0], (losses[0], losses[-5:])       # measured: 0.689 -> 0.456


def test_product_fails_loudly_without_extension(monkeypatch):
    from gedepth_b200 import kernels
    monkeypatch.setattr(kernels, "_lib", None)
    monkeypatch.setattr(kernels, "LIB_PATH", "/nonexistent/libgedepth_sm100.so")
    with pytest.raises(RuntimeError, match="no fallback"):
        kernels.has("linear")
