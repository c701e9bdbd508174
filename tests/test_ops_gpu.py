"""Per-kernel parity on the B200: every sm_100a kernel, called through the C-ABI (kernels.py ->
ctypes -> libgedepth_sm100.so), against (i) the oracle's numpy/torch-CPU restatement and (ii) the
plain PyTorch fp32 statement of the same op (ops_lib) on the GPU.  Tolerances are stated per test:
bit-exact for integer work, ~1e-5 for fp32 SIMT kernels, 2e-3 relative for TF32 tensor-core GEMMs.
"""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


@pytest.fixture(scope="module", autouse=True)
def _gpu():
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    from gedepth_b200 import kernels
    kernels.load()
    yield
    torch.cuda.synchronize()


def _close(a, b, rtol, atol, msg=""):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    err = (a - b).abs()
    tol = atol + rtol * b.abs()
    bad = int((err > tol).sum())
    assert bad == 0, f"{msg}: {bad}/{a.numel()} mismatches, max abs err {float(err.max()):.3e}, max ref {float(b.abs().max()):.3e}"


# ------------------------------------------------------------------------------------------------
# a1: ground plane + integer grid
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("H,W,u0,v0", [(375, 1242, 0, 0), (352, 1120, 61, 23), (37, 53, 5, 3)])
def test_ground_plane_bit_exact_vs_numpy(H, W, u0, v0):
    from gedepth_b200 import kernels as K
    from oracle import ground as og
    coef = og.plane_coefficients(og.kitti_projection(), og.KITTI_CAM_HEIGHT)
    pe = og.ground_plane(coef, H, W, u0, v0)
    ch3, ch4 = og.load_channels(pe, 200.0)
    ch3 = og.normalize_pe(ch3, 200.0)
    out = K.ground_plane(coef, H, W, DEV, 3, u0, v0, 200.0, 200.0, 1.0, 1.0).cpu().numpy()
    for b in range(3):
        assert np.array_equal(out[b, 1], ch4), "raw ground depth must equal float32(numpy float64 result) bit for bit"
        assert np.array_equal(out[b, 0], ch3)
    u, v = K.pixel_grid(H, W, DEV, u0, v0)
    ur, vr = og.pixel_grid(H, W, u0, v0)
    assert u.dtype == torch.int64 and np.array_equal(u.cpu().numpy(), ur) and np.array_equal(v.cpu().numpy(), vr)


def test_find_k_matches_numpy():
    from gedepth_b200 import kernels as K
    from oracle import ground as og
    coef = og.plane_coefficients(og.kitti_projection(), og.KITTI_CAM_HEIGHT)
    pe = og.ground_plane(coef, 96, 320, 400, 150).astype(np.float32)
    rng = np.random.default_rng(1)
    gt = np.where(rng.random((2, 96, 320)) < 0.3, np.abs(pe) * (1 + 0.2 * rng.standard_normal((2, 96, 320))), 0).astype(np.float32)
    got = K.find_k(torch.from_numpy(gt).to(DEV), torch.from_numpy(pe).to(DEV), 1.65, False).cpu().numpy()
    ref = np.stack([og.find_k_kitti(gt[b].astype(np.float64), pe) for b in range(2)])
    assert np.array_equal(got, ref)             # integer labels: bit-exact (same arithmetic types as the script)
    assert np.all(got[gt == 0] == 255)
    got_t = K.find_k(torch.from_numpy(gt).to(DEV), torch.from_numpy(pe).to(DEV), 1.56, True).cpu().numpy()
    ref_t = np.stack([og.find_k_ddad(gt[b].astype(np.float64), pe.astype(np.float64), 1.56) for b in range(2)])
    assert np.array_equal(got_t, ref_t)


def test_ground_plane_and_find_k_equal_the_reference_script():
    """tests/golden/ref_preprocess_kitti.npz holds what the reference's own tools/preprocess_data_kitti.py wrote when it was
    executed verbatim on a synthetic data/kitti tree (oracle/run_ref_preprocess.py): pe_165.npy (a1) and one slope-label
    file (f1).  The generator kernel reproduces float32(pe) and ged_find_k reproduces every label, bit for bit."""
    import hashlib
    from gedepth_b200 import kernels as K
    from oracle import ground as og
    from oracle.run_ref_preprocess import H, W, synth_gt
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_preprocess_kitti.npz"))
    coef = og.plane_coefficients(og.kitti_projection(), og.KITTI_CAM_HEIGHT)
    pe = K.ground_plane(coef, H, W, DEV, 1, 0, 0, 200.0, 200.0, 1.0, 1.0)[0, 1].contiguous()
    assert hashlib.sha256(pe.cpu().numpy().tobytes()).hexdigest() == str(g["pe_f32_sha256"])
    gt16 = synth_gt()
    assert hashlib.sha256(gt16.tobytes()).hexdigest() == str(g["gt_sha256"])
    gt = torch.from_numpy((gt16.astype(np.float64) / 256).astype(np.float32)).to(DEV)       # k/256, k < 2^16: exact in fp32
    k = K.find_k(gt[None], pe, 1.65, False)[0].cpu().numpy()
    assert np.array_equal(k, g["k_img"].astype(np.float32)), int((k != g["k_img"]).sum())


# ------------------------------------------------------------------------------------------------
# a14 / a15 / a17: ground embedding, forward and backward, vs the library statement and the oracle
# ------------------------------------------------------------------------------------------------
def _ge_inputs(B, H, W, adaptive, seed=0):
    from gedepth_b200.synth import synth_batch
    b = synth_batch(B, H, W, seed=seed, adaptive=adaptive)
    g = torch.Generator().manual_seed(seed)
    h2, w2 = (H + 1) // 2, (W + 1) // 2
    y_half = torch.rand(B, 1, h2, w2, generator=g)
    logits_half = torch.randn(B, 11, h2, w2, generator=g) * 2
    return torch.from_numpy(b["img"]), y_half, logits_half


@pytest.mark.parametrize("B,H,W", [(2, 64, 160), (1, 70, 166), (3, 35, 83), (2, 352, 1120), (1, 4, 8), (2, 6, 12), (1, 384, 640)])
def test_ge_vanilla_fwd_bwd(B, H, W):
    from gedepth_b200 import kernels as K
    from tests import ops_lib as L
    from oracle import model as om
    img, y_half, _ = _ge_inputs(B, H, W, False)
    y_o, pm_o = om.ground_embed_vanilla(img, y_half)                       # oracle (CPU)
    img_d = img.to(DEV)
    yh1 = y_half.to(DEV).requires_grad_(True)
    yh2 = y_half.to(DEV).requires_grad_(True)
    y, pm = K.ge_vanilla(img_d, yh1)
    y_l, pm_l = L.ge_vanilla(img_d, yh2)
    _close(y, y_o, 1e-6, 1e-6, "y vs oracle")
    _close(pm, pm_o, 1e-5, 1e-5, "pe_mask vs oracle")
    _close(y, y_l, 1e-6, 1e-6, "y vs torch")
    gy, gpm = torch.randn_like(y), torch.randn_like(pm) * 0.1
    (y * gy + pm * gpm).sum().backward()
    (y_l * gy + pm_l * gpm).sum().backward()
    _close(yh1.grad, yh2.grad, 1e-4, 1e-4 * float(yh2.grad.abs().max()), "g_y_half")


@pytest.mark.parametrize("B,H,W", [(2, 64, 160), (1, 4, 8), (2, 6, 12), (2, 352, 1120), (1, 384, 640), (1, 66, 516), (3, 34, 1032), (1, 2, 4)])
def test_ge_vanilla_x2_streaming_equals_tiled(B, H, W):
    """The streaming (warp per 128-column strip) x2 kernels, forward and backward, against the tiled ones (ged_set_ge_x2(2)) and
    the generic bilinear ones (0): strips that end inside a warp, maps narrower than one strip, chunks that end inside the map,
    a missing g_y / g_pe_mask."""
    from gedepth_b200 import kernels as K
    img, y_half, _ = _ge_inputs(B, H, W, False)
    img_d, yh = img.to(DEV), y_half.to(DEV)
    gy, gp = torch.randn(B, 1, H, W, device=DEV), torch.randn(B, 1, H, W, device=DEV) * 0.1
    res = {}
    for mode in (1, 2, 0):
        prev = K.set_ge_x2(mode)
        try:
            with torch.no_grad():
                y, pm = K.ge_vanilla(img_d, yh)
            outs = []
            for a, bb in ((gy, gp), (None, gp), (gy, None)):
                o = torch.full((B, 1, H // 2, W // 2), 3.0, device=DEV)
                if mode == 0:
                    o.zero_()
                K._call("ged_ge_vanilla_bwd", K._p(img_d[:, 3]), img_d.stride(0), K._p(a), K._p(bb), K._p(o), B, H, W, H // 2, W // 2,
                        K._stream())
                outs.append(o)
            torch.cuda.synchronize()
        finally:
            K.set_ge_x2(prev)
        res[mode] = (y, pm, *outs)
    for mode in (2, 0):
        for name, a, bb in zip(("y", "pe_mask", "g_y_half", "g_y_half (no g_y)", "g_y_half (no g_pe_mask)"), res[1], res[mode]):
            _close(a, bb, 1e-5, 1e-6 * max(1.0, float(bb.abs().max())), f"{name} vs mode {mode}")


@pytest.mark.parametrize("B,H,W,per_sample_h", [(2, 64, 160, False), (1, 70, 166, True), (2, 352, 1120, False), (1, 4, 8, False),
                                                 (2, 6, 12, True), (3, 36, 84, True), (1, 14, 520, False), (1, 384, 640, True),
                                                 (2, 30, 244, False)])
def test_ge_adaptive_fwd_bwd(B, H, W, per_sample_h):
    from gedepth_b200 import kernels as K
    from tests import ops_lib as L
    from oracle import model as om
    img, y_half, logits_half = _ge_inputs(B, H, W, True)
    height = torch.tensor([1.56, 1.57, 1.53][:B]) if per_sample_h else 1.65
    y_o, pm_o, lf_o = om.ground_embed_adaptive(img, y_half, logits_half, 200.0, height)
    img_d = img.to(DEV)
    hd = height.to(DEV) if per_sample_h else height
    a1 = [t.to(DEV).requires_grad_(True) for t in (y_half, logits_half)]
    a2 = [t.to(DEV).requires_grad_(True) for t in (y_half, logits_half)]
    y, pm, lf = K.ge_adaptive(img_d, a1[0], a1[1], hd, 200.0)
    y_l, pm_l, lf_l = L.ge_adaptive(img_d, a2[0], a2[1], hd, 200.0)
    _close(y, y_o, 1e-6, 1e-6, "y vs oracle")
    _close(lf, lf_o, 1e-5, 1e-5, "logits vs oracle")
    # the range mask is a hard threshold: exclude pixels whose offset sits within 1e-3 of a threshold
    ok = ((pm_o - pm_l.cpu()).abs() <= 1e-3 + 1e-3 * pm_o.abs())
    assert ok.float().mean() > 0.9999
    err = ((pm.cpu() - pm_o).abs() > 2e-3 + 2e-4 * pm_o.abs()) & ok
    assert err.float().mean() < 1e-5, f"pe_mask mismatch fraction {err.float().mean():.2e}"
    gy, gpm, glf = torch.randn_like(y), torch.randn_like(pm) * 0.1, torch.randn_like(lf) * 0.01
    ((y * gy).sum() + (pm * gpm).sum() + (lf * glf).sum()).backward()
    ((y_l * gy).sum() + (pm_l * gpm).sum() + (lf_l * glf).sum()).backward()
    for n, p, q in (("g_y_half", a1[0], a2[0]), ("g_logits_half", a1[1], a2[1])):
        d = (p.grad - q.grad).abs()
        tol = 2e-3 * float(q.grad.abs().max()) + 1e-3 * q.grad.abs()
        assert float((d > tol).float().mean()) < 1e-4, (n, float(d.max()), float(q.grad.abs().max()))


@pytest.mark.parametrize("B,H,W", [(2, 64, 160), (1, 352, 1120), (2, 30, 244)])
def test_ge_adaptive_x2_equals_generic(B, H, W):
    """The closed-form x2 kernels (csrc/ge_adaptive_x2.cu) against the generic bilinear ones on the same inputs, including
    logits whose neighbours differ by hundreds (the forward's bound on the softmax maximum underflows there and the exact
    per-pixel path takes over) and camera-plane values at / beyond the thresholds (0, inf, negative)."""
    from gedepth_b200 import kernels as K
    img, y_half, logits_half = _ge_inputs(B, H, W, True, seed=3)
    g = torch.Generator().manual_seed(11)
    logits_half[:, :, : H // 4] *= 150.0                       # wild logits in the upper half
    img[:, 4, :, : W // 8] = 0.0
    img[:, 4, :, W // 8: W // 4] = float("inf")
    img[:, 4, :, W // 4: W // 3] *= -1.0
    img_d = img.to(DEV)
    outs = {}
    for x2 in (1, 2, 0):
        prev = K.set_ge_x2(x2)
        try:
            a = [t.to(DEV).requires_grad_(True) for t in (y_half, logits_half)]
            y, pm, lf = K.ge_adaptive(img_d, a[0], a[1], 1.65, 200.0)
            gy, gpm, glf = (torch.randn(y.shape, generator=g).to(DEV), torch.randn(pm.shape, generator=g).to(DEV) * 0.1,
                            torch.randn(lf.shape, generator=g).to(DEV) * 0.01)
            g.manual_seed(11)
            ((y * gy).sum() + (pm * gpm).sum() + (lf * glf).sum()).backward()
            with torch.no_grad():
                y_i, pm_i, _ = K.ge_adaptive(img_d, a[0], a[1], 1.65, 200.0, want_logits=False)
            outs[x2] = (y, pm, lf, a[0].grad, a[1].grad, y_i, pm_i)
        finally:
            K.set_ge_x2(prev)
    y1, pm1, lf1, gy1, gl1, yi1, pmi1 = outs[1]
    y0, pm0, lf0, gy0, gl0, yi0, pmi0 = outs[0]
    # TMA staging vs per-thread asynchronous copies: the same arithmetic on the same staged values
    for a_, b_ in zip(outs[1], outs[2]):
        assert torch.equal(torch.nan_to_num(a_, 7.0), torch.nan_to_num(b_, 7.0))
    # inference variant (no logits written; column pairs (0,3)(1,2)) vs training variant: y identical, pe_mask to rounding
    _close(y1, yi1, 1e-6, 1e-6, "y, inference variant")
    badi = ((pm1 - pmi1).abs() > 2e-3 + 2e-4 * pm1.abs()) & ~torch.isnan(pm1)
    assert torch.equal(torch.isnan(pm1), torch.isnan(pmi1)) and float(badi.float().mean()) < 2e-5
    _close(y1, y0, 1e-6, 1e-6, "y")
    _close(lf1, lf0, 1e-5, 1e-4, "logits")
    assert torch.equal(torch.isnan(pm1), torch.isnan(pm0))
    bad = ((pm1 - pm0).abs() > 2e-3 + 2e-4 * pm0.abs()) & ~torch.isnan(pm0)
    assert float(bad.float().mean()) < 2e-5, float(bad.float().mean())      # hard range mask: threshold pixels may flip
    for n, p, q in (("g_y_half", gy1, gy0), ("g_logits_half", gl1, gl0)):
        fin = torch.isfinite(q) & torch.isfinite(p)
        assert float((torch.isfinite(q) != torch.isfinite(p)).float().mean()) < 1e-4
        d = (p - q).abs()[fin]
        tol = 2e-3 * float(q[fin].abs().max()) + 1e-3 * q[fin].abs()
        assert float((d > tol).float().mean()) < 1e-4, (n, float(d.max()), float(q[fin].abs().max()))


@pytest.mark.parametrize("B,H,W", [(2, 64, 160), (1, 70, 166), (2, 352, 1120)])
def test_fuse_head_fwd_bwd(B, H, W):
    from gedepth_b200 import kernels as K
    from tests import ops_lib as L
    g = torch.Generator().manual_seed(3)
    h2, w2 = (H + 1) // 2, (W + 1) // 2
    d0, pm0, y0 = torch.rand(B, 1, h2, w2, generator=g) * 20, torch.rand(B, 1, H, W, generator=g) * 80, torch.rand(B, 1, H, W, generator=g)
    a1 = [t.to(DEV).requires_grad_(True) for t in (d0, pm0, y0)]
    a2 = [t.to(DEV).requires_grad_(True) for t in (d0, pm0, y0)]
    o1, yh1 = K.fuse_head(*a1, 1e-3)
    o2, yh2 = L.fuse_head(*a2, 1e-3)
    _close(o1, o2, 1e-5, 1e-5, "fused depth")
    _close(yh1, yh2, 1e-6, 1e-6, "y_h")
    go = torch.randn_like(o1)
    (o1 * go).sum().backward()
    (o2 * go).sum().backward()
    for n, p, q in zip(("g_d", "g_pe_mask", "g_y"), a1, a2):
        _close(p.grad, q.grad, 1e-4, 1e-5 * float(q.grad.abs().max()) + 1e-6, n)


# ------------------------------------------------------------------------------------------------
# losses
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,H,W,up", [(2, 64, 160, True), (1, 70, 166, True), (2, 352, 1120, True), (2, 64, 160, False)])
def test_silog_fwd_bwd(B, H, W, up):
    from gedepth_b200 import kernels as K
    from tests import ops_lib as L
    from gedepth_b200.synth import synth_batch
    gt = torch.from_numpy(synth_batch(B, H, W, seed=5)["depth_gt"]).to(DEV)
    g = torch.Generator().manual_seed(4)
    shape = (B, 1, (H + 1) // 2, (W + 1) // 2) if up else (B, 1, H, W)
    p0 = (torch.rand(shape, generator=g) * 30 + 0.5)
    p1, p2 = p0.to(DEV).requires_grad_(True), p0.to(DEV).requires_grad_(True)
    l1 = K.silog(p1, gt, 1e-3, 0.15, None, up)
    l2 = L.silog(p2, gt, 1e-3, 0.15, None, up)
    assert abs(float(l1) - float(l2)) < 2e-5 * float(l2)
    (l1 * 1.7).backward()
    (l2 * 1.7).backward()
    _close(p1.grad, p2.grad, 1e-3, 1e-4 * float(p2.grad.abs().max()), "g_pred")


def test_cross_entropy_fwd_bwd():
    from gedepth_b200 import kernels as K
    from tests import ops_lib as L
    g = torch.Generator().manual_seed(6)
    l0 = torch.randn(2, 11, 64, 160, generator=g) * 2
    t = torch.randint(0, 11, (2, 64, 160), generator=g).float()
    t[torch.rand(2, 64, 160, generator=g) < 0.9] = 255
    t = t.to(DEV)
    l1, l2 = l0.to(DEV).requires_grad_(True), l0.to(DEV).requires_grad_(True)
    a, b = K.cross_entropy(l1, t, 255), L.cross_entropy(l2, t, 255)
    assert abs(float(a) - float(b)) < 1e-5 * abs(float(b)) + 1e-6
    (a * 0.08).backward()
    (b * 0.08).backward()
    _close(l1.grad, l2.grad, 1e-4, 1e-6 * float(l2.grad.abs().max()) + 1e-9, "g_logits")


# ------------------------------------------------------------------------------------------------
# Swin pieces
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("reg", [1, 0])
@pytest.mark.parametrize("rows,C", [(640, 96), (333, 384), (70, 1536), (50, 3072), (1237, 192), (9001, 768), (77, 512), (65, 640),
                                     (3, 4), (20011, 192)])
def test_layernorm_fwd_bwd(rows, C, reg):
    """reg = 1: rows of <= 768 channels in registers, dw / db out of the dx kernel (default); reg = 0: three-pass forward,
    two-kernel backward (what wider rows always take).  Also the residual-branch gradient added in the dx pass (g_add)."""
    from gedepth_b200 import kernels as K
    g = torch.Generator().manual_seed(7)
    x0 = torch.randn(2, rows, C, generator=g) * 3 + 1
    w0, b0 = torch.randn(C, generator=g), torch.randn(C, generator=g)
    a1 = [t.to(DEV).requires_grad_(True) for t in (x0, w0, b0)]
    a2 = [t.to(DEV).requires_grad_(True) for t in (x0, w0, b0)]
    prev = K.set_layernorm_reg(reg)
    try:
        y1, xid = K.layer_norm_fork(*a1, 1e-5)
        y2 = F.layer_norm(a2[0], (C,), a2[1], a2[2], 1e-5)
        _close(y1, y2, 1e-5, 1e-5, "layernorm")
        go, gi = torch.randn_like(y1), torch.randn_like(y1)
        ((y1 * go).sum() + (xid * gi).sum()).backward()
        ((y2 * go).sum() + (a2[0] * gi).sum()).backward()
        torch.cuda.synchronize()
    finally:
        K.set_layernorm_reg(prev)
    for n, p, q in zip(("dx", "dw", "db"), a1, a2):
        _close(p.grad, q.grad, 1e-4, 2e-5 * float(q.grad.abs().max()), n)


@pytest.mark.parametrize("core", ["path", "tcgen05", "tcgen05_fwd", "simt"])
@pytest.mark.parametrize("B,H,W,nH,shift", [(2, 16, 40, 3, 0), (2, 16, 40, 3, 3), (1, 9, 21, 6, 3), (2, 7, 7, 12, 0), (1, 18, 42, 3, 3),
                                            (2, 11, 35, 24, 3), (1, 11, 35, 48, 0), (3, 12, 20, 3, 3), (1, 88, 280, 3, 3)])
def test_window_attention_fwd_bwd(B, H, W, nH, shift, core, monkeypatch):
    """`path`: what the step runs - tcgen05 forward (3xTF32, csrc/winattn_tc.cu) and the mma.sync backward (one pass TF32 like
    every other backward GEMM, csrc/winattn.cu); `tcgen05`: forward AND backward on tcgen05 (optional, slower); `tcgen05_fwd`:
    tcgen05 forward with the fp32 SIMT backward (what GEDEPTH_BWD_GEMM_PASSES=3 selects); `simt`: both directions SIMT fp32.  11 x 35 (stage 3 at
    352 x 1120, padded to 14 x 35) and 12 x 20 maps exercise padding tokens whose q = k = v = the qkv bias; odd pair
    counts exercise the half-empty last work item."""
    from gedepth_b200 import kernels as K
    from tests import ops_lib as L
    from oracle import model as om
    monkeypatch.setattr(K, "WINATTN_TC", core != "simt")
    monkeypatch.setattr(K, "WINATTN_TC_BWD", core == "tcgen05")
    monkeypatch.setattr(K, "WINATTN_BWD_MMA", core == "path")
    C = nH * 32
    g = torch.Generator().manual_seed(8)
    qkv0 = torch.randn(B, H * W, 3 * C, generator=g)
    bias0 = torch.randn(3 * C, generator=g) * 0.5
    table0 = torch.randn(169, nH, generator=g) * 0.5
    index = om.relative_position_index(7).to(DEV)
    a1 = [t.to(DEV).requires_grad_(True) for t in (qkv0, bias0, table0)]
    a2 = [t.to(DEV).requires_grad_(True) for t in (qkv0, bias0, table0)]
    o1 = K.window_attention(a1[0], a1[1], a1[2], index, (H, W), nH, 7, shift, 32 ** -0.5)
    o2 = L.window_attention(a2[0], a2[1], a2[2], index, (H, W), nH, 7, shift, 32 ** -0.5)
    _close(o1, o2, 1e-4, 1e-5, "context")
    go = torch.randn_like(o1)
    (o1 * go).sum().backward()
    (o2 * go).sum().backward()
    for n, p, q in zip(("g_qkv", "g_bias", "g_table"), a1, a2):
        qg = q.grad if q.grad is not None else torch.zeros_like(q)      # no padded tokens -> bias unused
        # one pass TF32 (10-bit mantissa operands) through five chained products: ~2e-3 of the gradient's scale
        atol = (3e-3 if core in ("tcgen05", "path") else 2e-5) * float(qg.abs().max()) + 1e-6
        _close(p.grad, qg, 1e-3, atol, n)


def test_window_attention_nonstandard_index():
    """A relative-position index that is NOT Swin's standard one: the path must read the buffer (SIMT forward, mma.sync
    backward with index lookups) instead of the closed form."""
    from gedepth_b200 import kernels as K
    from tests import ops_lib as L
    from oracle import model as om
    B, H, W, nH, shift = 2, 14, 21, 3, 3
    C = nH * 32
    g = torch.Generator().manual_seed(18)
    qkv0 = torch.randn(B, H * W, 3 * C, generator=g)
    bias0 = torch.randn(3 * C, generator=g) * 0.5
    table0 = torch.randn(169, nH, generator=g) * 0.5
    index = om.relative_position_index(7)
    index = index[torch.randperm(49, generator=g)].contiguous().to(DEV)          # rows permuted: valid entries, not the closed form
    assert not K._standard_rel_index(index)
    a1 = [t.to(DEV).requires_grad_(True) for t in (qkv0, bias0, table0)]
    a2 = [t.to(DEV).requires_grad_(True) for t in (qkv0, bias0, table0)]
    o1 = K.window_attention(a1[0], a1[1], a1[2], index, (H, W), nH, 7, shift, 32 ** -0.5)
    o2 = L.window_attention(a2[0], a2[1], a2[2], index, (H, W), nH, 7, shift, 32 ** -0.5)
    _close(o1, o2, 1e-4, 1e-5, "context")
    go = torch.randn_like(o1)
    (o1 * go).sum().backward()
    (o2 * go).sum().backward()
    for n, p, q in zip(("g_qkv", "g_bias", "g_table"), a1, a2):
        qg = q.grad if q.grad is not None else torch.zeros_like(q)
        _close(p.grad, qg, 1e-3, 3e-3 * float(qg.abs().max()) + 1e-6, n)


# ------------------------------------------------------------------------------------------------
# tcgen05 GEMM / conv (TF32: 10-bit mantissa products, fp32 accumulate)
# ------------------------------------------------------------------------------------------------
@pytest.fixture(params=[3, 2, 1], ids=["3xtf32", "bf16x3", "tf32"])
def passes(request, monkeypatch):
    """GEMM arithmetic: 3 = error-compensated 3xTF32 (fp32-accurate), 2 = bf16 hi/lo split (three kind::f16 products,
    2^-16 per product; MN-major operands fall back to 3xTF32), 1 = single-pass TF32.
    The backward (dX) kernels follow the same setting in these tests."""
    from gedepth_b200 import kernels as Kn
    prev = Kn.set_gemm_precision(request.param)
    monkeypatch.setattr(Kn, "BACKWARD_PASSES", request.param)
    yield request.param
    Kn.set_gemm_precision(prev)


def _gemm_tol(ref, passes, K=512):
    # 1 pass: tcgen05 kind::tf32 drops the low 13 mantissa bits of both operands -> ~1e-3 of the output scale.
    # 3 passes: hi*hi + lo*hi + hi*lo; what remains is the 2^-22 lo*lo term and the tensor core's fp32
    # accumulation (truncating alignment), which grows with the contraction length K.
    # bf16 split: x = hi + lo to 2^-17, lo*lo dropped: ~2^-16 per product, random-sign accumulation
    return (1.5e-3 if passes == 1 else (6e-5 if passes == 2 else 1e-5 + 8e-9 * K)) * float(ref.abs().max()) + 1e-6


@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (256, 96, 96), (1000, 288, 96), (24640, 384, 96), (777, 512, 512),
                                   (3000, 1536, 384), (130, 64, 2304), (500, 256, 512), (260, 2, 512), (128, 32, 64)])
def test_gemm_plain(M, N, K, passes):
    from gedepth_b200 import kernels as Kn
    g = torch.Generator().manual_seed(9)
    a = torch.randn(M, K, generator=g).to(DEV)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(DEV)
    if N < 16:
        pytest.skip("N<16 goes to the library by design")
    out = Kn.gemm(a, w)
    ref = (a.double() @ w.double().t()).float()
    _close(out, ref, 0, _gemm_tol(ref, passes, K), f"gemm {M}x{N}x{K}")


@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (1000, 64, 96), (24640, 96, 384), (777, 32, 512), (50000, 64, 2304), (3000, 128, 100)])
def test_gemm_a_operand_from_tensor_memory(M, N, K):
    """3xTF32 with the A operand's hi / lo parts written to tensor memory by the splitter warps (tcgen05.st) and read by
    tcgen05.mma from there, against the shared-memory-operand kernel and fp64."""
    from gedepth_b200 import kernels as Kn
    g = torch.Generator().manual_seed(77)
    a = torch.randn(M, K, generator=g).to(DEV)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(DEV)
    bias = torch.randn(N, generator=g).to(DEV)
    prev_p = Kn.set_gemm_precision(3)
    outs = {}
    try:
        for on in (1, 0):
            prev = Kn.set_gemm_a_tmem(on)
            try:
                outs[on] = Kn.gemm(a, w, bias, "gelu")
            finally:
                Kn.set_gemm_a_tmem(prev)
    finally:
        Kn.set_gemm_precision(prev_p)
    from tests import ops_lib as L
    ref = L._act((a.double() @ w.double().t() + bias.double()).float(), "gelu")
    _close(outs[1], ref, 0, _gemm_tol(ref, 3, K), "A from TMEM vs fp64")
    _close(outs[1], outs[0], 0, 2e-6 * float(ref.abs().max()) + 1e-6, "A from TMEM vs A from shared memory")


@pytest.mark.parametrize("P,N,K", [(64, 32, 32), (600, 384, 96), (1000, 96, 288), (5000, 128, 128), (20011, 512, 64),
                                   (70000, 64, 256), (3000, 1536, 384), (777, 16, 2304), (24640, 2304, 768), (9000, 768, 192),
                                   (4100, 1024, 100)])
def test_gemm_dw(P, N, K, passes):
    """dW[n,k] = sum_p G[p,n] X[p,k] on tcgen05 with MN-major operands as stored; contraction split across SMs
    with atomic accumulation."""
    from gedepth_b200 import kernels as Kn
    g = torch.Generator().manual_seed(21)
    G = torch.randn(P, N, generator=g).to(DEV)
    X = torch.randn(P, K, generator=g).to(DEV)
    out = Kn.gemm_dw(G, X)
    ref = (G.double().t() @ X.double()).float()
    _close(out, ref, 0, _gemm_tol(ref, passes, P), f"dW {P}x{N}x{K}")
    prev = Kn.set_gemm_pair_dw(False)                      # single-CTA kernel (the pair kernel serves N >= 256 in one-pass TF32)
    try:
        out1 = Kn.gemm_dw(G, X)
    finally:
        Kn.set_gemm_pair_dw(prev)
    _close(out1, ref, 0, _gemm_tol(ref, passes, P), f"dW single-CTA {P}x{N}x{K}")
    # accumulation into a running gradient
    out2 = Kn.gemm_dw(G, X, out=ref.clone())
    _close(out2, 2 * ref, 0, 2 * _gemm_tol(ref, passes, P), "dW accumulate")


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(2, 11, 35, 64, 64), (1, 22, 70, 576, 192), (3, 9, 12, 96, 32), (2, 11, 35, 64, 512),
                                            (1, 22, 70, 192, 768)])
def test_conv3x3_dw_taps(B, H, W, Cin, Cout, passes):
    """3x3 weight gradient as nine row-shifted contractions over the zero-bordered NHWC operands."""
    from gedepth_b200 import kernels as Kn
    g = torch.Generator().manual_seed(22)
    x = torch.randn(B, H, W, Cin, generator=g).to(DEV)
    gy = torch.randn(B, H, W, Cout, generator=g).to(DEV)
    xp, gp = Kn.prep_conv_input(x, None, H, W), Kn.prep_conv_input(gy, None, H, W)
    taps = [(ky - 1) * (W + 2) + (kx - 1) for ky in range(3) for kx in range(3)]
    dwk = Kn.gemm_dw(gp.reshape(-1, Cout), xp.reshape(-1, Cin), None, taps)        # [Cout, 9, Cin]
    dw = dwk.reshape(Cout, 3, 3, Cin).permute(0, 3, 1, 2)
    xd = x.permute(0, 3, 1, 2).double().requires_grad_(False)
    wd = torch.zeros(Cout, Cin, 3, 3, dtype=torch.float64, device=DEV, requires_grad=True)
    (F.conv2d(xd, wd, padding=1) * gy.permute(0, 3, 1, 2).double()).sum().backward()
    ref = wd.grad.float()
    _close(dw, ref, 0, _gemm_tol(ref, passes, B * H * W) * (3.0 if passes == 1 else 1.0), "conv dW")


@pytest.mark.parametrize("M,N,K", [(24640, 384, 96), (40000, 512, 512), (33000, 256, 96), (20000, 64, 64), (19000, 1536, 384),
                                   (50001, 288, 100), (70000, 96, 384)])
def test_gemm_pair_kernel(M, N, K, passes):
    """CTA-pair (cta_group::2) kernel: large problems, all epilogue features, against fp64 and against the
    single-CTA kernel."""
    from gedepth_b200 import kernels as Kn
    from tests import ops_lib as L
    g = torch.Generator().manual_seed(41)
    a = torch.randn(M, K, generator=g).to(DEV)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(DEV)
    bias = torch.randn(N, generator=g).to(DEV)
    res = torch.randn(M, N, generator=g).to(DEV)
    rs = (torch.rand(4, generator=g) + 0.5).to(DEV)
    rpb = -(-M // 4)
    ref = L._act((a.double() @ w.double().t() + bias.double()).float(), "gelu")
    ref = ref * rs.repeat_interleave(rpb)[:M].unsqueeze(1) + res
    outs = {}
    for pair in (2, 0):
        prev = Kn.set_gemm_pair(pair)
        try:
            outs[pair] = Kn.gemm(a, w, bias, "gelu", 0.01, res, rs, rpb)
        finally:
            Kn.set_gemm_pair(prev)
        _close(outs[pair], ref, 0, _gemm_tol(ref, passes, K), f"pair={pair} gemm {M}x{N}x{K}")
    _close(outs[2], outs[0], 0, 2e-6 * float(ref.abs().max()) + 1e-6, "pair vs single-CTA")
    # dX form: B operand [K][N] read in place (MN-major)
    if N % 4 == 0 and K % 4 == 0:
        wt = (torch.randn(K, N, generator=g) / K ** 0.5).to(DEV)
        out = Kn.gemm_bt(a, wt)
        ref2 = (a.double() @ wt.double()).float()
        _close(out, ref2, 0, _gemm_tol(ref2, passes, K), f"gemm_bt {M}x{N}x{K}")
        # + residual: dense, accumulated in place, and a channel slice of a wider matrix (its own row pitch)
        wide = torch.randn(M, N + 12, generator=g).to(DEV)
        out_r = Kn.gemm_bt(a, wt, residual=res)
        _close(out_r, ref2 + res, 0, _gemm_tol(ref2, passes, K), "gemm_bt + residual")
        out_s = Kn.gemm_bt(a, wt, residual=wide[:, 4:4 + N])
        _close(out_s, ref2 + wide[:, 4:4 + N], 0, _gemm_tol(ref2, passes, K), "gemm_bt + strided residual")
        acc = res.clone()
        Kn.gemm_bt(a, wt, residual=acc, out=acc)
        _close(acc, ref2 + res, 0, _gemm_tol(ref2, passes, K), "gemm_bt accumulated in place")


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(4, 64, 160, 64, 64), (2, 44, 140, 96, 192), (2, 88, 280, 128, 32)])
def test_conv3x3_pair_kernel(B, H, W, Cin, Cout, passes):
    from gedepth_b200 import kernels as Kn
    g = torch.Generator().manual_seed(42)
    x = torch.randn(B, H, W, Cin, generator=g).to(DEV)
    wk = (torch.randn(Cout, 3, 3, Cin, generator=g) / (9 * Cin) ** 0.5).to(DEV)
    bias = torch.randn(Cout, generator=g).to(DEV)
    xp = Kn.prep_conv_input(x, None, H, W)
    ref = F.conv2d(x.permute(0, 3, 1, 2).double(), wk.permute(0, 3, 1, 2).double(), bias.double(), padding=1).permute(0, 2, 3, 1).float()
    y = Kn.conv3x3_padded(xp, wk, bias, None, 0.0)
    _close(y, ref, 0, _gemm_tol(ref, passes, 9 * Cin), "conv fwd (pair)")
    # dX from the forward weights in place
    gy = torch.randn(B, H, W, Cout, generator=g).to(DEV)
    gp = Kn.prep_conv_input(gy, None, H, W)
    dx = Kn.conv3x3_dx(gp, wk)
    xd = x.permute(0, 3, 1, 2).double().requires_grad_(True)
    (F.conv2d(xd, wk.permute(0, 3, 1, 2).double(), None, padding=1) * gy.permute(0, 3, 1, 2).double()).sum().backward()
    ref_dx = xd.grad.permute(0, 2, 3, 1).float()
    _close(dx, ref_dx, 0, _gemm_tol(ref_dx, passes, 9 * Cout), "conv dX (pair)")


def test_linear_epilogue_dropout_fwd_bwd(passes):
    """residual + dropout_0.1(x w^T + b) with the mask drawn in the GEMM epilogue (counter-based hash of seed, device
    step counter and element index) and re-drawn by ged_dropout_bwd: keep rate, 1/(1-p) scaling, fwd/bwd mask
    consistency, seed / step dependence."""
    from gedepth_b200 import kernels as Kn
    g = torch.Generator().manual_seed(51)
    M, K, N, p = 3000, 128, 256, 0.1
    x0 = torch.randn(M, K, generator=g).to(DEV)
    w0 = (torch.randn(N, K, generator=g) / K ** 0.5).to(DEV)
    b0 = torch.randn(N, generator=g).to(DEV)
    r0 = torch.randn(M, N, generator=g).to(DEV)
    y0 = (x0.double() @ w0.double().t() + b0.double()).float()
    step = torch.zeros(1, dtype=torch.int32, device=DEV)
    out = Kn.gemm(x0, w0, b0, None, 0.01, r0, dropout=(p, 1234, step))
    keep = (out - r0) != 0
    assert abs(float(keep.float().mean()) - (1 - p)) < 4e-3
    pq = round(p * 65536) / 65536            # the kernels quantise p to 16 bits and scale by 1/(1-pq)
    _close(torch.where(keep, out - r0, torch.zeros_like(out)), torch.where(keep, y0 / (1 - pq), torch.zeros_like(y0)), 0,
           _gemm_tol(y0, passes, K) * 1.2 + 2e-6 * float(r0.abs().max()), "kept entries")
    assert torch.equal(out, Kn.gemm(x0, w0, b0, None, 0.01, r0, dropout=(p, 1234, step)))          # deterministic
    assert not torch.equal(out, Kn.gemm(x0, w0, b0, None, 0.01, r0, dropout=(p, 1235, step)))      # seed
    step.add_(1)
    out2 = Kn.gemm(x0, w0, b0, None, 0.01, r0, dropout=(p, 1234, step))                            # step counter
    keep2 = (out2 - r0) != 0
    assert 0.75 < float((keep == keep2).float().mean()) < 0.9           # independent masks agree on ~0.82 of entries
    # autograd through kernels.linear: the backward re-draws the forward's mask
    Kn.RNG_STEP = step
    xs, ws, bs, rs = (t.clone().requires_grad_(True) for t in (x0, w0, b0, r0))
    y = Kn.linear(xs, ws, bs, None, rs, None, p)
    m = ((y.detach() - r0) != 0).float() / (1 - pq)
    go = torch.randn_like(y)
    (y * go).sum().backward()
    gz = (go * m).double()
    _close(xs.grad, (gz @ w0.double()).float(), 0, 2 * _gemm_tol(xs.grad, passes, N), "dx")
    _close(ws.grad, (gz.t() @ x0.double()).float(), 0, 2 * _gemm_tol(ws.grad, passes, M), "dw")
    _close(bs.grad, gz.sum(0).float(), 1e-5, 1e-4 * float(bs.grad.abs().max()), "db")
    assert torch.equal(rs.grad, go)
    Kn.RNG_STEP = None


@pytest.mark.parametrize("act", [None, "relu", "gelu", "leaky_relu", "sigmoid"])
def test_gemm_epilogue(act, passes):
    from gedepth_b200 import kernels as Kn
    from tests import ops_lib as L
    g = torch.Generator().manual_seed(10)
    B, T, K, N = 3, 215, 192, 160
    a = torch.randn(B * T, K, generator=g).to(DEV)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(DEV)
    bias = torch.randn(N, generator=g).to(DEV)
    res = torch.randn(B * T, N, generator=g).to(DEV)
    rs = torch.tensor([0.0, 1.0 / 0.7, 1.0 / 0.7]).to(DEV)
    out = Kn.gemm(a, w, bias, act, 0.01, res, rs, T)
    ref = L._act((a.double() @ w.double().t() + bias.double()).float(), act)
    ref = ref * rs.repeat_interleave(T).unsqueeze(1) + res
    _close(out, ref, 0, _gemm_tol(ref, passes) + (2e-6 if act == "sigmoid" else 0), f"epilogue {act}")
    assert torch.equal(out[:T], res[:T])          # dropped sample: exactly the residual


@pytest.mark.parametrize("act", [None, "gelu"])
def test_linear_autograd(act, passes):
    from gedepth_b200 import kernels as Kn
    from tests import ops_lib as L
    g = torch.Generator().manual_seed(11)
    x0, w0, b0 = torch.randn(2, 300, 96, generator=g), torch.randn(384, 96, generator=g) / 10, torch.randn(384, generator=g)
    r0 = torch.randn(2, 300, 384, generator=g)
    a1 = [t.to(DEV).requires_grad_(True) for t in (x0, w0, b0, r0)]
    a2 = [t.to(DEV).double().requires_grad_(True) for t in (x0, w0, b0, r0)]
    y1 = Kn.linear(a1[0], a1[1], a1[2], act, a1[3], None)
    y2 = L.linear(a2[0], a2[1], a2[2], act, a2[3], None)
    _close(y1, y2, 0, _gemm_tol(y2, passes), "linear fwd")
    go = torch.randn_like(y1)
    (y1 * go).sum().backward()
    (y2 * go.double()).sum().backward()
    for n, p, q in zip(("dx", "dw", "db", "dres"), a1, a2):
        # dw is still a cuBLAS GEMM (fp32 here: allow_tf32 is off in this module)
        _close(p.grad, q.grad, 0, 2 * _gemm_tol(q.grad, passes), n)


@pytest.mark.parametrize("B,H,W,Cin,Cout,act", [(2, 11, 35, 64, 64, "leaky_relu"), (1, 22, 70, 576, 192, "relu"),
                                                (2, 16, 40, 96, 11, None), (1, 9, 12, 2304, 768, "leaky_relu"),
                                                (2, 32, 80, 64, 1, "sigmoid"), (2, 20, 48, 64, 1, "relu"),
                                                (1, 17, 33, 64, 11, None)])
def test_conv3x3(B, H, W, Cin, Cout, act, passes):
    from gedepth_b200 import kernels as Kn
    g = torch.Generator().manual_seed(12)
    x0 = torch.randn(B, Cin, H, W, generator=g)
    w0 = torch.randn(Cout, Cin, 3, 3, generator=g) / (9 * Cin) ** 0.5
    b0 = torch.randn(Cout, generator=g)
    a1 = [t.to(DEV).requires_grad_(True) for t in (x0, w0, b0)]
    a1[0] = x0.to(DEV).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    a2 = [t.to(DEV).double().requires_grad_(True) for t in (x0, w0, b0)]
    y1 = Kn.conv2d(a1[0], a1[1], a1[2], 1, 1, act, 0.01)
    from tests import ops_lib as L
    pre2 = F.conv2d(a2[0], a2[1], a2[2], padding=1)
    if act in ("relu", "leaky_relu"):
        # piecewise-linear activations: take the branch the kernel took (pre-activations within TF32
        # round-off of 0 would otherwise flip the derivative and dominate the gradient comparison)
        pos = (y1.detach() > 0)
        y2 = torch.where(pos, pre2, pre2 * (0.01 if act == "leaky_relu" else 0.0))
        _close(y1, L._act(pre2, act, 0.01), 0, _gemm_tol(pre2, passes, 9 * Cin), "conv fwd")
    else:
        y2 = L._act(pre2, act, 0.01)
        _close(y1, y2, 0, _gemm_tol(pre2, passes, 9 * Cin), "conv fwd")
    assert y1.shape == y2.shape
    go = torch.randn(y2.shape, generator=g).to(DEV)
    (y1 * go).sum().backward()
    (y2 * go.double()).sum().backward()
    for n, p, q in zip(("dx", "dw", "db"), a1, a2):
        tol = 2 * _gemm_tol(q.grad, passes, 9 * max(Cin, Cout))
        if n == "db" and act == "sigmoid":
            # a column sum of B*H*W terms g * y(1-y) whose y carries the forward's GEMM round-off: the error scales with
            # sqrt(#terms), not with the (cancelling) sum itself
            tol += (1.5e-3 if passes == 1 else 1e-5) * 0.1 * (B * H * W) ** 0.5
        _close(p.grad, q.grad, 0, tol, n)


def test_conv1x1_and_folded_bn():
    from gedepth_b200 import kernels as Kn
    g = torch.Generator().manual_seed(13)
    x = torch.randn(2, 96, 16, 40, generator=g).to(DEV).contiguous(memory_format=torch.channels_last)
    w = (torch.randn(192, 96, 1, 1, generator=g) / 10).to(DEV)
    bn = torch.nn.BatchNorm2d(192).to(DEV)
    with torch.no_grad():
        bn.running_mean.normal_(0, 0.1, generator=None)
        bn.running_var.uniform_(0.5, 1.5)
        bn.weight.normal_(1, 0.1)
        bn.bias.normal_(0, 0.1)
    bn.eval()
    y = Kn.conv_bn_act(x, w, None, bn, 1, 0, "relu")
    ref = F.relu(bn(F.conv2d(x.double(), w.double()).float()))
    _close(y, ref, 0, 2e-3 * float(ref.abs().max()), "1x1 conv + folded BN + relu")
    bn.train()
    y = Kn.conv_bn_act(x, w, None, bn, 1, 0, "relu")
    bn2 = torch.nn.BatchNorm2d(192).to(DEV)
    bn2.load_state_dict(bn.state_dict())
    bn2.train()
    ref = F.relu(bn2(F.conv2d(x.double(), w.double()).float()))
    _close(y, ref, 0, 3e-3 * float(ref.abs().max()), "1x1 conv + train BN + relu")


# ------------------------------------------------------------------------------------------------
# deformable attention sampling
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,shapes,Q,ref_b", [(2, [(16, 40), (8, 20), (4, 10), (2, 5)], 850, 1),
                                               (1, [(18, 42), (9, 21), (5, 11), (3, 6)], 333, 1),
                                               (2, [(16, 40), (8, 20), (4, 10), (2, 5)], 1280, 2),
                                               (3, [(44, 140), (22, 70), (11, 35), (6, 18)], 6160, 1)])
@pytest.mark.parametrize("impl", ["path", "tc", "tile", "round1-q8xh1", "round1-q1xh8", "round1-q8xh1-split", "round1-q1xh8-split"])
def test_msda_fwd_bwd(B, shapes, Q, ref_b, impl):
    """Sampling + all four gradients against the grid_sample statement (ops_lib, which equals the oracle's msda_core).
    `path` = the defaults of kernels.py (round-1 forward + tcgen05 backward); `tc` = tcgen05 both ways (csrc/msda_tc.cu:
    forward 3xTF32, backward one-pass TF32 like the other backward GEMMs, hence the 4e-3 gradient tolerance); `tile` =
    the fp32 sorted-tile kernels (csrc/msda_tile.cu, what GEDEPTH_BWD_GEMM_PASSES=3 selects for the backward);
    round1-* = the four work mappings of the round-1 kernels.  ref_b == 1 with B > 1 is the cross-attention case:
    learnable reference points shared by the batch, so g_ref is the sum over the batch."""
    from gedepth_b200 import kernels as Kn
    saved = (Kn.MSDA_FWD, Kn.MSDA_BWD, Kn.set_msda_variant(-1))
    try:
        if impl.startswith("round1"):
            Kn.MSDA_FWD = Kn.MSDA_BWD = "round1"
            Kn.set_msda_variant(["round1-q8xh1", "round1-q1xh8", "round1-q8xh1-split", "round1-q1xh8-split"].index(impl))
        elif impl != "path":
            Kn.MSDA_FWD = Kn.MSDA_BWD = impl
        _msda_case(B, shapes, Q, ref_b, tf32=Kn._msda_bwd_impl() == "tc")
    finally:
        Kn.MSDA_FWD, Kn.MSDA_BWD = saved[0], saved[1]
        Kn.set_msda_variant(saved[2])


def _msda_inputs(B, shapes, Q, ref_b, clustered):
    g = torch.Generator().manual_seed(14)
    S = sum(h * w for h, w in shapes)
    v0 = torch.randn(B, S, 512, generator=g)
    ref0 = torch.rand(ref_b, Q, 2, generator=g) * 1.1 - 0.05
    off0 = torch.randn(B, Q, 8 * 4 * 8 * 2, generator=g) * (1.5 if clustered else 2.5)
    if clustered:
        off0[:, : Q // 50] *= 12.0          # a few queries throw their points far outside every window / the map
    lg0 = torch.randn(B, Q, 8 * 32, generator=g)
    return v0, ref0, off0, lg0


def _msda_case(B, shapes, Q, ref_b, clustered=False, tf32=False):
    from gedepth_b200 import kernels as Kn
    from tests import ops_lib as L
    v0, ref0, off0, lg0 = _msda_inputs(B, shapes, Q, ref_b, clustered)
    a1 = [t.to(DEV).requires_grad_(True) for t in (v0, ref0, off0, lg0)]
    a2 = [t.to(DEV).requires_grad_(True) for t in (v0, ref0, off0, lg0)]
    o1 = Kn.msda_sample(a1[0], shapes, a1[1], a1[2], a1[3], 8, 8)
    o2 = L.msda_sample(a2[0], shapes, a2[1], a2[2], a2[3], 8, 8)
    _close(o1, o2, 1e-4, 2e-5 * float(o2.abs().max()), "sampled")
    go = torch.randn_like(o1)
    (o1 * go).sum().backward()
    (o2 * go).sum().backward()
    names = ["g_value", "g_ref", "g_off", "g_logit"]
    for n, p, q in zip(names, a1, a2):
        assert p.grad.shape == q.grad.shape, n
        d = (p.grad - q.grad).abs()
        rel, ab = (4e-3, 2e-3) if tf32 else (1e-3, 5e-5)
        tol = rel * q.grad.abs() + ab * float(q.grad.abs().max())
        # bilinear kinks: a sample landing within fp32 round-off of a pixel boundary may pick the other cell
        assert float((d > tol).float().mean()) < 2e-4, (n, float(d.max()), float(q.grad.abs().max()), float((d > tol).float().mean()))


def test_linear_small_fwd_bwd():
    """HAHIHeteroNeck.reference_points: Linear 512 -> 2 + sigmoid (csrc/small.cu) against torch fp64."""
    from gedepth_b200 import kernels as Kn
    g = torch.Generator().manual_seed(21)
    x0, w0, b0 = torch.randn(1, 3001, 512, generator=g), torch.randn(2, 512, generator=g) / 20, torch.randn(2, generator=g)
    for act in ("sigmoid", None):
        a1 = [t.to(DEV).requires_grad_(True) for t in (x0, w0, b0)]
        a2 = [t.to(DEV).double().requires_grad_(True) for t in (x0, w0, b0)]
        y1 = Kn.linear_small(a1[0], a1[1], a1[2], act)
        y2 = F.linear(a2[0], a2[1], a2[2])
        y2 = torch.sigmoid(y2) if act else y2
        _close(y1, y2, 1e-5, 1e-5, "linear_small fwd")
        go = torch.randn(y2.shape, generator=g).to(DEV)
        (y1 * go).sum().backward()
        (y2 * go.double()).sum().backward()
        for n, p, q in zip(("dx", "dw", "db"), a1, a2):
            _close(p.grad, q.grad, 1e-4, 1e-5 * float(q.grad.abs().max()) + 1e-6, "linear_small " + n)


@pytest.mark.parametrize("shared_value", [True, False], ids=["self", "cross"])
def test_msda_module_fwd_bwd(shared_value):
    """The fused deformable-attention node (query + pos + level embedding, three input GEMMs, sampling, output_proj +
    identity; gradient fan-in summed in GEMM epilogues) against the library statement built from the same module."""
    from gedepth_b200 import hahi, ops
    from tests import ops_lib as L
    torch.manual_seed(3)
    shapes = [(16, 40), (8, 20), (4, 10), (2, 5)]
    S = sum(h * w for h, w in shapes)
    B, Q = 2, (S if shared_value else 1280)
    starts = [0]
    for h, w in shapes:
        starts.append(starts[-1] + h * w)
    mods = []
    for _ in range(2):
        torch.manual_seed(5)
        m = hahi.MultiScaleDeformableAttention(512, num_levels=4, num_heads=8, num_points=8, batch_first=True).to(DEV)
        with torch.no_grad():
            for p_ in m.parameters():
                p_.add_(torch.randn_like(p_) * 0.05)
        mods.append(m)
    g = torch.Generator().manual_seed(4)
    q0 = torch.randn(B, Q, 512, generator=g)
    v0 = torch.randn(B, S, 512, generator=g)
    pos0 = torch.randn(1, Q, 512, generator=g)
    le0 = torch.randn(4, 512, generator=g)
    ref0 = torch.rand(1, Q, 2, generator=g)
    outs, grads = [], []
    for m, fn in zip(mods, (ops.msda_module, L.msda_module)):
        q, v, le, ref = [t.to(DEV).requires_grad_(True) for t in (q0, v0, le0, ref0)]
        use_le = shared_value
        out = fn(q, None if shared_value else v, pos0.to(DEV), le if use_le else None, starts if use_le else None, ref,
                 shapes, m, 0.0)
        go = torch.randn(out.shape, generator=torch.Generator().manual_seed(6)).to(DEV)
        (out * go).sum().backward()
        outs.append(out.detach())
        gr = dict(query=q.grad, ref=ref.grad)
        if not shared_value:
            gr["value"] = v.grad
        if use_le:
            gr["level_embed"] = le.grad
        for n, p_ in m.named_parameters():
            gr[n] = p_.grad
        grads.append(gr)
    _close(outs[0], outs[1], 0, 2e-4 * float(outs[1].abs().max()), "msda_module out")
    for n in grads[1]:
        a, b = grads[0][n], grads[1][n]
        assert a is not None, n
        d = (a - b).abs()
        tol = 1e-2 * b.abs() + 4e-3 * float(b.abs().max())        # one-pass TF32 backward GEMMs
        assert float((d > tol).float().mean()) < 1e-3, (n, float(d.max()), float(b.abs().max()))


def test_msda_tile_outliers_and_order_invariance():
    """Corners far outside their window (and outside the map) take the direct path; the result does not depend on the
    query order (identity, reversed and the sorted order agree: bit-identical for the fp32 tile forward, to 1e-5 for the
    3xTF32 tensor-core forward whose rounding depends on the window)."""
    from gedepth_b200 import kernels as Kn
    shapes = [(44, 140), (22, 70), (11, 35), (6, 18)]
    saved = (Kn.MSDA_FWD, Kn.MSDA_BWD)
    try:
        for impl in ("tile", "tc"):
            Kn.MSDA_FWD = Kn.MSDA_BWD = impl
            _msda_case(2, shapes, 3000, 1, clustered=True, tf32=impl == "tc")
        v0, ref0, off0, lg0 = _msda_inputs(2, shapes, 3000, 1, True)
        v, ref, off, lg = [t.to(DEV) for t in (v0, ref0, off0, lg0)]
        order = Kn.msda_query_order(ref, shapes)
        assert sorted(order.cpu().tolist()) == list(range(3000)), "the sort must return a permutation"
        Kn.MSDA_FWD = "tile"
        outs = []
        for o in (order, torch.arange(3000, dtype=torch.int32, device=DEV), torch.arange(2999, -1, -1, dtype=torch.int32, device=DEV)):
            ref._ged_order = (ref._version, o.contiguous())
            outs.append(Kn.msda_sample(v, shapes, ref, off, lg, 8, 8))
        assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2]), "forward must not depend on the query order"
        Kn.MSDA_FWD = "tc"
        for o in (order, torch.arange(2999, -1, -1, dtype=torch.int32, device=DEV)):
            ref._ged_order = (ref._version, o.contiguous())
            _close(Kn.msda_sample(v, shapes, ref, off, lg, 8, 8), outs[0], 1e-5, 1e-5 * float(outs[0].abs().max()), "tc forward vs fp32 tile forward")
    finally:
        Kn.MSDA_FWD, Kn.MSDA_BWD = saved


# ------------------------------------------------------------------------------------------------
# optimizer
# ------------------------------------------------------------------------------------------------
def test_adamw_and_clip_match_torch():
    from gedepth_b200 import kernels as Kn
    g = torch.Generator().manual_seed(15)
    n = 100003
    p0, g0 = torch.randn(n, generator=g), torch.randn(n, generator=g) * 3
    p_ref = torch.nn.Parameter(p0.clone().to(DEV))
    opt = torch.optim.AdamW([p_ref], lr=1e-4, betas=(0.9, 0.999), weight_decay=0.01)
    p = p0.clone().to(DEV)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    ss = torch.zeros(1, dtype=torch.float64, device=DEV)
    for step in range(1, 4):
        grad = (g0 * step).to(DEV)
        p_ref.grad = grad.clone()
        torch.nn.utils.clip_grad_norm_([p_ref], 35.0)
        opt.step()
        Kn.sumsq(grad, ss)
        Kn.adamw_step(p, grad, m, v, None, ss, 35.0, 1.0, 1e-4, 0.9, 0.999, 1e-8, 0.01, step,
                      torch.tensor([step], dtype=torch.int32, device=DEV) if step == 2 else None)
        assert abs(float(ss.sqrt()) - float(grad.double().norm())) < 1e-6 * float(grad.norm())
    _close(p, p_ref.data, 1e-6, 1e-6, "adamw")


# ------------------------------------------------------------------------------------------------
# data-movement kernels around the convs
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("rows_mode", [1, 0])
@pytest.mark.parametrize("B,h0,w0,H,W,C0,C1", [(2, 11, 35, 22, 70, 64, 32), (1, 9, 12, 9, 12, 32, 96), (2, 22, 70, 44, 139, 96, 0),
                                               (1, 88, 280, 176, 560, 192, 64), (2, 5, 7, 16, 22, 8, 4), (1, 3, 4, 11, 15, 4, 0),
                                               (1, 7, 9, 7, 9, 4, 4), (1, 4, 9, 32, 72, 8, 0), (2, 11, 35, 44, 140, 64, 0),
                                               (1, 2, 3, 40, 50, 4, 0)])
def test_prep_conv_input_and_adjoint(B, h0, w0, H, W, C0, C1, rows_mode):
    """rows_mode = 1: one CTA per output row, bilinear taps tabulated in shared memory (default); 0: the flat grid-stride
    kernels.  x4 / x8 ratios (the PE necks' resize_add adjoints) use longer tap tables; (1, 2, 3, 40, 50) is a x20+ ratio,
    which the row form of the adjoint hands to the flat kernel by itself."""
    from gedepth_b200 import kernels as Kn
    g = torch.Generator().manual_seed(20)
    x0 = torch.randn(B, h0, w0, C0, generator=g).to(DEV)
    x1 = torch.randn(B, H, W, C1, generator=g).to(DEV) if C1 else None
    prev = Kn.set_layout_rows(rows_mode)
    try:
        xp = Kn.prep_conv_input(x0, x1, H, W)
        gfull = torch.randn(B, H, W, C0 + C1, generator=g).to(DEV)
        out = torch.empty(B, h0, w0, C0, device=DEV)
        Kn._call("ged_upsample_nhwc_bwd", Kn._p(gfull), C0 + C1, Kn._p(out), C0, B, H, W, h0, w0, Kn._stream())
        torch.cuda.synchronize()
    finally:
        Kn.set_layout_rows(prev)
    up = F.interpolate(x0.permute(0, 3, 1, 2), size=(H, W), mode="bilinear", align_corners=True) if (h0, w0) != (H, W) else x0.permute(0, 3, 1, 2)
    ref = torch.cat([up] + ([x1.permute(0, 3, 1, 2)] if C1 else []), 1)
    ref = F.pad(ref, (1, 1, 1, 1)).permute(0, 2, 3, 1)
    _close(xp, ref, 1e-5, 1e-5, "prep")
    # adjoint of the resize on the first C0 channels
    x0r = x0.clone().requires_grad_(True)
    upr = F.interpolate(x0r.permute(0, 3, 1, 2), size=(H, W), mode="bilinear", align_corners=True)
    (upr * gfull[..., :C0].permute(0, 3, 1, 2)).sum().backward()
    _close(out, x0r.grad, 1e-4, 1e-5, "resize adjoint")


def test_prep_conv_input_batch_strided_sources(layout_rows):
    """Sources that are one level's slice of a (B, S, C) token tensor (hahi.py:338-353: samples dense, batch stride S * C) are
    read in place; the result equals the one from dense copies."""
    from gedepth_b200 import kernels as Kn
    g = torch.Generator().manual_seed(26)
    B, H, W, C0, C1, h0, w0 = 3, 12, 20, 32, 16, 6, 10
    tok0 = torch.randn(B, h0 * w0 + 37, C0, generator=g).to(DEV)
    tok1 = torch.randn(B, 11 + H * W, C1, generator=g).to(DEV)
    x0 = tok0[:, 5:5 + h0 * w0].reshape(B, h0, w0, C0)
    x1 = tok1[:, 11:].reshape(B, H, W, C1)
    assert not x0.is_contiguous() and Kn._batch_dense(x0) and Kn._batch_dense(x1)
    got = Kn.prep_conv_input(x0, x1, H, W)
    ref = Kn.prep_conv_input(x0.contiguous(), x1.contiguous(), H, W)
    assert torch.equal(got, ref)
    same = Kn.prep_conv_input(x1, None, H, W)
    assert torch.equal(same, Kn.prep_conv_input(x1.contiguous(), None, H, W))


@pytest.mark.parametrize("N", [96, 64, 24, 200])
@pytest.mark.parametrize("act", [None, "relu", "leaky_relu", "gelu", "sigmoid"])
def test_act_bwd(act, N):
    """N = 64 / 24: the 16- and 8-lane thread mappings of narrow matrices; 200: a partial last column block."""
    from gedepth_b200 import kernels as Kn
    from tests import ops_lib as L
    g = torch.Generator().manual_seed(21)
    rows, T = 3 * 217, 217          # 651 rows: five 128-row chunks + a ragged one (the two-row prefetch loop's tail)
    pre = torch.randn(rows, N, generator=g).to(DEV).requires_grad_(True)
    go = torch.randn(rows, N, generator=g).to(DEV)
    rs = torch.tensor([0.0, 1.4, 1.4]).to(DEV)
    y = L._act(pre, act, 0.01)
    (y * rs.repeat_interleave(T).unsqueeze(1) * go).sum().backward()
    ref = pre.detach() if act == "gelu" else (y.detach() if act else None)
    gz, db = Kn.act_bwd(go, ref, act, 0.01, rs, T, True)
    _close(gz, pre.grad, 1e-5, 1e-6, "gz")
    _close(db, pre.grad.sum(0), 1e-4, 1e-4, "db")
    gz2, db2 = Kn.act_bwd(go, None, None, 0.01, None, 1, True)
    assert gz2 is go
    _close(db2, go.sum(0), 1e-4, 1e-4, "db only")


@pytest.mark.parametrize("act", [None, "relu", "gelu"])
def test_act_bwd_long_row_chunks(act):
    """Tall matrices take 512-row chunks per CTA (column sums: 4x fewer atomics); same numbers as torch."""
    from gedepth_b200 import kernels as Kn
    from tests import ops_lib as L
    g = torch.Generator().manual_seed(22)
    B, T, N = 4, 20011, 1024                    # 80044 rows: (rows / 512) * 8 column blocks >= 148 * 8; a ragged last chunk
    rows = B * T
    pre = torch.randn(rows, N, generator=g).to(DEV).requires_grad_(True)
    go = torch.randn(rows, N, generator=g).to(DEV)
    rs = torch.tensor([0.0, 1.4, 1.4, 0.7]).to(DEV)
    y = L._act(pre, act, 0.01)
    (y * rs.repeat_interleave(T).unsqueeze(1) * go).sum().backward()
    ref = pre.detach() if act == "gelu" else (y.detach() if act else None)
    gz, db = Kn.act_bwd(go, ref, act, 0.01, rs, T, True)
    _close(gz, pre.grad, 1e-5, 2e-6, "gz")
    _close(db, pre.grad.sum(0), 1e-4, 1e-4 * float(pre.grad.sum(0).abs().max()), "db")


def test_resize_add_fwd_bwd():
    from gedepth_b200 import kernels as Kn
    g = torch.Generator().manual_seed(22)
    t0, a0 = torch.randn(2, 64, 11, 35, generator=g), torch.randn(2, 64, 44, 140, generator=g)
    a1 = [x.to(DEV).contiguous(memory_format=torch.channels_last).requires_grad_(True) for x in (t0, a0)]
    a2 = [x.to(DEV).requires_grad_(True) for x in (t0, a0)]
    o1 = Kn.resize_add(a1[0], (44, 140), a1[1])
    o2 = a2[1] + F.interpolate(a2[0], size=(44, 140), mode="bilinear", align_corners=True)
    _close(o1, o2, 1e-5, 1e-5, "resize_add")
    go = torch.randn_like(o2)
    (o1 * go).sum().backward()
    (o2 * go).sum().backward()
    _close(a1[0].grad, a2[0].grad, 1e-4, 1e-5, "g_t")
    _close(a1[1].grad, a2[1].grad, 0, 0, "g_acc")


def test_conv2d_cat_upsample_fwd_bwd(passes):
    from gedepth_b200 import kernels as Kn
    g = torch.Generator().manual_seed(23)
    B, h0, w0, H, W, C0, C1, Co = 2, 11, 35, 22, 70, 64, 32, 64
    x0, x1 = torch.randn(B, C0, h0, w0, generator=g), torch.randn(B, C1, H, W, generator=g)
    w0_, b0 = torch.randn(Co, C0 + C1, 3, 3, generator=g) / (9 * (C0 + C1)) ** 0.5, torch.randn(Co, generator=g)
    a1 = [t.to(DEV).contiguous(memory_format=torch.channels_last).requires_grad_(True) for t in (x0, x1)] + \
         [t.to(DEV).requires_grad_(True) for t in (w0_, b0)]
    a2 = [t.to(DEV).double().requires_grad_(True) for t in (x0, x1, w0_, b0)]
    y1 = Kn.conv2d_cat(a1[0], a1[1], a1[2], a1[3], None, 0.0)
    up = F.interpolate(a2[0], size=(H, W), mode="bilinear", align_corners=True)
    y2 = F.conv2d(torch.cat([up, a2[1]], 1), a2[2], a2[3], padding=1)
    _close(y1, y2, 0, _gemm_tol(y2, passes, 9 * (C0 + C1)), "fwd")
    go = torch.randn_like(y2).float()
    (y1 * go).sum().backward()
    (y2 * go.double()).sum().backward()
    for n, p, q in zip(("dx_low", "dx_skip", "dw", "db"), a1, a2):
        _close(p.grad, q.grad, 0, 2 * _gemm_tol(q.grad, passes, 9 * (C0 + C1)), n)


@pytest.mark.parametrize("H,W", [(64, 160), (70, 166)])
def test_patch_embed_as_gemm(H, W, passes):
    from gedepth_b200 import kernels as Kn
    from tests import ops_lib as L
    g = torch.Generator().manual_seed(24)
    img = torch.randn(2, 5, H, W, generator=g).to(DEV)
    w = (torch.randn(96, 4, 4, 4, generator=g) / 8).to(DEV).requires_grad_(True)
    b = torch.randn(96, generator=g).to(DEV).requires_grad_(True)
    t1, hw1 = Kn.patch_embed(img[:, 0:4], w, b, 4)
    w2, b2 = w.detach().clone().requires_grad_(True), b.detach().clone().requires_grad_(True)
    t2, hw2 = L.patch_embed(img[:, 0:4], w2, b2, 4)
    assert tuple(hw1) == tuple(hw2)
    _close(t1, t2, 0, _gemm_tol(t2, passes, 64) + 1e-5, "tokens")
    go = torch.randn_like(t2)
    (t1 * go).sum().backward()
    (t2 * go).sum().backward()
    _close(w.grad, w2.grad, 0, 2e-3 * float(w2.grad.abs().max()), "dw")     # dW: library GEMM in both
    _close(b.grad, b2.grad, 1e-4, 1e-3, "db")


@pytest.fixture(params=[1, 0], ids=["rows", "flat"])
def layout_rows(request):
    """Both forms of the data-movement kernels: one CTA per output row (default) and the flat grid-stride kernels."""
    from gedepth_b200 import kernels as Kn
    prev = Kn.set_layout_rows(request.param)
    yield request.param
    Kn.set_layout_rows(prev)


@pytest.mark.parametrize("H,W", [(64, 160), (35, 83)])
def test_stem_conv_as_im2col_gemm(H, W, passes, layout_rows):
    """7x7/s2/p3 stem conv on channels 0-2 of the 5-channel batch (depthformer_swin.py:1032-1039,1152) as
    im2col + tcgen05 GEMM, forward and weight gradient, vs F.conv2d in fp64."""
    from gedepth_b200 import kernels as Kn
    g = torch.Generator().manual_seed(31)
    img = torch.randn(2, 5, H, W, generator=g).to(DEV)
    w0 = (torch.randn(64, 3, 7, 7, generator=g) / 147 ** 0.5)
    w1 = w0.to(DEV).requires_grad_(True)
    w2 = w0.to(DEV).double().requires_grad_(True)
    x = img[:, 0:3]
    assert Kn.conv_im2col_supported(x, w1, 2, 3)
    y1 = Kn.conv_im2col(x, w1, None, 2, 3, "relu")
    pre2 = F.conv2d(x.double(), w2, None, stride=2, padding=3)
    y2 = torch.where(y1.detach() > 0, pre2, torch.zeros_like(pre2))       # the branch the kernel took
    assert y1.shape == y2.shape
    _close(y1, F.relu(pre2), 0, _gemm_tol(pre2, passes, 147), "stem fwd")
    go = torch.randn_like(y2).float()
    (y1 * go).sum().backward()
    (y2 * go.double()).sum().backward()
    _close(w1.grad, w2.grad, 0, 2 * _gemm_tol(w2.grad, passes, 2 * (H // 2) * (W // 2)), "stem dW")


@pytest.mark.parametrize("H,W,C", [(16, 40, 96), (9, 21, 192), (5, 11, 384)])
def test_merge_patches_fwd_bwd(H, W, C, layout_rows):
    from gedepth_b200 import kernels as Kn
    from tests import ops_lib as L
    g = torch.Generator().manual_seed(25)
    x0 = torch.randn(2, H * W, C, generator=g)
    a, c = x0.to(DEV).requires_grad_(True), x0.to(DEV).requires_grad_(True)
    m1, m2 = Kn.merge_patches(a, H, W), L.merge_patches(c, H, W)
    assert torch.equal(m1, m2)
    go = torch.randn_like(m2)
    (m1 * go).sum().backward()
    (m2 * go).sum().backward()
    assert torch.equal(a.grad, c.grad)


def test_clamp_resize():
    from gedepth_b200 import kernels as Kn
    from tests import ops_lib as L
    g = torch.Generator().manual_seed(26)
    x = (torch.rand(2, 1, 35, 83, generator=g) * 120 - 10).to(DEV)
    with torch.no_grad():
        _close(Kn.clamp_resize(x, 1e-3, 80.0, (70, 166)), L.clamp_resize(x, 1e-3, 80.0, (70, 166), True), 1e-5, 1e-5, "clamp_resize")


@pytest.mark.parametrize("B,C,H,W,relu", [(2, 64, 32, 80, True), (3, 192, 9, 21, False), (2, 1536, 2, 5, True), (2, 24, 16, 20, True), (1, 200, 7, 9, False)])
def test_batchnorm_train_fwd_bwd(B, C, H, W, relu):
    from gedepth_b200 import kernels as Kn
    g = torch.Generator().manual_seed(27)
    x0 = torch.randn(B, C, H, W, generator=g) * 2 + 0.5
    bn1, bn2 = torch.nn.BatchNorm2d(C).to(DEV), torch.nn.BatchNorm2d(C).to(DEV)
    with torch.no_grad():
        bn1.weight.normal_(1, 0.2); bn1.bias.normal_(0, 0.2)
        bn2.load_state_dict(bn1.state_dict())
    bn1.train(); bn2.train()
    a = x0.to(DEV).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    c = x0.to(DEV).requires_grad_(True)
    y1 = Kn.bn_act_train(a, bn1, relu)
    y2 = bn2(c)
    y2 = F.relu(y2) if relu else y2
    _close(y1, y2, 1e-4, 1e-5, "bn fwd")
    _close(bn1.running_mean, bn2.running_mean, 1e-5, 1e-6, "running_mean")
    _close(bn1.running_var, bn2.running_var, 1e-4, 1e-6, "running_var")
    assert int(bn1.num_batches_tracked) == 1
    go = torch.randn_like(y2)
    (y1 * go).sum().backward()
    (y2 * go).sum().backward()
    _close(a.grad, c.grad, 1e-3, 2e-5 * float(c.grad.abs().max()), "dx")
    _close(bn1.weight.grad, bn2.weight.grad, 1e-3, 1e-4 * float(bn2.weight.grad.abs().max()), "dw")
    _close(bn1.bias.grad, bn2.bias.grad, 1e-3, 1e-4 * float(bn2.bias.grad.abs().max()), "db")


# ------------------------------------------------------------------------------------------------
# evaluation on the device
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,H,W,garg", [(3, 352, 1216, True), (2, 70, 166, False), (2, 384, 640, True)])
def test_depth_metrics_match_reference_numpy(B, H, W, garg):
    """Nine metrics + Garg crop + min/max mask + nan-mean over images vs the numpy restatement of
    depth/core/evaluation/metrics.py / kitti.py:366-385 (oracle.ground.depth_metrics)."""
    from gedepth_b200 import metrics as Mx
    from oracle import ground as og
    g = torch.Generator().manual_seed(61)
    gt = torch.rand(B, H, W, generator=g) * 90 + 0.5
    gt[torch.rand(B, H, W, generator=g) < 0.8] = 0          # LiDAR-sparse
    gt[B - 1] = 0                                           # an image without valid pixels -> NaN row, skipped by nanmean
    pred = (gt * (1 + 0.1 * torch.randn(B, H, W, generator=g))).clamp(1e-3, 80) + (gt == 0) * 10.0
    dm = Mx.DepthMetrics(1e-3, 80.0, garg_crop=garg)
    per_image = dm.update(pred.to(DEV), gt.to(DEV)).cpu().numpy()
    rect = Mx.crop_rect(H, W, garg)
    rows = []
    for b in range(B):
        gtb, pb = gt[b].numpy(), pred[b].numpy()
        mask = np.ones((H, W), bool)
        if rect is not None:
            mask[:] = False
            mask[rect[0]:rect[1], rect[2]:rect[3]] = True
        valid = mask & (gtb > 1e-3) & (gtb < 80.0)
        rows.append(og.depth_metrics(gtb[valid], pb[valid], 1e-3, 80.0))
    ref = np.array(rows, dtype=np.float64)
    assert np.isnan(per_image[B - 1]).all() and np.isnan(ref[B - 1]).all()
    np.testing.assert_allclose(per_image[:B - 1], ref[:B - 1], rtol=2e-5, atol=1e-6)
    out = dm.compute()
    assert list(out) == ["a1", "a2", "a3", "abs_rel", "rmse", "log_10", "rmse_log", "silog", "sq_rel"]
    np.testing.assert_allclose(list(out.values()), np.nanmean(ref, 0), rtol=2e-5, atol=1e-6)
    assert abs(out["abs_rel"] - np.nanmean(ref, 0)[3]) < 1e-4        # the north star's Abs-Rel bar


def test_tta_merge_is_unflip_and_average():
    from gedepth_b200 import kernels as Kn
    a = torch.randn(2, 1, 37, 53, device=DEV)
    b = torch.randn(2, 1, 37, 53, device=DEV)
    _close(Kn.tta_merge(a, b), (a + b.flip(3)) / 2, 0, 1e-7, "tta")


@pytest.mark.parametrize("H0,W0", [(375, 1242), (370, 1226)])
def test_test_time_views_bit_exact_vs_reference_pipeline(H0, W0):
    """KBCrop -> (flip) -> Normalize of the reference's test pipeline (transforms.py:40-48,176-197; mmcv.imnormalize via
    cv2) on the 5-channel image of loading.py:490-527, restated with numpy / cv2, vs the device-side builder."""
    import cv2
    from gedepth_b200 import inputs as I
    from oracle import ground as og
    rng = np.random.default_rng(5)
    bgr = rng.integers(0, 256, (H0, W0, 3), dtype=np.uint8)
    coef = og.plane_coefficients(og.kitti_projection(), og.KITTI_CAM_HEIGHT)
    imgs, metas, pts = I.build_test_views(torch.from_numpy(bgr).to(DEV), coef)
    # reference restatement
    pe = og.ground_plane(coef, H0, W0)
    ch3, ch4 = og.load_channels(pe, 200.0)
    img5 = np.concatenate([bgr.astype(np.float32), ch3[..., None], ch4[..., None]], -1)          # loading.py:524-527
    top, left = int(H0 - 352), int((W0 - 1216) / 2)
    img5 = img5[top:top + 352, left:left + 1216]
    assert abs(float(pts[0]) - float(ch4[-1, -1])) <= 1e-6 * abs(float(ch4[-1, -1]))
    for view, flip in zip(imgs, (False, True)):
        x = img5[:, ::-1].copy() if flip else img5.copy()                                      # mmcv.imflip horizontal
        rgb = x[:, :, 0:3].copy().astype(np.uint8).astype(np.float32)
        mean = np.float64(np.array(I.KITTI_MEAN, np.float32).reshape(1, -1))
        stdinv = 1 / np.float64(np.array(I.KITTI_STD, np.float32).reshape(1, -1))
        cv2.cvtColor(rgb, cv2.COLOR_BGR2RGB, rgb); cv2.subtract(rgb, mean, rgb); cv2.multiply(rgb, stdinv, rgb)
        want = np.concatenate([rgb, og.normalize_pe(x[:, :, 3], 200.0)[..., None], x[:, :, 4:5]], -1).transpose(2, 0, 1)
        got = view[0].cpu().numpy()
        assert got.shape == (5, 352, 1216)
        assert np.array_equal(got[0:3], want[0:3]), "RGB planes must equal cv2's float32 result bit for bit"
        assert np.array_equal(got[3], want[3]) and np.array_equal(got[4], want[4]), "ground-plane channels"
    assert metas[1][0]["flip"] and metas[1][0]["flip_direction"] == "horizontal"
