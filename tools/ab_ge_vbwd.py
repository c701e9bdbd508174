"""ge_vanilla_bwd x2: streaming (warp-per-column-strip) kernel vs the tiled one (ged_set_ge_x2(2)), equality + bandwidth.
usage: [GEDEPTH_VB_VARIANT=k] python tools/ab_ge_vbwd.py"""
import json, os, sys, torch
sys.path.insert(0, '.')
from gedepth_b200 import kernels as K
DEV = 'cuda:0'
peak = json.load(open('MEASURED_PEAKS.json'))['hbm_gbs']
flush = torch.empty(256 * 1024 * 1024 // 4, device=DEV)


def run(pe, stride, gy, gp, out, B, H, W):
    K._call("ged_ge_vanilla_bwd", K._p(pe), stride, K._p(gy), K._p(gp), K._p(out), B, H, W, H // 2, W // 2, K._stream())


def timed(fn, reps=7):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2] * 1e3


print("variant", os.environ.get("GEDEPTH_VB_VARIANT", "0"))
for (B, H, W) in [(2, 64, 160), (1, 4, 8), (2, 6, 12), (3, 352, 1120), (1, 384, 640), (2, 130, 516), (1, 34, 1032)]:
    img = torch.randn(B, 5, H, W, device=DEV)
    gy, gp = torch.randn(B, 1, H, W, device=DEV), torch.randn(B, 1, H, W, device=DEV)
    a = torch.full((B, 1, H // 2, W // 2), 7.0, device=DEV); b = torch.full_like(a, -3.0)
    pe = img[:, 3]
    K.set_ge_x2(1); run(pe, img.stride(0), gy, gp, a, B, H, W)
    K.set_ge_x2(2); run(pe, img.stride(0), gy, gp, b, B, H, W)
    K.set_ge_x2(1)
    # optional operands
    c = torch.empty_like(a); d = torch.empty_like(a)
    run(pe, img.stride(0), None, gp, c, B, H, W)
    K.set_ge_x2(2); run(pe, img.stride(0), None, gp, d, B, H, W); K.set_ge_x2(1)
    torch.cuda.synchronize()
    print(f"equal {B}x{H}x{W}: max|diff| {float((a - b).abs().max()):.3g} bitwise {bool(torch.equal(a, b))}; no g_y: {bool(torch.equal(c, d))}")
for (B, H, W) in [(16, 352, 1120), (32, 1024, 2048), (64, 1024, 2048)]:
    img = torch.randn(B, 5, H, W, device=DEV)
    gy, gp = torch.randn(B, 1, H, W, device=DEV), torch.randn(B, 1, H, W, device=DEV)
    out = torch.empty(B, 1, H // 2, W // 2, device=DEV)
    pe = img[:, 3]
    mb = 13 * B * H * W / 1e6
    for mode, name in ((1, "stream"), (2, "tiled")):
        K.set_ge_x2(mode)
        us = timed(lambda: run(pe, img.stride(0), gy, gp, out, B, H, W))
        print(f"{name:7s} {B}x{H}x{W}: {us:8.1f} us  {mb / us * 1e3:6.0f} GB/s  frac {mb / us * 1e3 / peak:.3f}")
    K.set_ge_x2(1)
    del img, gy, gp, out
