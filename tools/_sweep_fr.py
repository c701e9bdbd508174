import json, os, sys, torch
sys.path.insert(0, '.')
from gedepth_b200 import kernels as K
DEV = 'cuda:0'
peak = json.load(open('MEASURED_PEAKS.json'))['hbm_gbs']
flush = torch.empty(256 * 1024 * 1024 // 4, device=DEV)
def timed(fn, reps=7):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    ts.sort(); return ts[len(ts) // 2] * 1e3
print("FR", os.environ.get("GEDEPTH_VF_FR"), "CH", os.environ.get("GEDEPTH_VB_CH"))
for (B, H, W) in [(8, 352, 1120), (16, 352, 1120), (32, 352, 1120), (64, 352, 1120), (64, 384, 640), (8, 1024, 2048), (32, 1024, 2048), (64, 1024, 2048)]:
    img = torch.randn(B, 5, H, W, device=DEV); yh = torch.rand(B, 1, H // 2, W // 2, device=DEV)
    y, pm = torch.empty(B, 1, H, W, device=DEV), torch.empty(B, 1, H, W, device=DEV)
    gy, gp = torch.randn(B, 1, H, W, device=DEV), torch.randn(B, 1, H, W, device=DEV)
    out = torch.empty(B, 1, H // 2, W // 2, device=DEV)
    mb = 13 * B * H * W / 1e6
    f = timed(lambda: K._call("ged_ge_vanilla_fwd", K._p(img[:, 3]), img.stride(0), K._p(yh), K._p(y), K._p(pm), B, H, W, H // 2, W // 2, K._stream()))
    b = timed(lambda: K._call("ged_ge_vanilla_bwd", K._p(img[:, 3]), img.stride(0), K._p(gy), K._p(gp), K._p(out), B, H, W, H // 2, W // 2, K._stream()))
    print(f"{B}x{H}x{W}: fwd {f:7.1f} us {mb / f * 1e3 / peak:.3f} | bwd {b:7.1f} us {mb / b * 1e3 / peak:.3f}")
    del img, yh, y, pm, gy, gp, out
