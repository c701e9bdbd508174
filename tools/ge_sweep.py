"""BASELINE config 5: ground-embedding kernel HBM-bandwidth sweep (256x512 -> 1024x2048, batch 1-64) on one GPU.
CUDA-event timing, L2 flushed (256 MB write) before every launch; algorithmic bytes per SURVEY.md §8(d):
Vanilla fwd 13 B/px, Vanilla bwd 13 B/px, Adaptive fwd 24 B/px (inference) / 68 B/px (training), Adaptive bwd 62 B/px,
ground_plane 8 B/px.
Writes CSV to stdout:  kernel,B,H,W,MB,us,GBs,frac_of_peak,l2_resident
"""
import json, os, sys, torch
sys.path.insert(0, '.')
from gedepth_b200 import kernels as K
from gedepth_b200.synth import kitti_plane_coef
DEV = 'cuda:0'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
peak = 6650.0
p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
if os.path.exists(p):
    peak = json.load(open(p))['hbm_gbs']
flush = torch.empty(256 * 1024 * 1024 // 4, device=DEV)


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        tot += s.elapsed_time(e)
    return tot / reps * 1e3      # us


print(f"# peak {peak} GB/s (MEASURED_PEAKS.json hbm_gbs); L2 flushed before every timed launch")
print("kernel,B,H,W,MB,us,GBs,frac_of_peak,l2_resident")
coef = kitti_plane_coef()
for (H, W) in [(256, 512), (352, 1120), (384, 640), (512, 1024), (768, 1536), (1024, 2048)]:
    for B in (1, 2, 4, 8, 16, 32, 64):
        if B * H * W * 68 > 40e9 or (B not in (1, 8, 64) and (H, W) not in ((352, 1120), (1024, 2048))):
            continue
        img = torch.randn(B, 5, H, W, device=DEV)
        yh = torch.rand(B, 1, H // 2, W // 2, device=DEV)
        lh = torch.randn(B, 11, H // 2, W // 2, device=DEV)
        gy, gp = torch.randn(B, 1, H, W, device=DEV), torch.randn(B, 1, H, W, device=DEV)
        px = B * H * W
        cases = []
        with torch.no_grad():
            cases.append(("ge_vanilla_fwd", 13, lambda: K.ge_vanilla(img, yh)))
            cases.append(("ge_adaptive_fwd_infer", 24, lambda: K.ge_adaptive(img, yh, lh, 1.65, 200.0)))
            cases.append(("ground_plane", 8, lambda: K.ground_plane_into(img, coef)))

            def adaptive_train():
                with torch.enable_grad():       # grad mode: the 11 full-resolution logits are written too (68 B/px)
                    K.ge_adaptive(img, yh, lh, 1.65, 200.0)
            cases.append(("ge_adaptive_fwd_train", 68, adaptive_train))
            for name, bpp, fn in cases:
                us = timed(fn)
                mb = bpp * px / 1e6
                print(f"{name},{B},{H},{W},{mb:.1f},{us:.1f},{mb / us * 1e3:.0f},{mb / us * 1e3 / peak:.3f},{int(mb < 126)}", flush=True)
        # backward of the vanilla kernel through the C ABI (13 B/px)
        g_half = torch.empty(B, 1, H // 2, W // 2, device=DEV)
        pe = img[:, 3]

        def bwd():        # (the C entry point zero-fills by itself where the generic kernel needs it; the x2 kernels write every element)
            K._call("ged_ge_vanilla_bwd", K._p(pe), img.stride(0), K._p(gy), K._p(gp), K._p(g_half), B, H, W, H // 2, W // 2, K._stream())
        us = timed(bwd)
        mb = 13 * px / 1e6
        print(f"ge_vanilla_bwd,{B},{H},{W},{mb:.1f},{us:.1f},{mb / us * 1e3:.0f},{mb / us * 1e3 / peak:.3f},{int(mb < 126)}", flush=True)
        # backward of the adaptive kernel through the C ABI (62 B/px: g_y, g_pe_mask, pe + 44 B/px g_logits + half-res in / out)
        glf = torch.randn(B, 11, H, W, device=DEV)
        g_lh = torch.empty(B, 11, H // 2, W // 2, device=DEV)
        pe4 = img[:, 4]

        def abwd():
            K._call("ged_ge_adaptive_bwd", K._p(pe4), img.stride(0), K._p(yh), K._p(lh), None, 1.65, 200.0, K._p(gy), K._p(gp),
                    K._p(glf), K._p(g_half), K._p(g_lh), B, H, W, H // 2, W // 2, K._stream())
        us = timed(abwd)
        mb = 62 * px / 1e6
        print(f"ge_adaptive_bwd,{B},{H},{W},{mb:.1f},{us:.1f},{mb / us * 1e3:.0f},{mb / us * 1e3 / peak:.3f},{int(mb < 126)}", flush=True)
        del img, yh, lh, gy, gp, g_half, glf, g_lh
