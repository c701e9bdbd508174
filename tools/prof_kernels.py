"""Representative single launches of the hot kernels at BASELINE-config-2 shapes (for ncu)."""
import sys, torch
sys.path.insert(0, '.')
from gedepth_b200 import kernels as K
from gedepth_b200.swin import WindowMSA
DEV = 'cuda:0'
torch.manual_seed(0)
B = 8
T0 = B * 88 * 280
which = sys.argv[1] if len(sys.argv) > 1 else 'all'
def rnd(*s): return torch.randn(*s, device=DEV)
if which in ('all', 'gemm'):
    for (M, N, Kd, res) in [(T0, 288, 96, False), (T0, 96, 384, True), (T0 // 16, 1536, 384, False), (B * 98560, 512, 512, True), (B * 32725, 256, 512, False)]:
        a, w, bias = rnd(M, Kd), rnd(N, Kd) / Kd ** .5, rnd(N)
        r = rnd(M, N) if res else None
        K.gemm(a, w, bias, None, 0.01, r)
        torch.cuda.synchronize()
if which in ('all', 'dw'):
    K.set_gemm_precision(1)
    for (P, N, Kd) in [(B * 98560, 512, 512), (T0, 384, 96)]:
        g, x = rnd(P, N), rnd(P, Kd)
        K.gemm_dw(g, x)
        torch.cuda.synchronize()
    x = rnd(B, 88, 280, 576); gy = rnd(B, 88, 280, 192)
    xp, gp = K.prep_conv_input(x, None, 88, 280), K.prep_conv_input(gy, None, 88, 280)
    taps = [(ky - 1) * 282 + (kx - 1) for ky in range(3) for kx in range(3)]
    K.gemm_dw(gp.reshape(-1, 192), xp.reshape(-1, 576), None, taps)
    torch.cuda.synchronize()
    K.set_gemm_precision(3)
if which in ('all', 'conv'):
    for (Bc, H, W, Ci, Co) in [(B, 176, 560, 256, 64), (B, 22, 70, 2304, 768), (B, 88, 280, 576, 192)]:
        x = rnd(Bc, H, W, Ci); wk = rnd(Co, 3, 3, Ci) / (9 * Ci) ** .5
        K.conv3x3_raw(x, wk, rnd(Co), 'leaky_relu', 0.01)
        torch.cuda.synchronize()
if which in ('all', 'msda'):
    shapes = [(88, 280), (44, 140), (22, 70), (11, 35)]
    S = sum(h * w for h, w in shapes)
    Bm = 2
    for Q, refb in [(176 * 560, Bm), (S, 1)]:
        v = rnd(Bm, S, 512).requires_grad_(True)
        ref = torch.rand(refb, Q, 2, device=DEV)
        if refb == 1:
            refs = []
            for h, w in shapes:
                ry, rx = torch.meshgrid(torch.linspace(0.5, h - 0.5, h, device=DEV), torch.linspace(0.5, w - 0.5, w, device=DEV), indexing='ij')
                refs.append(torch.stack((rx.reshape(-1) / w, ry.reshape(-1) / h), -1))
            ref = torch.cat(refs, 0)[None].contiguous()
        else:
            # smooth learned reference points (sigmoid of a smooth function of position)
            yy, xx = torch.meshgrid(torch.linspace(0, 1, 176, device=DEV), torch.linspace(0, 1, 560, device=DEV), indexing='ij')
            ref = torch.stack((0.1 + 0.8 * xx + 0.02 * torch.sin(6 * yy), 0.1 + 0.8 * yy + 0.02 * torch.cos(5 * xx)), -1).reshape(1, Q, 2).expand(Bm, -1, -1).contiguous()
        off = (rnd(Bm, Q, 512) * 3).requires_grad_(True)
        lg = rnd(Bm, Q, 256).requires_grad_(True)
        out = K.msda_sample(v, shapes, ref, off, lg, 8, 8)
        out.backward(rnd(*out.shape))
        torch.cuda.synchronize()
if which in ('all', 'attn'):
    qkv = rnd(B, 88 * 280, 288).requires_grad_(True)
    idx = WindowMSA(96, 3, (7, 7)).relative_position_index.to(DEV)
    o = K.window_attention(qkv, rnd(288), rnd(169, 3), idx, (88, 280), 3, 7, 3, 32 ** -.5)
    o.backward(rnd(*o.shape))
    torch.cuda.synchronize()
if which in ('all', 'ge'):
    img = rnd(32, 5, 1024, 2048); yh = torch.rand(32, 1, 512, 1024, device=DEV)
    with torch.no_grad():
        K.ge_vanilla(img, yh)
    img = rnd(B, 5, 352, 1120); yh = torch.rand(B, 1, 176, 560, device=DEV); lh = rnd(B, 11, 176, 560)
    with torch.no_grad():
        K.ge_vanilla(img, yh); K.ge_adaptive(img, yh, lh, 1.65, 200.0)
    torch.cuda.synchronize()
print('done')
