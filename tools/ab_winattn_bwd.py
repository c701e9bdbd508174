"""A/B of the window-attention core backward: SIMT fp32 (csrc/winattn.cu) vs tcgen05 one-pass TF32 (csrc/winattn_tc.cu) at
the four Swin-L stage shapes of 352 x 1120, through the C ABI, L2 flushed between repetitions.
usage: python tools/ab_winattn_bwd.py [B] [stage]"""
import sys, torch
sys.path.insert(0, '.')
from gedepth_b200 import kernels as K
from gedepth_b200.kernels import _call, _p, _stream
DEV = 'cuda:0'
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
ONLY = int(sys.argv[2]) if len(sys.argv) > 2 else -1
K.load()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
_c = torch.arange(7)
_yy, _xx = torch.meshgrid(_c, _c, indexing="ij")
_y, _x = _yy.reshape(-1), _xx.reshape(-1)
index = ((_y[:, None] - _y[None, :] + 6) * 13 + (_x[:, None] - _x[None, :] + 6)).to(DEV)     # depthformer_swin.py:168-172


def t_ms(fn, reps=5):
    fn(); torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        tot += s.elapsed_time(e)
    return tot / reps


tot = {"simt": 0.0, "tcgen05": 0.0}
for st, (hh, ww) in enumerate(((88, 280), (44, 140), (22, 70), (11, 35))):
    if ONLY >= 0 and st != ONLY:
        continue
    nH = (6, 12, 24, 48)[st]
    blocks = (2, 2, 18, 2)[st]
    C = nH * 32
    qkv = torch.randn(B, hh * ww, 3 * C, device=DEV)
    bias = torch.randn(3 * C, device=DEV) * 0.1
    table = torch.randn(169, nH, device=DEV) * 0.2
    g = torch.randn(B, hh * ww, C, device=DEV)
    g_qkv, g_bias, g_table = torch.empty_like(qkv), torch.zeros(3 * C, device=DEV), torch.zeros_like(table)
    simt = lambda: _call("ged_winattn_bwd", _p(qkv), _p(bias), _p(table), _p(index), _p(g), _p(g_qkv), _p(g_bias), _p(g_table),
                         B, hh, ww, C, nH, 7, 3, 32 ** -0.5, _stream())
    tc = lambda: _call("ged_winattn_tc_bwd", _p(qkv), _p(bias), _p(table), _p(g), _p(g_qkv), _p(g_bias), _p(g_table),
                       B, hh, ww, C, nH, 7, 3, 32 ** -0.5, _stream())
    mma = lambda: _call("ged_winattn_bwd_mma", _p(qkv), _p(bias), _p(table), _p(index), _p(g), _p(g_qkv), _p(g_bias), _p(g_table),
                        B, hh, ww, C, nH, 7, 3, 32 ** -0.5, 1, _stream())
    a, b, m = t_ms(simt), t_ms(tc), t_ms(mma)
    tot["simt"] += a * blocks; tot["tcgen05"] += b * blocks; tot["mma"] = tot.get("mma", 0.0) + m * blocks
    print(f"   mma.sync {m:.3f} ms")
    pairs = (-(-hh // 7)) * (-(-ww // 7)) * B * nH
    print(f"stage {st} B={B} {hh}x{ww} heads {nH}: simt {a:.3f} ms, tcgen05 {b:.3f} ms ({b * 1e3 / (pairs / 2) * 148:.2f} us per duo per SM)", flush=True)
print(f"swin_l backward cores per step (blocks weighted): simt {tot['simt']:.2f} ms, tcgen05 {tot['tcgen05']:.2f} ms, mma.sync {tot.get('mma', 0):.2f} ms")
