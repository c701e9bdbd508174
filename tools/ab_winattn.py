"""A/B of the window-attention core forward: SIMT fp32 (csrc/winattn.cu) vs tcgen05 3xTF32 (csrc/winattn_tc.cu) at the
four Swin stage shapes of 352 x 1120 (Swin-L and Swin-T head counts), L2 flushed between repetitions."""
import sys, torch
sys.path.insert(0, '.')
from gedepth_b200 import kernels as K
DEV = 'cuda:0'
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
ONLY = sys.argv[2] if len(sys.argv) > 2 else ""
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
for name, e0, heads, blocks in (("swin_l", 192, (6, 12, 24, 48), (2, 2, 18, 2)), ("swin_t", 96, (3, 6, 12, 24), (2, 2, 6, 2))):
    if ONLY and name != ONLY:
        continue
    for st, (hh, ww) in enumerate(((88, 280), (44, 140), (22, 70), (11, 35))):
        nH = heads[st]
        C = nH * 32
        qkv = torch.randn(B, hh * ww, 3 * C, device=DEV)
        bias = torch.randn(3 * C, device=DEV) * 0.1
        table = torch.randn(169, nH, device=DEV) * 0.2
        res = {}
        for core in ("simt", "tcgen05"):
            K.WINATTN_TC = core == "tcgen05"
            with torch.no_grad():
                res[core] = t_ms(lambda: K.window_attention(qkv, bias, table, index, (hh, ww), nH, 7, 3, 32 ** -0.5))
            tot[core] += res[core] * blocks[st] if name == "swin_l" else 0.0
        nW = (-(-hh // 7)) * (-(-ww // 7)) * B
        gb = nW * nH * 49 * 32 * 4 * 4 / 1e9
        print(f"{name} stage {st} B={B} {hh}x{ww} heads {nH}: simt {res['simt']:.3f} ms, tcgen05 {res['tcgen05']:.3f} ms "
              f"({gb / res['tcgen05'] * 1e3:.0f} GB/s of q,k,v,ctx)", flush=True)
print(f"swin_l forward cores per step (blocks weighted): simt {tot['simt']:.2f} ms, tcgen05 {tot['tcgen05']:.2f} ms")
K.WINATTN_TC = True
