"""Narrow (N = 64) 3xTF32 conv tiles: A operand from shared memory (default) vs from tensor memory, interleaved repetitions.
(The same script, with temporary switches, measured 8 splitter warps, two alternating splitter groups, 3 / 2 ring stages,
a probe that skipped a third of the MMAs and CTA pairing: 3.87 / 3.87 / 3.89 / 4.90 / 3.86 / 4.19 ms against 3.87 for the
256 -> 64 conv at 16 x 176 x 560.)"""
import sys, torch
sys.path.insert(0, '.')
from gedepth_b200 import kernels as K
DEV = 'cuda:0'
torch.manual_seed(0)
K.set_gemm_precision(3)


def t_ms(fn, reps=10):
    fn(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / reps


for (Bc, H, W, Ci, Co) in [(16, 176, 560, 256, 64), (16, 176, 560, 576, 64), (16, 176, 560, 64, 64)]:
    x = torch.randn(Bc, H, W, Ci, device=DEV); wk = torch.randn(Co, 3, 3, Ci, device=DEV) / (9 * Ci) ** .5
    xp = K.prep_conv_input(x, None, H, W)
    res = {0: [], 1: []}
    for rep in range(3):
        for mode in (0, 1):
            K.set_gemm_a_tmem(mode)
            res[mode].append(t_ms(lambda: K.conv3x3_padded(xp, wk, None, 'leaky_relu', 0.01)))
    K.set_gemm_a_tmem(0)
    print(Bc, H, W, Ci, Co, "A in shared memory", [round(v, 3) for v in res[0]], "| A in tensor memory", [round(v, 3) for v in res[1]], flush=True)
