"""A/B: CTA-pair (cta_group::2) vs single-CTA tcgen05 GEMM / conv at BASELINE-config-2 shapes."""
import sys, torch
sys.path.insert(0, '.')
from gedepth_b200 import kernels as K
DEV = 'cuda:0'
torch.manual_seed(0)
B = 8
T0 = B * 88 * 280


def t_ms(fn, reps=5):
    fn(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / reps


shapes = [(T0, 288, 96), (T0, 384, 96), (T0, 96, 384), (T0, 96, 96), (T0 // 4, 576, 192), (T0 // 4, 768, 192), (T0 // 4, 192, 768),
          (T0 // 16, 1536, 384), (T0 // 16, 384, 1536), (B * 98560, 512, 512), (B * 98560, 256, 512), (B * 32725, 512, 512),
          (B * 98560, 512, 64), (B * 98560, 64, 64)]
print("GEMM  M N K | passes: single ms TF | pair ms TF")
for (M, N, Kd) in shapes:
    a, w = torch.randn(M, Kd, device=DEV), torch.randn(N, Kd, device=DEV) / Kd ** .5
    out = torch.empty(M, N, device=DEV)
    for passes in (3, 1):
        K.set_gemm_precision(passes)
        row = []
        for pair in (0, 1):
            K.set_gemm_pair(pair)
            ms = t_ms(lambda: K.gemm(a, w, out=out))
            row.append(f"pair={pair}: {ms:.3f} ms {2 * M * N * Kd / ms / 1e9:.0f} TF")
        print(M, N, Kd, f"passes={passes}", " | ".join(row), flush=True)
    del a, w, out
print("conv3x3 fwd  B H W Cin Cout")
for (Bc, H, W, Ci, Co) in [(B, 176, 560, 256, 64), (B, 22, 70, 2304, 768), (B, 88, 280, 576, 192), (B, 44, 140, 1152, 384),
                           (B, 176, 560, 576, 64), (B, 176, 560, 64, 64), (B, 88, 280, 192, 192), (B, 44, 140, 384, 384)]:
    x = torch.randn(Bc, H, W, Ci, device=DEV); wk = torch.randn(Co, 3, 3, Ci, device=DEV) / (9 * Ci) ** .5
    xp = K.prep_conv_input(x, None, H, W)
    for passes in (3, 1):
        K.set_gemm_precision(passes)
        row = []
        for pair in (0, 1):
            K.set_gemm_pair(pair)
            ms = t_ms(lambda: K.conv3x3_padded(xp, wk, None, 'leaky_relu', 0.01))
            row.append(f"pair={pair}: {ms:.3f} ms {2 * Bc * H * W * Ci * Co * 9 / ms / 1e9:.0f} TF")
        print(Bc, H, W, Ci, Co, f"passes={passes}", " | ".join(row), flush=True)
    del x, wk, xp
K.set_gemm_pair(1)
print('done')
