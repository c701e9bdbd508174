"""Forward GEMM / 3x3 conv arithmetic modes (3 = 3xTF32, 2 = bf16 hi/lo split, 1 = TF32) at Swin-L / config-3 shapes."""
import sys, torch
sys.path.insert(0, '.')
from gedepth_b200 import kernels as K
DEV = 'cuda:0'
torch.manual_seed(0)
B = 16
T0 = B * 88 * 280


def t_ms(fn, reps=5):
    fn(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / reps


for (M, N, Kd) in [(T0, 576, 192), (T0, 768, 192), (T0, 192, 768), (T0 // 4, 1536, 384), (T0 // 16, 2304, 768), (T0 // 16, 3072, 768),
                   (T0 // 16, 768, 3072), (B * 98560 // 2, 512, 512), (T0, 64, 192), (T0, 96, 192), (T0 // 4, 128, 384)]:
    a, w = torch.randn(M, Kd, device=DEV), torch.randn(N, Kd, device=DEV) / Kd ** .5
    out = torch.empty(M, N, device=DEV)
    row = []
    for passes in (3, 2, 1):
        K.set_gemm_precision(passes)
        ms = t_ms(lambda: K.gemm(a, w, out=out))
        row.append(f"p{passes}: {ms:.3f} ms {2 * M * N * Kd / ms / 1e9:.0f} TF")
        if passes == 3:
            prev = K.set_gemm_a_tmem(0)
            ms = t_ms(lambda: K.gemm(a, w, out=out))
            K.set_gemm_a_tmem(prev)
            row.append(f"p3 (A in smem): {ms:.3f} ms")
    print("gemm", M, N, Kd, " | ".join(row), flush=True)
    del a, w, out
for (Bc, H, W, Ci, Co) in [(B, 176, 560, 256, 64), (B, 22, 70, 2304, 768), (B, 88, 280, 576, 192), (B, 44, 140, 1152, 384)]:
    x = torch.randn(Bc, H, W, Ci, device=DEV); wk = torch.randn(Co, 3, 3, Ci, device=DEV) / (9 * Ci) ** .5
    xp = K.prep_conv_input(x, None, H, W)
    row = []
    for passes in (3, 2, 1):
        K.set_gemm_precision(passes)
        ms = t_ms(lambda: K.conv3x3_padded(xp, wk, None, 'leaky_relu', 0.01))
        row.append(f"p{passes}: {ms:.3f} ms {2 * Bc * H * W * Ci * Co * 9 / ms / 1e9:.0f} TF")
        if passes == 3:
            prev = K.set_gemm_a_tmem(0)
            ms = t_ms(lambda: K.conv3x3_padded(xp, wk, None, 'leaky_relu', 0.01))
            K.set_gemm_a_tmem(prev)
            row.append(f"p3 (A in smem): {ms:.3f} ms")
    print("conv3x3", Bc, H, W, Ci, Co, " | ".join(row), flush=True)
K.set_gemm_precision(3)
