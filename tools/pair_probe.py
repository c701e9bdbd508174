"""First contact with the CTA-pair GEMM kernel: one launch, compared with torch; run under `timeout`."""
import sys, torch
sys.path.insert(0, '.')
from gedepth_b200 import kernels as K
DEV = 'cuda:0'
torch.manual_seed(0)
for passes in (1, 3):
    K.set_gemm_precision(passes)
    for (M, N, Kd) in [(20000, 64, 64), (40000, 256, 128), (40000, 512, 512)]:
        a, w = torch.randn(M, Kd, device=DEV), torch.randn(N, Kd, device=DEV) / Kd ** .5
        K.set_gemm_pair(1)
        out = K.gemm(a, w)
        torch.cuda.synchronize()
        ref = (a.double() @ w.double().t()).float()
        err = float((out - ref).abs().max()) / float(ref.abs().max())
        print(f"pair passes={passes} {M}x{N}x{Kd}: max rel err {err:.3e}", flush=True)
        wt = w.t().contiguous()
        out2 = K.gemm_bt(a, wt)
        torch.cuda.synchronize()
        err2 = float((out2 - ref).abs().max()) / float(ref.abs().max())
        print(f"pair b_mn passes={passes} {M}x{N}x{Kd}: max rel err {err2:.3e}", flush=True)
print("probe ok")
