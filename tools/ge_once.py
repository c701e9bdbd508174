"""One launch of each ground-embedding kernel at a large shape (for ncu).  usage: python tools/ge_once.py [B H W]"""
import sys, torch
sys.path.insert(0, '.')
from gedepth_b200 import kernels as K
from gedepth_b200.kernels import _call, _p, _stream
B, H, W = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (16, 1024, 2048)
dev = 'cuda:0'
K.load()
img = torch.randn(B, 5, H, W, device=dev); img[:, 4] = img[:, 4].abs() * 30 + 2
yh = torch.rand(B, 1, H // 2, W // 2, device=dev); lh = torch.randn(B, 11, H // 2, W // 2, device=dev)
y, pm = torch.empty(B, 1, H, W, device=dev), torch.empty(B, 1, H, W, device=dev)
lf = torch.empty(B, 11, H, W, device=dev)
gy, gpm, glf = torch.randn_like(y), torch.randn_like(pm), torch.randn_like(lf)
g_yh, g_lh = torch.empty_like(yh), torch.empty_like(lh)
h2, w2, bs = H // 2, W // 2, 5 * H * W
for _ in range(2):
    _call("ged_ge_vanilla_fwd", _p(img[:, 3]), bs, _p(yh), _p(y), _p(pm), B, H, W, h2, w2, _stream())
    _call("ged_ge_adaptive_fwd", _p(img[:, 4]), bs, _p(yh), _p(lh), None, 1.65, 200.0, _p(y), _p(pm), None, B, H, W, h2, w2, _stream())
    _call("ged_ge_adaptive_fwd", _p(img[:, 4]), bs, _p(yh), _p(lh), None, 1.65, 200.0, _p(y), _p(pm), _p(lf), B, H, W, h2, w2, _stream())
    _call("ged_ge_vanilla_bwd", _p(img[:, 3]), bs, _p(gy), _p(gpm), _p(g_yh), B, H, W, h2, w2, _stream())
    _call("ged_ge_adaptive_bwd", _p(img[:, 4]), bs, _p(yh), _p(lh), None, 1.65, 200.0, _p(gy), _p(gpm), _p(glf), _p(g_yh), _p(g_lh), B, H, W, h2, w2, _stream())
torch.cuda.synchronize()
