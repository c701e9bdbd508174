import sys, os, torch, time
sys.path.insert(0, '.')
from gedepth_b200 import kernels as K
DEV='cuda:0'
torch.manual_seed(0)
def rnd(*s): return torch.randn(*s, device=DEV)
shapes = [(88, 280), (44, 140), (22, 70), (11, 35)]
S = sum(h * w for h, w in shapes); Bm = 4; Q = 176 * 560
yy, xx = torch.meshgrid(torch.linspace(0, 1, 176, device=DEV), torch.linspace(0, 1, 560, device=DEV), indexing='ij')
ref = torch.stack((0.1 + 0.8 * xx + 0.02 * torch.sin(6 * yy), 0.1 + 0.8 * yy + 0.02 * torch.cos(5 * xx)), -1).reshape(1, Q, 2).expand(Bm, -1, -1).contiguous()
v = rnd(Bm, S, 512).requires_grad_(True)
for name, offs in (('smooth offsets (bias only)', (rnd(1, 1, 512) * 3).expand(Bm, Q, 512).contiguous()),
                   ('bias + 0.5px per-query noise', (rnd(1, 1, 512) * 3 + rnd(Bm, Q, 512) * 0.5)),
                   ('iid 3px offsets', rnd(Bm, Q, 512) * 3)):
    off = offs.clone().requires_grad_(True); lg = rnd(Bm, Q, 256).requires_grad_(True)
    out = K.msda_sample(v, shapes, ref, off, lg, 8, 8)
    go = rnd(*out.shape)
    for _ in range(2): out.backward(go, retain_graph=True)
    torch.cuda.synchronize(); s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(3): out.backward(go, retain_graph=True)
    e.record(); torch.cuda.synchronize()
    print(os.environ.get('GEDEPTH_MSDA_BWD', '1'), name, 'bwd ms (B=4):', s.elapsed_time(e) / 3)
