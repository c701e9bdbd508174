import sys, torch
sys.path.insert(0, '.')
from gedepth_b200 import kernels as K, ops_lib as L
from gedepth_b200.synth import synth_batch
DEV = 'cuda:0'
for (B, H, W) in [(1, 16, 32), (2, 16, 32), (1, 64, 160), (1, 16, 300), (2, 64, 160)]:
    b = synth_batch(B, H, W, seed=0, adaptive=True)
    g = torch.Generator().manual_seed(0)
    h2, w2 = (H + 1) // 2, (W + 1) // 2
    img = torch.from_numpy(b['img']).to(DEV)
    yh0 = torch.rand(B, 1, h2, w2, generator=g); lh0 = torch.randn(B, 11, h2, w2, generator=g) * 2
    for mode in ('gy', 'gpm', 'glf', 'all'):
        a1 = [t.to(DEV).requires_grad_(True) for t in (yh0, lh0)]
        a2 = [t.to(DEV).requires_grad_(True) for t in (yh0, lh0)]
        y, pm, lf = K.ge_adaptive(img, a1[0], a1[1], 1.65, 200.0)
        y2, pm2, lf2 = L.ge_adaptive(img, a2[0], a2[1], 1.65, 200.0)
        gg = torch.Generator().manual_seed(1)
        gy, gpm, glf = torch.randn(y.shape, generator=gg).to(DEV), torch.randn(pm.shape, generator=gg).to(DEV) * 0.1, torch.randn(lf.shape, generator=gg).to(DEV) * 0.01
        l1 = {'gy': (y * gy).sum(), 'gpm': (pm * gpm).sum(), 'glf': (lf * glf).sum(), 'all': (y * gy + pm * gpm + lf * glf).sum()}[mode]
        l2 = {'gy': (y2 * gy).sum(), 'gpm': (pm2 * gpm).sum(), 'glf': (lf2 * glf).sum(), 'all': (y2 * gy + pm2 * gpm + lf2 * glf).sum()}[mode]
        l1.backward(); l2.backward()
        out = []
        for n, p, q in (('g_yh', a1[0], a2[0]), ('g_lh', a1[1], a2[1])):
            pg = p.grad if p.grad is not None else torch.zeros_like(p)
            qg = q.grad if q.grad is not None else torch.zeros_like(q)
            d = (pg - qg).abs()
            nan = int(torch.isnan(pg).sum()), int(torch.isnan(qg).sum())
            d = torch.nan_to_num(d, 0.0)
            idx = int(d.argmax())
            out.append(f"{n}: maxdiff {float(d.max()):.3e} @ {idx} refmax {float(torch.nan_to_num(qg).abs().max()):.3e} nan {nan}")
        print((B, H, W), mode, ' | '.join(out), 'pm nan', int(torch.isnan(pm).sum()), int(torch.isnan(pm2).sum()))
