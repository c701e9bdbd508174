"""A/B timing of the deformable-attention kernels at BASELINE-config-2 sizes (cross-attention: B=8, Q=176*560,
S=32725; self-attention: Q=S).  Reference points / offsets distributed as in the bench model (noisy learned
reference points, ~1.5 px offsets)."""
import sys, torch
sys.path.insert(0, '.')
from gedepth_b200 import kernels as K
DEV = 'cuda:0'
torch.manual_seed(0)
shapes = [(88, 280), (44, 140), (22, 70), (11, 35)]
S = sum(h * w for h, w in shapes)
Bm = 8


def rnd(*s): return torch.randn(*s, device=DEV)


def t_ms(fn, reps=3):
    fn(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / reps


for name, Q in (("cross", 176 * 560), ("self", S)):
    if name == "cross":
        ref = torch.sigmoid(rnd(1, Q, 2) * 0.7).expand(Bm, -1, -1).contiguous()
    else:
        refs = []
        for h, w in shapes:
            ry, rx = torch.meshgrid(torch.linspace(0.5, h - 0.5, h, device=DEV), torch.linspace(0.5, w - 0.5, w, device=DEV), indexing='ij')
            refs.append(torch.stack((rx.reshape(-1) / w, ry.reshape(-1) / h), -1))
        ref = torch.cat(refs, 0)[None].contiguous()
    v = rnd(Bm, S, 512).requires_grad_(True)
    off = (rnd(1, 1, 512) * 1.5 + rnd(Bm, Q, 512) * 0.5).requires_grad_(True)
    lg = rnd(Bm, Q, 256).requires_grad_(True)
    for variant in (0, 1, 2, 3):
        K.set_msda_variant(variant)
        with torch.no_grad():
            f = t_ms(lambda: K.msda_sample(v, shapes, ref, off, lg, 8, 8))
        out = K.msda_sample(v, shapes, ref, off, lg, 8, 8)
        go = rnd(*out.shape)
        b = t_ms(lambda: torch.autograd.grad(out, (v, off, lg), go, retain_graph=True))
        print(f"{name} Q={Q} variant={variant}: fwd {f:.2f} ms, bwd {b:.2f} ms (incl. zero-fill of g_value)")
        del out, go
    del v, off, lg
K.set_msda_variant(0)
print("done")
