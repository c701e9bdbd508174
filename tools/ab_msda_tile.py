"""A/B of the MSDA kernels on the tensors the bench model really produces (BASELINE config 2: Swin-T + Vanilla, B = 8,
352 x 1120, synthetic weights): one forward of the model with a tap on kernels.msda_sample, then the round-1 kernels
(csrc/msda.cu) against the sorted-tile kernels (csrc/msda_tile.cu) on exactly those (value, ref, offsets, logits), L2
flushed between repetitions, plus the agreement of the two implementations."""
import sys, torch
sys.path.insert(0, '.')
import gedepth_b200.models as M
from gedepth_b200 import kernels as K
from gedepth_b200.presets import model_cfg
from gedepth_b200.synth import synth_batch, synth_state_dict

DEV = 'cuda:0'
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
ONLY_TILE = len(sys.argv) > 2 and sys.argv[2] == "tile"      # for ncu: tile kernels only
ONLY = sys.argv[3] if len(sys.argv) > 3 else ""               # "self" / "cross"
cfg = model_cfg('v', 'kitti', 'swin_t', pretrained=None)
model = M.build_depther(cfg)
model.load_state_dict(synth_state_dict(model.state_dict(), 0))
model.to(DEV).train()
b = synth_batch(B, 352, 1120, seed=1234)
img, gt = torch.from_numpy(b['img']).to(DEV), torch.from_numpy(b['depth_gt']).to(DEV)
taps = []
orig = K.msda_sample


def tap(v, shapes, ref, off, logit, nH, P):
    taps.append((v.detach().clone(), list(shapes), ref.detach().clone(), off.detach().clone(), logit.detach().clone()))
    return orig(v, shapes, ref, off, logit, nH, P)


K.msda_sample = tap
with torch.no_grad():
    model.train_step(dict(img=img, img_metas=[{}] * B, depth_gt=gt), None)
K.msda_sample = orig
del model
torch.cuda.empty_cache()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)


def t_ms(fn, reps=5):
    fn(); torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        tot += s.elapsed_time(e)
    return tot / reps


for name, (v, shapes, ref, off, lg) in zip(("self", "cross"), taps):
    if ONLY and name != ONLY:
        continue
    Q = off.shape[1]
    go = torch.randn(B, Q, 512, device=DEV)
    res = {}
    for label, fwd_impl, bwd_impl in (("round1", "round1", "round1"), ("tile fp32", "tile", "tile"), ("tc", "tc", "tc")):
        if ONLY_TILE and label != "tc":
            continue
        K.MSDA_FWD, K.MSDA_BWD = fwd_impl, bwd_impl
        vv, oo, ll = v.clone().requires_grad_(True), off.clone().requires_grad_(True), lg.clone().requires_grad_(True)
        rr = ref.clone().requires_grad_(ref.shape[0] == 1 and name == "cross")
        with torch.no_grad():
            f = t_ms(lambda: K.msda_sample(vv, shapes, rr, oo, ll, 8, 8))
        out = K.msda_sample(vv, shapes, rr, oo, ll, 8, 8)
        ins = (vv, oo, ll) + ((rr,) if rr.requires_grad else ())
        bw = t_ms(lambda: torch.autograd.grad(out, ins, go, retain_graph=True))
        grads = torch.autograd.grad(out, ins, go)
        res[label] = (out.detach(), [g_.detach() for g_ in grads])
        print(f"{name} B={B} Q={Q} {label:14s}: fwd {f:7.3f} ms   bwd {bw:7.3f} ms (incl. zero-fill of g_value)", flush=True)
    if ONLY_TILE:
        continue
    for other in ("tile fp32", "tc"):
      a, t = res["round1"], res[other]
      print(f"   {other} vs round-1: out max|d| {float((a[0] - t[0]).abs().max()):.3e} (max {float(a[0].abs().max()):.3e}); " +
            "; ".join(f"g{i} {float((x - y).abs().max()):.3e}/{float(x.abs().max()):.3e}" for i, (x, y) in enumerate(zip(a[1], t[1]))))
print("done")
