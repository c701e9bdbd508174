"""CPU trace of every GEMM-shaped op of one forward pass (library statements injected, batch 1) -> table of
(M, N, K, taps), flops and share, grouped by shape: where the tensor-core time of a configuration can go.
Rows scale linearly with the batch (M = rows per frame x batch).

    python tools/gemm_shapes.py [v|a] [swin_t|swin_l] [H W]
"""
import sys
from collections import defaultdict

import torch

sys.path.insert(0, ".")
import gedepth_b200.models as M                      # noqa: E402
from gedepth_b200 import ops                         # noqa: E402
from tests import ops_lib as L                       # noqa: E402  (library statements: test infrastructure)
from gedepth_b200.presets import model_cfg           # noqa: E402
from gedepth_b200.synth import synth_batch, synth_state_dict   # noqa: E402

variant = sys.argv[1] if len(sys.argv) > 1 else "v"
backbone = sys.argv[2] if len(sys.argv) > 2 else "swin_t"
H, W = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (352, 1120)
rec = defaultdict(lambda: [0, 0.0])


def note(kind, Mr, N, K, taps=1):
    r = rec[(kind, Mr, N, K, taps)]
    r[0] += 1
    r[1] += 2.0 * Mr * N * K * taps


_lin, _conv, _cba = L.linear, L.conv2d, L.conv_bn_act


def linear(x, w, b=None, act=None, residual=None, row_scale=None):
    note("linear", x.numel() // x.shape[-1], w.shape[0], w.shape[1])
    return _lin(x, w, b, act, residual, row_scale)


def conv2d(x, w, b=None, stride=1, padding=0, act=None, slope=0.01):
    y = _conv(x, w, b, stride, padding, act, slope)
    note("conv%dx%d" % (w.shape[2], w.shape[3]), y.shape[0] * y.shape[2] * y.shape[3], w.shape[0], w.shape[1], w.shape[2] * w.shape[3])
    return y


def conv_bn_act(x, w, b, bn, stride=1, padding=0, act=None):
    y = _cba(x, w, b, bn, stride, padding, act)
    note("conv%dx%d" % (w.shape[2], w.shape[3]), y.shape[0] * y.shape[2] * y.shape[3], w.shape[0], w.shape[1], w.shape[2] * w.shape[3])
    return y


L.linear, L.conv2d, L.conv_bn_act = linear, conv2d, conv_bn_act
L.install(ops)
model = M.build_depther(model_cfg(variant, "kitti", backbone, pretrained=None))
model.load_state_dict(synth_state_dict(model.state_dict(), 0))
model.eval()
b = synth_batch(1, H, W, seed=1, adaptive=variant == "a")
with torch.no_grad():
    model.encode_decode(torch.from_numpy(b["img"]), [dict(ori_shape=(H, W, 3), flip=False)], rescale=True)
tot = sum(v[1] for v in rec.values())
print(f"# {backbone} {variant} {H}x{W}, batch 1: {tot / 1e9:.1f} GFLOP forward in {sum(v[0] for v in rec.values())} GEMM-shaped launches")
print("kind,rows_per_frame,N,K,taps,launches,gflop,share_pct")
for k, v in sorted(rec.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[0]},{k[1]},{k[2]},{k[3]},{k[4]},{v[0]},{v[1] / 1e9:.2f},{100 * v[1] / tot:.1f}")
