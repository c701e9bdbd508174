"""Numpy prototype of the planned MSDA backward: queries sorted by reference point once per step,
tiles of T sorted queries, the tile's (value row, weight, query) corner records sorted by row and reduced per row -
ONE global add per distinct row instead of one per record.  Checks that the result equals the plain scatter-add of
msda_bwd_kernel and prints what a CTA would have to hold: records, distinct rows, shared-memory bytes.

    python tools/msda_sorted_bwd_proto.py [T]
"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import gedepth_b200.models as M                      # noqa: E402
from gedepth_b200.presets import model_cfg           # noqa: E402
from gedepth_b200.synth import synth_state_dict      # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 64
torch.manual_seed(0)
rng = np.random.default_rng(0)
model = M.build_depther(model_cfg("v", "kitti", "swin_t", pretrained=None))
model.load_state_dict(synth_state_dict(model.state_dict(), 0))
neck = model.neck
h, w = 44, 140                                        # a quarter-size query grid keeps the check fast
shapes = [(22, 70), (11, 35), (6, 18), (3, 9)]
starts = np.cumsum([0] + [a * b for a, b in shapes])[:-1]
S = sum(a * b for a, b in shapes)
qpe = neck.conv_positional_encoding.tokens(h, w, "cpu")
ref = torch.sigmoid(torch.nn.functional.linear(qpe, neck.reference_points.weight, neck.reference_points.bias))[0].detach().numpy()
Q, HD = ref.shape[0], 16                              # 16 of the 64 head channels are enough for the equality check
bias = neck.multi_att.sampling_offsets.bias.detach().view(8, 4, 8, 2).numpy()[0]          # head 0
off = bias[None] + 0.3 * rng.standard_normal((Q, 4, 8, 2))
logit = rng.standard_normal((Q, 32))
aw = np.exp(logit - logit.max(1, keepdims=True)); aw /= aw.sum(1, keepdims=True)
go = rng.standard_normal((Q, HD)).astype(np.float64)


def records(qidx):
    """(row, weight, query) of every valid bilinear corner of the queries in qidx."""
    rows, wts, qs = [], [], []
    for l, (H, W) in enumerate(shapes):
        for p in range(8):
            x = (ref[qidx, 0] + off[qidx, l, p, 0] / W) * W - 0.5
            y = (ref[qidx, 1] + off[qidx, l, p, 1] / H) * H - 0.5
            x0, y0 = np.floor(x).astype(int), np.floor(y).astype(int)
            lx, ly = x - x0, y - y0
            for dy, wy in ((0, 1 - ly), (1, ly)):
                for dx, wx in ((0, 1 - lx), (1, lx)):
                    xx, yy = x0 + dx, y0 + dy
                    ok = (xx >= 0) & (xx < W) & (yy >= 0) & (yy < H)
                    rows.append((starts[l] + yy * W + xx)[ok])
                    wts.append((aw[qidx, l * 8 + p] * wy * wx)[ok])
                    qs.append(qidx[ok])
    return np.concatenate(rows), np.concatenate(wts), np.concatenate(qs)


# (a) the kernel of round 1: scatter-add of every record
r, wt, q = records(np.arange(Q))
g_ref = np.zeros((S, HD))
np.add.at(g_ref, r, wt[:, None] * go[q])

# (b) sorted tiles, per-row reduction
key = (np.floor(ref[:, 1] * 64).astype(np.int64) << 16) | np.floor(ref[:, 0] * 64 * 64).astype(np.int64)
order = np.argsort(key, kind="stable")
g_new = np.zeros((S, HD))
n_rec, n_rows, flushes = [], [], 0
for t0 in range(0, Q, T):
    r, wt, q = records(order[t0:t0 + T])
    o = np.argsort(r, kind="stable")
    r, wt, q = r[o], wt[o], q[o]
    uniq, first = np.unique(r, return_index=True)
    sums = np.add.reduceat(wt[:, None] * go[q], first, axis=0)        # segmented reduction: registers in the kernel
    g_new[uniq] += sums                                               # one global red per distinct row
    n_rec.append(len(r)); n_rows.append(len(uniq)); flushes += len(uniq)
assert np.allclose(g_new, g_ref, rtol=1e-12, atol=1e-12)
n_rec, n_rows = np.array(n_rec), np.array(n_rows)
print(f"T={T}: records/tile mean {n_rec.mean():.0f} max {n_rec.max()}, distinct rows/tile mean {n_rows.mean():.0f} max {n_rows.max()}")
print(f"global adds: {n_rec.sum()} (scatter) -> {flushes} (sorted tiles): {n_rec.sum() / flushes:.1f}x fewer")
print(f"shared memory per tile at 64 channels: rows {n_rows.max() * 256 / 1024:.0f} KB max, records {n_rec.max() * 12 / 1024:.0f} KB, g_out {T * 256 / 1024:.0f} KB")
print("equal to the scatter-add result: OK")

# (c) the exact logic of tools/experimental/msda_bwd_sorted.cu: T = 32, per-level 12 x 12 windows anchored at the mean
# floor coordinate of the tile's points, window cells reduced and flushed once, everything else added directly
WIN, TT = 12, 32
g_k = np.zeros((S, HD))
n_in = n_out = n_flush = 0
for t0 in range(0, Q, TT):
    idx = order[t0:t0 + TT]
    for l, (H, W) in enumerate(shapes):
        x = (ref[idx, None, 0] + off[idx, l, :, 0] / W) * W - 0.5            # (T, 8)
        y = (ref[idx, None, 1] + off[idx, l, :, 1] / H) * H - 0.5
        x0, y0 = np.floor(x).astype(int), np.floor(y).astype(int)
        lx, ly = x - x0, y - y0
        near = (x0 >= -1) & (x0 < W) & (y0 >= -1) & (y0 < H)
        n = max(int(near.sum()), 1)
        wx0 = int(np.floor(x0[near].sum() / n + 0.5)) - WIN // 2 + 1
        wy0 = int(np.floor(y0[near].sum() / n + 0.5)) - WIN // 2 + 1
        cells = np.zeros((WIN, WIN, HD))
        touched = np.zeros((WIN, WIN), bool)
        a = aw[idx, l * 8:(l + 1) * 8]
        for dy in (0, 1):
            for dx in (0, 1):
                xx, yy = x0 + dx, y0 + dy
                wgt = a * (lx if dx else 1 - lx) * (ly if dy else 1 - ly)
                valid = (xx >= 0) & (xx < W) & (yy >= 0) & (yy < H)
                cx, cy = xx - wx0, yy - wy0
                inside = valid & (cx >= 0) & (cx < WIN) & (cy >= 0) & (cy < WIN)
                qq = np.broadcast_to(idx[:, None], xx.shape)
                np.add.at(cells, (cy[inside], cx[inside]), wgt[inside][:, None] * go[qq[inside]])
                touched[cy[inside], cx[inside]] = True
                outside = valid & ~inside
                np.add.at(g_k, starts[l] + yy[outside] * W + xx[outside], wgt[outside][:, None] * go[qq[outside]])
                n_in += int(inside.sum()); n_out += int(outside.sum())
        cy, cx = np.nonzero(touched)
        g_k[starts[l] + (wy0 + cy) * W + (wx0 + cx)] += cells[cy, cx]
        n_flush += len(cy)
assert np.allclose(g_k, g_ref, rtol=1e-12, atol=1e-12)
print(f"draft-kernel logic (T=32, 12x12 windows, mean anchor): equal; {100 * n_out / (n_in + n_out):.2f} % of the records "
      f"outside their window, global adds {n_in + n_out} -> {n_flush + n_out} ({(n_in + n_out) / (n_flush + n_out):.1f}x fewer)")
