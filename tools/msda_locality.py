"""CPU analysis for the next MSDA backward: how many corner hits land on the same value row inside a tile of T queries,
in image order vs with the queries sorted (Morton order) by their reference point - which is static per step and shared
by batch and heads.  Uses the bench model's synthetic weights.  Result (round 1): T=64: 2-20 hits/row in image order,
22-70 sorted; T=512: 5-94 vs 72-322."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
import gedepth_b200.models as M
from gedepth_b200.presets import model_cfg
from gedepth_b200.synth import synth_state_dict
torch.manual_seed(0)
model = M.build_depther(model_cfg('v', 'kitti', 'swin_t', pretrained=None))
model.load_state_dict(synth_state_dict(model.state_dict(), 0))
neck = model.neck
h, w = 176, 560
qpe = neck.conv_positional_encoding.tokens(h, w, 'cpu')            # (1, Q, 512)
ref = torch.sigmoid(torch.nn.functional.linear(qpe, neck.reference_points.weight, neck.reference_points.bias))[0].detach().numpy()
print('ref range x', ref[:,0].min(), ref[:,0].max(), 'y', ref[:,1].min(), ref[:,1].max(), 'std', ref.std(0))
bias = neck.multi_att.sampling_offsets.bias.detach().view(8, 4, 8, 2).numpy()
shapes = [(88, 280), (44, 140), (22, 70), (11, 35)]
rng = np.random.default_rng(0)
Q = ref.shape[0]
def morton(ix, iy):
    def part(v):
        v = v.astype(np.uint64); v = (v | (v << 16)) & 0x0000FFFF0000FFFF; v = (v | (v << 8)) & 0x00FF00FF00FF00FF
        v = (v | (v << 4)) & 0x0F0F0F0F0F0F0F0F; v = (v | (v << 2)) & 0x3333333333333333; v = (v | (v << 1)) & 0x5555555555555555
        return v
    return part(ix) | (part(iy) << np.uint64(1))
order_img = np.arange(Q)
order_sorted = np.argsort(morton((ref[:,0]*4096).astype(np.int64), (ref[:,1]*4096).astype(np.int64)))
head = 0
for T in (64, 512, 4096):
    for name, order in (('image order', order_img), ('sorted by ref', order_sorted)):
        res = []
        for l, (H, W) in enumerate(shapes):
            uniq, tot = 0, 0
            for t0 in range(0, Q - T, max(T, Q // 40)):
                idx = order[t0:t0 + T]
                rows = []
                for p in range(8):
                    off = bias[head, l, p] + 0.3 * rng.standard_normal((T, 2))      # per-query variation of the offsets
                    x = (ref[idx, 0] + off[:, 0] / W) * W - 0.5
                    y = (ref[idx, 1] + off[:, 1] / H) * H - 0.5
                    x0, y0 = np.floor(x).astype(int), np.floor(y).astype(int)
                    for dy in (0, 1):
                        for dx in (0, 1):
                            xx, yy = x0 + dx, y0 + dy
                            ok = (xx >= 0) & (xx < W) & (yy >= 0) & (yy < H)
                            rows.append((yy * W + xx)[ok])
                rows = np.concatenate(rows)
                uniq += len(np.unique(rows)); tot += len(rows)
            res.append(tot / max(uniq, 1))
        print(f'T={T:5d} {name:14s} hits per distinct row, levels 0-3:', ' '.join(f'{r:7.1f}' for r in res))
