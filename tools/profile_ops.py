"""Which torch (library) ops are still inside the step, by input shape and call site."""
import sys, torch
sys.path.insert(0, '.')
import gedepth_b200.models as M
from gedepth_b200.presets import model_cfg
from gedepth_b200.synth import synth_batch, synth_state_dict
from gedepth_b200.train import Trainer
from torch.profiler import profile, ProfilerActivity
dev = 'cuda:0'
torch.backends.cuda.matmul.allow_tf32 = True; torch.backends.cudnn.allow_tf32 = True
VAR, BB, B = (sys.argv[1], sys.argv[2], int(sys.argv[3])) if len(sys.argv) > 3 else ('v', 'swin_t', 8)
H, W = 352, 1120
model = M.build_depther(model_cfg(VAR, 'kitti', BB, pretrained=None))
model.load_state_dict(synth_state_dict(model.state_dict(), 0)); model.to(dev).train()
tr = Trainer(model)
b = synth_batch(B, H, W, seed=1, adaptive=VAR == 'a')
d = {k: torch.from_numpy(v).to(dev) for k, v in b.items()}
def step(): tr.step(dict(img=d['img'], img_metas=[{}] * B, depth_gt=d['depth_gt'], **({'pe_k_gt': d['pe_k_gt']} if VAR == 'a' else {})))
for _ in range(3): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True, with_stack=True) as prof:
    step(); torch.cuda.synchronize()
rows = []
for e in prof.key_averages(group_by_input_shape=True, group_by_stack_n=6):
    if e.key.startswith('aten::') and e.self_device_time_total > 150:
        stack = [s for s in e.stack if 'gedepth_b200' in s or 'bench' in s][:2]
        rows.append((e.self_device_time_total / 1e3, e.count, e.key, str(e.input_shapes)[:90], ' <- '.join(s.split('/')[-1][:60] for s in stack)))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print(f'library ops with device time: {tot:.1f} ms')
for r in rows[:60]: print(f'{r[0]:7.2f} ms x{r[1]:3d} {r[2]:32s} {r[3]:90s} {r[4]}')
