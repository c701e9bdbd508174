"""torch.profiler view of one training step (BASELINE config 2): top CUDA kernels and CPU-side time."""
import sys, os, torch
sys.path.insert(0, '.')
import gedepth_b200.models as M
from gedepth_b200 import kernels
from gedepth_b200.presets import model_cfg
from gedepth_b200.synth import synth_batch, synth_state_dict
from gedepth_b200.train import Trainer
from torch.profiler import profile, ProfilerActivity
dev = 'cuda:0'
torch.backends.cuda.matmul.allow_tf32 = True; torch.backends.cudnn.allow_tf32 = True
variant = sys.argv[1] if len(sys.argv) > 1 else 'v'
B, H, W = 8, 352, 1120
model = M.build_depther(model_cfg(variant, 'kitti', 'swin_t', pretrained=None))
model.load_state_dict(synth_state_dict(model.state_dict(), 0)); model.to(dev).train()
tr = Trainer(model)
b = synth_batch(B, H, W, seed=1, adaptive=variant == 'a')
d = {k: torch.from_numpy(v).to(dev) for k, v in b.items()}
metas = [{}] * B
def step():
    tr.step(dict(img=d['img'], img_metas=metas, depth_gt=d['depth_gt'], **({'pe_k_gt': d['pe_k_gt']} if variant == 'a' else {})))
for _ in range(3): step()
torch.cuda.synchronize()
import time
t0 = time.time(); step(); t_cpu = time.time() - t0; torch.cuda.synchronize(); t_all = time.time() - t0
print(f'cpu launch time {t_cpu*1e3:.1f} ms, step wall {t_all*1e3:.1f} ms')
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=45, max_name_column_width=70))
