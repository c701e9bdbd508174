"""The five ground-embedding kernels alone against the measured HBM copy peak (bench.py's `ground_embed` block) at the
workload shape and at the large end of the sweep.  usage: python tools/ge_probe.py [B H W]"""
import json, sys, torch
sys.path.insert(0, '.')
import bench
from gedepth_b200 import kernels
c = bench.Ctx()
c.torch, c.kernels, c.dev = torch, kernels, torch.device('cuda', 0)
kernels.load()
B, H, W = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (16, 352, 1120)
r = bench.ground_embed_probe(c, bench.peaks(), B, H, W)
for where in ("at_workload", "at_sweep_max"):
    print(where, r[where]["shape"])
    for k, v in r[where]["kernels"].items():
        print(f"  {k:28s} {v['bytes_per_px']:5.1f} B/px  {v['us']:9.1f} us  {v['achieved']:7.0f} GB/s  frac {v['frac']:.3f}")
