#!/usr/bin/env python
"""bench.py - frames/s (352x1120) fwd+bwd of the GEDepth path.

Headline workload (what `--gpus N` measures): BASELINE configs[2], the configuration the reference ships -
DepthFormer-Swin-L + GEDepth-Adaptive (learned slope), batch 16 per GPU, 352x1120 synthetic KITTI frames,
deterministic synthetic weights.  One "step" = forward + SiLog + 0.08 CE + backward (+ one NCCL all-reduce of the flat
gradient arena when N > 1) + clip + AdamW on one batch, replayed as one CUDA graph.  At N = 1 the same line also carries
BASELINE configs[1] (Swin-T + Vanilla, batch 8) and the per-GPU shard of configs[3] (Swin-L + Adaptive, DDAD 384x640,
batch 4) under `other_configs`, and `gpu_library_baseline`: the same step through the cuDNN / cuBLAS-TF32 /
grid_sample statement of every op (the reference's own design on this GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload config3|config2|config4]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Prints ONE JSON line (rank 0).  `value` = device-resident throughput; `e2e` = the same metric through the public
train-step call with HOST (pinned) buffers copied H2D every step and the loss read back; `roofline` = the dominant
kernel of the step (by summed device time, measured with CUDA events around every C-ABI launch of one extra step);
`ground_embed` = the HBM roofline of the kernels the metric names; `cpu_baseline` = the oracle port of the same step on
this box's host cores.  --impl reference times that CPU port alone (the reference itself cannot be installed here:
its setup.py imports mmcv, which is absent and there is no network).
"""
from __future__ import annotations

import argparse
import gc
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "config3": dict(variant="a", backbone="swin_l", dataset="kitti", batch=16, H=352, W=1120,
                    label="DepthFormer-Swin-L + GEDepth-Adaptive, batch 16/GPU, 352x1120 synthetic KITTI (BASELINE configs[2])"),
    "config2": dict(variant="v", backbone="swin_t", dataset="kitti", batch=8, H=352, W=1120,
                    label="DepthFormer-Swin-T + GEDepth-Vanilla, batch 8/GPU, 352x1120 synthetic KITTI (BASELINE configs[1])"),
    "config4": dict(variant="a", backbone="swin_l", dataset="ddad", batch=4, H=384, W=640,
                    label="DepthFormer-Swin-L + GEDepth-Adaptive, batch 4/GPU, DDAD 384x640 synthetic (per-GPU shard of BASELINE configs[3])"),
}
HEADLINE = "config3"
METRIC = "frames/sec (352x1120) fwd+bwd"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], bf16_burst=d["bf16_tflops"], bf16_sustained=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, bf16_burst=1590.0, bf16_sustained=1400.0, src="fallback")


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=reasons,
                    samples=len(sm))


# ---------------------------------------------------------------------------------------------------------------------
# CPU legs (the only places that execute oracle/)
# ---------------------------------------------------------------------------------------------------------------------
def cpu_step_times(torch, spec, n_frames=1, budget_s=200.0, steps=1, warmup=0):
    """Oracle port (plain PyTorch CPU restatement of the reference path) fwd+bwd on `n_frames` frames of `spec`."""
    import numpy as np
    import gedepth_b200.models as M
    from gedepth_b200.presets import SWIN_L, SWIN_T, model_cfg
    from gedepth_b200.synth import synth_batch, synth_state_dict
    from oracle import model as om
    adaptive, ddad = spec["variant"] == "a", spec["dataset"] == "ddad"
    tmpl = M.build_depther(model_cfg(spec["variant"], spec["dataset"], spec["backbone"], pretrained=None)).state_dict()
    sd = synth_state_dict(tmpl, 0)
    skip = ("running_mean", "running_var", "num_batches_tracked", "relative_position_index")
    sd = {k: v.requires_grad_(not k.endswith(skip)) for k, v in sd.items()}
    b = synth_batch(n_frames, spec["H"], spec["W"], seed=1234, adaptive=adaptive, depth_scale=250.0 if ddad else 200.0,
                    max_depth=200.0 if ddad else 80.0)
    img, gt = torch.from_numpy(b["img"]), torch.from_numpy(b["depth_gt"])
    kgt = torch.from_numpy(b["pe_k_gt"]) if adaptive else None
    sw = SWIN_L if spec["backbone"] == "swin_l" else SWIN_T
    cfg = om.PathConfig(embed_dims=sw["embed_dims"], depths=tuple(sw["depths"]), num_heads=tuple(sw["num_heads"]),
                        adaptive=adaptive, depth_scale=250.0 if ddad else 200.0, max_depth=200.0 if ddad else 80.0,
                        train_bn=True)
    height = torch.from_numpy(np.array(([1.56, 1.57, 1.53, 1.53] * n_frames)[:n_frames], dtype=np.float32)) if ddad else 1.65
    times = []
    t_begin = time.time()
    for i in range(warmup + steps):
        t0 = time.time()
        r = om.forward_train(sd, cfg, img, gt, kgt, height)
        r["loss"].backward()
        for v in sd.values():
            v.grad = None
        dt = time.time() - t0
        if i >= warmup:
            times.append(dt)
        if time.time() - t_begin > budget_s and len(times) >= 1:
            break
    return times


def ground_plane_numpy_rate(H, W):
    from oracle import ground as og
    coef = og.plane_coefficients(og.kitti_projection(), og.KITTI_CAM_HEIGHT)
    best = 1e9
    for _ in range(5):
        t0 = time.perf_counter()
        og.ground_plane(coef, H, W, 61, 23)
        best = min(best, time.perf_counter() - t0)
    return H * W / best / 1e6


def run_reference(args):
    """The reference arm: the reference's path on the host cores (oracle port; the reference package cannot be installed
    here).  One step = fwd + loss + bwd on ONE frame of the headline batch (a bounded sample: frames/s is per frame);
    the run is capped at ~4 minutes, so fewer than --steps steps may be timed - `steps` reports what was."""
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    spec = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    wu = max(1, min(args.warmup, 3))
    t0 = time.time()
    times = cpu_step_times(torch, spec, 1, budget_s=240.0, steps=max(1, args.steps), warmup=wu)
    ms = 1e3 * sum(times) / len(times)
    val = 1.0 / (ms / 1e3)
    sample = (f"1 frame of the batch per step ({spec['H']}x{spec['W']} fwd+bwd), {wu} warm-up + {len(times)} timed steps "
              f"(240 s cap; {args.steps} requested), torch-CPU {torch.get_num_threads()} threads, {time.time() - t0:.0f} s wall")
    line = dict(metric=METRIC, value=val, unit="frames/s", n_gpus=args.gpus, steps=len(times), warmup=wu, ms_per_step=ms,
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic", impl="reference",
                frames_per_step=1,
                config=dict(workload=spec["label"], frames_per_step=1,
                            note="oracle port of the reference path on the host cores, one frame of the batch per step; "
                                 "the reference package itself needs mmcv-full (absent, no network)"),
                cpu_baseline=dict(value=val, unit="frames/s", cores=cores, kind="port", sample=sample),
                e2e=dict(value=val, unit="frames/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
# one workload on the GPU(s)
# ---------------------------------------------------------------------------------------------------------------------
class Ctx:
    pass


def algorithmic(name, a):
    """Algorithmic flops (tensor kernels) or compulsory HBM bytes (streaming kernels) of one C-ABI launch."""
    if name in ("ged_gemm_tf32", "ged_gemm_tf32_bt"):
        return 2.0 * a[6] * a[7] * a[8]
    if name in ("ged_conv3x3_tf32", "ged_conv3x3_dx_tf32"):
        return 2.0 * a[4] * a[5] * a[6] * a[7] * a[8] * 9
    if name == "ged_gemm_dw_tf32":            # N x K x P per tap
        return 2.0 * a[6] * a[7] * a[8] * a[10]
    if name == "ged_msda_fwd":                # value + offsets + logits read once, output written once
        Bq, Sq, Qq, nHq = a[8], a[9], a[10], a[11]
        return 4.0 * (Bq * Sq * nHq * 64 + Bq * Qq * nHq * 96 + Bq * Qq * nHq * 64)
    if name in ("ged_msda_tc_bwd", "ged_msda_tile_bwd"):   # + g_out read, g_value read-modify-write, g_off / g_logit written
        Bq, Sq, Qq, nHq = a[13], a[14], a[15], a[16]
        return 4.0 * (3 * Bq * Sq * nHq * 64 + 2 * Bq * Qq * nHq * 96 + Bq * Qq * nHq * 64)
    if name == "ged_act_bwd":                 # g [, ref] -> gz (+ column sums)
        return 4.0 * a[7] * a[8] * (1 + int(a[2] is not None) + int(a[3] is not None))
    if name == "ged_bn_train_fwd":            # statistics pass + normalise pass: x read twice, y written
        return 4.0 * a[9] * a[10] * 3
    if name == "ged_bn_train_bwd":            # sums pass (g, x [, y]) + dx pass (g, x [, y]) + dx written
        return 4.0 * a[10] * a[11] * (2 * (3 if a[2] is not None else 2) + 1)
    if name == "ged_layernorm_fwd":
        return 4.0 * a[6] * a[7] * 2
    if name == "ged_layernorm_bwd":           # g, x [, g_add] read once, dx written (dw / db come out of the same pass for
        # rows of <= 768 channels; wider rows take a second pass over g and x)
        return 4.0 * a[9] * a[10] * (3 + int(a[5] is not None) + (2 if a[10] > 768 else 0))
    if name == "ged_prep_conv_input":         # sources once, bordered tensor written
        C0, h0, w0, C1, Bq, Hq, Wq = a[1], a[2], a[3], a[5], a[7], a[8], a[9]
        return 4.0 * Bq * (h0 * w0 * C0 + Hq * Wq * C1 + (Hq + 2) * (Wq + 2) * (C0 + C1))
    if name == "ged_adamw_step":              # p, g, m, v read; p, m, v written; 1-byte decay mask
        return 29.0 * a[5]
    return 0.0


def measure(c: Ctx, spec: dict, batch: int, steps: int, warmup: int, want_profile: bool, use_graph: bool = True):
    """Build the model of `spec`, run the captured training step `steps` times on device-resident inputs and again end to
    end from pinned host buffers.  Returns a dict; frees everything it allocated."""
    import numpy as np
    torch, dist, kernels = c.torch, c.dist, c.kernels
    import gedepth_b200.models as M
    from gedepth_b200.presets import model_cfg
    from gedepth_b200.synth import synth_batch, synth_state_dict
    from gedepth_b200.train import Trainer
    H, W = spec["H"], spec["W"]
    adaptive, ddad = spec["variant"] == "a", spec["dataset"] == "ddad"
    model = M.build_depther(model_cfg(spec["variant"], spec["dataset"], spec["backbone"], pretrained=None))   # drop_path 0.3 as configured
    model.load_state_dict(synth_state_dict(model.state_dict(), 0))
    model.to(c.dev).train()
    trainer = Trainer(model)
    host = []
    for i in range(2):          # two distinct pinned host batches so the e2e copies are real
        b = synth_batch(batch, H, W, seed=1234 + c.rank * 17 + i, adaptive=adaptive, depth_scale=250.0 if ddad else 200.0,
                        max_depth=200.0 if ddad else 80.0)
        if ddad and adaptive:
            b["height"] = np.array(([1.56, 1.57, 1.53, 1.53] * batch)[:batch], dtype=np.float32)
        host.append({k: torch.from_numpy(v).pin_memory() for k, v in b.items()})
    metas = [dict(ori_shape=(H, W, 3), img_shape=(H, W, 3), pad_shape=(H, W, 3), flip=False)] * batch
    resident = [{k: v.to(c.dev) for k, v in hb.items()} for hb in host]

    def as_batch(d):
        extra = {k: d[k] for k in ("pe_k_gt", "height") if k in d}
        return dict(img=d["img"], img_metas=metas, depth_gt=d["depth_gt"], **extra)

    if use_graph:
        trainer.capture(as_batch(resident[0]), warmup=max(3, warmup))

    def step_eager(i):
        return trainer.step(as_batch(resident[i % 2]))

    def step_resident(i):
        if use_graph:
            return trainer.step_graph(as_batch(resident[i % 2]))          # device->device into the static inputs
        return step_eager(i)

    def step_e2e(i):
        hb = host[i % 2]
        if use_graph:
            loss = trainer.step_graph(as_batch(hb))                       # pinned host -> static device inputs, replay
        else:
            loss, _ = trainer.step(as_batch({k: v.to(c.dev, non_blocking=True) for k, v in hb.items()}))
        return float(loss.detach())                                      # device->host read of the step's result

    def barrier():
        if c.world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = kernels.LAUNCHES
        e0.record()
        for i in range(n):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=c.dev)
        if c.world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / n, kernels.LAUNCHES - n0

    for i in range(max(3, warmup)):
        step_resident(i)
    clocks = ClockSampler(c.local)
    if c.rank == 0:
        clocks.start()
    ms_step, launches = timed(step_resident, steps)
    for i in range(2):
        step_e2e(i)
    ms_e2e, _ = timed(step_e2e, steps)
    res = dict(ms_step=ms_step, ms_e2e=ms_e2e, launches=launches, clocks=clocks.stop() if c.rank == 0 else None,
               h2d=sum(v.numel() * v.element_size() for v in host[0].values()), batch=batch,
               mem_gb=torch.cuda.max_memory_allocated(c.dev) / 2 ** 30, graph=use_graph, prof=None)

    if want_profile:
        # per-kernel device time of ONE extra eager step (CUDA events around every C-ABI launch)
        if use_graph:
            trainer.release_graph()
        if c.rank != 0:
            step_eager(1); step_eager(0)          # same collectives as rank 0's two eager steps below
        else:
            orig_call, records = kernels._call, []

            def prof_call(name, *a):
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                orig_call(name, *a)
                e.record()
                records.append((name, s, e, algorithmic(name, a)))

            step_eager(1)                      # re-warm the eager path (allocator pools differ from the graph's)
            kernels._call = prof_call
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            step_eager(0)
            e1.record()
            torch.cuda.synchronize()
            kernels._call = orig_call
            prof = {}
            for name, s, e, fl in records:
                d = prof.setdefault(name, dict(ms=0.0, calls=0, flops=0.0))
                d["ms"] += s.elapsed_time(e)
                d["calls"] += 1
                d["flops"] += fl
            res["prof"], res["step_ms_profiled"] = prof, e0.elapsed_time(e1)
        if c.world > 1:
            dist.barrier()
    # free everything before the next workload is built
    trainer.release_graph() if use_graph and trainer._graph is not None else None
    del trainer, model, resident, host
    kernels.RNG_STEP = None
    gc.collect()
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    torch.cuda.reset_peak_memory_stats(c.dev)
    return res


def measure_with_oom_fallback(c, spec, batch, steps, warmup, want_profile, use_graph=True):
    """The headline batch fills the 180 GB of one B200 almost completely; if the allocator refuses, halve the batch
    rather than lose the measurement (the line then says so)."""
    note = None
    while True:
        oom = False
        try:
            r = measure(c, spec, batch, steps, warmup, want_profile, use_graph)
            r["oom_note"] = note
            return r
        except c.torch.cuda.OutOfMemoryError:
            if c.world > 1 or batch <= 1:
                raise
            oom = True
        if oom:          # outside the handler: the traceback (and the frames holding the tensors) is gone now
            gc.collect()
            c.torch.cuda.synchronize()
            c.torch.cuda.empty_cache()
            c.kernels.RNG_STEP = None
            note = f"batch {batch} did not fit in device memory next to the CUDA-graph pool; measured at batch {batch // 2}"
            batch //= 2


def library_baseline(c: Ctx, spec: dict, batch: int, steps: int):
    """The reference's own design on this GPU: the same model and step with every op routed to its cuDNN / cuBLAS
    (allow_tf32=True, PyTorch 1.8's default on Ampere+) / ATen statement and grid_sample deformable attention, eager
    launches, torch autograd - what `tools/benchmark.py`-style timing of the reference would run.  A small batch: the
    grid_sample MSDA materialises (B*8, 64, Q, 8) per level (26 GB per level at batch 16)."""
    torch = c.torch
    import gedepth_b200.models as M
    from gedepth_b200 import ops
    from gedepth_b200.presets import model_cfg
    from gedepth_b200.synth import synth_batch, synth_state_dict
    H, W = spec["H"], spec["W"]
    adaptive, ddad = spec["variant"] == "a", spec["dataset"] == "ddad"
    from tests import ops_lib          # the library statement of every op: comparator only, never the product path
    restore = ops_lib.install(ops)
    try:
        model = M.build_depther(model_cfg(spec["variant"], spec["dataset"], spec["backbone"], pretrained=None))
        model.load_state_dict(synth_state_dict(model.state_dict(), 0))
        model.to(c.dev).train()
        opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=0.01)
        b = synth_batch(batch, H, W, seed=99, adaptive=adaptive, depth_scale=250.0 if ddad else 200.0,
                        max_depth=200.0 if ddad else 80.0)
        d = {k: torch.from_numpy(v).to(c.dev) for k, v in b.items()}
        metas = [dict(ori_shape=(H, W, 3), img_shape=(H, W, 3), pad_shape=(H, W, 3), flip=False)] * batch
        extra = {k: d[k] for k in ("pe_k_gt",) if k in d}

        def step():
            opt.zero_grad(set_to_none=True)
            losses = model(img=d["img"], img_metas=metas, depth_gt=d["depth_gt"], **extra)
            loss, _ = model._parse_losses(losses, sync=False)
            loss.backward()
            torch.nn.utils.clip_grad_norm_(model.parameters(), 35.0)
            opt.step()
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        out = dict(frames_s=batch / (ms / 1e3), ms_per_step=ms, batch=batch, steps=steps, launch="eager",
                   note="same model / step through tests/ops_lib.py: cuDNN convs and cuBLAS linears with allow_tf32=True, ATen "
                        "elementwise + LayerNorm + BatchNorm, grid_sample deformable attention, torch AdamW + clip_grad_norm_")
        del model, opt, d
    finally:
        restore()
        gc.collect()
        torch.cuda.synchronize()
        torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=HEADLINE, choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default: the workload's)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip other_configs / gpu_library_baseline / probes (N = 1 only anyway)")
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of the captured CUDA graph")
    ap.add_argument("--passes", type=int, default=3, choices=[1, 2, 3], help="forward GEMM arithmetic: 3 = 3xTF32 (fp32-accurate, default), 2 = bf16 hi/lo split (3 kind::f16 products), 1 = TF32")
    ap.add_argument("--ncu-step", action="store_true",
                    help="for `ncu --profile-from-start off`: warm up, bracket ONE eager step with cudaProfilerStart/Stop, exit")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from gedepth_b200 import kernels, ops

    c = Ctx()
    c.torch, c.dist, c.kernels = torch, dist, kernels
    c.world = int(os.environ.get("WORLD_SIZE", "1"))
    c.rank = int(os.environ.get("RANK", "0"))
    c.local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(c.local)
    c.dev = torch.device("cuda", c.local)
    if c.world > 1:
        dist.init_process_group("nccl", device_id=c.dev)
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    kernels.load()
    kernels.set_gemm_precision(args.passes)
    world, rank, dev = c.world, c.rank, c.dev

    spec = WORKLOADS[args.workload]
    Bn = args.batch or spec["batch"]
    H, W = spec["H"], spec["W"]
    use_graph = not args.no_graph

    if args.ncu_step:
        ncu_step(c, spec, Bn, args)
        return

    main_res = measure_with_oom_fallback(c, spec, Bn, args.steps, args.warmup, want_profile=True, use_graph=use_graph)
    Bn = main_res["batch"]
    extras = world == 1 and not args.no_extras
    others, lib = {}, None
    if extras:
        for name in ("config2", "config4"):
            if name == args.workload:
                continue
            r = measure_with_oom_fallback(c, WORKLOADS[name], WORKLOADS[name]["batch"], args.steps, args.warmup, False, use_graph)
            others[name] = dict(workload=WORKLOADS[name]["label"], frames_s=r["batch"] / (r["ms_step"] / 1e3), ms_per_step=r["ms_step"],
                                e2e_frames_s=r["batch"] / (r["ms_e2e"] / 1e3), batch=r["batch"], gpu_launches=r["launches"],
                                device_mem_gb=round(r["mem_gb"], 1))
        try:
            lib = library_baseline(c, spec, 2, 3)
        except Exception as ex:          # the comparator must never cost the headline
            lib = dict(error=f"{type(ex).__name__}: {str(ex)[:200]}")

    def finish():
        """Leave without tearing NCCL down: communicators captured in CUDA graphs can block
        destroy_process_group(); every rank meets at one last barrier, flushes and exits."""
        sys.stdout.flush()
        sys.stderr.flush()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
            os._exit(0)

    if rank != 0:
        return finish()

    pk = peaks()
    ncu_traffic = {}
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")     # dram bytes per launch from ncu --set full captures
    if os.path.exists(tpath):
        ncu_traffic = json.load(open(tpath))
    ms_step, ms_e2e, prof = main_res["ms_step"], main_res["ms_e2e"], main_res["prof"]
    frames = Bn * world
    value = frames / (ms_step / 1e3)
    e2e_val = frames / (ms_e2e / 1e3)

    # ---- dominant kernel of the step ---------------------------------------------------------------------------------
    kern = prof
    native_ms = sum(v["ms"] for v in kern.values())
    dom = max(kern, key=lambda k: kern[k]["ms"])
    tensor_names = ("ged_gemm_tf32", "ged_conv3x3_tf32", "ged_gemm_tf32_bt", "ged_conv3x3_dx_tf32", "ged_gemm_dw_tf32")
    fwd_names = ("ged_gemm_tf32", "ged_conv3x3_tf32")
    tf32_peak = pk["bf16_sustained"] / 2.0
    # tensor-pipe time of one algorithmic pass, in units of a TF32 pass: 3xTF32 issues 3 kind::tf32 MMAs per k-step, the bf16
    # hi/lo split 3 kind::f16 MMAs at twice the TF32 rate, single-pass TF32 one
    pipe_passes = {1: 1.0, 2: 1.5, 3: 3.0}
    tkey = f"{dom}@{args.workload}"
    if dom in tensor_names:
        tf = kern[dom]["flops"] / (kern[dom]["ms"] / 1e3) / 1e12
        passes = args.passes if dom in fwd_names else kernels.BACKWARD_PASSES
        roof = dict(kernel=dom, bound="tensor", achieved=tf, peak=tf32_peak, unit="TFLOP/s", frac=tf / tf32_peak,
                    traffic=ncu_traffic.get(tkey, ncu_traffic.get(dom)),
                    peak_note=f"TF32 dense = 1/2 of the {pk['src']} sustained bf16 cuBLAS figure ({pk['bf16_sustained']})",
                    calls_per_step=kern[dom]["calls"], avg_launch_ms=kern[dom]["ms"] / kern[dom]["calls"],
                    share_of_step=kern[dom]["ms"] / main_res["step_ms_profiled"], mma_passes=passes,
                    issued_mma_frac=tf * pipe_passes[passes] / tf32_peak,
                    note="achieved counts ALGORITHMIC flops (2MNK) of all launches of this kernel in one step over their "
                         "summed device time (CUDA events on the launching stream); issued_mma_frac = tensor-pipe time of the "
                         "MMAs actually issued: mma_passes 3 = 3xTF32 (3 kind::tf32 MMAs per k-step), 2 = bf16 hi/lo split (3 "
                         "kind::f16 MMAs at twice the TF32 rate = 1.5 TF32 passes), 1 = single-pass TF32")
    else:
        gbs = kern[dom]["flops"] / (kern[dom]["ms"] / 1e3) / 1e9 if kern[dom]["flops"] else None
        roof = dict(kernel=dom, bound="hbm", achieved=gbs, peak=pk["hbm"], unit="GB/s",
                    frac=(gbs / pk["hbm"]) if gbs else None, traffic=ncu_traffic.get(tkey, ncu_traffic.get(dom)),
                    calls_per_step=kern[dom]["calls"], avg_launch_ms=kern[dom]["ms"] / kern[dom]["calls"],
                    share_of_step=kern[dom]["ms"] / main_res["step_ms_profiled"],
                    note="achieved = COMPULSORY HBM bytes (each operand once) / summed device time of the kernel's launches")
    tens = [kern[k] for k in tensor_names if k in kern]
    tens_ms = sum(t["ms"] for t in tens)
    tens_tf = sum(t["flops"] for t in tens) / (tens_ms / 1e3) / 1e12 if tens_ms else None

    def group(names, passes):
        ks = [kern[k] for k in names if k in kern]
        ms = sum(t["ms"] for t in ks)
        if not ms:
            return None
        alg = sum(t["flops"] for t in ks) / (ms / 1e3) / 1e12
        return dict(kernels=[k for k in names if k in kern], ms=round(ms, 3), algorithmic_tflops=alg, mma_passes=passes,
                    issued_mma_tflops=alg * pipe_passes[passes], tensor_pipe_frac=alg * pipe_passes[passes] / tf32_peak)
    roof_tensor = dict(kernels=list(tensor_names), bound="tensor", achieved=tens_tf, peak=tf32_peak, unit="TFLOP/s",
                       frac=(tens_tf / tf32_peak) if tens_tf else None, share_of_step=tens_ms / main_res["step_ms_profiled"],
                       forward=group(fwd_names, args.passes),
                       backward=group(tuple(k for k in tensor_names if k not in fwd_names), kernels.BACKWARD_PASSES),
                       note="ALGORITHMIC flops (2MNK) of every tcgen05 GEMM / conv launch of the step over their summed device "
                            "time, against TF32 dense = 1/2 of the measured sustained bf16 figure")
    msda = {k: dict(ms=round(kern[k]["ms"], 3), calls=kern[k]["calls"]) for k in kern if "msda" in k}
    hbm_kernels = {}
    for k in ("ged_act_bwd", "ged_bn_train_fwd", "ged_bn_train_bwd", "ged_layernorm_fwd", "ged_layernorm_bwd",
              "ged_prep_conv_input", "ged_adamw_step"):
        if k in kern and kern[k]["flops"] > 0:
            gb = kern[k]["flops"] / (kern[k]["ms"] / 1e3) / 1e9
            hbm_kernels[k] = dict(ms=round(kern[k]["ms"], 3), calls=kern[k]["calls"], achieved_gbs=round(gb, 1),
                                  frac=round(gb / pk["hbm"], 3))

    ge = swin_attn = cpu = aug = None
    if extras:
        ge = ground_embed_probe(c, pk, Bn, H, W)
        swin_attn = swin_attention_probe(c, pk, spec, Bn, H, W, args.passes)
        aug = train_augment_probe(c, pk)
    if not args.no_cpu_baseline and world == 1:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        t = cpu_step_times(torch, spec, 1, budget_s=90.0, steps=1, warmup=0)
        cpu = dict(value=1.0 / t[0], unit="frames/s", cores=cores, kind="port",
                   sample=f"1 frame {H}x{W} fwd+bwd of the headline model through the oracle port (plain PyTorch CPU), all host threads",
                   ground_plane_numpy_mpx_s=ground_plane_numpy_rate(H, W))

    cfg = dict(workload=spec["label"], global_batch=frames, parallelism=f"dp{world}",
               step="fwd + SiLog (+0.08 CE) + bwd + allreduce(N>1) + clip + AdamW", drop_path_rate=0.3,
               launch="one captured CUDA graph per step" if use_graph else "eager",
               l2="inputs + activations per step (tens of GB) exceed the 126 MB L2",
               device_mem_gb=round(main_res["mem_gb"], 1))
    if main_res.get("oom_note"):
        cfg["note"] = main_res["oom_note"]
    line = dict(metric=METRIC, value=value, unit="frames/s", n_gpus=world, steps=args.steps, warmup=max(3, args.warmup),
                ms_per_step=ms_step, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype=("f32 (fp32 storage and accumulation; forward GEMMs/convs error-compensated 3xTF32 on tcgen05, backward GEMMs "
                       "single-pass TF32 = what the reference's PyTorch 1.8 runs on Ampere+)") if args.passes == 3
                else ("f32 storage and accumulation; forward GEMMs/convs bf16 hi/lo split (hi*hi + lo*hi + hi*lo as kind::f16, 2^-16 per "
                      "product) on tcgen05, backward GEMMs single-pass TF32") if args.passes == 2
                else "tf32 (single-pass TF32 forward and backward, fp32 storage and accumulation)", data="synthetic",
                config=cfg,
                e2e=dict(value=e2e_val, unit="frames/s", h2d_bytes_per_step=main_res["h2d"], d2h_bytes_per_step=4, ms_per_step=ms_e2e),
                gpu_launches=main_res["launches"], clocks=main_res["clocks"], roofline=roof, roofline_tensor=roof_tensor,
                msda_kernels=msda, ground_embed=ge, swin_window_attention=swin_attn, train_augment=aug,
                roofline_hbm_kernels=hbm_kernels,
                cpu_baseline=cpu, other_configs=others or None, gpu_library_baseline=lib,
                vs_library_gpu=(value / lib["frames_s"]) if lib and "frames_s" in lib else None,
                native_ops=ops.native_table(),
                kernel_ms={k: round(v["ms"], 3) for k, v in sorted(kern.items(), key=lambda kv: -kv[1]["ms"])},
                native_ms_of_step=[round(native_ms, 2), round(main_res["step_ms_profiled"], 2)])
    print(json.dumps(line), flush=True)
    finish()


def ncu_step(c, spec, Bn, args):
    import numpy as np
    torch, kernels = c.torch, c.kernels
    import gedepth_b200.models as M
    from gedepth_b200.presets import model_cfg
    from gedepth_b200.synth import synth_batch, synth_state_dict
    from gedepth_b200.train import Trainer
    adaptive, ddad = spec["variant"] == "a", spec["dataset"] == "ddad"
    model = M.build_depther(model_cfg(spec["variant"], spec["dataset"], spec["backbone"], pretrained=None))
    model.load_state_dict(synth_state_dict(model.state_dict(), 0))
    model.to(c.dev).train()
    trainer = Trainer(model)
    b = synth_batch(Bn, spec["H"], spec["W"], seed=1234, adaptive=adaptive, depth_scale=250.0 if ddad else 200.0,
                    max_depth=200.0 if ddad else 80.0)
    if ddad and adaptive:
        b["height"] = np.array(([1.56, 1.57, 1.53, 1.53] * Bn)[:Bn], dtype=np.float32)
    d = {k: torch.from_numpy(v).to(c.dev) for k, v in b.items()}
    batch = dict(img=d["img"], img_metas=[{}] * Bn, depth_gt=d["depth_gt"], **{k: d[k] for k in ("pe_k_gt", "height") if k in d})
    for _ in range(max(3, args.warmup)):
        trainer.step(batch)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    trainer.step(batch)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print(json.dumps(dict(ncu_step=True, workload=spec["label"], batch=Bn, gpu_launches=kernels.LAUNCHES)), flush=True)


def ground_embed_probe(c, pk, Bn, H, W):
    """The kernels the metric names, alone, against the measured HBM copy peak (L2 flushed between repetitions):
    at the workload's shape and at the large end of the sweep (BASELINE configs[4])."""
    torch, kernels = c.torch, c.kernels
    flush = torch.empty(256 * 1024 * 1024 // 4, device=c.dev)

    def timeit(fn, reps):
        for _ in range(3):
            fn()
        tot = 0.0
        for _ in range(reps):
            flush.zero_()                         # evict L2 (126 MB) between iterations
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); fn(); e.record()
            torch.cuda.synchronize()
            tot += s.elapsed_time(e)
        return tot / reps / 1e3

    def one(Bx, Hx, Wx, reps):
        img = torch.randn(Bx, 5, Hx, Wx, device=c.dev)
        img[:, 4] = img[:, 4].abs() * 30 + 2
        yh = torch.rand(Bx, 1, Hx // 2, Wx // 2, device=c.dev)
        lh = torch.randn(Bx, 11, Hx // 2, Wx // 2, device=c.dev)
        px = Bx * Hx * Wx
        out = {}
        with torch.no_grad():
            t = timeit(lambda: kernels.ge_vanilla(img, yh), reps)
            out["ge_vanilla_fwd"] = dict(bytes_per_px=13, us=t * 1e6, achieved=13.0 * px / t / 1e9)
            t = timeit(lambda: kernels.ge_adaptive(img, yh, lh, 1.65, 200.0), reps)
            out["ge_adaptive_fwd_inference"] = dict(bytes_per_px=24, us=t * 1e6, achieved=24.0 * px / t / 1e9)
            t = timeit(lambda: kernels.ground_plane_into(img, (-1.578, -1.464e-5, -1.386e-3, 0.2589)), reps)
            out["ground_plane"] = dict(bytes_per_px=8, us=t * 1e6, achieved=8.0 * px / t / 1e9)
        # training-mode forward (full-resolution logits written) and the two backward kernels, straight through the C ABI
        # on preallocated tensors (through autograd the host-side launch overhead would be timed, not the kernel)
        from gedepth_b200.kernels import _call, _p, _stream
        y, pm = torch.empty(Bx, 1, Hx, Wx, device=c.dev), torch.empty(Bx, 1, Hx, Wx, device=c.dev)
        lf = torch.empty(Bx, 11, Hx, Wx, device=c.dev)
        gy, gpm, glf = torch.randn_like(y), torch.randn_like(pm), torch.randn_like(lf)
        g_yh, g_lh = torch.empty_like(yh), torch.empty_like(lh)
        h2, w2, bs = Hx // 2, Wx // 2, 5 * Hx * Wx
        t = timeit(lambda: _call("ged_ge_adaptive_fwd", _p(img[:, 4]), bs, _p(yh), _p(lh), None, 1.65, 200.0, _p(y), _p(pm), _p(lf),
                                 Bx, Hx, Wx, h2, w2, _stream()), reps)
        out["ge_adaptive_fwd_train"] = dict(bytes_per_px=68, us=t * 1e6, achieved=68.0 * px / t / 1e9)
        t = timeit(lambda: _call("ged_ge_vanilla_bwd", _p(img[:, 3]), bs, _p(gy), _p(gpm), _p(g_yh), Bx, Hx, Wx, h2, w2, _stream()), reps)
        out["ge_vanilla_bwd"] = dict(bytes_per_px=13, us=t * 1e6, achieved=13.0 * px / t / 1e9)
        t = timeit(lambda: _call("ged_ge_adaptive_bwd", _p(img[:, 4]), bs, _p(yh), _p(lh), None, 1.65, 200.0, _p(gy), _p(gpm), _p(glf),
                                 _p(g_yh), _p(g_lh), Bx, Hx, Wx, h2, w2, _stream()), reps)
        out["ge_adaptive_bwd"] = dict(bytes_per_px=62.0, us=t * 1e6, achieved=62.0 * px / t / 1e9,
                                      note="12 B/px of g_y, g_pe_mask, pe + 44 B/px of g_logits + 3 + 3 B/px of half-resolution operands / outputs")
        for v in out.values():
            v["frac"] = v["achieved"] / pk["hbm"]
        return out
    res = dict(bound="hbm", unit="GB/s", peak=pk["hbm"], peak_src=pk["src"], l2_flushed_between_iterations=True,
               at_workload=dict(shape=[Bn, H, W], kernels=one(Bn, H, W, 10)),
               at_sweep_max=dict(shape=[32, 1024, 2048], kernels=one(32, 1024, 2048, 4)),
               note="algorithmic bytes per full-resolution pixel as in SURVEY 8(d); the workload shape is tens of MB: launch / L2 "
                    "latency dominated")
    del flush
    torch.cuda.empty_cache()
    return res


def train_augment_probe(c, pk, frames=16):
    """the KITTI train-time augmentation of the 5-channel input on the device (csrc/augment.cu) next to the
    same chain through cv2 / numpy on the host (what the reference's data-loader workers run), same drawn parameters."""
    import random
    import numpy as np
    torch = c.torch
    from gedepth_b200 import augment as ga
    from gedepth_b200.synth import synth_raw_frame
    aug = ga.TrainAugmenter(c.dev)
    np.random.seed(7)
    random.seed(7)
    img5, depth, lab = synth_raw_frame(7)
    bgr = torch.from_numpy(np.ascontiguousarray(img5[:, :, 0:3].astype(np.uint8))).to(c.dev)
    f = aug.frame_planes(bgr, torch.from_numpy(np.ascontiguousarray(img5[:, :, 3])).to(c.dev),
                         torch.from_numpy(np.ascontiguousarray(img5[:, :, 4])).to(c.dev))
    d, l = torch.from_numpy(depth).to(c.dev), torch.from_numpy(lab).to(c.dev)
    params = [ga.draw_params() for _ in range(frames)]
    run = lambda: aug([f] * frames, [d] * frames, [l] * frames, params)
    for _ in range(3):
        run()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    s.record()
    for _ in range(reps):
        run()
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / reps
    # algorithmic bytes per frame: KB window read (7 planes), canvas written and read back (7 planes), 7 output planes
    by = sum(4.0 * 7 * (352 * 1216 + 2 * p["canvas_w"] * p["canvas_h"] + 352 * 704) for p in params)
    out = dict(frames=frames, ms=round(ms, 3), frames_s=round(frames / ms * 1e3, 1), achieved_gbs=round(by / ms / 1e6, 1),
               frac_of_hbm_peak=round(by / ms / 1e6 / pk["hbm"], 3),
               note="resize + pad + rotate + flip + crop + ColorAug + Normalize: one descriptor copy + two kernels per BATCH; "
                    "bit-identical to the reference transforms for the same drawn parameters")
    try:
        from oracle import augment as oa           # CPU baseline leg: the same chain through cv2 / numpy on the host
        t = sum(oa.cv2_reference_seconds(img5, depth, lab, p, reps=2) for p in params[:4]) / 4
        out["cpu_cv2"] = dict(frames_s=round(1.0 / t, 1), cores=1, kind="port",
                              sample="4 frames through cv2.resize / cv2.warpAffine / numpy as the reference's loader workers run them, one thread")
    except Exception as ex:      # cv2 missing on the box
        out["cpu_cv2"] = dict(unavailable=str(ex)[:80])
    return out


def swin_attention_probe(c, pk, spec, Bn, H, W, passes):
    """Swin window attention (QKV GEMM -> 49x49 core -> proj GEMM), the north star's tensor-pipe target."""
    torch, kernels = c.torch, c.kernels
    from gedepth_b200.presets import SWIN_L, SWIN_T
    from gedepth_b200.swin import WindowMSA
    sw = SWIN_L if spec["backbone"] == "swin_l" else SWIN_T
    Bx = min(Bn, 8)

    def one(tokens_hw, Cc, nH, shift):
        hh, ww = tokens_hw
        T = Bx * hh * ww
        x = torch.randn(Bx, hh * ww, Cc, device=c.dev)
        wq, bq = torch.randn(3 * Cc, Cc, device=c.dev) / Cc ** 0.5, torch.randn(3 * Cc, device=c.dev) * 0.02
        wp, bp = torch.randn(Cc, Cc, device=c.dev) / Cc ** 0.5, torch.randn(Cc, device=c.dev) * 0.02
        table = torch.randn(169, nH, device=c.dev) * 0.2
        index = WindowMSA(Cc, nH, (7, 7)).relative_position_index.to(c.dev)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        with torch.no_grad():
            def run(record):
                if record: ev[0].record()
                qkv = kernels.linear(x, wq, bq)
                if record: ev[1].record()
                ctx = kernels.window_attention(qkv, bq, table, index, (hh, ww), nH, 7, shift, 32 ** -0.5)
                if record: ev[2].record()
                out = kernels.linear(ctx, wp, bp, None, x)
                if record: ev[3].record()
                return out
            for _ in range(3):
                run(False)
            run(True)
            torch.cuda.synchronize()
        t = [ev[i].elapsed_time(ev[i + 1]) for i in range(3)]
        nW = (-(-hh // 7)) * (-(-ww // 7)) * Bx
        f_gemm = 2.0 * T * Cc * 3 * Cc + 2.0 * T * Cc * Cc
        f_core = 2.0 * nW * nH * 49 * 49 * 32 * 2
        core_tc = kernels.WINATTN_TC if hasattr(kernels, "WINATTN_TC") else False
        ms = sum(t)
        # tensor-pipe time in TF32 passes: GEMMs per the forward arithmetic (3xTF32 = 3, bf16 split = 1.5, TF32 = 1); the core is
        # always 3xTF32 when it runs on the tensor cores
        issued = {1: 1.0, 2: 1.5, 3: 3.0}[passes] * f_gemm + (3.0 * f_core if core_tc else 0.0)
        return dict(shape=dict(tokens=[Bx, hh, ww], C=Cc, heads=nH, shift=shift), ms=dict(qkv=t[0], core=t[1], proj=t[2]),
                    algorithmic_tflops=(f_gemm + f_core) / ms / 1e9, issued_mma_tflops=issued / ms / 1e9,
                    tensor_pipe_frac=issued / ms / 1e9 / (pk["bf16_sustained"] / 2.0), core_on_tensor_cores=bool(core_tc))
    e0, h0 = sw["embed_dims"], sw["num_heads"][0]
    return [one((-(-H // 4), -(-W // 4)), e0, h0, 3), one((-(-H // 16), -(-W // 16)), 4 * e0, 4 * h0, 0)]


if __name__ == "__main__":
    main()
