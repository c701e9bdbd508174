#!/usr/bin/env python
"""bench.py - frames/s (352x1120) fwd+bwd of the GEDepth path, BASELINE config 2:
DepthFormer-Swin-T + GEDepth-Vanilla, batch 8 per GPU, synthetic KITTI-shape frames, random-init
(deterministic synthetic) weights.  One "step" = forward + SiLog + backward (+ one NCCL all-reduce
of the flat gradient arena when N>1) + clip + AdamW on one batch.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Prints ONE JSON line (rank 0).  `value` = device-resident throughput; `e2e` = the same metric through
the public train-step call with HOST (pinned) buffers copied H2D every step and the loss read back;
`roofline` = the dominant kernel of the step (by summed device time, measured with CUDA events around
every C-ABI launch of one extra step); `ground_embed` = the HBM roofline of the kernel the metric
names; `cpu_baseline` = the oracle port of the same step on this box's host cores.
--impl reference times that CPU port alone (the reference's design cannot run here: mmcv is absent).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W, B_PER_GPU = 352, 1120, 8
WORKLOAD = "DepthFormer-Swin-T + GEDepth-Vanilla, batch 8/GPU, 352x1120 synthetic KITTI (BASELINE configs[1])"


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


def cpu_step_time(torch, n_frames=1, budget_s=200.0, steps=1, warmup=0):
    """Oracle port (plain PyTorch CPU restatement of the reference path) fwd+bwd on `n_frames` frames."""
    import gedepth_b200.models as M
    from gedepth_b200.presets import model_cfg
    from gedepth_b200.synth import synth_batch, synth_state_dict
    from oracle import model as om
    tmpl = M.build_depther(model_cfg("v", "kitti", "swin_t", pretrained=None)).state_dict()
    sd = synth_state_dict(tmpl, 0)
    skip = ("running_mean", "running_var", "num_batches_tracked", "relative_position_index")
    sd = {k: v.requires_grad_(not k.endswith(skip)) for k, v in sd.items()}
    b = synth_batch(n_frames, H, W, seed=1234)
    img, gt = torch.from_numpy(b["img"]), torch.from_numpy(b["depth_gt"])
    cfg = om.PathConfig(train_bn=True)
    times = []
    t_begin = time.time()
    for i in range(warmup + steps):
        t0 = time.time()
        r = om.forward_train(sd, cfg, img, gt)
        r["loss"].backward()
        for v in sd.values():
            v.grad = None
        dt = time.time() - t0
        if i >= warmup:
            times.append(dt)
        if time.time() - t_begin > budget_s and times:
            break
    return times


def ground_plane_numpy_rate():
    import numpy as np
    from oracle import ground as og
    coef = og.plane_coefficients(og.kitti_projection(), og.KITTI_CAM_HEIGHT)
    best = 1e9
    for _ in range(5):
        t0 = time.perf_counter()
        og.ground_plane(coef, H, W, 61, 23)
        best = min(best, time.perf_counter() - t0)
    return H * W / best / 1e6


def run_reference(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    times = cpu_step_time(torch, 1, budget_s=200.0, steps=max(1, args.steps), warmup=min(args.warmup, 1))
    ms = 1e3 * sum(times) / len(times)
    val = 1.0 / (ms / 1e3)
    sample = f"1 frame of the batch per step (352x1120 fwd+bwd), {len(times)} timed steps, torch-CPU {torch.get_num_threads()} threads"
    line = dict(metric="frames/sec (352x1120) fwd+bwd", value=val, unit="frames/s", n_gpus=args.gpus, steps=len(times),
                warmup=min(args.warmup, 1), ms_per_step=ms, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f32", data="synthetic", impl="reference",
                config=dict(workload=WORKLOAD, note="oracle port of the reference path on host cores; the reference's "
                            "own GPU path needs mmcv-full (absent, no network)"),
                cpu_baseline=dict(value=val, unit="frames/s", cores=cores, kind="port", sample=sample),
                e2e=dict(value=val, unit="frames/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=B_PER_GPU)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--variant", default="v", choices=["v", "a"])
    ap.add_argument("--backbone", default="swin_t", choices=["swin_t", "swin_l"])
    ap.add_argument("--dataset", default="kitti", choices=["kitti", "ddad"], help="ddad: 384x640, depth_scale 250 (BASELINE configs[3])")
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of the captured CUDA graph")
    ap.add_argument("--passes", type=int, default=3, choices=[1, 3], help="GEMM arithmetic: 3 = 3xTF32 (fp32-accurate), 1 = TF32")
    ap.add_argument("--ncu-step", action="store_true",
                    help="for `ncu --profile-from-start off`: warm up, bracket ONE eager step with cudaProfilerStart/Stop, exit")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    import gedepth_b200.models as M
    from gedepth_b200 import kernels, ops
    from gedepth_b200.presets import model_cfg
    from gedepth_b200.synth import synth_batch, synth_state_dict
    from gedepth_b200.train import Trainer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    kernels.load()
    kernels.set_gemm_precision(args.passes)

    global H, W
    if args.dataset == "ddad":
        H, W = 384, 640
    Bn = args.batch
    adaptive = args.variant == "a"
    workload = WORKLOAD
    if (args.variant, args.backbone, args.dataset, Bn) != ("v", "swin_t", "kitti", B_PER_GPU):
        workload = (f"DepthFormer-{args.backbone} + GEDepth-{'Adaptive' if adaptive else 'Vanilla'}, batch {Bn}/GPU, "
                    f"{H}x{W} synthetic {args.dataset.upper()} (not the headline configuration)")
    model = M.build_depther(model_cfg(args.variant, args.dataset, args.backbone, pretrained=None))   # drop_path 0.3 as configured
    model.load_state_dict(synth_state_dict(model.state_dict(), 0))
    model.to(dev).train()
    trainer = Trainer(model)

    # synthetic host batches (pinned) - a few distinct ones so e2e copies are real
    host = []
    for i in range(2):
        b = synth_batch(Bn, H, W, seed=1234 + rank * 17 + i, adaptive=adaptive,
                        depth_scale=250.0 if args.dataset == "ddad" else 200.0,
                        max_depth=200.0 if args.dataset == "ddad" else 80.0)
        if args.dataset == "ddad" and adaptive:
            b["height"] = np.array(([1.56, 1.57, 1.53, 1.53] * Bn)[:Bn], dtype=np.float32)
        host.append({k: torch.from_numpy(v).pin_memory() for k, v in b.items()})
    metas = [dict(ori_shape=(H, W, 3), img_shape=(H, W, 3), pad_shape=(H, W, 3), flip=False)] * Bn
    resident = [{k: v.to(dev) for k, v in hb.items()} for hb in host]

    def as_batch(d):
        extra = {k: d[k] for k in ("pe_k_gt", "height") if k in d}
        return dict(img=d["img"], img_metas=metas, depth_gt=d["depth_gt"], **extra)

    if args.ncu_step:
        for i in range(max(3, args.warmup)):
            trainer.step(as_batch(resident[i % len(resident)]))
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        trainer.step(as_batch(resident[0]))
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        print(json.dumps(dict(ncu_step=True, workload=workload, gpu_launches=kernels.LAUNCHES)), flush=True)
        return

    use_graph = not args.no_graph
    if use_graph:
        trainer.capture(as_batch(resident[0]), warmup=max(3, args.warmup))

    def step_eager(i):
        return trainer.step(as_batch(resident[i % len(resident)]))

    def step_resident(i):
        if use_graph:
            return trainer.step_graph(as_batch(resident[i % len(resident)]))     # device->device into the static inputs
        return step_eager(i)

    def step_e2e(i):
        hb = host[i % len(host)]
        if use_graph:
            loss = trainer.step_graph(as_batch(hb))                # pinned host -> static device inputs, then replay
        else:
            d = {k: v.to(dev, non_blocking=True) for k, v in hb.items()}
            loss, _ = trainer.step(as_batch(d))
        return float(loss.detach())          # device->host read of the step's result

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = kernels.LAUNCHES
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / steps, kernels.LAUNCHES - n0

    for i in range(max(3, args.warmup)):
        step_resident(i)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ms_step, launches = timed(step_resident, args.steps)
    for i in range(2):
        step_e2e(i)
    ms_e2e, _ = timed(step_e2e, args.steps)
    clk = clocks.stop() if rank == 0 else None

    # ---- per-kernel device time of ONE extra step (CUDA events around every C-ABI launch) ----------
    prof = {}
    if use_graph:
        trainer.release_graph()         # timing is done: give the graph's activation pool back before the eager passes
    if rank != 0:
        step_eager(1)          # same collectives as rank 0's two eager steps below
        step_eager(0)
    if rank == 0:
        orig_call = kernels._call
        records = []

        def prof_call(name, *a):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            orig_call(name, *a)
            e.record()
            flops = 0.0          # algorithmic flops (tensor kernels) or compulsory HBM bytes (the rest)
            if name == "ged_gemm_tf32":
                flops = 2.0 * a[6] * a[7] * a[8]
            elif name == "ged_conv3x3_tf32":
                flops = 2.0 * a[4] * a[5] * a[6] * a[7] * a[8] * 9
            elif name == "ged_gemm_tf32_bt":
                flops = 2.0 * a[6] * a[7] * a[8]
            elif name == "ged_conv3x3_dx_tf32":
                flops = 2.0 * a[4] * a[5] * a[6] * a[7] * a[8] * 9
            elif name == "ged_gemm_dw_tf32":   # N x K x P per tap
                flops = 2.0 * a[6] * a[7] * a[8] * a[10]
            elif name == "ged_msda_fwd":       # value + offsets + logits read once, output written once
                Bq, Sq, Qq, nHq = a[8], a[9], a[10], a[11]
                flops = 4.0 * (Bq * Sq * nHq * 64 + Bq * Qq * nHq * 96 + Bq * Qq * nHq * 64)
            elif name == "ged_msda_bwd":       # + g_out read, g_value read-modify-write, g_off / g_logit written
                Bq, Sq, Qq, nHq = a[12], a[13], a[14], a[15]
                flops = 4.0 * (3 * Bq * Sq * nHq * 64 + 2 * Bq * Qq * nHq * 96 + Bq * Qq * nHq * 64)
            # compulsory HBM bytes of the streaming kernels (every operand once per pass the kernel makes)
            if name == "ged_act_bwd":          # g [, ref] -> gz (+ column sums)
                has_ref, writes = a[2] is not None, a[3] is not None
                flops = 4.0 * a[7] * a[8] * (1 + int(has_ref) + int(writes))
            elif name == "ged_bn_train_fwd":   # statistics pass + normalise pass: x read twice, y written
                flops = 4.0 * a[9] * a[10] * 3
            elif name == "ged_bn_train_bwd":   # sums pass (g, x [, y]) + dx pass (g, x [, y]) + dx written
                flops = 4.0 * a[10] * a[11] * (2 * (3 if a[2] is not None else 2) + 1)
            elif name == "ged_layernorm_fwd":
                flops = 4.0 * a[6] * a[7] * 2
            elif name == "ged_layernorm_bwd":  # dx pass (g, x [, g_add] -> dx) + weight/bias pass (g, x)
                flops = 4.0 * a[9] * a[10] * (5 + int(a[5] is not None))
            elif name == "ged_prep_conv_input":   # sources once, bordered tensor written
                C0, h0, w0, C1, Bq, Hq, Wq = a[1], a[2], a[3], a[5], a[7], a[8], a[9]
                flops = 4.0 * Bq * (h0 * w0 * C0 + Hq * Wq * C1 + (Hq + 2) * (Wq + 2) * (C0 + C1))
            elif name == "ged_adamw_step":     # p, g, m, v read; p, m, v written; 1-byte decay mask
                flops = 29.0 * a[5]
            atom = 0.0
            if name == "ged_msda_bwd":
                atom = 32.0 * 4 * 256 * a[12] * a[14] * a[15]          # B * Q * nH rows of 128 corner segments
            records.append((name, s, e, flops, atom))

        step_eager(1)                      # re-warm the eager path (allocator pools differ from the graph's)
        kernels._call = prof_call
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        step_eager(0)
        e1.record()
        torch.cuda.synchronize()
        kernels._call = orig_call
        step_ms = e0.elapsed_time(e1)
        for name, s, e, fl, _atom in records:
            d = prof.setdefault(name, dict(ms=0.0, calls=0, flops=0.0))
            d["ms"] += s.elapsed_time(e)
            d["calls"] += 1
            d["flops"] += fl
        prof["_step_ms_profiled"] = step_ms
    if world > 1:
        dist.barrier()

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
    frames = Bn * world
    value = frames / (ms_step / 1e3)
    e2e_val = frames / (ms_e2e / 1e3)
    h2d = sum(v.numel() * v.element_size() for v in host[0].values())

    # dominant kernel of the step
    kern = {k: v for k, v in prof.items() if not k.startswith("_")}
    native_ms = sum(v["ms"] for v in kern.values())
    dom = max(kern, key=lambda k: kern[k]["ms"])
    tensor_names = ("ged_gemm_tf32", "ged_conv3x3_tf32", "ged_gemm_tf32_bt", "ged_conv3x3_dx_tf32", "ged_gemm_dw_tf32")
    fwd_names = ("ged_gemm_tf32", "ged_conv3x3_tf32")
    if dom in tensor_names:
        tf = kern[dom]["flops"] / (kern[dom]["ms"] / 1e3) / 1e12
        peak = pk["bf16_sustained"] / 2.0
        roof = dict(kernel=dom, bound="tensor", achieved=tf, peak=peak, unit="TFLOP/s", frac=tf / peak, traffic=None,
                    peak_note=f"TF32 dense = 1/2 of the {pk['src']} sustained bf16 cuBLAS figure ({pk['bf16_sustained']})",
                    calls_per_step=kern[dom]["calls"], share_of_step=kern[dom]["ms"] / ms_step,
                    mma_passes=args.passes,
                    note="achieved counts ALGORITHMIC flops (2MNK); the 3xTF32 split issues 3 tcgen05.mma per k-step")
    else:
        # gather/atomics-bound kernel (MSDA): algorithmic bytes = 32 points x 4 corners x 256 B per (query, head),
        # served by L1/L2, not HBM - reported against the HBM peak for scale only
        gbs = kern[dom]["flops"] / (kern[dom]["ms"] / 1e3) / 1e9 if kern[dom]["flops"] else None
        l2_atomic = None
        if dom == "ged_msda_bwd":
            # what actually bounds it: L2 atomic throughput.  Payload = 32 points x 4 corners x 256 B per (query, head)
            # of every MSDA call of the step, against the measured rate of the bare scatter pattern.
            peak_atom = kernels.msda_atomic_probe(device=dev)
            payload = sum(r[4] for r in records if r[0] == "ged_msda_bwd")
            ach = payload / (kern[dom]["ms"] / 1e3) / 1e9
            l2_atomic = dict(achieved=ach, peak=peak_atom, unit="GB/s of red.global.add.v4.f32 payload", frac=ach / peak_atom,
                             peak_src="ged_msda_atomic_probe (same 256-byte row scatter, no gathers), measured in this run",
                             note="the fused kernel also gathers the same rows for the offset / weight gradients")
        roof = dict(kernel=dom, bound="hbm", achieved=gbs, peak=pk["hbm"], unit="GB/s",
                    frac=(gbs / pk["hbm"]) if gbs else None, traffic=ncu_traffic.get(dom), l2_atomic=l2_atomic,
                    calls_per_step=kern[dom]["calls"], share_of_step=kern[dom]["ms"] / ms_step,
                    note="achieved = COMPULSORY HBM bytes (each operand once) / time. The deformable-attention "
                         "kernels are bound by the L1 gather of 32x4 corner segments per (query, head) and by L2 "
                         "atomics, not by HBM (ncu: l1tex 81 %, lts 70 %, DRAM 2 %; profiles/r01b_ncu_msda.csv)")
    # second view: all tcgen05 launches of the step together (the tensor-bound share)
    tens = [kern[k] for k in tensor_names if k in kern]
    tens_ms = sum(t["ms"] for t in tens)
    tens_tf = sum(t["flops"] for t in tens) / (tens_ms / 1e3) / 1e12 if tens_ms else None
    tf32_peak = pk["bf16_sustained"] / 2.0

    def group(names, passes):
        ks = [kern[k] for k in names if k in kern]
        ms = sum(t["ms"] for t in ks)
        if not ms:
            return None
        alg = sum(t["flops"] for t in ks) / (ms / 1e3) / 1e12
        return dict(kernels=[k for k in names if k in kern], ms=round(ms, 3), algorithmic_tflops=alg, mma_passes=passes,
                    issued_mma_tflops=alg * passes, tensor_pipe_frac=alg * passes / tf32_peak)
    roof_tensor = dict(kernels=list(tensor_names), bound="tensor", achieved=tens_tf, peak=tf32_peak,
                       unit="TFLOP/s", frac=(tens_tf / tf32_peak) if tens_tf else None,
                       share_of_step=tens_ms / ms_step,
                       forward=group(fwd_names, args.passes),
                       backward=group(tuple(k for k in tensor_names if k not in fwd_names), kernels.BACKWARD_PASSES),
                       note="achieved/frac: ALGORITHMIC flops (2MNK) of every tcgen05 launch of the step over their summed "
                            "device time, against TF32 dense = 1/2 of the measured sustained bf16 figure.  The forward is "
                            "error-compensated 3xTF32 (3 tcgen05.mma per k-step): issued_mma_tflops / tensor_pipe_frac "
                            "count those; the backward (dX, dW) is single-pass TF32")

    # third view: the streaming (HBM-bound) helper kernels of the step against the measured copy peak
    hbm_kernels = {}
    for k in ("ged_act_bwd", "ged_bn_train_fwd", "ged_bn_train_bwd", "ged_layernorm_fwd", "ged_layernorm_bwd",
              "ged_prep_conv_input", "ged_adamw_step"):
        if k in kern and kern[k]["flops"] > 0:
            gb = kern[k]["flops"] / (kern[k]["ms"] / 1e3) / 1e9
            hbm_kernels[k] = dict(ms=round(kern[k]["ms"], 3), calls=kern[k]["calls"], achieved_gbs=round(gb, 1),
                                  frac=round(gb / pk["hbm"], 3))

    # ---- the kernel the metric names: ground embedding, HBM roofline ---------------------------------
    def ge_bw(Bx, Hx, Wx, reps=20):
        img = torch.randn(Bx, 5, Hx, Wx, device=dev)
        yh = torch.rand(Bx, 1, Hx // 2, Wx // 2, device=dev)
        flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
        with torch.no_grad():
            for _ in range(3):
                kernels.ge_vanilla(img, yh)
            tot = 0.0
            for _ in range(reps):
                flush.zero_()                         # evict L2 (126 MB) between iterations
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                kernels.ge_vanilla(img, yh)
                e.record()
                torch.cuda.synchronize()
                tot += s.elapsed_time(e)
        t = tot / reps / 1e3
        gb = 13.0 * Bx * Hx * Wx / 1e9                # 13 B per full-resolution pixel
        return gb / t, t * 1e6
    bw_work, us_work = ge_bw(Bn, H, W)
    bw_big, us_big = ge_bw(32, 1024, 2048, reps=5)
    ge = dict(kernel="ge_vanilla_fwd_kernel", bound="hbm", unit="GB/s", peak=pk["hbm"], peak_src=pk["src"],
              bytes_per_pixel=13, at_workload=dict(shape=[Bn, H, W], achieved=bw_work, frac=bw_work / pk["hbm"], us=us_work,
                                                   note="41 MB: launch-latency/L2 dominated"),
              at_sweep_max=dict(shape=[32, 1024, 2048], achieved=bw_big, frac=bw_big / pk["hbm"], us=us_big),
              traffic=ncu_traffic.get("ge_vanilla_fwd_kernel"), l2_flushed_between_iterations=True)

    # ---- Swin window attention (QKV GEMM -> 49x49 core -> proj GEMM), the north star's tensor-pipe target -----
    def swin_attention(tokens_hw, Cc, nH, shift):
        hh, ww = tokens_hw
        T = Bn * hh * ww
        x = torch.randn(Bn, hh * ww, Cc, device=dev)
        wq, bq = torch.randn(3 * Cc, Cc, device=dev) / Cc ** 0.5, torch.randn(3 * Cc, device=dev) * 0.02
        wp, bp = torch.randn(Cc, Cc, device=dev) / Cc ** 0.5, torch.randn(Cc, device=dev) * 0.02
        table = torch.randn(169, nH, device=dev) * 0.2
        from gedepth_b200.swin import WindowMSA
        index = WindowMSA(Cc, nH, (7, 7)).relative_position_index.to(dev)
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
        nW = (-(-hh // 7)) * (-(-ww // 7)) * Bn
        f_gemm = 2.0 * T * Cc * 3 * Cc + 2.0 * T * Cc * Cc
        f_core = 2.0 * nW * nH * 49 * 49 * 32 * 2
        ms = sum(t)
        return dict(shape=dict(tokens=[Bn, hh, ww], C=Cc, heads=nH, shift=shift), ms=dict(qkv=t[0], core=t[1], proj=t[2]),
                    algorithmic_tflops=(f_gemm + f_core) / ms / 1e9,
                    issued_mma_tflops=args.passes * f_gemm / ms / 1e9,
                    tensor_pipe_frac=args.passes * f_gemm / ms / 1e9 / (pk["bf16_sustained"] / 2.0),
                    note="QKV and proj run on tcgen05 (3xTF32 when passes=3); the 49x49 core is a SIMT fp32 kernel, so its "
                         "flops do not count towards the tensor pipe")
    e0 = 96 if args.backbone == "swin_t" else 192
    h0 = 3 if args.backbone == "swin_t" else 6
    swin_attn = [swin_attention((-(-H // 4), -(-W // 4)), e0, h0, 3),
                 swin_attention((-(-H // 16), -(-W // 16)), 4 * e0, 4 * h0, 0)]

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        t = cpu_step_time(torch, 1, budget_s=60.0, steps=1, warmup=0)
        cpu = dict(value=1.0 / t[0], unit="frames/s", cores=cores, kind="port",
                   sample="1 frame 352x1120 fwd+bwd through the oracle port (plain PyTorch CPU), all host threads",
                   ground_plane_numpy_mpx_s=ground_plane_numpy_rate())

    line = dict(metric="frames/sec (352x1120) fwd+bwd", value=value, unit="frames/s", n_gpus=world, steps=args.steps,
                warmup=max(3, args.warmup), ms_per_step=ms_step, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype=("f32 (fp32 storage and accumulation; forward GEMMs/convs error-compensated 3xTF32 on tcgen05, backward GEMMs "
                       "single-pass TF32 = what the reference's PyTorch 1.8 runs on Ampere+)") if args.passes == 3
                else "tf32 (single-pass TF32 forward and backward, fp32 storage and accumulation)", data="synthetic",
                config=dict(workload=workload, global_batch=frames, parallelism=f"dp{world}",
                            step="fwd + SiLog + bwd + allreduce(N>1) + clip + AdamW", drop_path_rate=0.3,
                            launch="one captured CUDA graph per step" if use_graph else "eager",
                            l2="inputs + activations per step (>2 GB) exceed the 126 MB L2"),
                e2e=dict(value=e2e_val, unit="frames/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=4, ms_per_step=ms_e2e),
                gpu_launches=launches, clocks=clk, roofline=roof, roofline_tensor=roof_tensor, ground_embed=ge,
                swin_window_attention=swin_attn, roofline_hbm_kernels=hbm_kernels,
                cpu_baseline=cpu,
                native_ops=ops.native_table(),
                kernel_ms={k: round(v["ms"], 3) for k, v in sorted(kern.items(), key=lambda kv: -kv[1]["ms"])},
                native_ms_of_step=[round(native_ms, 2), round(prof["_step_ms_profiled"], 2)])
    print(json.dumps(line), flush=True)
    finish()


if __name__ == "__main__":
    main()
