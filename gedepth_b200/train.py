"""Training step of the path: forward + SiLog(/CE) + backward + ONE NCCL all-reduce over a flat fp32
gradient arena + fused clip/AdamW.

Replaces, for this path, the reference's runner plumbing around ``train_step``:
MMDistributedDataParallel bucketed all-reduce (depth/apis/train.py:58-67), OptimizerHook grad-clip
(configs/depthformer/depthformer_v.py:148) and AdamW with paramwise decay_mult
(depthformer_v.py:128-139, mmcv DefaultOptimizerConstructor: custom_keys match as substrings).
One process per GPU; batch sharded only (no collective in the forward; BatchNorm statistics stay
per-GPU exactly as in the reference).
"""
from __future__ import annotations

import os
from typing import Dict, Optional

import torch
import torch.distributed as dist

from . import kernels

NO_DECAY_KEYS = ("absolute_pos_embed", "relative_position_bias_table", "norm")


def cosine_warmup_lr(step: int, base_lr: float = 1e-4, max_iters: int = 1600 * 48, warmup_iters: int = 16 * 1600,
                     warmup_ratio: float = 1e-3, min_lr_ratio: float = 1e-8) -> float:
    """Learning rate of 0-based iteration `step` under the GE configs' lr_config (depthformer_v.py:141-147):
    mmcv CosineAnnealingLrUpdaterHook(by_epoch=False) with linear warm-up [external, mmcv 1.3.x]:
    regular = target + (base - target) (1 + cos(pi t / T)) / 2, target = base * min_lr_ratio; during warm-up
    lr = regular * (1 - (1 - t / warmup_iters) (1 - warmup_ratio))."""
    import math
    target = base_lr * min_lr_ratio
    lr = target + 0.5 * (base_lr - target) * (1.0 + math.cos(math.pi * min(step, max_iters) / max_iters))
    if step < warmup_iters:
        lr *= 1.0 - (1.0 - step / warmup_iters) * (1.0 - warmup_ratio)
    return lr


def completion_group(name: str) -> int:
    """Order in which the gradients of the path's parameters become FINAL during the backward (it runs head -> PE necks
    -> neck -> Swin stage 3 -> 2 -> 1 -> 0 / patch embedding / stem): 0 decode head + PE necks, 1 HAHI neck, 2..4 Swin
    stages 3..1 (with their output LayerNorms), 5 the rest of the backbone.  The arena keeps each group contiguous so
    that a group can be all-reduced as soon as the backward has passed it."""
    parts = name.split(".")
    if parts[0] == "neck":
        return 1
    if parts[0] == "backbone":
        if parts[1] == "stages" and parts[2].isdigit() and int(parts[2]) in (1, 2, 3):
            return 2 + (3 - int(parts[2]))
        if parts[1].startswith("norm") and parts[1][4:].isdigit() and int(parts[1][4:]) in (1, 2, 3):
            return 2 + (3 - int(parts[1][4:]))
        return 5
    return 0


class FlatArena:
    """All parameters (and their gradients) of a model as views into two contiguous fp32 buffers, laid out group by
    group in the order their gradients complete (``completion_group``); ``group_ranges`` lists (group, start, end)."""

    def __init__(self, model: torch.nn.Module):
        params = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
        params = [np_ for _, np_ in sorted(enumerate(params), key=lambda t: (completion_group(t[1][0]), t[0]))]
        self.names = [n for n, _ in params]
        self.params = [p for _, p in params]
        dev = self.params[0].device
        sizes = [((p.numel() + 3) // 4) * 4 for p in self.params]        # 16-byte aligned segments
        self.total = sum(sizes)
        self.flat_p = torch.zeros(self.total, dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros(self.total, dtype=torch.float32, device=dev)
        self.wd_mask = torch.ones(self.total, dtype=torch.uint8, device=dev)
        off = 0
        self.group_ranges = []
        for (n, p), sz in zip(params, sizes):
            gidx = completion_group(n)
            if self.group_ranges and self.group_ranges[-1][0] == gidx:
                self.group_ranges[-1][2] = off + sz
            else:
                self.group_ranges.append([gidx, off, off + sz])
            view = self._segment(self.flat_p, off, p)
            view.copy_(p.data)
            p.data = view
            p.grad = self._segment(self.flat_g, off, p)
            p._ged_sink = True            # backward kernels accumulate straight into p.grad (kernels._sink)
            if any(k in n for k in NO_DECAY_KEYS):
                self.wd_mask[off:off + sz] = 0
            off += sz

    @staticmethod
    def _segment(flat, off, p):
        """View of one parameter inside a flat buffer.  4-D conv weights are kept channels-last in memory
        ([Cout][kh][kw][Cin] - the layout the implicit-GEMM conv and its weight-gradient kernel use), exposed with
        the logical (Cout, Cin, kh, kw) shape, so state_dict() shapes are the reference's."""
        seg = flat[off:off + p.numel()]
        if p.dim() == 4:
            co, ci, kh, kw = p.shape
            return seg.view(co, kh, kw, ci).permute(0, 3, 1, 2)
        return seg.view_as(p)

    def zero_grad(self):
        self.flat_g.zero_()


class Trainer:
    def __init__(self, model, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01, max_norm=35.0,
                 native_optimizer: bool = True, lr_schedule=None, overlap_allreduce: bool = True):
        """lr_schedule: None (constant lr) or a callable step -> lr (e.g. ``cosine_warmup_lr``).  The step's learning
        rate lives in a device scalar that is refreshed before every step, eager or replayed, so a captured CUDA graph
        follows the schedule (and ``trainer.lr = x`` takes effect on the next step)."""
        self.model = model
        self.lr_schedule = lr_schedule
        self.arena = FlatArena(model)
        self.m = torch.zeros_like(self.arena.flat_p)
        self.v = torch.zeros_like(self.arena.flat_p)
        self.sumsq = torch.zeros(1, dtype=torch.float64, device=self.arena.flat_p.device)
        self.lr, self.betas, self.eps, self.wd, self.max_norm = lr, betas, eps, weight_decay, max_norm
        self.step_idx = 0
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=self.arena.flat_p.device)
        self.lr_dev = torch.full((1,), float(lr), dtype=torch.float32, device=self.arena.flat_p.device)
        self._lr_host = torch.full((1,), float(lr), dtype=torch.float32)
        if self.arena.flat_p.is_cuda:
            self._lr_host = self._lr_host.pin_memory()
        kernels.RNG_STEP = self.step_dev     # epilogue dropout mixes the device step counter into its seed
        # optional: weight-gradient GEMMs on a side stream, overlapping the rest of the backward (A/B switch)
        self._dw_side = None
        if os.environ.get("GEDEPTH_DW_STREAM", "0") == "1" and self.arena.flat_p.is_cuda:
            self._dw_side = dict(stream=torch.cuda.Stream(), keep=[])
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        # gradient all-reduce sliced by completion group and launched from autograd hooks while the backward is still
        # running (replaces the bucketed overlap of MMDistributedDataParallel, depth/apis/train.py:58-67)
        self.overlap = bool(overlap_allreduce) and self.world > 1 and os.environ.get("GEDEPTH_OVERLAP_ALLREDUCE", "1") != "0"
        self._works, self._launched = [], set()
        if self.overlap:
            self._install_overlap_hooks()
        self._graph = None
        self._static = None
        self._static_loss = None

    # ---- overlapped all-reduce ---------------------------------------------------------------------------------------
    def _reduce_group(self, gidx: int):
        """All-reduce the arena slice of one completion group on NCCL's stream (async): everything the backward has
        launched so far is ordered before it; the optimizer waits for all slices."""
        if gidx in self._launched:
            return
        self._launched.add(gidx)
        for g, s_, e_ in self.arena.group_ranges:
            if g == gidx:
                self._works.append(dist.all_reduce(self.arena.flat_g[s_:e_], async_op=True))

    def _watch(self, tensors, gidx: int):
        """When every tensor of `tensors` has received its gradient, the groups up to `gidx` are final."""
        if not (self.overlap and torch.is_grad_enabled()):
            return
        ts = [t for t in tensors if torch.is_tensor(t) and t.requires_grad]
        if not ts:
            return
        state = {"left": len(ts)}

        def hook(grad):
            state["left"] -= 1
            if state["left"] == 0:
                for g in range(gidx + 1):
                    self._reduce_group(g)
            return None
        for t in ts:
            t.register_hook(hook)

    def _install_overlap_hooks(self):
        m = self.model
        if not (hasattr(m, "neck") and hasattr(m, "backbone")):
            self.overlap = False
            return
        m.neck.register_forward_hook(lambda mod, inp, out: self._watch(out, 0))            # head + PE necks done
        m.backbone.register_forward_hook(lambda mod, inp, out: self._watch(out, 1))        # + neck done
        stages = getattr(m.backbone, "stages", None)
        if stages is not None and len(stages) == 4:
            for i in (3, 2, 1):                                                            # + Swin stage i done
                stages[i].register_forward_pre_hook(lambda mod, inp, i=i: self._watch(inp[:1], 2 + (3 - i)))

    def _finish_allreduce(self):
        if self.world <= 1:
            return
        if not self.overlap:
            dist.all_reduce(self.arena.flat_g)            # ONE collective per step, NCCL over NVLink
            return
        for g in sorted({r[0] for r in self.arena.group_ranges}):
            self._reduce_group(g)                          # whatever the hooks have not launched yet (at least group 5)
        for w in self._works:
            w.wait()
        self._works, self._launched = [], set()

    def _refresh_lr(self):
        """Host -> device copy of this step's learning rate (a memcpy on the current stream, never captured)."""
        if self.lr_schedule is not None:
            self.lr = float(self.lr_schedule(self.step_idx))
        if float(self._lr_host[0]) != float(self.lr):
            self._lr_host[0] = float(self.lr)
            self.lr_dev.copy_(self._lr_host, non_blocking=True)

    def step(self, data_batch: Dict, sync_logs: bool = False):
        """One optimisation step.  Returns the loss tensor (device) and, if ``sync_logs``, the
        rank-averaged log_vars of the reference's ``_parse_losses`` (one device->host copy)."""
        if not torch.cuda.is_current_stream_capturing() if self.arena.flat_p.is_cuda else True:
            self._refresh_lr()
        self.arena.zero_grad()
        self._works, self._launched = [], set()
        losses = self.model(**data_batch)
        loss, log_vars = self.model._parse_losses(losses, sync=sync_logs)
        kernels.DW_SIDE = self._dw_side
        try:
            loss.backward()
        finally:
            kernels.DW_SIDE = None
        if self._dw_side is not None:               # join: every dW has landed in the arena before it is reduced / read
            torch.cuda.current_stream().wait_stream(self._dw_side["stream"])
            self._dw_side["keep"].clear()
        self.early_groups = sorted(self._launched)      # groups whose all-reduce was launched from inside the backward
        self._finish_allreduce()
        self.step_idx += 1
        self.step_dev.add_(1)
        kernels.sumsq(self.arena.flat_g, self.sumsq)
        kernels.adamw_step(self.arena.flat_p, self.arena.flat_g, self.m, self.v, self.arena.wd_mask, self.sumsq,
                           self.max_norm, 1.0 / self.world, self.lr, self.betas[0], self.betas[1], self.eps,
                           self.wd, self.step_idx, self.step_dev, self.lr_dev)
        return loss, log_vars

    # ---- CUDA-graph path: the whole step (forward, loss, backward, all-reduce, clip, AdamW) is captured
    # once and replayed; ~1400 kernel launches per step stop costing CPU time ------------------------
    def capture(self, example_batch: Dict, warmup: int = 3):
        """Capture ``step`` on static copies of ``example_batch``'s tensors.  Shapes are then fixed."""
        self._static = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in example_batch.items()}
        # the warm-up below runs REAL optimisation steps: snapshot everything they mutate and put it back afterwards, so
        # that capture() leaves parameters, Adam moments, BatchNorm statistics and the step counter untouched
        snap = [t.clone() for t in (self.arena.flat_p, self.m, self.v, self.step_dev)]
        bufs = [b for b in self.model.buffers()]
        snap_bufs = [b.clone() for b in bufs]
        step0 = self.step_idx
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self.step(self._static)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        for t, s_ in zip((self.arena.flat_p, self.m, self.v, self.step_dev), snap):
            t.copy_(s_)
        for b, s_ in zip(bufs, snap_bufs):
            b.copy_(s_)
        self.step_idx = step0
        del snap, snap_bufs
        torch.cuda.empty_cache()      # the eager warm-up's activation blocks would otherwise sit beside the graph's pool
        self._graph = torch.cuda.CUDAGraph()
        n0 = kernels.LAUNCHES
        with torch.cuda.graph(self._graph):
            loss, _ = self.step(self._static)
            self._static_loss = loss.detach()
        self.launches_per_step = kernels.LAUNCHES - n0
        self.step_idx -= 1            # the capture pass itself does not execute (neither step_dev.add_ nor BN updates ran)
        return self

    def release_graph(self):
        """Drop the captured graph and its private memory pool (shapes may change afterwards)."""
        self._graph = None
        self._static_loss = None
        torch.cuda.synchronize()
        torch.cuda.empty_cache()

    def step_graph(self, data_batch: Optional[Dict] = None):
        """Replay the captured step.  Tensors of ``data_batch`` are copied into the static inputs first
        (host tensors: asynchronous H2D from pinned memory).  Returns the (static) loss tensor."""
        if data_batch is not None:
            for k, v in data_batch.items():
                if torch.is_tensor(v):
                    self._static[k].copy_(v, non_blocking=True)
        self._refresh_lr()
        self._graph.replay()
        self.step_idx += 1
        kernels.LAUNCHES += self.launches_per_step
        return self._static_loss
