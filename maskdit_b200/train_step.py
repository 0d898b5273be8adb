"""Data-parallel training step on the B200 engine (reference: train.py:200-230 inner step, :178 DDP, :141 optimizer,
train_utils/helper.py:47-58 EMA).

One process per GPU.  A step is:
    zero flat grad  ->  fused EDM loss forward/backward (engine)  ->  NCCL all-reduce (SUM) of the flat fp32 gradient
    buffer over NVLink (the 1/world factor is folded into the optimizer kernel)  ->  fused AdamW + EMA + bf16-shadow
    kernel over the flat buffers.
With `overlap=True` the flat buffer is reduced and stepped in per-block contiguous ranges on a side stream
while the backward of the earlier blocks is still running (the role DDP's bucketed hooks play in the reference);
with `overlap=False` (default: measured equal or faster up to 2 GPUs, see profiles/README.md) it is literally one
all-reduce and one optimizer launch.  No other collective is issued in the
step; the loss is returned as a device tensor (no per-step `.item()` host sync as at train.py:227).
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist

from . import ops
from .loss import EDMLoss
from .maskdit import EDMPrecond


class DataParallelB200:
    """What the reference's loss expects from the DDP wrapper: `.module`, `.training`, callable (loss.py:41,47,52).
    Gradient synchronisation is NOT hooked into autograd; `TrainStep` all-reduces the flat gradient buffer once."""

    def __init__(self, module: EDMPrecond):
        self.module = module

    @property
    def training(self):
        return self.module.training

    def train(self, mode=True):
        self.module.train(mode)
        return self

    def eval(self):
        return self.train(False)

    def parameters(self):
        return self.module.parameters()

    def __call__(self, *a, **k):
        return self.module(*a, **k)


def shard_batch(global_batch: int, world_size: int, rank: int):
    """Even batch split by rank (train.py:72-75: global = per-GPU batch x world)."""
    if global_batch % world_size:
        raise ValueError(f"global batch {global_batch} not divisible by world size {world_size}")
    per = global_batch // world_size
    return rank * per, (rank + 1) * per


def lr_at(step: int, base_lr: float, global_batch: int, rampup_kimg: float):
    """train.py:223 — note lr = 0 at step 0 when rampup_kimg > 0; with rampup 0 it is base_lr from step 1 on."""
    return base_lr * min(step * global_batch / max(rampup_kimg * 1000, 1e-8), 1)


def ar_chunk_bounds(n, k):
    """[lo, hi) element ranges of the k all-reduce chunks of a flat buffer of n elements (4 KiB aligned starts)."""
    if k <= 1 or n < k * 1024:
        return [(0, n)]
    step = -(-(-(-n // k)) // 1024) * 1024
    return [(lo, min(n, lo + step)) for lo in range(0, n, step)]


class TrainStep:
    def __init__(self, net: EDMPrecond, ema: EDMPrecond | None = None, lr=1e-4, betas=(0.9, 0.999), eps=1e-8,
                 weight_decay=0.0, ema_decay=0.9999, loss_fn: EDMLoss | None = None, process_group=None,
                 lr_rampup_kimg=0.0, global_batch=None, device=None, overlap=False, graph=None):
        self.net, self.ema = net, ema
        self.lr, self.betas, self.eps, self.wd, self.ema_decay = lr, betas, eps, weight_decay, ema_decay
        self.loss_fn = loss_fn or EDMLoss()
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        self.rampup, self.global_batch = lr_rampup_kimg, global_batch
        self.step_count = 0
        dev = device or next(net.parameters()).device
        self.st = net.prepare(dev)
        self.st.ensure_grad()
        n = self.st.n_train
        self.m = torch.zeros(n, dtype=torch.float32, device=dev)
        self.v = torch.zeros(n, dtype=torch.float32, device=dev)
        self.ema_st = None
        if ema is not None:
            self.ema_st = ema.prepare(dev)
            assert self.ema_st.n_train == n and self.ema_st.offsets == self.st.offsets
        for k, p in net.named_parameters():  # .grad views into the flat buffer (optimizer-compatible)
            if p.requires_grad:
                p.grad = self.st.gview(k)
        # Overlap: gradient ranges are reduced + stepped on a side stream as soon as a block's backward is enqueued.
        self.overlap = overlap
        self.graph = (os.environ.get("MDT_TRAIN_GRAPH", "0") == "1") if graph is None else bool(graph)
        self._graphs = {}
        self.bg_blocks = 24
        # gradient all-reduce in this many chunks, pipelined against the optimizer pass (world > 1).  Default 1 = one
        # flat call: measured on 2 x B200 (same box) 132.0 ms/step flat vs 133.3 ms with 8 chunks - the all-reduce and
        # the AdamW/EMA pass are both HBM-bound, so running them side by side buys nothing.
        self.ar_chunks = int(os.environ.get("MDT_AR_CHUNKS", "1"))
        self.side = torch.cuda.Stream(device=dev) if overlap else None
        self._done = []          # [lo, hi) ranges already handled in the current step
        self._lr_now = lr
        net._grad_ready_hook = self._on_grads_ready if overlap else None

    # -- optimizer state for checkpoints (reference: train.py:259-270 stores optimizer.state_dict() under 'opt') ------
    def state_dict(self):
        """Per-parameter AdamW state in named_parameters order, laid out like torch.optim.AdamW.state_dict()['state']
        (exp_avg / exp_avg_sq / step), plus the hyper-parameters; tensors are copies on the current device."""
        state, names = {}, []
        for i, (k, p) in enumerate((k, p) for k, p in self.net.named_parameters() if p.requires_grad):
            lo, _, shape = self.st.offsets[k]
            n = p.numel()
            state[i] = {"step": torch.tensor(float(self.step_count)), "exp_avg": self.m[lo:lo + n].view(shape).clone(),
                        "exp_avg_sq": self.v[lo:lo + n].view(shape).clone()}
            names.append(k)
        return {"state": state, "param_names": names,
                "param_groups": [{"lr": self.lr, "betas": self.betas, "eps": self.eps, "weight_decay": self.wd,
                                  "params": list(range(len(names)))}]}

    def load_state_dict(self, sd):
        names = sd.get("param_names") or [k for k, p in self.net.named_parameters() if p.requires_grad]
        if len(names) != len(sd["state"]):
            raise ValueError(f"optimizer state holds {len(sd['state'])} tensors, the model has {len(names)}")
        for i, k in enumerate(names):
            e = sd["state"][i] if i in sd["state"] else sd["state"][str(i)]
            lo, _, shape = self.st.offsets[k]
            n = e["exp_avg"].numel()
            if tuple(e["exp_avg"].shape) != tuple(shape):
                raise ValueError(f"optimizer state of {k}: shape {tuple(e['exp_avg'].shape)} != {tuple(shape)}")
            self.m[lo:lo + n].copy_(e["exp_avg"].reshape(-1))
            self.v[lo:lo + n].copy_(e["exp_avg_sq"].reshape(-1))
            self.step_count = int(float(e["step"]))
        g = sd["param_groups"][0]
        self.lr, self.betas, self.eps, self.wd = g["lr"], tuple(g["betas"]), g["eps"], g["weight_decay"]

    # -- one gradient range: (all-reduce) + fused AdamW/EMA/bf16-shadow, on the current stream -----------------------
    def _reduce_and_step(self, lo, hi, max_blocks=0):
        st, n = self.st, hi - lo
        if n <= 0:
            return
        if self.world > 1:
            dist.all_reduce(st.grad[lo:hi], op=dist.ReduceOp.SUM, group=self.pg)
        self._step_range(lo, hi, max_blocks)

    def _step_range(self, lo, hi, max_blocks=0):
        st, n = self.st, hi - lo
        if n <= 0:
            return
        g = st.grad[lo:hi]
        ops.adamw_ema(st.w32[lo:hi], g, self.m[lo:hi], self.v[lo:hi],
                      self.ema_st.w32[lo:hi] if self.ema_st is not None else None, st.w16[lo:hi], n, self._lr_now,
                      self.step_count, self.betas[0], self.betas[1], self.eps, self.wd, self.ema_decay,
                      1.0 / self.world, max_blocks)

    def _on_grads_ready(self, lo, hi):
        main = torch.cuda.current_stream()
        self.side.wait_stream(main)
        with torch.cuda.stream(self.side):
            # background launch: a block's optimizer pass needs <2 % of the HBM bandwidth to finish before the backward
            # does, so it is capped to a few CTAs and leaves the SMs to the tensor-core GEMMs
            self._reduce_and_step(lo, hi, max_blocks=self.bg_blocks)
        self._done.append((lo, hi))

    def _fwd_bwd_graphed(self, images, labels, mask_ratio, mae_loss_coef):
        """Gradient zeroing + loss forward + engine backward (~770 launches, 70 ms of host time) replayed from a CUDA
        graph captured once per (shapes, mask_ratio, mae_loss_coef); the all-reduce and the optimizer pass stay eager
        (their scalars change every step).  Opt-in (`TrainStep(graph=True)` / MDT_TRAIN_GRAPH=1): written at the end of
        round 1 and NOT yet run on hardware."""
        key = (tuple(images.shape), tuple(labels.shape), float(mask_ratio), float(mae_loss_coef))
        ent = self._graphs.get(key)
        if ent is None:
            gx, gy = images.clone(), labels.clone()

            def body():
                self.st.grad.zero_()
                loss = self.loss_fn(self.net, gx, gy, mask_ratio=mask_ratio, mae_loss_coef=mae_loss_coef)
                loss.mean().backward()
                return loss.detach()

            cur = torch.cuda.current_stream()
            side = torch.cuda.Stream()
            side.wait_stream(cur)
            with torch.cuda.stream(side):   # warm-up outside the capture (lazy kernel attributes, allocator pools)
                body()
            cur.wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            n0 = ops.L.LAUNCHES
            with torch.cuda.graph(graph):
                out = body()
            ent = (graph, gx, gy, out, ops.L.LAUNCHES - n0)
            self._graphs[key] = ent
        graph, gx, gy, out, n_launch = ent
        gx.copy_(images), gy.copy_(labels)
        graph.replay()
        ops.L.LAUNCHES += n_launch
        return out.clone()

    def step(self, images, labels, mask_ratio=0.5, mae_loss_coef=0.1):
        """One optimisation step on this rank's shard.  Returns the per-sample loss [B] (device tensor)."""
        st = self.st
        self.step_count += 1
        gb = self.global_batch or images.shape[0] * self.world
        self._lr_now = lr_at(self.step_count, self.lr, gb, self.rampup) if self.rampup > 0 else self.lr
        self._done = []
        if self.graph and not self.overlap and ops.L.GEMM_PROFILE is None:
            loss = self._fwd_bwd_graphed(images, labels, mask_ratio, mae_loss_coef)
        else:
            st.grad.zero_()
            loss = self.loss_fn(self.net, images, labels, mask_ratio=mask_ratio, mae_loss_coef=mae_loss_coef)
            loss.mean().backward()   # engine backward; with overlap=True block ranges are already being reduced/stepped
        main = torch.cuda.current_stream()
        if self.overlap:
            self.side.wait_stream(main)
            with torch.cuda.stream(self.side):
                cur = 0
                for lo, hi in sorted(self._done) + [(st.n_train, st.n_train)]:   # the complement of the block ranges
                    self._reduce_and_step(cur, lo)
                    cur = max(cur, hi)
            main.wait_stream(self.side)
        elif self.world > 1 and self.ar_chunks > 1:
            # The all-reduce is exposed after the backward (overlapping it with the persistent GEMMs costs more than
            # it hides), but it need not also serialise with the optimizer: the flat gradient is reduced in chunks on
            # NCCL's stream and the fused AdamW/EMA pass of chunk k runs while chunk k+1 is still on the wire.
            bounds = ar_chunk_bounds(st.n_train, self.ar_chunks)
            works = [dist.all_reduce(st.grad[lo:hi], op=dist.ReduceOp.SUM, group=self.pg, async_op=True)
                     for lo, hi in bounds]
            for (lo, hi), w in zip(bounds, works):
                w.wait()   # the current stream waits for this chunk only
                self._step_range(lo, hi)
        else:
            self._reduce_and_step(0, st.n_train)   # one flat all-reduce + one optimizer pass
        st.mark_shadow_fresh(self.net._params())   # the kernel refreshed the bf16 shadow itself
        if self.ema_st is not None:
            self.ema_st._versions = None           # EMA weights changed behind PyTorch's back: shadow is stale
        return loss.detach()
