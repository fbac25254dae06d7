"""Data-parallel training step with the semantics of the reference's ``train.py:126-133,155-167``:
SGD(lr, momentum 0.9, weight_decay 1e-4); ``strategy='dp'`` with ``sync_batchnorm=False`` => every replica
keeps its own BatchNorm batch statistics and its own loss, and the replicas' gradients are averaged.

B200 form: one process per GPU, parameters / gradients / momentum in three flat fp32 buffers (the
nn.Parameters are views), ONE sum all-reduce of the flat gradient over NCCL (NVLink 5 / NVSwitch) per step,
followed by one fused SGD kernel that applies the 1/world scale.  Forward+backward(+SGD on one GPU) can be
captured into a CUDA graph: the step is ~1.5k small launches otherwise.
"""
import os
from typing import Dict, Optional

import torch
import torch.distributed as dist
from torch import nn


def flatten_parameters(model: nn.Module, device=None):
    """Move all parameters into one flat fp32 buffer (parameters become views); returns (flat_p, flat_g)."""
    params = [p for p in model.parameters()]
    n = sum(p.numel() for p in params)
    device = params[0].device if device is None else device
    flat_p = torch.empty(n, dtype=torch.float32, device=device)
    flat_g = torch.zeros(n, dtype=torch.float32, device=device)
    off = 0
    for p in params:
        k = p.numel()
        flat_p[off:off + k].copy_(p.data.reshape(-1))
        p.data = flat_p[off:off + k].view(p.shape)
        p.grad = flat_g[off:off + k].view(p.shape)
        off += k
    return flat_p, flat_g


def allreduce_mean_(flat_g: torch.Tensor, group=None) -> float:
    """Sum all-reduce in place; returns the scale (1/world) the optimiser must apply (fused into SGD)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat_g, op=dist.ReduceOp.SUM, group=group)
        return 1.0 / dist.get_world_size(group)
    return 1.0


class FusionTrainer:
    def __init__(self, model: nn.Module, criterion, lr: float = 0.1, momentum: float = 0.9,
                 weight_decay: float = 1e-4, group=None, accumulate_grad_batches: int = 1):
        self.model, self.criterion = model, criterion
        self.lr, self.momentum, self.weight_decay = lr, momentum, weight_decay
        self.group = group
        self.accumulate = max(1, int(accumulate_grad_batches))     # train.py:161
        self.flat_p, self.flat_g = flatten_parameters(model)
        self.mom = torch.zeros_like(self.flat_p)
        self.steps = 0
        self.micro = 0
        self._graph = None
        self._static: Optional[Dict[str, torch.Tensor]] = None
        self._static_loss = None
        self._stage = None             # prefetch(): staging buffers + copy stream
        # gradient sink: backward kernels write straight into flat_g (valid while every parameter is used once per
        # backward and flat_g is zeroed once per optimisation step; BatchNorm gradients are plain writes)
        # packed-weight arena: the first forward+backward records every bf16 weight image, later steps regenerate all of
        # them with one kernel instead of one tiny packing launch per conv call (~170 per step)
        self._arena = None
        self._arena_state = 0          # 0 not started, 1 recording, 2 sealed
        self._weights_dirty = False    # set by mark_weights_dirty(): repack before the next forward
        self._sink = {}
        if self.accumulate == 1:
            for p in model.parameters():
                if p.dim() >= 1 and p.grad is not None:
                    self._sink[p.data_ptr()] = p.grad

    # -- pieces ----------------------------------------------------------------------------------------
    def forward_backward(self, batch):
        from . import functional, ops
        from .functional import get_compute_dtype
        if get_compute_dtype() == torch.bfloat16 and not os.environ.get('FFPN_NO_ARENA'):
            if self._arena_state == 0:
                self._arena = torch.empty(96 << 20, dtype=torch.uint8, device=self.flat_p.device)
                ops.weight_arena_begin(self._arena)
                self._arena_state = 1
            elif self._arena_state == 2 and self._weights_dirty:
                ops.weight_arena_pack(self.flat_p)     # weights were changed behind the trainer's back (load_state_dict, ...)
                self._weights_dirty = False
        functional.set_grad_sink(self._sink)
        try:
            out = self.model(batch)
            loss, _ = self.criterion(batch, out)
            (loss / self.accumulate if self.accumulate > 1 else loss).backward()
            functional.join_side_streams()     # the sink is written from every branch stream
        finally:
            functional.set_grad_sink(None)
            if self._arena_state == 1:
                ops.weight_arena_seal(self.flat_p.device.index if self.flat_p.device.index is not None else 0)
                self._arena_state = 2
        return loss.detach()

    def mark_weights_dirty(self):
        """Call after modifying parameters outside optimizer_step() (e.g. load_state_dict): the bf16 weight images of the
        arena are regenerated before the next forward."""
        self._weights_dirty = True

    def close(self):
        """Detach the packed-weight arena from the library context (the context is shared per device)."""
        if self._arena_state:
            from . import ops
            ops.weight_arena_end(self.flat_p.device.index if self.flat_p.device.index is not None else 0)
            self._arena_state, self._arena = 0, None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def optimizer_step(self):
        from . import ops
        scale = allreduce_mean_(self.flat_g, self.group)
        ops.sgd_step(self.flat_p, self.flat_g, self.mom, self.lr, self.momentum, self.weight_decay, scale,
                     self.steps == 0)
        if self._arena_state == 2:
            ops.weight_arena_pack(self.flat_p)         # keep the arena images in step with the weights (eval forwards use them)
        self.steps += 1
        self.flat_g.zero_()

    def step(self, batch):
        """Eager training step (forward, loss, backward, all-reduce, SGD)."""
        loss = self.forward_backward(batch)
        self.micro += 1
        if self.micro % self.accumulate == 0:
            self.optimizer_step()
        return loss

    # -- CUDA graph --------------------------------------------------------------------------------------
    def capture(self, example_batch, warmup: int = 3, include_optimizer: Optional[bool] = None):
        """Capture forward+loss+backward (and SGD when there is no collective) on static input buffers."""
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        self._fused_opt = (world == 1 and self.accumulate == 1) if include_optimizer is None else include_optimizer
        self._static = {k: v.clone() for k, v in example_batch.items()}
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self.forward_backward(self._static)
                self.flat_g.zero_()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        first = self.steps == 0
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self._static_loss = self.forward_backward(self._static)
            if self._fused_opt:
                from . import ops
                ops.sgd_step(self.flat_p, self.flat_g, self.mom, self.lr, self.momentum, self.weight_decay, 1.0, False)
                if self._arena_state == 2:
                    ops.weight_arena_pack(self.flat_p)
                self.flat_g.zero_()
        self._first_graph_step = first
        return self

    def prefetch(self, batch):
        """Start the host->device copy of the NEXT step's batch (pinned host tensors) on a copy stream, into staging
        buffers; it overlaps the step that is running.  ``replay(prefetched=True)`` consumes it."""
        if self._stage is None:
            self._stage = {k: torch.empty_like(v) for k, v in self._static.items()}
            self._copy_stream = torch.cuda.Stream()
            self._staged, self._consumed = torch.cuda.Event(), torch.cuda.Event()
            self._consumed.record()
        self._copy_stream.wait_event(self._consumed)       # the previous staging -> static copy has read the buffers
        with torch.cuda.stream(self._copy_stream):
            for k, v in batch.items():
                if k in self._stage:
                    self._stage[k].copy_(v, non_blocking=True)
            self._staged.record()

    def replay(self, batch=None, prefetched: bool = False):
        """One captured step; ``batch`` (device or pinned host tensors) is copied into the static buffers first, or
        (``prefetched``) the batch staged by ``prefetch()`` is."""
        if prefetched:
            cur = torch.cuda.current_stream()
            cur.wait_event(self._staged)
            for k, v in self._stage.items():
                self._static[k].copy_(v, non_blocking=True)
            self._consumed.record()
        elif batch is not None:
            for k, v in batch.items():
                if k in self._static:
                    self._static[k].copy_(v, non_blocking=True)
        self._graph.replay()
        if self._fused_opt:
            self.steps += 1
        else:
            self.micro += 1
            if self.micro % self.accumulate == 0:
                self.optimizer_step()
        return self._static_loss
