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


def broadcast_replica_state(model: nn.Module, flat_p: torch.Tensor, mom: torch.Tensor, group=None) -> None:
    """Make every rank start from rank 0's parameters, momentum and BatchNorm buffers.  The reference's DataParallel
    re-replicates device 0's module every step (torch/nn/parallel/replicate.py), so its replicas can never differ; with one
    process per GPU the equivalent guarantee is one broadcast at construction (ranks that seeded differently, or loaded
    different checkpoints, would otherwise average gradients taken at different weights and drift apart silently)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    src = dist.get_global_rank(group, 0) if group is not None else 0
    dist.broadcast(flat_p, src=src, group=group)
    dist.broadcast(mom, src=src, group=group)
    bufs = [b for b in model.buffers()]
    fl = [b for b in bufs if b.is_floating_point()]
    if fl:
        flat = torch.cat([b.reshape(-1).float() for b in fl])
        dist.broadcast(flat, src=src, group=group)
        off = 0
        for b in fl:
            b.copy_(flat[off:off + b.numel()].view(b.shape))
            off += b.numel()
    ints = [b for b in bufs if not b.is_floating_point()]
    if ints:
        flat = torch.stack([b.reshape(()).to(torch.int64) for b in ints])          # num_batches_tracked counters
        dist.broadcast(flat, src=src, group=group)
        for i, b in enumerate(ints):
            b.copy_(flat[i])


# The packed-weight arena lives in the per-device library context, so it has ONE owner at a time: the trainer whose token is
# registered here.  A second trainer on the same device runs without an arena (per-call weight packing: slower, never wrong)
# until the owner is closed or collected; a trainer can only ever end its own arena.
_ARENA_OWNER: Dict[int, int] = {}
_ARENA_TOKENS = [0]


class FusionTrainer:
    def __init__(self, model: nn.Module, criterion, lr: float = 0.1, momentum: float = 0.9,
                 weight_decay: float = 1e-4, group=None, accumulate_grad_batches: int = 1):
        self.model, self.criterion = model, criterion
        self.lr, self.momentum, self.weight_decay = lr, momentum, weight_decay
        self.group = group
        self.accumulate = max(1, int(accumulate_grad_batches))     # train.py:161
        self.flat_p, self.flat_g = flatten_parameters(model)
        self.mom = torch.zeros_like(self.flat_p)
        broadcast_replica_state(model, self.flat_p, self.mom, group)
        # weights replaced behind the trainer's back (load_state_dict / load_checkpoint copy in place, same pointers): the
        # packed bf16 images of the arena are regenerated before the next forward or replay
        self._hook = model.register_load_state_dict_post_hook(lambda _m, _k: self.mark_weights_dirty())
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
        _ARENA_TOKENS[0] += 1
        self._arena_token = _ARENA_TOKENS[0]
        self._weights_dirty = False    # set by mark_weights_dirty(): repack before the next forward
        # With accumulate_grad_batches > 1 (train.py:161) the sink stays off: BatchNorm gradients are plain writes, so the
        # micro-batches are summed by autograd's in-place accumulation into the same flat views instead.
        self._sink = {}
        if self.accumulate == 1:
            for p in model.parameters():
                if p.dim() >= 1 and p.grad is not None:
                    self._sink[p.data_ptr()] = p.grad

    # -- pieces ----------------------------------------------------------------------------------------
    def forward_backward(self, batch):
        from . import functional, ops
        from .functional import get_compute_dtype
        dev = self._dev()
        if get_compute_dtype() == torch.bfloat16 and not os.environ.get('FFPN_NO_ARENA'):
            if self._arena_state == 0 and _ARENA_OWNER.get(dev) is None:
                self._arena = torch.empty(96 << 20, dtype=torch.uint8, device=self.flat_p.device)
                ops.weight_arena_begin(self._arena)
                _ARENA_OWNER[dev] = self._arena_token
                self._arena_state = 1
            elif self._arena_state == 2:
                self._repack_if_dirty()
        functional.set_grad_sink(self._sink)
        # the arena is only consulted while THIS forward+backward runs: any other conv call on the device (an eval forward,
        # another model) packs its weights from the current fp32 values and cannot see a stale image
        if self._arena_state:
            ops.weight_arena_enable(dev, True)
        try:
            out = self.model(batch)
            loss, _ = self.criterion(batch, out)
            (loss / self.accumulate if self.accumulate > 1 else loss).backward()
            functional.join_side_streams()     # the sink is written from every branch stream
        finally:
            functional.set_grad_sink(None)
            if self._arena_state == 1:
                ops.weight_arena_seal(dev)
                self._arena_state = 2
            if self._arena_state:
                ops.weight_arena_enable(dev, False)
        return loss.detach()

    def _dev(self) -> int:
        return self.flat_p.device.index if self.flat_p.device.index is not None else 0

    def _repack_if_dirty(self):
        if self._weights_dirty and self._arena_state == 2:
            from . import ops
            ops.weight_arena_pack(self.flat_p)
            self._weights_dirty = False

    def mark_weights_dirty(self):
        """Call after modifying parameters outside optimizer_step() (e.g. load_state_dict): the bf16 weight images of the
        arena are regenerated before the next forward."""
        self._weights_dirty = True

    def close(self):
        """Detach the packed-weight arena from the library context (the context is shared per device)."""
        if self._arena_state and _ARENA_OWNER.get(self._dev()) == self._arena_token:
            from . import ops
            ops.weight_arena_end(self._dev())
            del _ARENA_OWNER[self._dev()]
        self._arena_state, self._arena = 0, None
        if getattr(self, '_hook', None) is not None:
            self._hook.remove()
            self._hook = None

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
        # The warm-up passes (>= 1: the packed-weight arena must be sealed before the capture, its cudaMalloc / memcpy cannot be
        # captured) are real training-mode forwards: they would momentum-update every BatchNorm running statistic and bump
        # num_batches_tracked.  Snapshot the buffers and put them back, so that a graph-trained model's buffers equal an
        # eagerly trained one's.
        warmup = max(1, int(warmup))
        saved = [b.detach().clone() for b in self.model.buffers()]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self.forward_backward(self._static)
                self.flat_g.zero_()
            for b, s in zip(self.model.buffers(), saved):
                b.copy_(s)
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
        self._repack_if_dirty()                # load_state_dict since the last step: the graph reads the arena images
        self._graph.replay()
        if self._fused_opt:
            self.steps += 1
        else:
            self.micro += 1
            if self.micro % self.accumulate == 0:
                self.optimizer_step()
        return self._static_loss
