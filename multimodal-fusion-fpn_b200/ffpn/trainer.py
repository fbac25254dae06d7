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


LATE_STAGES = ('conv1', 'conv2', 'zdimRed1', 'zdimRed2')     # backward reaches these last (fusion3D2D.py forward order)


def late_parameter_names(model: nn.Module):
    """Parameters whose gradients are produced AFTER the bucket marker of the fusion bodies (the first two 3-D encoder levels
    and their projective blocks): everything else can be all-reduced and stepped while their backward still runs."""
    late = set()
    for name, _ in model.named_parameters():
        parts = name.split('.')
        if len(parts) > 1 and parts[0] == 'resensnet' and parts[1] in LATE_STAGES:
            late.add(name)
    return late


def flatten_parameters(model: nn.Module, device=None, late_names=None):
    """Move all parameters into one flat fp32 buffer (parameters become views); returns (flat_p, flat_g).  ``late_names``:
    these parameters are laid out at the END of the buffers (second gradient bucket); ``flat_p.n_early`` tells where it starts."""
    named = list(model.named_parameters())
    if late_names:
        named = [(k, p) for k, p in named if k not in late_names] + [(k, p) for k, p in named if k in late_names]
    params = [p for _, p in named]
    n = sum(p.numel() for p in params)
    n_early = sum(p.numel() for k, p in named if not (late_names and k in late_names))
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
    flat_p.n_early = n_early
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
        # Optional two gradient buckets (FFPN_BUCKETS=1): [everything else | conv1, conv2, zdimRed1,
        # zdimRed2]; the first one is all-reduced and stepped on a communication stream while the backward of the second still
        # runs.  Measured on 8 x B200 (same box, 40 steps): 9.70 ms/step with the overlap vs 9.62 ms with ONE all-reduce at the
        # end of the captured step (2 GPUs: 9.53 vs 9.50) -- the NCCL kernel competes with the level-1/2 backward for SMs and HBM
        # and the collective is only ~1 % of the step, so the single bucket is the default; what did pay was capturing the
        # all-reduce + SGD + re-packing inside the CUDA graph (8 GPUs: 10.13 -> 9.62 ms/step).
        late = late_parameter_names(model) if self.accumulate == 1 and os.environ.get('FFPN_BUCKETS') == '1' else set()
        self.flat_p, self.flat_g = flatten_parameters(model, late_names=late)
        self.n_early = self.flat_p.n_early if late else 0       # 0: single bucket
        self.mom = torch.zeros_like(self.flat_p)
        self._comm_stream = None
        self._early_done = False       # the early bucket's all-reduce + SGD were issued by the bucket hook of this backward
        self._bucket_checked = False   # first step: verify that no early gradient is written after the marker
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
        self._early_done = False
        if self.n_early and self._optimizer_in_backward:
            functional.set_bucket_hook(self._on_early_bucket_ready)
        try:
            out = self.model(batch)
            loss, _ = self.criterion(batch, out)
            (loss / self.accumulate if self.accumulate > 1 else loss).backward()
            functional.join_side_streams()     # the sink is written from every branch stream
            if self._early_done:
                torch.cuda.current_stream().wait_stream(self._comm_stream)
        finally:
            functional.set_bucket_hook(None)
            functional.set_grad_sink(None)
            if self._arena_state == 1:
                ops.weight_arena_seal(dev)
                self._arena_state = 2
            if self._arena_state:
                ops.weight_arena_enable(dev, False)
        return loss.detach()

    # -- gradient buckets ---------------------------------------------------------------------------------
    _optimizer_in_backward = False     # set by step() / capture(): forward_backward() alone must not touch the weights

    def _on_early_bucket_ready(self):
        """Bucket marker (runs inside backward, when the gradient has flowed back through encoder level 3): every kernel that
        writes a gradient of the first bucket has been issued.  All-reduce + SGD of that bucket go to the communication
        stream, ordered after all of those kernels, and overlap the backward of levels 2 and 1."""
        from . import functional
        if self._early_done:
            return
        cur = torch.cuda.current_stream()
        if self._comm_stream is None:
            self._comm_stream = torch.cuda.Stream(device=self.flat_p.device)
        if not self._bucket_checked:
            # verification pass (first step): snapshot the first bucket after everything ISSUED so far has run; at the end of the
            # step the snapshot must equal the final gradients, i.e. nothing issued later writes into this bucket
            for st in functional.used_side_streams(self.flat_p.device):
                cur.wait_stream(st)
            self._early_snapshot = (self.flat_g[:self.n_early].clone(), cur)
            return
        comm = self._comm_stream
        comm.wait_stream(cur)
        for st in functional.used_side_streams(self.flat_p.device):
            comm.wait_stream(st)
        with torch.cuda.stream(comm):
            self._reduce_and_step(0, self.n_early)
        self._early_done = True

    def _reduce_and_step(self, a: int, b: int):
        from . import ops
        g = self.flat_g[a:b]
        scale = allreduce_mean_(g, self.group)
        ops.sgd_step(self.flat_p[a:b], g, self.mom[a:b], self.lr, self.momentum, self.weight_decay, scale, self.steps == 0 and not self._in_capture)
        g.zero_()

    _in_capture = False

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
        """All-reduce (sum; 1/world folded into the SGD kernel) + SGD + re-packing of the bf16 weight images.  When the bucket
        hook already handled the first bucket during backward, only the second one (the first two encoder levels) is left."""
        from . import ops
        if self._early_done:
            self._reduce_and_step(self.n_early, self.flat_p.numel())
            self._early_done = False
        else:
            self._reduce_and_step(0, self.flat_p.numel())
        if self._arena_state == 2:
            ops.weight_arena_pack(self.flat_p)         # keep the arena images in step with the weights
        self.steps += 1

    def _verify_buckets(self):
        """First step only: the gradients of the first bucket must not change after the marker fired (else the layout
        assumption does not hold for this model and the trainer falls back to a single bucket)."""
        snap = getattr(self, '_early_snapshot', None)
        self._bucket_checked = True
        if snap is None:
            self.n_early = 0                          # the marker never fired (a body without one): single bucket
            return
        self._early_snapshot = None
        if not torch.equal(snap[0], self.flat_g[:self.n_early]):
            self.n_early = 0

    def step(self, batch):
        """Eager training step (forward, loss, backward, all-reduce, SGD)."""
        self._optimizer_in_backward = self.accumulate == 1
        try:
            loss = self.forward_backward(batch)
        finally:
            self._optimizer_in_backward = False
        if self.n_early and not self._bucket_checked and self.accumulate == 1:
            self._verify_buckets()
        self.micro += 1
        if self.micro % self.accumulate == 0:
            self.optimizer_step()
        return loss

    # -- CUDA graph --------------------------------------------------------------------------------------
    def capture(self, example_batch, warmup: int = 3, include_optimizer: Optional[bool] = None):
        """Capture the whole optimisation step on static input buffers: forward + loss + backward, the bucketed gradient
        all-reduce (NCCL, captured on the communication stream, overlapping the backward of the first two encoder levels), SGD
        and the re-packing of the bf16 weight images.  ``include_optimizer=False`` (or gradient accumulation) captures
        forward + backward only and leaves the optimiser to ``replay()``."""
        self._fused_opt = (self.accumulate == 1) if include_optimizer is None else (include_optimizer and self.accumulate == 1)
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
                # bucket hook in "verify" mode on the first pass (it only snapshots the first bucket's gradients)
                self._optimizer_in_backward = self._fused_opt and not self._bucket_checked
                try:
                    self.forward_backward(self._static)
                finally:
                    self._optimizer_in_backward = False
                if self.n_early and not self._bucket_checked and self._fused_opt:
                    self._verify_buckets()
                self.flat_g.zero_()
            for b, s in zip(self.model.buffers(), saved):
                b.copy_(s)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        first = self.steps == 0
        self._graph = torch.cuda.CUDAGraph()
        self._in_capture = True
        self._optimizer_in_backward = self._fused_opt
        try:
            with torch.cuda.graph(self._graph):
                self._static_loss = self.forward_backward(self._static)
                if self._fused_opt:
                    self.optimizer_step()
        finally:
            self._in_capture = False
            self._optimizer_in_backward = False
        if self._fused_opt:
            self.steps -= 1                  # optimizer_step() counted the capture itself; replay() counts the real steps
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
        from . import ops
        ops.note_bn_statistics_update()        # the replay updates the running statistics behind torch's version counters
        self._graph.replay()
        if self._fused_opt:
            self.steps += 1
        else:
            self.micro += 1
            if self.micro % self.accumulate == 0:
                self.optimizer_step()
        return self._static_loss
