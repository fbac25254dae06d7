"""GPU side of the input pipeline and of the training metrics: what the reference does per
step on the host -- numpy z-scoring of every B-scan in the dataloader workers (``training_config.py:60`` ->
``mytransforms.ZScoreNormalization(axis=(2,3))``, ``mytransforms.py:277-296``) and ``.cpu().numpy()`` metric updates with a
device synchronisation each (``pl_model_wrapper.py:251-253``, ``metrics.py:216-253``) -- as kernels of libfusionfpn.so on the
training stream / a copy stream, so that a step never waits for the host."""
from typing import Dict, Union

import torch

from . import ops

Key = Union[int, str]


class GpuInputPipeline:
    """Pinned host batch -> device batch with the volume z-scored per B-scan on the GPU.

    ``prepare(batch)`` enqueues the host->device copies and the normalisation on a copy stream and returns immediately;
    ``get()`` makes the caller's stream wait for them and hands out the device batch.  Two buffer sets alternate, so the copy
    of batch i+1 overlaps the training step of batch i."""

    def __init__(self, device='cuda', normalize_keys=('image',), eps: float = 1e-8):
        self.device = torch.device(device)
        self.normalize_keys = tuple(normalize_keys)
        self.eps = eps
        self.stream = torch.cuda.Stream(device=self.device)
        self._slots = [{}, {}]
        self._ready = [torch.cuda.Event(), torch.cuda.Event()]
        self._free = [torch.cuda.Event(), torch.cuda.Event()]
        self._next, self._pending, self._last = 0, None, None

    def prepare(self, batch: Dict[str, torch.Tensor]) -> None:
        i = self._next
        self._next ^= 1
        slot = self._slots[i]
        self.stream.wait_event(self._free[i])                      # the consumer of this slot's previous batch is done with it
        with torch.cuda.stream(self.stream):
            for k, v in batch.items():
                if k not in slot or slot[k].shape != v.shape or slot[k].dtype != v.dtype:
                    slot[k] = torch.empty(v.shape, dtype=v.dtype, device=self.device)
                slot[k].copy_(v, non_blocking=True)
                if k in self.normalize_keys:
                    ops.zscore_bscans(slot[k], self.eps, out=slot[k])
            self._ready[i].record()
        self._pending = i

    def get(self) -> Dict[str, torch.Tensor]:
        i = self._pending
        if i is None:
            raise RuntimeError('GpuInputPipeline.get() without a prepared batch')
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(self._ready[i])
        self._pending, self._last = None, i
        return dict(self._slots[i])

    def release(self) -> None:
        """Call after the step that consumed the last ``get()`` has been enqueued: its buffers may then be overwritten."""
        if self._last is not None:
            self._free[self._last].record(torch.cuda.current_stream(self.device))
            self._last = None


class DeviceDice:
    """Dice metric with the interface of the reference's ``metrics.Dice`` (``update(ground, predict)``, ``get()``,
    ``reset()``; ``metrics.py:216-253``): per-sample Dice of the thresholded prediction and mask, averaged with nanmean at
    ``get()``.  The per-sample values stay on the device; the only synchronisation is the one ``get()`` call per epoch."""

    def __init__(self, output_key: Key = 0, target_key: Key = 0, slice: int = 0, output_threshold: float = 0.5,
                 target_threshold: float = 0.5):
        self.output_key, self.target_key, self.slice = output_key, target_key, slice
        self.output_threshold, self.target_threshold = output_threshold, target_threshold
        self.accumulator = []

    def calculate_batch(self, ground, predict) -> torch.Tensor:
        pred, gr = predict[self.output_key].detach(), ground[self.target_key].detach()
        assert gr[:, self.slice].shape == pred[:, self.slice].shape, f'GT: {gr.shape}, Pred.: {pred.shape}'
        return ops.dice_metric(pred, gr, self.slice, self.output_threshold, self.target_threshold)

    def update(self, ground, predict) -> None:
        self.accumulator.append(self.calculate_batch(ground, predict))

    def get(self) -> float:
        if not self.accumulator:
            return float('nan')
        return float(torch.nanmean(torch.cat(self.accumulator)).item())

    def reset(self) -> None:
        self.accumulator = []
