"""Checkpoint I/O in the reference's layout and the top-k ensemble of its evaluation script (host side only).

Reference behaviour restated here:
* ``train.py:106-114``  -- ``ModelCheckpoint(save_top_k=5, save_weights_only=True)`` of pytorch-lightning 1.5.10 writes
  ``{'epoch', 'global_step', 'pytorch-lightning_version', 'state_dict'}`` where the keys of ``state_dict`` are those of
  the ``pl_model_wrapper.Model`` wrapper, i.e. the network's keys prefixed with ``model.``;
* ``train.py:146-153``  -- ``--model-weights``: ``checkpoint['state_dict']`` if present, else the object itself is the
  state dict; loaded with ``strict=True``;
* ``validate_ensemble.py:221-263`` -- the five ``epoch=*.ckpt`` files are loaded into five wrappers (legacy key fix
  ``resensenet`` -> ``resensnet``, ``strict=True``), put in ``eval()`` mode, and
* ``test_utils.py:21-38, 351-357`` -- every batch goes through all of them and the outputs are averaged
  (``average_outputs``: dicts key by key, tensors ``sum / n``, strings: the first).
"""
from typing import Dict, Iterable, List, Mapping, Union

import torch
from torch import nn

PL_VERSION = '1.5.10'          # requirements.txt:47 of the reference
_WRAPPER_PREFIX = 'model.'


def average_outputs(outputs, dtype):
    """``test_utils.average_outputs``: average a list of model outputs (dict of tensors / tensors / strings)."""
    if isinstance(outputs, list) and dtype == dict:
        return {key: average_outputs([d[key] for d in outputs], dtype=type(outputs[0][key])) for key in outputs[0].keys()}
    if isinstance(outputs, list) and dtype == str:
        return outputs[0]
    if isinstance(outputs, list) and issubclass(dtype, torch.Tensor):
        return sum(outputs) / len(outputs)
    raise AssertionError(f'cannot average outputs of type {dtype}')


def _is_wrapper(module: nn.Module) -> bool:
    """True for ``pl_model_wrapper.Model`` (its network lives under ``.model``), False for a bare network."""
    return isinstance(getattr(module, 'model', None), nn.Module) and any(k.startswith(_WRAPPER_PREFIX) for k in module.state_dict())


def normalise_state_dict(state_dict: Mapping[str, torch.Tensor], for_wrapper: bool) -> Dict[str, torch.Tensor]:
    """Legacy key fix of ``validate_ensemble.py:251-256`` plus adding / stripping the wrapper's ``model.`` prefix so that
    a checkpoint written from the wrapper also loads into a bare network and vice versa."""
    out = {}
    for k, v in state_dict.items():
        k = k.replace('resensenet', 'resensnet')
        has = k.startswith(_WRAPPER_PREFIX)
        if for_wrapper and not has:
            k = _WRAPPER_PREFIX + k
        elif not for_wrapper and has:
            k = k[len(_WRAPPER_PREFIX):]
        out[k] = v
    return out


def load_checkpoint(module: nn.Module, checkpoint: Union[str, Mapping], strict: bool = True, map_location='cpu'):
    """Load a reference ``.ckpt`` (or a bare ``state_dict``) into the wrapper or the bare network; returns the checkpoint
    dict.  A trainer that keeps packed bf16 weight images must be told afterwards (``FusionTrainer.mark_weights_dirty``)."""
    if isinstance(checkpoint, (str, bytes)) or hasattr(checkpoint, '__fspath__'):
        checkpoint = torch.load(checkpoint, map_location=map_location)
    try:
        state_dict = checkpoint['state_dict']                  # train.py:148-151
    except KeyError:
        state_dict = checkpoint
    module.load_state_dict(normalise_state_dict(state_dict, _is_wrapper(module)), strict=strict)
    return checkpoint


def save_checkpoint(module: nn.Module, path: str, epoch: int = 0, global_step: int = 0) -> None:
    """Write the ``save_weights_only`` checkpoint of pytorch-lightning 1.5.10 (keys of the wrapper: ``model.`` prefix)."""
    sd = {k: v.detach().cpu() for k, v in normalise_state_dict(module.state_dict(), for_wrapper=True).items()}
    torch.save({'epoch': int(epoch), 'global_step': int(global_step), 'pytorch-lightning_version': PL_VERSION, 'state_dict': sd},
               path)


class InferenceArena:
    """Packed bf16 weight images for eval-mode forwards of one or several models on a device: recorded during the first
    forward, looked up afterwards (no per-conv weight-packing launch), regenerated when any parameter changed (tensor version
    counters).  Uses the per-device packed-weight arena of the library, which has one owner at a time (a live FusionTrainer on
    the same device keeps it: the session then simply runs without an arena)."""

    def __init__(self, modules: Iterable[nn.Module], device):
        from . import trainer as _t
        self._t = _t
        self.params = [p for m in modules for p in m.parameters()]
        self.device = torch.device(device)
        self.dev = self.device.index if self.device.index is not None else torch.cuda.current_device()
        _t._ARENA_TOKENS[0] += 1
        self.token = _t._ARENA_TOKENS[0]
        self.state, self.buf, self.versions = 0, None, None        # 0 none, 1 recording, 2 sealed

    def _version(self):
        return sum(p._version for p in self.params)

    def __enter__(self):
        from . import ops
        from .functional import get_compute_dtype
        if get_compute_dtype() != torch.bfloat16 or torch.cuda.is_current_stream_capturing():
            return self
        if self.state == 0 and self._t._ARENA_OWNER.get(self.dev) is None:
            self.buf = torch.empty(96 << 20, dtype=torch.uint8, device=self.device)
            ops.weight_arena_begin(self.buf)
            self._t._ARENA_OWNER[self.dev] = self.token
            self.state, self.versions = 1, self._version()
        elif self.state == 2:
            if self._version() != self.versions:
                ops.weight_arena_pack(self.buf)
                self.versions = self._version()
            ops.weight_arena_enable(self.dev, True)
        return self

    def __exit__(self, *exc):
        from . import ops
        if self.state == 1:
            ops.weight_arena_seal(self.dev)
            self.state = 2
        if self.state == 2:
            ops.weight_arena_enable(self.dev, False)

    def close(self):
        if self.state and self._t._ARENA_OWNER.get(self.dev) == self.token:
            from . import ops
            ops.weight_arena_end(self.dev)
            del self._t._ARENA_OWNER[self.dev]
        self.state, self.buf = 0, None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Ensemble(nn.Module):
    """The evaluation ensemble of ``test_utils.run_evaluation_instance``: all members in ``eval()`` mode (BatchNorm on
    running statistics), one forward each under ``no_grad``, outputs averaged."""

    def __init__(self, members: Iterable[nn.Module]):
        super().__init__()
        self.members = nn.ModuleList(list(members))
        if len(self.members) == 0:
            raise ValueError('an ensemble needs at least one member')
        self.eval()

    @classmethod
    def from_checkpoints(cls, build, paths: List[str], device='cuda'):
        """``build()`` constructs one network (``model_factory[config.model]()``); one member per checkpoint file."""
        members = []
        for path in paths:
            net = build()
            load_checkpoint(net, path, strict=True)
            members.append(net.to(device))
        return cls(members)

    def train(self, mode: bool = True):                        # the ensemble is inference-only, as in the reference
        return super().train(False)

    @torch.no_grad()
    def forward(self, batch):
        first = next(self.members[0].parameters())
        if first.is_cuda:
            if getattr(self, '_arena', None) is None:
                self._arena = InferenceArena(self.members, first.device)
            with self._arena:
                outputs = [m(batch) for m in self.members]
        else:
            outputs = [m(batch) for m in self.members]
        return average_outputs(outputs, type(outputs[0]))
