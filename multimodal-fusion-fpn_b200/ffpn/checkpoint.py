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
        outputs = [m(batch) for m in self.members]
        return average_outputs(outputs, type(outputs[0]))
