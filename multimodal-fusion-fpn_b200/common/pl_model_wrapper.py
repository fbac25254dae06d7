"""Training wrapper with the API of the reference's ``common/pl_model_wrapper.py`` (``Model``, :108-329)
without pytorch-lightning (not installable here): same constructor, ``forward``, ``training_step``,
``validation_step``, ``*_epoch_end`` and ``configure_optimizers``.  The debug-PNG / matplotlib side effects
of the reference (:169-241, :269-297) are visual debugging and are left out; ``self.log`` records into
``self.logged`` instead of Lightning's logger.  ``load_state_dict`` keys keep the ``model.`` prefix."""
import gc
import json
from os.path import join
from typing import Dict, List

import torch
from torch import nn


class ReleaseMemCache:
    def __call__(self):
        gc.collect()
        torch.cuda.empty_cache()


class DoNotReleaseMemCache:
    def __call__(self):
        pass


class StaticFactory:
    classes: List[type] = []

    def __init__(self):
        self.classes_names = {c.__name__: c for c in self.classes}

    def create_class(self, class_name, *args, **kwargs):
        return self.classes_names[class_name](*args, **kwargs)


class MemCacheStrategies(StaticFactory):
    classes = [ReleaseMemCache, DoNotReleaseMemCache]


class Model(nn.Module):
    def __init__(self, model, losses, training_metrics, metrics, metametrics, optim,
                 force_mem_cache_release="DoNotReleaseMemCache", validation=None, _log_file=None, model_path: str = ''):
        super().__init__()
        self.model = model
        self.loss = losses
        self.metrics = metrics
        self.metametrics = metametrics
        self.optim = optim
        self.training_metrics = training_metrics
        self.validation = validation
        self.force_mem_cache_release = MemCacheStrategies().create_class(force_mem_cache_release)
        self.curves: Dict[str, list] = {}
        for tm in (self.training_metrics or {}):
            self.curves[f'{tm} (train)'] = []
        for vm in (self.metrics or {}):
            self.curves[f'{vm} (val)'] = []
        self.model_path = model_path
        self.logged: Dict[str, list] = {}
        self.sync_logging = True      # False: keep logged losses on the device (no .item() sync per step)

    def log(self, name, value, **_kw):
        self.logged.setdefault(name, []).append(value)

    def forward(self, x, **kwargs):  # type: ignore
        self.force_mem_cache_release()
        return self.model(x, **kwargs)

    def training_step(self, batch, _batch_idx):  # type: ignore
        res = self(batch)
        loss, values = self.loss(batch, res)
        for k in values:
            self.log('Training/' + str(k), values[k].item() if self.sync_logging else values[k].detach(),
                     on_step=True, on_epoch=False)
        with torch.no_grad():
            for k in (self.training_metrics or {}):
                self.training_metrics[k].update(batch, res)
        return loss

    def training_epoch_end(self, _outputs) -> None:  # type: ignore
        for k in (self.training_metrics or {}):
            value = self.training_metrics[k].get()
            self.log('Training/' + str(k), value, on_epoch=True)
            self.training_metrics[k].reset()
            self.curves[k + ' (train)'].append(value)
        if self.model_path:
            with open(join(self.model_path, 'curves.json'), 'w') as f:
                json.dump(self.curves, f)

    def validation_step(self, batch, _batch_idx):  # type: ignore
        with torch.no_grad():
            res = self(batch)
            for k in (self.metrics or {}):
                self.metrics[k].update(batch, res)
        return res

    def validation_epoch_end(self, _validation_step_outputs):  # type: ignore
        results = {k: self.metrics[k].get() for k in (self.metrics or {})}
        for k in results:
            self.log('Validation/' + str(k), results[k], on_epoch=True)
            self.metrics[k].reset()
            self.curves[k + ' (val)'].append(results[k])
        for k in (self.metametrics or {}):
            self.log(str(k), self.metametrics[k].get(results), on_epoch=True)

    def configure_optimizers(self):
        return self.optim
