"""Training losses with the reference's interface (``common/loss.py``): every loss is called as
``loss(target_dict, predict_dict)`` and looks its tensors up by key.  They stay in PyTorch: the tensors are
the (B,1,S,1,W) prediction and mask, a few thousand elements (component 6)."""
from typing import Optional, Union

import torch
import torch.nn as nn
import torch.nn.functional as F

Key = Union[int, str]


class Mix(nn.Module):
    """Weighted sum of named losses divided by their count (reference loss.py:9-28).
    Returns ``(total, {name: value})``."""

    def __init__(self, losses, coefficients: Optional[dict] = None):
        super().__init__()
        self.losses = losses
        self.coefficients = {k: 1 for k in losses} if coefficients is None else coefficients

    def forward(self, target, predict):
        fused = self._fused_dice_bce(target, predict)
        if fused is not None:
            return fused
        results = {k: fn(target, predict) for k, fn in self.losses.items()}
        live = [results[k] * self.coefficients[k] for k in results if results[k] is not None]
        return sum(live) / len(results), results

    def _fused_dice_bce(self, target, predict):
        """The training configuration of the reference -- {Dice_loss_jointv2, BCE_Lossv2} on the same tensors, unit
        coefficients (train.py:135-139) -- on CUDA fp32 tensors runs as three kernels of libfusionfpn.so (two forward, one
        backward) instead of ~45 ATen launches.  Anything else takes the generic path below."""
        if len(self.losses) != 2 or any(self.coefficients.get(k, 1) != 1 for k in self.losses):
            return None
        dice = [k for k, f in self.losses.items() if type(f) is Dice_loss_jointv2 and not f.force_binary]
        bce = [k for k, f in self.losses.items() if type(f) is BCE_Lossv2]
        if len(dice) != 1 or len(bce) != 1:
            return None
        fd, fb = self.losses[dice[0]], self.losses[bce[0]]
        if (fd.output_key, fd.target_key) != (fb.output_key, fb.target_key):
            return None
        t, p = fd._pair(target, predict)
        if not (p.is_cuda and t.is_cuda and p.dtype == torch.float32 and p.dim() >= 3):
            return None
        from ffpn.functional import MixDiceBCEFunction
        total, d, b = MixDiceBCEFunction.apply(p, t)
        return total, {k: (d if k == dice[0] else b) for k in self.losses}

    @staticmethod
    def normalize_data(data):
        lo, hi = torch.min(data), torch.max(data)
        return (data - lo) / (hi - lo)


class _KeyedLoss(nn.Module):
    def __init__(self, output_key: Key = 0, target_key: Key = 0):
        super().__init__()
        self.output_key = output_key
        self.target_key = target_key

    def _pair(self, target, predict):
        t, p = target[self.target_key], predict[self.output_key]
        assert t.shape == p.shape, f'{t.shape} != {p.shape}'
        return t, p


class BCE_Lossv2(_KeyedLoss):
    """Mean binary cross entropy over all elements (reference loss.py:35-56)."""

    def __init__(self, output_key: Key = 0, target_key: Key = 0, bg_weight=1):
        super().__init__(output_key, target_key)
        self.bg_weight = bg_weight

    def forward(self, target, predict):
        t, p = self._pair(target, predict)
        return F.binary_cross_entropy(p.reshape(-1), t.reshape(-1), reduction='mean')


class Dice_loss_jointv2(_KeyedLoss):
    """1 - mean_c 2(sum p*g + 1e-6)/(sum(p^2 + g) + 2e-6), sums over batch and space (reference loss.py:59-90)."""

    def __init__(self, output_key: Key = 0, target_key: Key = 0, force_binary: bool = False, threshold: float = 0.5):
        super().__init__(output_key, target_key)
        self.force_binary = force_binary
        self.threshold = threshold

    def forward(self, target, predict):
        t, p = self._pair(target, predict)
        b, c = t.shape[0], t.shape[1]
        p, t = p.reshape(b, c, -1), t.reshape(b, c, -1)
        if self.force_binary:
            t = (t > self.threshold).float()
        inter = (p * t).sum(dim=(0, 2)) + 1e-6
        union = (p ** 2 + t).sum(dim=(0, 2)) + 2e-6
        return 1.0 - torch.mean(2.0 * inter / union)
