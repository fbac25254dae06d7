"""Task wrappers and model factory, API-compatible with the reference's ``models/fusion_nets.py``: the same
eight registered names, ``forward(x: dict) -> {'prediction': Tensor}``, ``.resensnet`` attributes and
``state_dict`` keys.  ``FPNHybridFusion`` (reference :84-121) is the hot path of this repository."""
import configparser
import os

import torch
from torch import nn
from torch import Tensor
from torch.nn import functional as F

from config import config
from ffpn import functional as FF
from models.fpn.unets3D import ModifiedUnet3D
from models.fpn.unets2D import ModifiedUnet2DLevel5
from models.fpn.fusion3D2D import ModifiedUnet3D2DLevel5
from utils import get_factory_adder

add_class, factory_classes = get_factory_adder()

_INI = 'modifiedUnet3D_red-convPlusFully_dropout00'


class FPNConfig(nn.Module):
    """Reads the architecture .ini (reference :21-26).  The reference opens it relative to the CWD; this one
    falls back to the copy shipped next to this file so the factory also works from another directory."""

    def __init__(self):
        super().__init__()
        self.config = configparser.ConfigParser()
        rel = os.path.join('models', 'fpn', _INI + '.ini')
        if not self.config.read(rel):
            self.config.read(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'fpn', _INI + '.ini'))


class _fused_sigmoid:
    """Context: run ``head`` (a HeadConv3d) with the sigmoid fused in when ``wrapper`` still has the stock sigmoid
    ``last_activation`` (the Regression subclasses override it) -> ``.on`` tells the caller to skip its own activation."""

    def __init__(self, wrapper, stock, head):
        self.head = head
        self.on = head is not None and FF.head_activation_fused() and type(wrapper).last_activation is stock

    def __enter__(self):
        if self.on:
            self.head.fused_activation = 'sigmoid'
        return self

    def __exit__(self, *exc):
        if self.on:
            self.head.fused_activation = None


def _interpolate_mode():
    """config.crop -> how 2-D features reach the en-face grid (reference :102-107)."""
    mode = '2d' if 'relative_2d' in config.crop else None
    if 'max' in config.crop and mode is not None:
        mode += '_max'
    return mode


@add_class
class FPN(FPNConfig):
    def __init__(self):
        super().__init__()
        self.resensnet = ModifiedUnet3D(self.config)

    def last_activation(self, x): return torch.sigmoid(x)

    def forward(self, x):
        oct = x['image'].permute(0, 1, 2, 4, 3)               # (B,1,S,H,W) -> Z x W x H
        with _fused_sigmoid(self, FPN.last_activation, self.resensnet.final1 if self.resensnet.use_1x1 else None) as f:
            seg = self.resensnet(oct).permute(0, 1, 2, 4, 3)
        return {'prediction': seg if f.on else self.last_activation(seg)}


@add_class
class FPNRegression(FPN):
    def last_activation(self, x): return x


@add_class
class FPNClassification(FPN, FPNConfig):
    def __init__(self):
        FPNConfig.__init__(self)
        self.resensnet = ModifiedUnet3D(self.config, classification=True)
        self.one_one = nn.Conv3d(256, config.number_of_outputs, kernel_size=1, stride=1, padding=0, bias=False)
        self.adaptive_pool = nn.AdaptiveAvgPool3d((1, 1, 1))

    def last_activation(self, _x):
        raise NotImplementedError

    def forward(self, x):
        oct = x['image'].permute(0, 1, 2, 4, 3)
        feat = self.resensnet(oct)                             # (B,256,S/4,W/16,H/16) from the CUDA encoder
        # 1x1x1 conv and global average commute; the classifier tail is a (B,256)x(256,n) product
        pooled = feat.float().mean(dim=(2, 3, 4))
        pred = pooled @ self.one_one.weight.flatten(1).t()
        return {'prediction': torch.softmax(pred, dim=-1)}


@add_class
class FPNHybridFusion(FPNConfig):
    """Hybrid fusion FPN: 3-D OCT encoder + projective 3-D->2-D blocks, 2-D SLO/FAF encoder, shared 2-D
    decoder fed by both at every level.  crop 'oct': 2-D image already on the en-face grid; 'relative_2d':
    2-D features bilinearly resized; 'relative_2d_max': adaptive-max resized."""

    def __init__(self):
        super().__init__()
        self.interpolate = _interpolate_mode()
        self.resensnet = ModifiedUnet3D2DLevel5(self.config, self.interpolate)

    def last_activation(self, x): return torch.sigmoid(x)

    def forward(self, x):
        oct = x['image'].permute(0, 1, 2, 4, 3)               # Z x W x H
        slo = x[config.fusion_modality][:, :, :, 0, :]
        with _fused_sigmoid(self, FPNHybridFusion.last_activation, self.resensnet.final1) as f:
            seg = self.resensnet(oct, slo).permute(0, 1, 2, 4, 3)
        return {'prediction': seg if f.on else self.last_activation(seg)}


@add_class
class FPNHybridFusionRegression(FPNHybridFusion):
    def last_activation(self, x): return x


@add_class
class FPN2D(FPNConfig):
    def __init__(self):
        super().__init__()
        self.resensnet = ModifiedUnet2DLevel5(self.config)

    def forward(self, x):
        fused = x[config.fusion_modality][:, :, :, 0, :]
        head = self.resensnet.final1[0]
        head.fused_activation = 'sigmoid' if FF.head_activation_fused() else None
        try:
            seg = self.resensnet(fused).permute(0, 1, 2, 4, 3)
        finally:
            head.fused_activation = None
        if not FF.head_activation_fused():
            seg = torch.sigmoid(seg)
        if seg.shape != x['mask'].shape:
            seg = F.interpolate(seg, size=x['mask'].shape[2:], mode='trilinear')
        return {'prediction': seg}


@add_class
class FPNLateFusion(FPNConfig):
    """Late fusion: an OCT-only FPN and a 2-D FPN run separately; their 16-channel feature maps are resized
    to a common grid, concatenated and mixed by a 1x1x1 conv (reference :150-216)."""

    def __init__(self):
        super().__init__()
        self.resensnet3d = ModifiedUnet3D(self.config)
        self.resensnet2d = ModifiedUnet2DLevel5(self.config, output_features=True)
        self.resensnet3d.use_1x1 = False                       # features, not logits
        self.fusion_module = nn.Conv3d(32, config.number_of_outputs, (1, 1, 1))
        self.interpolate = _interpolate_mode()

    def last_activation(self, x): return torch.sigmoid(x)

    def forward(self, x):
        oct = x['image'].permute(0, 1, 2, 4, 3)
        oct_feat = self.resensnet3d(oct)                       # (B,16,S,W,1)
        fused = x[config.fusion_modality][:, :, :, 0, :]
        f2d = self.resensnet2d(fused)                          # (B,16,S',W',1)
        f2d = FF.Resize2DFunction.apply(f2d[:, :, :, :, 0], tuple(oct_feat.shape[2:4]), self.interpolate)
        fuse = FF.head_activation_fused() and type(self).last_activation is FPNLateFusion.last_activation
        seg = self.fuse_features(oct_feat, f2d, 'sigmoid' if fuse else None)
        return {'prediction': seg if fuse else self.last_activation(seg)}

    def fuse_features(self, oct_seg: Tensor, fused_seg: Tensor, act=None):
        cat = FF.cat(oct_seg, fused_seg)
        seg = FF.HeadFunction.apply(cat, self.fusion_module.weight, self.fusion_module.bias, act)
        return seg.permute(0, 1, 2, 4, 3)


@add_class
class FPNLateFusionRegression(FPNLateFusion):
    def last_activation(self, x): return x
