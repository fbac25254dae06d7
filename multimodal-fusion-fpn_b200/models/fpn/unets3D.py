"""OCT-only FPN body, API-compatible with the reference's ``models/fpn/unets3D.py`` (ModifiedUnet3D,
:8-485): the same 3-D encoder and projective blocks as the fusion body, a two-input decoder, and the
``original`` / ``classification`` / ``use_1x1`` switches.  Runs on libfusionfpn.so."""
import torch
from torch import nn

from config import config as global_config
from ffpn import functional as FF
from models.fpn import components as K
from models.fpn.components import SegmentationNetwork, unet3dConvX, unet3dUp2modified
from models.fpn.fusion3D2D import ModifiedUnet3D2D, POOLS, UPFACTORS


class ModifiedUnet3D(SegmentationNetwork):
    '''3-D encoder -> projective blocks -> 2-D decoder (skip connections from the projected features).'''

    def __init__(self, config, original=False, classification=False):
        super().__init__(n_classes=global_config.number_of_outputs,
                         is_batchnorm=config.getboolean('architecture', 'is-batchnorm'), in_channels=1,
                         is_deconv=config.getboolean('architecture', 'is-deconv'))
        self.use_1x1 = True
        self.original = original              # final projection kernel 8 and no depth mean (:79-82, :458-471)
        self.classification = classification  # return conv5, freeze projection + decoder (:175-185, :453-454)
        ModifiedUnet3D2D._read_architecture(self, config)
        ch, bn = self.channels, self.is_batchnorm
        for l in range(5):
            setattr(self, f'conv{l + 1}', self._make_layer_2plus3(self.in_channels if l == 0 else ch[l - 1], ch[l],
                                                                  is_batchnorm=bn, is_residual=True,
                                                                  dropout=self.dropout[l]))
        for l, k in enumerate(POOLS):
            setattr(self, f'pool{l + 1}', K.MaxPool3d(kernel_size=k))
        final_kernel_size = 8 if self.original else 4
        for l in range(5):
            setattr(self, f'zdimRed{l + 1}', self._make_zdimReductionConvPlusFully(
                channels_in=ch[l], channels_out=ch[l], num_convreductions=4 - l, final_kernelsize=final_kernel_size,
                is_batchnorm=bn, is_residual=True, dropout=0.0))
        for l in (4, 3, 2, 1):
            setattr(self, f'up_concat{l}', unet3dUp2modified(ch[l], ch[l - 1], upfactor=UPFACTORS[l],
                                                             is_deconv=self.is_deconv, is_residual=True,
                                                             dropout=self.dropout[9 - l], is_batchnorm=bn))
        self.final1 = K.HeadConv3d(in_channels=ch[0], out_channels=self.n_classes, kernel_size=1)
        if self.classification:
            frozen = [getattr(self, f'zdimRed{l}') for l in range(1, 6)] + [getattr(self, f'up_concat{l}') for l in (4, 3, 2, 1)]
            for module in frozen:
                for param in module.parameters():
                    param.requires_grad = False

    def _make_layer_2plus3(self, channels_in, channels_out, is_batchnorm, is_residual, dropout):
        return ModifiedUnet3D2D._two_blocks(unet3dConvX, nn.Conv3d, nn.BatchNorm3d, 3, channels_in, channels_out,
                                            is_batchnorm, is_residual, dropout)

    def _make_zdimReductionConvPlusFully(self, channels_in, channels_out, num_convreductions, final_kernelsize,
                                         is_batchnorm, is_residual, dropout):
        n = num_convreductions
        downsample = None
        if (channels_in != channels_out) or (n > 0 and is_residual):
            sc = nn.Conv3d(channels_in, channels_out, kernel_size=(1, 1, 1), stride=(1, 1, 2 ** n), bias=not is_batchnorm)
            downsample = nn.Sequential(sc, nn.BatchNorm3d(channels_out)) if is_batchnorm else sc
        layers = []
        if n > 0:
            layers.append(unet3dConvX(channels_in, channels_out, kernel_size=[(1, 1, 3)] * n, stride=[(1, 1, 2)] * n,
                                      padding=[(0, 0, 1)] * n, is_batchnorm=is_batchnorm, is_residual=is_residual,
                                      dropout=dropout, downsample=downsample))
            channels_in = channels_out
        layers.append(unet3dConvX(channels_in, channels_out, kernel_size=[(1, 1, final_kernelsize)], stride=[(1, 1, 1)],
                                  padding=[(0, 0, 0)], is_batchnorm=is_batchnorm, is_residual=False, dropout=dropout,
                                  downsample=None))
        return nn.Sequential(*layers)

    _bump_bn_counters = ModifiedUnet3D2D._bump_bn_counters
    _level = staticmethod(ModifiedUnet3D2D._level)
    _encode_3d = ModifiedUnet3D2D._encode_3d
    _project = ModifiedUnet3D2D._project

    def forward(self, x):
        self._bump_bn_counters()
        f3d = self._encode_3d(x)
        if self.classification:
            return f3d[4]
        proj = [self._project(f3d[l - 1], l, take_mean=not self.original) for l in range(1, 6)]
        deeper = proj[4]
        for l in (4, 3, 2, 1):
            deeper = getattr(self, f'up_concat{l}')(proj[l - 1], deeper)
        return self.final1(deeper) if self.use_1x1 else deeper
