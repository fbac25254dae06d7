"""Hybrid 3-D/2-D fusion FPN bodies, API-compatible with the reference's ``models/fpn/fusion3D2D.py``:
``ModifiedUnet3D2D`` (:10-469), ``ModifiedUnet3D2DLevel5`` (:473-581), ``unet3dConvX`` (:585-732),
``unet2dConvX`` (:735-893), 3-input ``unet3dUp2modified`` (:897-966) and ``unet3dUp2modifiedAdd`` (:969-1039).
Same constructors, attribute names, registration order (hence ``state_dict`` order and ``weight_init``
RNG order); the arithmetic runs on libfusionfpn.so.

Layout: 3-D features are logical (B, C, S, W, H) -- depth H last -- held channels-last; 2-D features
(B, C, S', W').
"""
from typing import Union

import torch
import torch.nn as nn

from config import config as global_config
from ffpn import functional as FF
from models.fpn import components as K
from models.fpn.components import SegmentationNetwork, Upsample_Custom3d_nearest  # noqa: F401  (re-exported)

ENC_KERNELS = ([(1, 3, 3), (1, 3, 3)], [(1, 3, 3), (1, 3, 3), (3, 1, 1)])      # within B-scan x2 (+ across B-scans)
ENC_PADS = ([(0, 1, 1), (0, 1, 1)], [(0, 1, 1), (0, 1, 1), (1, 0, 0)])
POOLS = ((1, 2, 2), (1, 2, 2), (2, 2, 2), (2, 2, 2))
UPFACTORS = {4: (2, 2, 1), 3: (2, 2, 1), 2: (1, 2, 1), 1: (1, 2, 1)}


class unet3dConvX(K.ConvXBase):
    '''Convolution-Block with X convolutions in 3D: [X-1 x (Conv3d, BN, ReLU)] + (Conv3d, BN), shortcut, ReLU.'''
    wrap_lone_conv = False


class unet2dConvX(K.ConvXBase):
    '''Convolution-Block with X convolutions in 2D.'''
    conv_cls = nn.Conv2d
    bn_cls = nn.BatchNorm2d
    wrap_lone_conv = False
    ndim = 4


class unet3dUp2modified(K.unet3dUp2modified):
    '''ModifiedUpsampling-Block 3D fed by BOTH encoders: cat([3-D skip, 2-D skip, upsampled deeper]).'''
    n_skips = 2
    convx_cls = unet3dConvX

    @staticmethod
    def _wrap_bias_shortcut(conv):
        return conv                                  # fusion3D2D.py:935-941: bare conv when is_batchnorm=False

    def forward(self, inputs1, inputs1_b, inputs2):
        """inputs1: projected 3-D features, inputs1_b: resized 2-D features, inputs2: deeper level."""
        return self.conv(self._cat([inputs1, inputs1_b], inputs2))


class unet3dUp2modifiedAdd(unet3dUp2modified):
    '''Additive fusion: cat([3-D skip + 2-D skip, upsampled deeper]) (reference :969-1039).'''
    n_skips = 1

    def forward(self, inputs1, inputs1_b, inputs2):
        return self.conv(self._cat([inputs1 + inputs1_b], inputs2))


class ModifiedUnet3D2D(SegmentationNetwork):
    '''Four-level hybrid fusion FPN: 3-D encoder + projective blocks, 2-D encoder, shared 2-D decoder.'''

    def __init__(self, config, interpolate: Union[str, None] = None, feature_fusion: str = 'concat'):
        """config: ConfigParser with an [architecture] section; interpolate: None | '2d' | '2d_max' (how 2-D
        features reach the en-face grid); feature_fusion: 'concat' | 'add'."""
        super().__init__(n_classes=global_config.number_of_outputs,
                         is_batchnorm=config.getboolean('architecture', 'is-batchnorm'), in_channels=1,
                         is_deconv=config.getboolean('architecture', 'is-deconv'))
        self.interpolate = interpolate
        self.feature_fusion = feature_fusion
        self._read_architecture(config)
        ch, bn = self.channels, self.is_batchnorm

        for l in range(5):                                                   # conv1..conv5
            setattr(self, f'conv{l + 1}', self._make_layer_2plus3(self.in_channels if l == 0 else ch[l - 1], ch[l],
                                                                  is_batchnorm=bn, is_residual=True,
                                                                  dropout=self.dropout[l]))
        for l, k in enumerate(POOLS):
            setattr(self, f'pool{l + 1}', K.MaxPool3d(kernel_size=k))
        for l in range(5):                                                   # zdimRed1..5: 4,3,2,1,0 halvings
            setattr(self, f'zdimRed{l + 1}', self._make_zdimReductionConvPlusFully(
                channels_in=ch[l], channels_out=ch[l], num_convreductions=4 - l, final_kernelsize=4, is_batchnorm=bn,
                is_residual=True, dropout=0.0))
        self._build_2d_encoder(levels=4)
        if self.feature_fusion == 'concat':
            self.upsampling_module = unet3dUp2modified
        elif self.feature_fusion == 'add':
            self.upsampling_module = unet3dUp2modifiedAdd
        else:
            raise ValueError('Unknown feature_fusion parameter: {}'.format(self.feature_fusion))
        for l in (4, 3, 2, 1):                                               # up_concat4..1
            setattr(self, f'up_concat{l}', self.upsampling_module(ch[l], ch[l - 1], upfactor=UPFACTORS[l],
                                                                  is_deconv=self.is_deconv, is_residual=True,
                                                                  dropout=self.dropout[9 - l], is_batchnorm=bn))
        self.final1 = K.HeadConv3d(in_channels=ch[0], out_channels=self.n_classes, kernel_size=1)

    # -- construction helpers -----------------------------------------------------------------------------
    def _read_architecture(self, config):
        self.channels = [int(i) for i in config.get('architecture', 'channels').split(',')]
        self.dropout = [float(i) for i in config.get('architecture', 'dropout').split(',')]
        self.model_name = config.get('architecture', 'architecture-name')
        assert len(self.channels) == 5
        assert len(self.dropout) == 9
        print('Channel-variable: ' + str(self.channels))

    def _build_2d_encoder(self, levels):
        ch = self.channels
        for l in range(levels):
            setattr(self, f'conv{l + 1}_2d', self._make_layer_2plus3_2d(1 if l == 0 else ch[l - 1], ch[l],
                                                                        is_batchnorm=self.is_batchnorm,
                                                                        is_residual=True, dropout=self.dropout[l]))
        for l, k in enumerate(POOLS):
            setattr(self, f'pool{l + 1}_2d', K.MaxPool2d(kernel_size=k[:2]))

    @staticmethod
    def _two_blocks(block_cls, conv_cls, bn_cls, cut, channels_in, channels_out, is_batchnorm, is_residual, dropout):
        downsample = None
        if channels_in != channels_out:
            downsample = nn.Sequential(conv_cls(channels_in, channels_out, kernel_size=1, stride=1, bias=False),
                                       bn_cls(channels_out))
        blocks = []
        for b, cin in ((0, channels_in), (1, channels_out)):
            ks = [k[:cut] for k in ENC_KERNELS[b]]
            blocks.append(block_cls(cin, channels_out, kernel_size=ks, stride=[(1,) * cut] * len(ks),
                                    padding=[p[:cut] for p in ENC_PADS[b]], is_batchnorm=is_batchnorm,
                                    is_residual=is_residual, dropout=dropout, downsample=downsample if b == 0 else None))
        return nn.Sequential(*blocks)

    def _make_layer_2plus3(self, channels_in, channels_out, is_batchnorm, is_residual, dropout):
        """Encoder level = ConvX[(1,3,3)x2] (+1x1x1 shortcut when widening) then ConvX[(1,3,3)x2,(3,1,1)]."""
        return self._two_blocks(unet3dConvX, nn.Conv3d, nn.BatchNorm3d, 3, channels_in, channels_out, is_batchnorm,
                                is_residual, dropout)

    def _make_layer_2plus3_2d(self, channels_in, channels_out, is_batchnorm, is_residual, dropout):
        """2-D twin on (S', W'): kernels (1,3),(1,3) then (1,3),(1,3),(3,1).  2-D kernels are the 3-D ones with
        the B-scan-depth axis dropped: (1,3,3)->(1,3), (3,1,1)->(3,1)."""
        return self._two_blocks(unet2dConvX, nn.Conv2d, nn.BatchNorm2d, 2, channels_in, channels_out, is_batchnorm,
                                is_residual, dropout)

    def _make_zdimReductionConvPlusFully(self, channels_in, channels_out, num_convreductions, final_kernelsize,
                                         is_batchnorm, is_residual, dropout):
        """Projective block: n x (1,1,3) stride-(1,1,2) convs with a stride-(1,1,2^n) 1x1x1 shortcut, then a
        non-residual (1,1,final_kernelsize) conv."""
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

    # -- forward stages -------------------------------------------------------------------------------------
    def _bump_bn_counters(self):
        """One multi-tensor launch for all num_batches_tracked instead of one per BatchNorm."""
        if not self.training:
            return
        blocks = [m for m in self.modules() if isinstance(m, K.ConvXBase)]
        nbt = []
        for m in blocks:
            m._nbt_managed = True
            if m.training:
                nbt += [b.num_batches_tracked for b in m._bn_modules()]
        if nbt:
            torch._foreach_add_(nbt, 1)

    @staticmethod
    def _level(seq, x, pool=None):
        """One encoder level (two ConvX blocks); the max-pool that follows is fused into the second block's
        block-end kernel (forward: pooled copy; backward: argmax routing), returning (features, pooled).  Whether the
        block computes an input gradient follows ``x.requires_grad`` (ConvXBase.forward)."""
        h = seq[0](x)
        if pool is None:
            return seq[1](h), None
        return seq[1](h, pool=pool)

    def _encode_2d(self, slo, levels):
        feats, x = [], FF.pack_image2d(slo)
        for l in range(1, levels + 1):
            f, x = self._level(getattr(self, f'conv{l}_2d'), x, getattr(self, f'pool{l}_2d') if l < levels else None)
            feats.append(f)
        return feats

    def _encode_3d(self, oct, mark_bucket=False):
        feats, x = [], FF.pack_oct(oct)
        for l in range(1, 6):
            f, x = self._level(getattr(self, f'conv{l}'), x, getattr(self, f'pool{l}') if l < 5 else None)
            if l == 2 and mark_bucket:
                x = FF.bucket_marker(x)
            feats.append(f)
        return feats

    def _project(self, feat, l, take_mean=True, slot=None):
        """zdimRed<l> then torch.mean(dim=4, keepdim=True); the mean is fused into the last block's tail and written
        straight into ``slot`` (a concat buffer's channel slice) when one is given."""
        seq = getattr(self, f'zdimRed{l}')
        x = feat
        for blk in list(seq)[:-1]:
            x = blk(x)
        return seq[-1](x, tail='mean' if take_mean else 'relu', out=slot if take_mean else None)

    def _resize_2d(self, f2d, size, slot=None):
        """conv_2d[:,:,:,:,None] then None | trilinear | adaptive max to the en-face grid ``size``."""
        return FF.resize2d(f2d, tuple(size), self.interpolate, slot)

    def _cat_slots(self, oct, levels2d):
        """Concat buffers of the decoder, allocated up front: level l holds [projected conv_l | resized conv_l_2d | upsampled
        deeper level] (fusion3D2D.py:956-966), level 5 of the Level5 body [conv5 | conv5_2d] (:572).  Their producers write
        into the slots, so no ``cat`` copy and no stand-alone projected tensor exist.  -> {level: (slot3d, slot2d)} and the
        en-face size per level; None for feature_fusion='add' (the sum is a new tensor)."""
        B, _, S, W, _H = oct.shape
        ch = self.channels
        sizes = [(S, W), (S, W // 2), (S, W // 4), (S // 2, W // 8), (S // 4, W // 16)]
        if self.feature_fusion != 'concat' or not oct.is_cuda:
            return None, sizes
        dt = FF.get_compute_dtype()
        slots = {}
        for l in (1, 2, 3, 4):
            deeper = ch[l] * (2 if (l == 4 and levels2d == 5) else 1)
            buf = torch.empty((B, sizes[l - 1][0], sizes[l - 1][1], 1, 2 * ch[l - 1] + deeper), dtype=dt, device=oct.device)
            slots[l] = (FF.CatSlot(buf, 0, ch[l - 1]), FF.CatSlot(buf, ch[l - 1], ch[l - 1]))
        if levels2d == 5:
            buf = torch.empty((B, sizes[4][0], sizes[4][1], 1, 2 * ch[4]), dtype=dt, device=oct.device)
            slots[5] = (FF.CatSlot(buf, 0, ch[4]), FF.CatSlot(buf, ch[4], ch[4]))
        return slots, sizes

    def _forward_serial(self, oct, slo, levels2d):
        slots, sizes = self._cat_slots(oct, levels2d)
        s3 = (lambda l: slots[l][0] if slots and l in slots else None)
        s2 = (lambda l: slots[l][1] if slots and l in slots else None)
        f3d = self._encode_3d(oct, mark_bucket=True)
        f2d = self._encode_2d(slo, levels2d)                # after the 3-D encoder: see _forward_branches
        proj = [self._project(f3d[l - 1], l, slot=s3(l)) for l in range(1, 6)]
        r2d = [self._resize_2d(f2d[l], sizes[l], s2(l + 1)) for l in range(levels2d)]
        deeper = FF.cat(proj[4], r2d[4]) if levels2d == 5 else proj[4]
        for l in (4, 3, 2, 1):
            deeper = getattr(self, f'up_concat{l}')(proj[l - 1], r2d[l - 1], deeper)
        return self.final1(deeper)

    def _forward_branches(self, oct, slo, levels2d):
        """The forward DAG on side streams: the 2-D encoder (+ resizes) on one, the projective block of each level 1-4
        on its own, the 3-D encoder -> level-5 projection -> decoder chain on the caller's stream.  Each up block waits
        only for the two skips it reads, so the large level-1 projection overlaps the deep (latency-bound) levels."""
        dev = oct.device
        slots, sizes = self._cat_slots(oct, levels2d)       # before any fork: every stream below is ordered after the allocation
        s3 = (lambda l: slots[l][0] if slots and l in slots else None)
        s2 = (lambda l: slots[l][1] if slots and l in slots else None)
        start = torch.cuda.Event()
        start.record()
        proj, x = [None] * 5, FF.pack_oct(oct)
        for l in range(1, 6):
            f, x = self._level(getattr(self, f'conv{l}'), x, getattr(self, f'pool{l}') if l < 5 else None)
            if l == 2:
                x = FF.bucket_marker(x)                     # backward: everything behind this point has issued its gradients
            if l < 5:
                sp = FF.fork(FF.side_stream(dev, l), f)
                with torch.cuda.stream(sp):
                    proj[l - 1] = self._project(f, l, slot=s3(l))
            else:
                proj[4] = self._project(f, 5, slot=s3(5))
            assert tuple(proj[l - 1].shape[2:4]) == sizes[l - 1], (tuple(proj[l - 1].shape), sizes[l - 1])
        # The 2-D encoder is ISSUED after the 3-D one but starts at `start` on its own stream: on the device (and in a captured
        # graph) it runs concurrently from the beginning, while autograd -- which replays nodes newest first -- issues its
        # backward before the 3-D encoder's, so that its gradients are complete when the bucket marker fires.
        s2d = FF.fork_from(FF.side_stream(dev, 0), start, slo)
        with torch.cuda.stream(s2d):
            f2d = self._encode_2d(slo, levels2d)
            r2d = [self._resize_2d(f2d[l], sizes[l], s2(l + 1)) for l in range(levels2d)]
        deeper = proj[4]
        if levels2d == 5:
            FF.join(s2d, r2d[4])
            deeper = FF.cat(proj[4], r2d[4])
        for l in (4, 3, 2, 1):
            FF.join(FF.side_stream(dev, l), proj[l - 1])
            if l == 4 and levels2d < 5:
                FF.join(s2d)
            r2d[l - 1].record_stream(torch.cuda.current_stream())
            deeper = getattr(self, f'up_concat{l}')(proj[l - 1], r2d[l - 1], deeper)
        return self.final1(deeper)

    def forward(self, oct, slo):
        self._bump_bn_counters()
        if oct.is_cuda and FF.streams_enabled():
            return self._forward_branches(oct, slo, 4)
        return self._forward_serial(oct, slo, 4)


class ModifiedUnet3D2DLevel5(ModifiedUnet3D2D):
    '''Five-level variant: a fifth 2-D encoder level whose features are concatenated with the projected conv5
    before the first up block (reference :473-581).'''

    def __init__(self, config, interpolate: Union[str, None] = None, feature_fusion: str = 'concat'):
        super().__init__(config, interpolate, feature_fusion)
        ch = self.channels
        self.conv5_2d = self._make_layer_2plus3_2d(ch[3], ch[4], is_batchnorm=self.is_batchnorm, is_residual=True,
                                                   dropout=self.dropout[4])
        self.up_concat4 = self.upsampling_module(ch[4] * 2, ch[3], upfactor=(2, 2, 1), is_deconv=self.is_deconv,
                                                 is_residual=True, dropout=self.dropout[5],
                                                 is_batchnorm=self.is_batchnorm)

    def forward(self, oct, slo):
        self._bump_bn_counters()
        if oct.is_cuda and FF.streams_enabled():
            return self._forward_branches(oct, slo, 5)
        return self._forward_serial(oct, slo, 5)
